// imrcd_host.hpp -- C++ host side above the C ABI (include/imrcd.h): the reference's CollisionDetection interface
// (IMR/include/CollisionDetection/CollisionDetection.h:11-35) with the engine's types factored out, so that it compiles and
// is testable without the engine tree.  CollisionDetection_drop_in.hpp instantiates it with the engine's own types.
//
//   Reset()                          IMR/src/CollisionDetection/CollisionDetection.cpp:28
//   AddCollisionDetectionEntry(e)    :33   (here the entry is written once, straight into the context's pinned staging)
//   ExecuteCollisionDetection()      :38-129: < 2 entries is a no-op (:40); colliding pairs come back from the GPU; both
//                                    directions of CollisionCallbackData are produced (:72-78); every ancestor of an
//                                    entity that is not shared with the other entity's chain receives it (:106-125);
//                                    MakeCallbacks hands one vector to every component (:131-141).
// Header-only, C++17, no exceptions thrown across the ABI (errors surface as std::runtime_error from this layer only).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#include "../../../include/imrcd.h"

namespace imrcd {

template <class EntityT>
struct CallbackData {                       // CollisionCallbackData, IMR/include/ECS/ECStypes.h:158-163
    EntityT familyEntity;
    EntityT collideWithEntity;
    float deltaVector[3] = {0.f, 0.f, 0.f};
};

// EcsPolicy must provide:
//   std::vector<EntityT> GetEntityAncestors(EntityT) const;                      (EntitiesHandler.cpp:186-204, root first)
//   void MakeCallbacks(const std::vector<std::pair<EntityT, std::vector<CallbackData<EntityT>>>>&);   (CollisionDetection.cpp:131-141)
template <class EntityT, class EcsPolicy>
class CollisionDetectionT {
public:
    using Callback = CallbackData<EntityT>;

    explicit CollisionDetectionT(EcsPolicy* ecs, int device = 0, void* cuda_stream = nullptr) : ecs_(ecs) {
        const int rc = imrcd_create(device, cuda_stream, &ctx_);
        if (rc != IMRCD_OK) throw std::runtime_error("imrcd_create failed (" + std::to_string(rc) + "): there is no CPU fallback");
    }
    // several GPUs of this process (imrcd_group_*): meshes replicated, every frame sharded by entity over the devices, the colliding
    // pairs merged by the library's end-of-frame NCCL all-gather; the interface stays the engine's
    CollisionDetectionT(EcsPolicy* ecs, const std::vector<int>& devices) : ecs_(ecs) {
        const int rc = imrcd_group_create(devices.data(), static_cast<uint32_t>(devices.size()), &group_);
        if (rc != IMRCD_OK) throw std::runtime_error("imrcd_group_create failed (" + std::to_string(rc) + "): there is no CPU fallback");
        ctx_ = imrcd_group_ctx(group_, 0);
    }
    ~CollisionDetectionT() { if (group_) imrcd_group_destroy(group_); else imrcd_destroy(ctx_); }
    CollisionDetectionT(const CollisionDetectionT&) = delete;
    CollisionDetectionT& operator=(const CollisionDetectionT&) = delete;

    imrcd_ctx* context() const { return ctx_; }

    // replaces OBBtree::OBBtree(std::vector<Triangle>&&) (IMR/src/Geometry/OBBtree.cpp:321): n_tri * 9 floats each, ids n_tri * 3
    uint32_t CreateOBBtree(const float* positions, const float* normals, const uint32_t* vertex_ids, uint64_t n_tri,
                           uint32_t build_mode = IMRCD_BUILD_MORTON) {
        uint32_t id = 0;
        if (group_) gcheck(imrcd_group_mesh_create(group_, positions, normals, vertex_ids, n_tri, build_mode, &id));
        else check(imrcd_mesh_create(ctx_, positions, normals, vertex_ids, n_tri, build_mode, &id));
        return id;
    }

    // the engine's own route, PrimitivesOfMeshes::StartRecordOBBtree / one PrimitiveOBBtreeData per primitive / GetOBBtreeAndReset
    // (IMR/src/Graphics/Meshes/PrimitivesOfMeshes.cpp:637-671,835-863): Triangle::CreateTriangleList runs on the device.
    // points / normals: n_points * stride floats (stride 4 = the engine's vec4 arrays), indices may be null, draw_mode = glTFmode.
    void StartRecordOBBtree() { for (uint32_t i = 0; i < n_ctx(); ++i) check_on(ctx_at(i), imrcd_mesh_begin(ctx_at(i))); }
    void RecordPrimitive(const float* points, uint64_t n_points, uint32_t stride, const float* normals, const uint32_t* indices, uint64_t n_indices, uint32_t draw_mode) {
        for (uint32_t i = 0; i < n_ctx(); ++i) check_on(ctx_at(i), imrcd_mesh_add_primitive(ctx_at(i), points, n_points, stride, normals, indices, n_indices, draw_mode));
    }
    uint32_t GetOBBtreeAndReset(uint32_t build_mode = IMRCD_BUILD_MORTON) {
        uint32_t id = 0;
        for (uint32_t i = 0; i < n_ctx(); ++i) check_on(ctx_at(i), imrcd_mesh_end(ctx_at(i), build_mode, &id));      // the same id on every GPU
        return id;
    }

    // A whole .gltf / .glb: one tree per glTF mesh, ids in file order (what MeshesOfNodes::AddMeshesOfModel keeps as
    // MeshInfo::boundBoxTree, IMR/src/Graphics/Meshes/MeshesOfNodes.cpp:34-53); the file is read by the library, no tinygltf model needed.
    std::vector<uint32_t> LoadMeshesOfModel(const std::string& path, uint32_t build_mode = IMRCD_BUILD_MORTON) {
        uint32_t n = 0;
        check(imrcd_gltf_load(ctx_, path.c_str(), build_mode, nullptr, 0, &n));
        std::vector<uint32_t> ids(n);
        if (n && group_) gcheck(imrcd_group_gltf_load(group_, path.c_str(), build_mode, ids.data(), n, &n));
        else if (n) check(imrcd_gltf_load(ctx_, path.c_str(), build_mode, ids.data(), n, &n));
        return ids;
    }

    void Reset() {                                                       // CollisionDetection.cpp:28
        if (group_) gcheck(imrcd_group_frame_reset(group_)); else check(imrcd_frame_reset(ctx_));
        n_ = 0; mapped_ = 0; pending_ = 0; any_previous_ = false;
    }

    // current / previous: glm::mat4 layout, 16 floats column-major (ECStypes.h:149-156)
    void AddCollisionDetectionEntry(const float* current, const float* previous, uint32_t mesh_id, bool should_callback, EntityT entity) {
        if (pending_ == mapped_) flush_and_map();
        std::memcpy(cur_ + 16 * pending_, current, 64);
        const bool moved = previous && std::memcmp(previous, current, 64) != 0;
        if (moved || any_previous_) {
            if (!any_previous_) { std::memcpy(prev_, cur_, 64 * pending_); any_previous_ = true; }   // earlier entries of this batch did not move
            std::memcpy(prev_ + 16 * pending_, previous ? previous : current, 64);
        }
        mesh_[pending_] = mesh_id; cb_[pending_] = should_callback ? 1 : 0; ent_[pending_] = static_cast<uint32_t>(entity);
        ++pending_; ++n_;
    }

    void ExecuteCollisionDetection() {                                   // CollisionDetection.cpp:38-129
        if (n_ < 2) return;                                              // :40
        commit();
        if (group_) gcheck(imrcd_group_frame_execute(group_)); else check(imrcd_frame_execute(ctx_));
        const imrcd_entity_pair* pairs = nullptr; uint64_t n_pairs = 0;
        check(imrcd_frame_results(ctx_, &pairs, &n_pairs, nullptr, nullptr));          // with a group: the merged pairs of all GPUs
        if (!ecs_) return;
        std::unordered_map<EntityT, std::vector<Callback>> to_make;
        for (uint64_t k = 0; k < n_pairs; ++k) {
            const imrcd_entity_pair& p = pairs[k];
            Callback first, second;                                      // :72-78
            first.familyEntity = static_cast<EntityT>(p.entity_first);  first.collideWithEntity = static_cast<EntityT>(p.entity_second);
            second.familyEntity = static_cast<EntityT>(p.entity_second); second.collideWithEntity = static_cast<EntityT>(p.entity_first);
            std::memcpy(first.deltaVector, p.delta_first, 12); std::memcpy(second.deltaVector, p.delta_second, 12);
            const std::vector<EntityT> fa = ecs_->GetEntityAncestors(first.familyEntity), sa = ecs_->GetEntityAncestors(second.familyEntity);
            for (size_t i = 0; i != fa.size(); ++i)                      // :109-116
                if (i >= sa.size() || fa[i] != sa[i]) to_make[fa[i]].emplace_back(first);
            for (size_t i = 0; i != sa.size(); ++i)                      // :118-125
                if (i >= fa.size() || fa[i] != sa[i]) to_make[sa[i]].emplace_back(second);
        }
        std::vector<std::pair<EntityT, std::vector<Callback>>> vec(std::make_move_iterator(to_make.begin()), std::make_move_iterator(to_make.end()));
        ecs_->MakeCallbacks(vec);                                        // :131-141
    }

    // results of the last frame, valid until the next Reset()
    void Results(const imrcd_entity_pair** pairs, uint64_t* n_pairs, const imrcd_tri_hit** hits = nullptr, uint64_t* n_hits = nullptr) {
        check(imrcd_frame_results(ctx_, pairs, n_pairs, hits, n_hits));
    }
    imrcd_frame_stats Stats() { imrcd_frame_stats s; check(imrcd_frame_get_stats(ctx_, &s)); return s; }

private:
    static constexpr uint64_t kBatch = 4096;       // entries mapped at a time (each batch's DMA starts at its commit)

    void check(int rc) { if (rc != IMRCD_OK) throw std::runtime_error(std::string("imrcd: ") + imrcd_last_error(ctx_)); }
    void check_on(imrcd_ctx* c, int rc) { if (rc != IMRCD_OK) throw std::runtime_error(std::string("imrcd: ") + imrcd_last_error(c)); }
    void gcheck(int rc) { if (rc != IMRCD_OK) throw std::runtime_error(std::string("imrcd: ") + imrcd_group_last_error(group_)); }
    uint32_t n_ctx() const { return group_ ? imrcd_group_size(group_) : 1u; }
    imrcd_ctx* ctx_at(uint32_t i) const { return group_ ? imrcd_group_ctx(group_, i) : ctx_; }
    void commit() {
        if (pending_ && group_)        // every GPU of the group is handed the batch and keeps its share
            gcheck(imrcd_group_frame_add_entries(group_, pending_, cur_, any_previous_ ? prev_ : nullptr, mesh_, cb_, ent_));
        else if (pending_) check(imrcd_frame_commit_entries(ctx_, pending_, any_previous_ ? 1 : 0));
        pending_ = 0; mapped_ = 0; any_previous_ = false;
    }
    void flush_and_map() {
        commit();
        if (group_) {
            g_cur_.resize(16 * kBatch); g_prev_.resize(16 * kBatch); g_mesh_.resize(kBatch); g_cb_.resize(kBatch); g_ent_.resize(kBatch);
            cur_ = g_cur_.data(); prev_ = g_prev_.data(); mesh_ = g_mesh_.data(); cb_ = g_cb_.data(); ent_ = g_ent_.data();
        } else check(imrcd_frame_map_entries(ctx_, kBatch, &cur_, &prev_, &mesh_, &cb_, &ent_));
        mapped_ = kBatch;
    }

    EcsPolicy* ecs_ = nullptr;
    imrcd_ctx* ctx_ = nullptr;
    imrcd_group* group_ = nullptr;
    std::vector<float> g_cur_, g_prev_; std::vector<uint32_t> g_mesh_, g_ent_; std::vector<uint8_t> g_cb_;      // group mode: the batch being written
    uint64_t n_ = 0, mapped_ = 0, pending_ = 0;
    bool any_previous_ = false;
    float* cur_ = nullptr; float* prev_ = nullptr; uint32_t* mesh_ = nullptr; uint8_t* cb_ = nullptr; uint32_t* ent_ = nullptr;
};

}  // namespace imrcd
