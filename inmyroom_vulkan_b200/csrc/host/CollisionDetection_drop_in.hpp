// CollisionDetection_drop_in.hpp -- the class a maintainer drops into the engine in place of
// IMR/include/CollisionDetection/CollisionDetection.h + IMR/src/CollisionDetection/CollisionDetection.cpp.
// Same name, same constructor and the same three calls, so Engine.cpp:50 and ModelCollisionComp.cpp:16-37 compile unchanged.
// Compiles only inside the engine tree (it includes the engine's ECS and Geometry headers); see INTEGRATION.md.
//
// Trees: the engine keeps building its OBBtree objects at load time (MeshesOfNodes.cpp:38,53).  The first time an entry
// refers to a tree (CollisionDetectionEntry::OBBtree_ptr, ECStypes.h:153) the adapter flattens it through the public
// OBBtreeTraveler (OBBtree.h:78-101) and uploads it with imrcd_mesh_import_tree: the GPU then traverses the engine's OWN tree,
// which is the configuration the parity tests prove bit-identical.  Define IMRCD_DROP_IN_REBUILD to hand the triangles to
// imrcd_mesh_create instead (GPU Morton build, tighter boxes, tree-independent parity).
#pragma once
#include <unordered_map>
#include <vector>

#include "ECS/ECStypes.h"
#include "ECS/ECSwrapper.h"
#include "Geometry/OBBtree.h"

#include "imrcd_host.hpp"

class CollisionDetection
{
    struct EcsPolicy {
        ECSwrapper* ecs;
        std::vector<Entity> GetEntityAncestors(Entity e) const { return ecs->GetEntitiesHandler()->GetEntityAncestors(e); }      // CollisionDetection.cpp:106-107
        void MakeCallbacks(const std::vector<std::pair<Entity, std::vector<imrcd::CallbackData<Entity>>>>& in) {
            // imrcd::CallbackData<Entity> and CollisionCallbackData have the same members; convert to the engine's type (glm::vec3)
            std::vector<std::pair<Entity, std::vector<CollisionCallbackData>>> out;
            out.reserve(in.size());
            for (const auto& kv : in) {
                std::vector<CollisionCallbackData> v;
                v.reserve(kv.second.size());
                for (const auto& c : kv.second) {
                    CollisionCallbackData d;
                    d.familyEntity = c.familyEntity; d.collideWithEntity = c.collideWithEntity;
                    d.deltaVector = glm::vec3(c.deltaVector[0], c.deltaVector[1], c.deltaVector[2]);
                    v.emplace_back(d);
                }
                out.emplace_back(kv.first, std::move(v));
            }
#ifdef IMRCD_DROP_IN_CALLBACK_SINK
            IMRCD_DROP_IN_CALLBACK_SINK(out);                   // tests: hand the callbacks to a checker instead of the components
#else
            for (const auto& id_component : ecs->GetComponentIDtoComponentBaseClassMap())                                       // CollisionDetection.cpp:136-140
                if (id_component.second != nullptr) id_component.second->CollisionCallback(out);
#endif
        }
    };

public:
    CollisionDetection(ECSwrapper* in_ECSwrapper_ptr) : policy{in_ECSwrapper_ptr}, impl(&policy) {}

    void Reset() { impl.Reset(); }

    void AddCollisionDetectionEntry(const CollisionDetectionEntry in_collisionDetectionEntry)
    {
        impl.AddCollisionDetectionEntry(&in_collisionDetectionEntry.currentGlobalMatrix[0][0], &in_collisionDetectionEntry.previousGlobalMatrix[0][0],
                                        MeshOf(in_collisionDetectionEntry.OBBtree_ptr), in_collisionDetectionEntry.shouldCallback, in_collisionDetectionEntry.entity);
    }

    void ExecuteCollisionDetection() { impl.ExecuteCollisionDetection(); }

    // Optional: trees that were made on the device (imrcd_gltf_load / LoadMeshesOfModel, imrcd_mesh_end) instead of flattened from the
    // engine's OBBtree objects.  After UseDeviceTree(&meshInfo.boundBoxTree, id) entries that point at that engine tree use mesh `id`.
    imrcd::CollisionDetectionT<Entity, EcsPolicy>& Adapter() { return impl; }
    void UseDeviceTree(const OBBtree* engine_tree, uint32_t mesh_id) { meshes[engine_tree] = mesh_id; }

private:
    struct Flat { std::vector<float> boxes; std::vector<int32_t> left, right; std::vector<uint32_t> off, cnt; size_t n_tri = 0; };
    static int32_t Walk(const OBBtree::OBBtreeTraveler& t, Flat& f)
    {
        const int32_t me = int32_t(f.left.size());
        const OBB b = t.GetOBB();
        const glm::vec3 c = b.GetCenter(), u = b.GetSideDirectionU(), v = b.GetSideDirectionV(), w = b.GetSideDirectionW();
        const float box[12] = {c.x, c.y, c.z, u.x, u.y, u.z, v.x, v.y, v.z, w.x, w.y, w.z};
        f.boxes.insert(f.boxes.end(), box, box + 12);
        f.left.push_back(-1); f.right.push_back(-1); f.off.push_back(0); f.cnt.push_back(0);
        if (t.IsLeaf()) {
            f.off[me] = uint32_t(t.GetTrianglesOffset()); f.cnt[me] = uint32_t(t.GetTrianglesCount());
            f.n_tri = std::max(f.n_tri, t.GetTrianglesOffset() + t.GetTrianglesCount());
        } else {
            const int32_t l = Walk(t.GetLeftChildTraveler(), f);
            const int32_t r = Walk(t.GetRightChildTraveler(), f);
            f.left[me] = l; f.right[me] = r;
        }
        return me;
    }
    uint32_t MeshOf(const OBBtree* tree)
    {
        const auto it = meshes.find(tree);
        if (it != meshes.end()) return it->second;
        Flat f;
        Walk(tree->GetRootTraveler(), f);
        std::vector<float> pos(9 * f.n_tri), nrm(9 * f.n_tri);
        std::vector<uint32_t> vid(3 * f.n_tri);
        for (size_t i = 0; i != f.n_tri; ++i) {
            const TrianglePosition p = tree->GetTrianglePosition(i); const TriangleNormal n = tree->GetTriangleNormal(i); const TriangleIndices x = tree->GetTriangleIndices(i);
            for (int k = 0; k != 3; ++k) {
                const glm::vec3 q = p.GetP(k), m = n.GetN(k);
                pos[9 * i + 3 * k] = q.x; pos[9 * i + 3 * k + 1] = q.y; pos[9 * i + 3 * k + 2] = q.z;
                nrm[9 * i + 3 * k] = m.x; nrm[9 * i + 3 * k + 1] = m.y; nrm[9 * i + 3 * k + 2] = m.z;
                vid[3 * i + k] = x.GetI(k);
            }
        }
        uint32_t id = 0;
#ifdef IMRCD_DROP_IN_REBUILD
        id = impl.CreateOBBtree(pos.data(), nrm.data(), vid.data(), f.n_tri);
#else
        if (imrcd_mesh_import_tree(impl.context(), f.left.size(), f.boxes.data(), f.left.data(), f.right.data(), f.off.data(), f.cnt.data(),
                                   f.n_tri, pos.data(), nrm.data(), vid.data(), nullptr, &id) != IMRCD_OK)
            throw std::runtime_error(imrcd_last_error(impl.context()));
#endif
        meshes.emplace(tree, id);
        return id;
    }

    EcsPolicy policy;
    imrcd::CollisionDetectionT<Entity, EcsPolicy> impl;
    std::unordered_map<const OBBtree*, uint32_t> meshes;
};
