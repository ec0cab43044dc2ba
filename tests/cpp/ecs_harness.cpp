// ecs_harness.cpp -- SURVEY 8f F3: the drop-in under the engine's REAL ECS code, end to end.
// One source, two builds (oracle/Makefile `harness`), both linking the reference's own ECSwrapper.cpp, EntitiesHandler.cpp,
// ComponentBaseClass.cpp and Geometry sources, compiled where they lie under /root/reference:
//   ecs_harness_ref     + the reference's CollisionDetection.cpp, SweepAndPrune.cpp, OBBtreesCollision.cpp, CreateUncollideRays.cpp, ShootUncollideRays.cpp
//   ecs_harness_dropin  -DIMRCD_DROP_IN: csrc/host/CollisionDetection_drop_in.hpp + libimrcd.so (GPU) instead of those five files
// Each builds the same scene (procedural meshes through Triangle::CreateTriangleList and OBBtree::OBBtree, entities in the engine's
// EntitiesHandler with parents inside and across instances), registers a component that overrides CollisionCallback exactly like a game
// component would (ComponentBaseClass.h:25, SnakePlayerComp.cpp:65-96), runs the engine's per-frame sequence Reset /
// AddCollisionDetectionEntry / ExecuteCollisionDetection (ModelCollisionComp.cpp:16-37) and prints what the component received.
// TEST INFRASTRUCTURE: built only where the reference checkout exists; the binaries live under oracle/_ref/.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "ECS/ECSwrapper.h"
#include "ECS/ECStypes.h"
#include "Geometry/OBBtree.h"
#include "Geometry/Triangle.h"
#include "glm/gtc/matrix_transform.hpp"
#include "glm/gtc/quaternion.hpp"
#ifdef IMRCD_DROP_IN
#include "CollisionDetection_drop_in.hpp"
#else
#include "CollisionDetection/CollisionDetection.h"
#endif

struct StubExported : ExportedFunctions {
    void BindCameraEntity(Entity) const override {}
    size_t GetSphereMeshIndex() const override { return 0; }
    size_t GetCylinderMeshIndex() const override { return 0; }
};

struct Received { Entity receiver, family, other; glm::vec3 delta; };

class RecorderComp : public ComponentBaseClass {
public:
    explicit RecorderComp(ECSwrapper* e) : ComponentBaseClass(e) {}
    void CollisionCallback(const std::vector<std::pair<Entity, std::vector<CollisionCallbackData>>>& in) override {
        for (const auto& kv : in)
            for (const auto& c : kv.second) got.push_back({kv.first, c.familyEntity, c.collideWithEntity, c.deltaVector});
    }
    componentID GetComponentID() const override { return 1000; }
    std::string GetComponentName() const override { return "Recorder"; }
    std::vector<Received> got;
};

static OBBtree make_torus(int nu, int nv, float R, float r) {
    std::vector<glm::vec3> pts, nrm; std::vector<uint32_t> idx;
    for (int i = 0; i < nu; ++i)
        for (int j = 0; j < nv; ++j) {
            const float a = 6.2831853f * i / nu, b = 6.2831853f * j / nv;
            const glm::vec3 n(std::cos(a) * std::cos(b), std::sin(a) * std::cos(b), std::sin(b));
            pts.emplace_back(glm::vec3(R * std::cos(a), R * std::sin(a), 0.f) + r * n); nrm.emplace_back(n);
        }
    for (int i = 0; i < nu; ++i)
        for (int j = 0; j < nv; ++j) {
            const uint32_t a = i * nv + j, b = ((i + 1) % nu) * nv + j, c = ((i + 1) % nu) * nv + (j + 1) % nv, d = i * nv + (j + 1) % nv;
            idx.insert(idx.end(), {a, b, c, a, c, d});
        }
    return OBBtree(Triangle::CreateTriangleList(pts, nrm, idx, glTFmode::triangles));
}

int main(int argc, char** argv) {
    const int n_bodies = argc > 1 ? std::atoi(argv[1]) : 48;
    const bool moved = argc > 2 ? std::atoi(argv[2]) != 0 : true;
    StubExported exported;
    ECSwrapper ecs(&exported);
    RecorderComp recorder(&ecs);
    ecs.AddComponent(&recorder);
    EntitiesHandler* eh = ecs.GetEntitiesHandler();

    // a fab of three entities: 0 (instance root) <- 1 <- 2; the collision entry lives on entity 2
    FabInfo fab; fab.fabName = "body"; fab.fabIndex = 0; fab.size = 3; fab.entitiesParents = {Entity(-1), 0, 1};
    FabInfo group; group.fabName = "group"; group.fabIndex = 1; group.size = 1; group.entitiesParents = {Entity(-1)};
    std::vector<Entity> group_root;
    for (int g = 0; g < 4; ++g) group_root.push_back(eh->AddInstanceEntities(&group, 0)->entityOffset);

    OBBtree torus = make_torus(28, 14, 1.0f, 0.35f);
    OBBtree small = make_torus(16, 8, 0.6f, 0.25f);
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    std::vector<CollisionDetectionEntry> entries;
    for (int k = 0; k < n_bodies; ++k) {
        // bodies of a group hang under the group's entity: pairs inside a group share ancestors (CollisionDetection.cpp:109-125)
        const Entity parent = (k % 3 == 0) ? Entity(0) : group_root[k % 4];
        const InstanceInfo* inst = eh->AddInstanceEntities(&fab, parent);
        glm::mat4 m = glm::translate(glm::mat4(1.f), glm::vec3(U(rng), U(rng), U(rng)) * 2.6f);
        m = m * glm::mat4_cast(glm::normalize(glm::quat(U(rng), U(rng), U(rng), U(rng))));
        m = glm::scale(m, glm::vec3(0.8f + 0.3f * U(rng)));
        CollisionDetectionEntry e;
        e.currentGlobalMatrix = m;
        e.previousGlobalMatrix = (moved && k % 5 != 0) ? glm::translate(glm::mat4(1.f), glm::vec3(U(rng), U(rng), U(rng)) * 0.03f) * m : m;
        e.OBBtree_ptr = (k % 2) ? &torus : &small;
        e.shouldCallback = (k % 7 != 3);
        e.entity = Entity(inst->entityOffset + 2);
        entries.push_back(e);
    }
    eh->AdditionsCompleted();

    CollisionDetection cd(&ecs);
    for (int frame = 0; frame < 2; ++frame) {          // two frames: the second must reproduce the first (Reset clears the entries)
        recorder.got.clear();
        cd.Reset();
        for (const auto& e : entries) cd.AddCollisionDetectionEntry(e);
        cd.ExecuteCollisionDetection();
    }
    std::vector<Received>& got = recorder.got;
    std::sort(got.begin(), got.end(), [](const Received& a, const Received& b) {
        return std::tie(a.receiver, a.family, a.other) < std::tie(b.receiver, b.family, b.other); });
    std::printf("callbacks %zu\n", got.size());
    for (const auto& r : got) std::printf("%u %u %u %.9g %.9g %.9g\n", unsigned(r.receiver), unsigned(r.family), unsigned(r.other), r.delta.x, r.delta.y, r.delta.z);
    return 0;
}
