// GPU test of the C++ host adapter (csrc/host/imrcd_host.hpp) over the C ABI: the reference's Reset / AddCollisionDetectionEntry /
// ExecuteCollisionDetection call pattern and its callback fan-out rule (CollisionDetection.cpp:106-141), with a toy ECS.
// Built and run by tests/test_gpu_host_adapter.py.  Exit code 0 = pass.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include "imrcd_host.hpp"

using Entity = uint16_t;                                    // IMR/include/ECS/ECStypes.h:18
struct ToyEcs {
    std::map<Entity, Entity> parent;                        // child -> parent (0 = none)
    std::vector<std::pair<Entity, std::vector<imrcd::CallbackData<Entity>>>> received;
    int calls = 0;
    std::vector<Entity> GetEntityAncestors(Entity e) const {   // root first, the entity itself last (EntitiesHandler.cpp:186-204)
        std::vector<Entity> chain;
        for (Entity x = e; x != 0; x = parent.count(x) ? parent.at(x) : 0) chain.insert(chain.begin(), x);
        return chain;
    }
    void MakeCallbacks(const std::vector<std::pair<Entity, std::vector<imrcd::CallbackData<Entity>>>>& v) { received = v; ++calls; }
};

static void box_mesh(std::vector<float>& pos, std::vector<uint32_t>& vid) {
    const float v[8][3] = {{-1,-1,-1},{1,-1,-1},{1,1,-1},{-1,1,-1},{-1,-1,1},{1,-1,1},{1,1,1},{-1,1,1}};
    const int f[12][3] = {{0,2,1},{0,3,2},{4,5,6},{4,6,7},{0,1,5},{0,5,4},{2,3,7},{2,7,6},{1,2,6},{1,6,5},{0,4,7},{0,7,3}};
    for (auto& t : f) for (int k = 0; k < 3; ++k) { for (int c = 0; c < 3; ++c) pos.push_back(v[t[k]][c]); vid.push_back(uint32_t(t[k])); }
}
static void translation(float* m, float x, float y, float z, float rot) {
    const float c = cosf(rot), s = sinf(rot);
    const float r[16] = {c, s, 0, 0, -s, c, 0, 0, 0, 0, 1, 0, x, y, z, 1};     // column-major, rotation about z
    memcpy(m, r, 64);
}
#define EXPECT(cond) do { if (!(cond)) { printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

using CD = imrcd::CollisionDetectionT<Entity, ToyEcs>;
static int run_all(CD& cd, ToyEcs& ecs, int argc, char** argv);

int main(int argc, char** argv) {
    // two separate families 1 -> 2 and 3 -> 4, and two siblings 5, 6 under the common parent 7; 8 is far away
    {
        ToyEcs ecs; ecs.parent = {{2, 1}, {4, 3}, {5, 7}, {6, 7}};
        CD cd(&ecs, 0);
        if (run_all(cd, ecs, argc, argv)) return 1;
    }
    // the same through N GPUs of this process (argv[2] = N): frames sharded by entity, merged by the library's NCCL all-gather
    const int n_gpus = argc > 2 ? atoi(argv[2]) : 1;
    if (n_gpus > 1) {
        ToyEcs ecs; ecs.parent = {{2, 1}, {4, 3}, {5, 7}, {6, 7}};
        std::vector<int> devs; for (int d = 0; d < n_gpus; ++d) devs.push_back(d);
        CD cd(&ecs, devs);
        if (run_all(cd, ecs, argc, argv)) return 1;
        printf("host adapter ok on %d gpus\n", n_gpus);
    }
    printf("host adapter ok\n");
    return 0;
}

static int run_all(CD& cd, ToyEcs& ecs, int argc, char** argv) {
    std::vector<float> pos; std::vector<uint32_t> vid;
    box_mesh(pos, vid);
    const uint32_t mesh = cd.CreateOBBtree(pos.data(), nullptr, vid.data(), 12);

    // < 2 entries: no-op (CollisionDetection.cpp:40)
    cd.Reset();
    float m[16]; translation(m, 0, 0, 0, 0);
    cd.AddCollisionDetectionEntry(m, m, mesh, true, 2);
    cd.ExecuteCollisionDetection();
    EXPECT(ecs.calls == 0);

    for (int frame = 0; frame < 3; ++frame) {               // buffers are reused across frames
        cd.Reset();
        float a[16], b[16], c[16], d[16], e[16];
        translation(a, 0.f, 0.f, 0.f, 0.1f); translation(b, 1.2f, 0.3f, 0.5f, 0.7f);           // 2 and 4 interpenetrate
        translation(c, 20.f, 0.f, 0.f, 0.2f); translation(d, 21.1f, 0.4f, -0.3f, 1.1f);          // 5 and 6 interpenetrate
        translation(e, 100.f, 50.f, 0.f, 0.f);                                                   // 8 touches nothing
        cd.AddCollisionDetectionEntry(a, a, mesh, true, 2);
        cd.AddCollisionDetectionEntry(b, a, mesh, false, 4);                                     // moved since last frame; no callback flag
        cd.AddCollisionDetectionEntry(c, c, mesh, true, 5);
        cd.AddCollisionDetectionEntry(d, d, mesh, true, 6);
        cd.AddCollisionDetectionEntry(e, e, mesh, true, 8);
        cd.ExecuteCollisionDetection();
        EXPECT(ecs.calls == frame + 1);
        const imrcd_entity_pair* pairs; uint64_t n_pairs;
        cd.Results(&pairs, &n_pairs);
        EXPECT(n_pairs == 2);
        std::set<std::pair<int, int>> coll;
        for (uint64_t k = 0; k < n_pairs; ++k) {
            coll.insert({std::min(pairs[k].entity_first, pairs[k].entity_second), std::max(pairs[k].entity_first, pairs[k].entity_second)});
            EXPECT(pairs[k].n_hits > 0 && pairs[k].n_rays_first + pairs[k].n_rays_second > 0);
        }
        EXPECT(coll.count({2, 4}) && coll.count({5, 6}));
        // fan-out: 2 and its ancestor 1 hear about 4; 4 and 3 hear about 2; 5 hears about 6 and 6 about 5; the shared parent 7 hears nothing
        std::map<Entity, std::vector<std::pair<Entity, Entity>>> got;
        for (auto& kv : ecs.received) for (auto& cb : kv.second) got[kv.first].push_back({cb.familyEntity, cb.collideWithEntity});
        EXPECT(got.size() == 6);
        EXPECT(got[1] == (std::vector<std::pair<Entity, Entity>>{{2, 4}}) && got[2] == (std::vector<std::pair<Entity, Entity>>{{2, 4}}));
        EXPECT(got[3] == (std::vector<std::pair<Entity, Entity>>{{4, 2}}) && got[4] == (std::vector<std::pair<Entity, Entity>>{{4, 2}}));
        EXPECT(got[5] == (std::vector<std::pair<Entity, Entity>>{{5, 6}}) && got[6] == (std::vector<std::pair<Entity, Entity>>{{6, 5}}));
        EXPECT(!got.count(7) || got[7].empty());
    }
    // a pair needs shouldCallback on either side (SweepAndPrune.cpp:60)
    cd.Reset();
    float a[16], b[16];
    translation(a, 0.f, 0.f, 0.f, 0.1f); translation(b, 1.2f, 0.3f, 0.5f, 0.7f);
    cd.AddCollisionDetectionEntry(a, a, mesh, false, 2);
    cd.AddCollisionDetectionEntry(b, b, mesh, false, 4);
    cd.ExecuteCollisionDetection();
    const imrcd_entity_pair* pairs; uint64_t n_pairs;
    cd.Results(&pairs, &n_pairs);
    EXPECT(n_pairs == 0);
    // a whole glTF file through the adapter (F4): one tree per mesh of tests/golden/gltf_scene.glb, triangle counts as the reference's loader gives them
    if (argc > 1) {
        const std::vector<uint32_t> ids = cd.LoadMeshesOfModel(argv[1], IMRCD_BUILD_REFERENCE);
        const uint64_t want[5] = {730, 341, 400, 0, 40};
        EXPECT(ids.size() == 5);
        for (size_t m = 0; m < ids.size() && m < 5; ++m) {
            uint64_t n_tri = 0, n_vert = 0;
            EXPECT(imrcd_mesh_info(cd.context(), ids[m], &n_tri, &n_vert) == IMRCD_OK && n_tri == want[m]);
        }
        bool threw = false;
        try { cd.LoadMeshesOfModel(std::string(argv[1]) + ".missing"); } catch (const std::runtime_error&) { threw = true; }
        EXPECT(threw);
    }
    return 0;
}
