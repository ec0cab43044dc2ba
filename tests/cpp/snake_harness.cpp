// snake_harness.cpp -- SURVEY 8f F3: SnakeGame's own SnakePlayerComp, the engine's own frame loop, headless.
// The game is the REAL game DLL (oracle/_ref/libsnake_game.so: testGames/SnakeGame/game_dll + the engine sources its CMakeLists lists,
// unmodified, -DGAME_DLL, loaded with dlopen and asked for its components through GetGameDLLComponents like Engine.cpp does).  The engine
// side is the reference's real ECS and general components (NodeData, AnimationComposer, AnimationActor, Early/LateNodeGlobalMatrix,
// CameraDefaultInput, Camera, ModelCollision), compiled where they lie; what is missing is only what needs a window: Graphics (a test
// double of MeshesOfNodes.h that holds the trees) and the glTF importer (the fabs of gameConfig.cfg are laid out by hand below, with the
// component maps GameImporter.cpp:186-196,274-327 would attach).  Two builds (oracle/Makefile `snake`):
//   snake_harness_ref     the reference's CollisionDetection
//   snake_harness_dropin  tests/cpp/stubs_dropin in front of the include path: ModelCollisionComp.cpp compiles UNCHANGED against the drop-in
//   snake_harness_shadow  -DIMRCD_SHADOW, tests/cpp/stubs_shadow: BOTH behind one class; the reference drives the game, the drop-in gets the
//                         same entries every frame and its callbacks are compared with the reference's in lock step ("shadow" lines)
//   snake_harness_record  -DIMRCD_SHADOW -DIMRCD_RECORD_ONLY: the tee with the reference alone (CPU only); SNAKE_DUMP=<file> writes the meshes,
//                         every frame's entries and the reference's verdict (tests/golden/make_golden.py snake -> tests/golden/snake_frames.npz)
// Frames run as Engine::Run does (ECSwrapper::Update, CompleteAddsAndRemoves) on a fixed 1/60 s clock (steady_clock::now is interposed,
// so both builds see the same delta times and the game's std::rand draws line up).  Every frame prints each snake's position: the loop
// is closed -- deltaVectors move the snakes (SnakePlayerCompEntity.cpp:223-251), the moved snakes make the next frame's entries.
// TEST INFRASTRUCTURE: built only where the reference checkout exists.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <time.h>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "ECS/ECSwrapper.h"
#include "ECS/ComponentsIDsEnum.h"
#include "ECS/GeneralComponents/AnimationActorComp.h"
#include "ECS/GeneralComponents/AnimationComposerComp.h"
#include "ECS/GeneralComponents/CameraComp.h"
#include "ECS/GeneralComponents/CameraDefaultInputComp.h"
#include "ECS/GeneralComponents/EarlyNodeGlobalMatrixComp.h"
#include "ECS/GeneralComponents/LateNodeGlobalMatrixComp.h"
#include "ECS/GeneralComponents/ModelCollisionComp.h"
#include "ECS/GeneralComponents/NodeDataComp.h"
#include "Graphics/Meshes/AnimationsDataOfNodes.h"
#include "Geometry/Triangle.h"

// ---- the fixed clock: one definition in the executable, seen by the game DLL too (ELF interposition, -rdynamic)
static long long g_now_ns = 1'000'000'000;
namespace std { namespace chrono { inline namespace _V2 {
steady_clock::time_point steady_clock::now() noexcept { return time_point(duration(g_now_ns)); }
} } }

struct Exported : ExportedFunctions {
    void BindCameraEntity(Entity) const override {}
    size_t GetSphereMeshIndex() const override { return 0; }
    size_t GetCylinderMeshIndex() const override { return 0; }
};

static FILE* g_dump = nullptr;               // SNAKE_DUMP=<file>: meshes, every frame's entries and the reference's callbacks, floats as %a
struct Got { Entity receiver, family, other; glm::vec3 delta; };
static bool got_less(const Got& a, const Got& b) { return std::tie(a.receiver, a.family, a.other) < std::tie(b.receiver, b.family, b.other); }
static std::vector<Got> g_reference_callbacks, g_drop_in_callbacks;
static void record(std::vector<Got>& to, const std::vector<std::pair<Entity, std::vector<CollisionCallbackData>>>& in) {
    for (const auto& kv : in) for (const auto& c : kv.second) to.push_back({kv.first, c.familyEntity, c.collideWithEntity, c.deltaVector});
}
#ifdef IMRCD_SHADOW
static const MeshesOfNodes* g_meshes = nullptr;
void imrcd_shadow_entry(const CollisionDetectionEntry& e) {
    if (!g_dump) return;
    size_t mesh = 0;
    while (mesh < g_meshes->meshes.size() && &g_meshes->meshes[mesh].boundBoxTree != e.OBBtree_ptr) ++mesh;
    std::fprintf(g_dump, "entry %u %zu %d", unsigned(e.entity), mesh, int(e.shouldCallback));
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) std::fprintf(g_dump, " %a", e.currentGlobalMatrix[c][r]);
    for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) std::fprintf(g_dump, " %a", e.previousGlobalMatrix[c][r]);
    std::fprintf(g_dump, "\n");
}
static size_t g_shadow_missing = 0, g_shadow_compared = 0;
static double g_shadow_worst = 0.0, g_shadow_frame_worst = 0.0, g_shadow_frame_worst_abs = 0.0;
static size_t g_shadow_over = 0;            // deltaVectors further apart than 1e-4 of their length + 2e-5 world units (a few ulps of the scene's coordinates)
void imrcd_shadow_sink(const std::vector<std::pair<Entity, std::vector<CollisionCallbackData>>>& callbacks) { record(g_drop_in_callbacks, callbacks); }
void imrcd_shadow_compare() {           // same entries went to both: same receivers and pairs, deltaVectors equal up to the order of the FP64 sums
    std::vector<Got>& a = g_reference_callbacks; std::vector<Got>& b = g_drop_in_callbacks;
    std::sort(a.begin(), a.end(), got_less); std::sort(b.begin(), b.end(), got_less);
    if (g_dump) {                            // the reference's verdict on this frame: one row per colliding (entity, other) with its deltaVector
        for (const Got& g : a) if (g.receiver == g.family) std::fprintf(g_dump, "callback %u %u %a %a %a\n", unsigned(g.family), unsigned(g.other), g.delta.x, g.delta.y, g.delta.z);
        std::fprintf(g_dump, "end\n");
    }
    g_shadow_frame_worst = 0.0; g_shadow_frame_worst_abs = 0.0;
#ifdef IMRCD_RECORD_ONLY
    a.clear(); return;
#endif
    static const bool verbose = std::getenv("SNAKE_SHADOW_VERBOSE") != nullptr;
    size_t i = 0, j = 0;
    while (i < a.size() || j < b.size()) {          // merge by (receiver, family, other)
        if (j == b.size() || (i < a.size() && got_less(a[i], b[j]))) {
            if (verbose) std::printf("  only reference: %u %u %u  %.9g %.9g %.9g\n", unsigned(a[i].receiver), unsigned(a[i].family), unsigned(a[i].other), a[i].delta.x, a[i].delta.y, a[i].delta.z);
            ++g_shadow_missing; ++i; continue;
        }
        if (i == a.size() || got_less(b[j], a[i])) {
            if (verbose) std::printf("  only drop-in:   %u %u %u  %.9g %.9g %.9g\n", unsigned(b[j].receiver), unsigned(b[j].family), unsigned(b[j].other), b[j].delta.x, b[j].delta.y, b[j].delta.z);
            ++g_shadow_missing; ++j; continue;
        }
        ++g_shadow_compared;
        const glm::vec3 da = a[i].delta, db = b[j].delta;
        const bool nan_a = std::isnan(da.x + da.y + da.z), nan_b = std::isnan(db.x + db.y + db.z);
        if (nan_a || nan_b) { if (nan_a != nan_b) ++g_shadow_missing; ++i; ++j; continue; }
        const double len = glm::length(da), err = glm::length(da - db);
        const double rel = len > 0.0 ? err / len : (err > 0.0 ? 1.0 : 0.0);
        if (verbose && rel > 1e-4) std::printf("  rel %.3g: %u %u %u  ref %.9g %.9g %.9g  drop-in %.9g %.9g %.9g\n", rel, unsigned(a[i].receiver), unsigned(a[i].family), unsigned(a[i].other), da.x, da.y, da.z, db.x, db.y, db.z);
        g_shadow_frame_worst = std::max(g_shadow_frame_worst, rel); g_shadow_frame_worst_abs = std::max(g_shadow_frame_worst_abs, err);
        if (err > 1e-4 * len + 2e-5) ++g_shadow_over;
        ++i; ++j;
    }
    g_shadow_worst = std::max(g_shadow_worst, g_shadow_frame_worst);
    a.clear(); b.clear();
}
#endif

struct CallbackCounter : ComponentBaseClass {                   // counts what every component is handed (CollisionDetection.cpp:136-140)
    explicit CallbackCounter(ECSwrapper* e) : ComponentBaseClass(e) {}
    void CollisionCallback(const std::vector<std::pair<Entity, std::vector<CollisionCallbackData>>>& in) override {
        record(g_reference_callbacks, in);
        for (const auto& kv : in) {
            n += kv.second.size();
            if (dump) for (const auto& c : kv.second)
                std::printf("cb %u %u %u %.9g %.9g %.9g\n", unsigned(kv.first), unsigned(c.familyEntity), unsigned(c.collideWithEntity), c.deltaVector.x, c.deltaVector.y, c.deltaVector.z);
        }
    }
    bool dump = std::getenv("SNAKE_DUMP_CALLBACKS") != nullptr;
    componentID GetComponentID() const override { return 20000; }
    std::string GetComponentName() const override { return "CallbackCounter"; }
    size_t n = 0;
};

// ---- meshes, through the reference's own Triangle::CreateTriangleList and OBBtree constructor
static OBBtree tree_of(const std::vector<glm::vec3>& pts, const std::vector<glm::vec3>& nrm, const std::vector<uint32_t>& idx) {
    if (g_dump) {
        std::fprintf(g_dump, "mesh %zu %zu\n", pts.size(), idx.size());
        for (size_t i = 0; i < pts.size(); ++i) std::fprintf(g_dump, "%a %a %a %a %a %a\n", pts[i].x, pts[i].y, pts[i].z, nrm[i].x, nrm[i].y, nrm[i].z);
        for (size_t i = 0; i < idx.size(); ++i) std::fprintf(g_dump, "%u%c", idx[i], i % 24 == 23 || i + 1 == idx.size() ? '\n' : ' ');
    }
    return OBBtree(Triangle::CreateTriangleList(pts, nrm, idx, glTFmode::triangles));
}
static OBBtree sphere_tree(int nu, int nv, float r) {          // one vertex per pole and a fan around it: no zero-area triangles (those send
    std::vector<glm::vec3> pts, nrm; std::vector<uint32_t> idx; // the reference's tri-tri test into its uninitialised-read path)
    for (int i = 1; i < nv; ++i)
        for (int j = 0; j < nu; ++j) {
            const float t = 3.14159265f * i / nv, p = 6.2831853f * j / nu;
            const glm::vec3 n(std::sin(t) * std::cos(p), std::cos(t), std::sin(t) * std::sin(p));
            pts.push_back(r * n); nrm.push_back(n);
        }
    const uint32_t north = uint32_t(pts.size()); pts.push_back(glm::vec3(0.f, r, 0.f)); nrm.push_back(glm::vec3(0.f, 1.f, 0.f));
    const uint32_t south = uint32_t(pts.size()); pts.push_back(glm::vec3(0.f, -r, 0.f)); nrm.push_back(glm::vec3(0.f, -1.f, 0.f));
    for (int j = 0; j < nu; ++j) {
        const uint32_t a = j, b = (j + 1) % nu, c = (nv - 2) * nu + j, d = (nv - 2) * nu + (j + 1) % nu;
        idx.insert(idx.end(), {north, b, a, south, c, d});
    }
    for (int i = 0; i + 2 < nv; ++i)
        for (int j = 0; j < nu; ++j) {
            const uint32_t a = i * nu + j, b = i * nu + (j + 1) % nu, c = (i + 1) * nu + (j + 1) % nu, d = (i + 1) * nu + j;
            idx.insert(idx.end(), {a, b, c, a, c, d});
        }
    return tree_of(pts, nrm, idx);
}
static OBBtree slab_tree(float hx, float hy, float hz, int sub) {       // a box, every face a sub x sub grid
    std::vector<glm::vec3> pts, nrm; std::vector<uint32_t> idx;
    for (int axis = 0; axis < 3; ++axis)
        for (int side = -1; side <= 1; side += 2) {
            const uint32_t base = uint32_t(pts.size());
            glm::vec3 n(0.f); n[axis] = float(side);
            const int u = (axis + 1) % 3, v = (axis + 2) % 3;
            const float h[3] = {hx, hy, hz};
            for (int i = 0; i <= sub; ++i)
                for (int j = 0; j <= sub; ++j) {
                    glm::vec3 p; p[axis] = side * h[axis]; p[u] = -h[u] + 2.f * h[u] * i / sub; p[v] = -h[v] + 2.f * h[v] * j / sub;
                    pts.push_back(p); nrm.push_back(n);
                }
            for (int i = 0; i < sub; ++i)
                for (int j = 0; j < sub; ++j) {
                    const uint32_t a = base + i * (sub + 1) + j, b = a + 1, c = a + sub + 2, d = a + sub + 1;
                    if (side > 0) idx.insert(idx.end(), {a, b, c, a, c, d}); else idx.insert(idx.end(), {a, c, b, a, d, c});
                }
        }
    return tree_of(pts, nrm, idx);
}

// ---- fabs, as GameImporter would hand them to ECSwrapper::AddFabs
static constexpr componentID ID(componentIDenum e) { return static_cast<componentID>(e); }
static constexpr componentID kSnakePlayerID = 1100;              // GameSpecificComponentsIDsEnum.inj (visible only with -DGAME_DLL)
static void spatial(Node* n, const CompEntityInitMap& node_data = CompEntityInitMap()) {      // GameImporter.cpp:284-287
    n->componentIDsToInitMaps.emplace(ID(componentIDenum::NodeData), node_data);
    n->componentIDsToInitMaps.emplace(ID(componentIDenum::EarlyNodeGlobalMatrix), CompEntityInitMap());
    n->componentIDsToInitMaps.emplace(ID(componentIDenum::LateNodeGlobalMatrix), CompEntityInitMap());
}
static Node* child(Node* parent, const std::string& name) {
    parent->children.emplace_back(std::make_unique<Node>());
    parent->children.back()->nodeName = name;
    return parent->children.back().get();
}
static CompEntityInitMap at(glm::vec3 t, glm::vec3 s = glm::vec3(1.f)) {
    CompEntityInitMap m;
    m.vec4Map["LocalTranslation"] = glm::vec4(t, 0.f); m.vec4Map["LocalScale"] = glm::vec4(s, 0.f);
    return m;
}
static std::unique_ptr<Node> snake_fab(const std::string& name, glm::vec3 start, glm::vec3 dir) {   // gameConfig.cfg snake_fabN over snake/Scene
    auto root = std::make_unique<Node>(); root->nodeName = name; spatial(root.get());
    Node* arm = child(root.get(), "SnakeArmature");
    CompEntityInitMap nd; nd.vec4Map["LocalScale"] = glm::vec4(4.f, 4.f, 4.f, 0.f); nd.vec4Map["GlobalTranslation"] = glm::vec4(start, 0.f);
    spatial(arm, nd);
    CompEntityInitMap sp;
    sp.floatMap["Speed"] = 4.0f; sp.floatMap["RotationSpeed"] = 1.5f; sp.stringMap["AnimationComposer"] = "../_animations/SnakeArmatureAction";
    sp.vec4Map["CameraOffset"] = glm::vec4(-2.f, -1.f, 0.f, 0.f); sp.vec4Map["InitDirection"] = glm::vec4(dir, 0.f);
    arm->componentIDsToInitMaps.emplace(kSnakePlayerID, sp);
    arm->componentIDsToInitMaps.emplace(ID(componentIDenum::Camera), CompEntityInitMap());
    Node* bone = child(arm, "Bone"); spatial(bone);
    CompEntityInitMap actor; actor.stringMap["Animation_0"] = "SnakeArmatureAction"; actor.intMap["SnakeArmatureAction_animationIndex"] = 0;
    bone->componentIDsToInitMaps.emplace(ID(componentIDenum::AnimationActor), actor);
    Node* ball = child(bone, "_collision_sphere"); spatial(ball);
    CompEntityInitMap mc; mc.intMap["MeshIndex"] = 0; mc.intMap["ShouldCallback"] = 1;                 // gameConfig.cfg "SnakeArmature/Bone/_collision_sphere"
    ball->componentIDsToInitMaps.emplace(ID(componentIDenum::ModelCollision), mc);
    Node* anims = child(root.get(), "_animations");
    Node* action = child(anims, "SnakeArmatureAction");
    CompEntityInitMap ac; ac.stringMap["AnimationName"] = "SnakeArmatureAction"; ac.stringMap["NodesRelativeName_0"] = "../../SnakeArmature/Bone";
    action->componentIDsToInitMaps.emplace(ID(componentIDenum::AnimationComposer), ac);
    return root;
}

int main(int argc, char** argv) {
    const int n_frames = argc > 1 ? std::atoi(argv[1]) : 90;
    const int n_snakes = argc > 2 ? std::atoi(argv[2]) : 6;
    const char* dll = argc > 3 ? argv[3] : "libsnake_game.so";

    Exported exported;
    ECSwrapper ecs(&exported);
    AnimationsDataOfNodes animations;
    {   // one animation: the bone (and the collision sphere under it) sways sideways
        const size_t index = animations.RegistAnimationsDataAndGetIndex();
        AnimationData d;
        for (int k = 0; k <= 8; ++k) d.timeToTranslationKey_map.emplace(0.25f * k, glm::vec3(0.f, 0.f, 0.12f * std::sin(0.785398f * k)));
        animations.AddAnimationData(index, d);
    }
    if (const char* path = std::getenv("SNAKE_DUMP")) g_dump = std::fopen(path, "w");
    MeshesOfNodes meshes;
    meshes.meshes.push_back({sphere_tree(20, 12, 0.16f)});          // 0: the snake's collision sphere (scaled x4 by the armature)
    meshes.meshes.push_back({slab_tree(16.f, 0.5f, 16.f, 24)});      // 1: the floor
    meshes.meshes.push_back({slab_tree(0.6f, 2.0f, 0.6f, 6)});       // 2: pillars
    meshes.meshes.push_back({sphere_tree(24, 16, 1.0f)});            // 3: boulders / apples
    meshes.meshes.push_back({slab_tree(16.f, 1.5f, 0.4f, 16)});      // 4: walls

#ifdef IMRCD_SHADOW
    g_meshes = &meshes;
#endif
    // Engine.cpp:26-52
    ecs.AddComponentAndOwnership(std::make_unique<AnimationComposerComp>(&ecs));
    ecs.AddComponentAndOwnership(std::make_unique<AnimationActorComp>(&ecs, &animations));
    ecs.AddComponentAndOwnership(std::make_unique<NodeDataComp>(&ecs));
    ecs.AddComponentAndOwnership(std::make_unique<EarlyNodeGlobalMatrixComp>(&ecs));
    ecs.AddComponentAndOwnership(std::make_unique<LateNodeGlobalMatrixComp>(&ecs));
    ecs.AddComponentAndOwnership(std::make_unique<CameraDefaultInputComp>(&ecs, 2.f));
    ecs.AddComponentAndOwnership(std::make_unique<CameraComp>(&ecs, 1.0f, 1.6f, 0.1f, 100.f));
    CollisionDetection collision_detection(&ecs);
    ecs.AddComponentAndOwnership(std::make_unique<ModelCollisionComp>(&ecs, &collision_detection, &meshes));
    CallbackCounter counter(&ecs);
    ecs.AddComponent(&counter);

    // the game DLL (Engine / GameImporter: dlopen + GetGameDLLComponents, game_dll.h)
    void* lib = dlopen(dll, RTLD_NOW | RTLD_LOCAL);
    if (!lib) { std::fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    using GetComponentsFn = std::vector<std::unique_ptr<ComponentBaseClass>> (*)(ECSwrapper*);
    auto get_components = reinterpret_cast<GetComponentsFn>(dlsym(lib, "_Z20GetGameDLLComponentsP10ECSwrapper"));
    if (!get_components) { std::fprintf(stderr, "GetGameDLLComponents not found\n"); return 2; }
    for (auto& c : get_components(&ecs)) ecs.AddComponentAndOwnership(std::move(c));

    // fabs
    std::vector<std::unique_ptr<Node>> fabs;
    {
        auto map = std::make_unique<Node>(); map->nodeName = "map_fab"; spatial(map.get());
        auto solid = [&](const std::string& name, int mesh, glm::vec3 t, glm::vec3 s = glm::vec3(1.f)) {
            Node* n = child(map.get(), name); spatial(n, at(t, s));
            CompEntityInitMap mc; mc.intMap["MeshIndex"] = mesh;                                       // GameImporter.cpp:321-327
            n->componentIDsToInitMaps.emplace(ID(componentIDenum::ModelCollision), mc);
        };
        solid("floor", 1, glm::vec3(0.f, 0.5f, 0.f));                                                  // +y is down (UpDirection 0,-1,0): top face at y = 0
        solid("wall_n", 4, glm::vec3(0.f, -1.f, 9.f)); solid("wall_s", 4, glm::vec3(0.f, -1.f, -9.f));
        for (int k = 0; k < 10; ++k) {
            const float a = 0.6283185f * k;
            solid("pillar_" + std::to_string(k), 2, glm::vec3(5.5f * std::cos(a), -1.8f, 5.5f * std::sin(a)));
            solid("apple_" + std::to_string(k), 3, glm::vec3(2.6f * std::cos(a + 0.3f), -0.7f, 2.6f * std::sin(a + 0.3f)), glm::vec3(0.8f));
        }
        fabs.push_back(std::move(map));
    }
    for (int s = 0; s < n_snakes; ++s) {
        const float a = 6.2831853f * s / n_snakes;
        fabs.push_back(snake_fab("snake_fab" + std::to_string(s), glm::vec3(1.2f * std::cos(a), -1.5f - 0.2f * s, 1.2f * std::sin(a)),
                                 glm::vec3(std::cos(a + 2.2f), 0.f, std::sin(a + 2.2f))));
    }
    std::vector<Node*> roots;
    for (auto& f : fabs) roots.push_back(f.get());
    ecs.AddFabs(roots);
    ecs.AddInstance("map_fab", "map");                                                                  // GameImporter.cpp:690-692
    ecs.CompleteAddsAndRemoves();
    std::vector<Entity> snakes;
    for (int s = 0; s < n_snakes; ++s) {
        const std::string name = "snake" + std::to_string(s);
        ecs.AddInstance("snake_fab" + std::to_string(s), name);
        ecs.CompleteAddsAndRemoves();
        snakes.push_back(ecs.GetEntitiesHandler()->FindEntityByPath(name, "snake_fab" + std::to_string(s) + "/SnakeArmature"));
    }
    auto* node_data = static_cast<NodeDataComp*>(ecs.GetComponentByID(ID(componentIDenum::NodeData)));

    ecs.RefreshUpdateDeltaTime();                                                                       // Engine.cpp:103
    auto wall = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; };    // the real clock (steady_clock is the fixed one)
    double wall_frames = 0.0;
    for (int frame = 0; frame < n_frames; ++frame) {
        g_now_ns += 16'666'667;
        if (g_dump) std::fprintf(g_dump, "frame %d\n", frame);
        const size_t before = counter.n;
        const double t0 = wall();
        ecs.Update();                                                                                   // Engine.cpp:133-135 (DrawFrame left out)
        ecs.CompleteAddsAndRemoves();
        if (frame >= 10) wall_frames += wall() - t0;                                                    // the first frames pay for tree uploads / warm-up
        std::printf("frame %d callbacks %zu\n", frame, counter.n - before);
#ifdef IMRCD_SHADOW
        std::printf("shadow %d compared %zu mismatched %zu worst %.3g abs %.3g over %zu\n", frame, g_shadow_compared, g_shadow_missing, g_shadow_frame_worst, g_shadow_frame_worst_abs, g_shadow_over);
#else
        g_reference_callbacks.clear();
#endif
        for (size_t s = 0; s < snakes.size(); ++s) {
            const glm::vec3 p = node_data->GetComponentEntity(snakes[s]).globalTranslation;
            std::printf("snake %zu %.9g %.9g %.9g\n", s, p.x, p.y, p.z);
        }
    }
    if (n_frames > 10) std::fprintf(stderr, "wall: %.3f ms per ECS frame over %d frames\n", 1e3 * wall_frames / (n_frames - 10), n_frames - 10);
    if (g_dump) std::fclose(g_dump);
    return 0;
}
