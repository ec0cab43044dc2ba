// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" shim around the *unmodified* reference collision sources,
// compiled where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libimr_ref.so (flags: -std=c++20 -O2 -ffp-contract=off, no
// -ffast-math; see SURVEY.md finding 2).  Nothing here re-implements the
// reference's arithmetic: every box, SAT verdict, triangle predicate and ray
// comes out of the reference's own translation units.  The shim only
//   * builds std::vector<Triangle> from flat arrays and calls OBBtree::OBBtree
//     (inMyRoom_vulkan/src/Geometry/OBBtree.cpp:321),
//   * flattens the result through the public OBBtreeTraveler
//     (inMyRoom_vulkan/include/Geometry/OBBtree.h:78-101),
//   * drives SweepAndPrune / OBBtreesCollision / CreateUncollideRays
//     (inMyRoom_vulkan/src/CollisionDetection/*.cpp) on flat entry tables, and
//   * re-walks the leaf combos with the reference's own
//     TrianglePosition::IntersectTriangles to list the individual hits that
//     CreateUncollideRays.cpp:74-115 consumes but never exposes.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
#include <cstdint>
#include <cstring>
#include <chrono>
#include <unordered_map>
#include <vector>
#include <string>
#include <algorithm>

#include "Geometry/OBBtree.h"
#include "Geometry/Triangle.h"
#include "Geometry/Plane.h"
#include "CollisionDetection/SweepAndPrune.h"
#include "CollisionDetection/OBBtreesCollision.h"
#include "CollisionDetection/CreateUncollideRays.h"
#include "CollisionDetection/ShootUncollideRays.h"

namespace {

struct RefTree {
    OBBtree tree;
    size_t n_tri = 0;
    std::vector<uint32_t> orig_index;  // leaf-order triangle -> input triangle index
};

struct TriKey {
    uint32_t w[12];
    bool operator==(const TriKey& o) const { return std::memcmp(w, o.w, sizeof(w)) == 0; }
};
struct TriKeyHash {
    size_t operator()(const TriKey& k) const {
        uint64_t h = 1469598103934665603ull;
        for (uint32_t x : k.w) { h ^= x; h *= 1099511628211ull; }
        return size_t(h);
    }
};

TriKey make_key(const TrianglePosition& p, const TriangleIndices& idx) {
    TriKey k;
    for (int v = 0; v < 3; ++v) {
        glm::vec3 q = p.GetP(v);
        std::memcpy(&k.w[3 * v], &q, 12);
        k.w[9 + v] = idx.GetI(v);
    }
    return k;
}

glm::mat4 load_mat(const float* m) {
    glm::mat4 r;
    std::memcpy(&r, m, 64);
    return r;
}

// Sweep axes exactly as the engine constructs them
// (inMyRoom_vulkan/src/CollisionDetection/CollisionDetection.cpp:9-13).
void sweep_axes(glm::vec3& U, glm::vec3& V, glm::vec3& W) {
    U = glm::normalize(glm::vec3(0.8f, -0.2f, 0.f));
    W = glm::normalize(glm::cross(U, glm::vec3(0.f, -1.f, 0.f)));
    V = glm::normalize(glm::cross(W, U));
}

struct Flattener {
    float* boxes; int32_t* left; int32_t* right; uint32_t* tri_off; uint32_t* tri_cnt;
    size_t n = 0;
    size_t walk(const OBBtree::OBBtreeTraveler& t, bool write) {
        size_t me = n++;
        if (write) {
            OBB b = t.GetOBB();
            glm::vec3 c = b.GetCenter(), u = b.GetSideDirectionU(), v = b.GetSideDirectionV(), w = b.GetSideDirectionW();
            float* o = boxes + 12 * me;
            std::memcpy(o + 0, &c, 12); std::memcpy(o + 3, &u, 12);
            std::memcpy(o + 6, &v, 12); std::memcpy(o + 9, &w, 12);
        }
        if (t.IsLeaf()) {
            if (write) {
                left[me] = -1; right[me] = -1;
                tri_off[me] = uint32_t(t.GetTrianglesOffset());
                tri_cnt[me] = uint32_t(t.GetTrianglesCount());
            }
        } else {
            size_t l = walk(t.GetLeftChildTraveler(), write);
            size_t r = walk(t.GetRightChildTraveler(), write);
            if (write) { left[me] = int32_t(l); right[me] = int32_t(r); tri_off[me] = 0; tri_cnt[me] = 0; }
        }
        return me;
    }
};

CollisionDetectionEntry make_entry(const float* cur, const float* prev, const RefTree* t, bool cb, uint16_t ent) {
    CollisionDetectionEntry e;
    e.currentGlobalMatrix = load_mat(cur);
    e.previousGlobalMatrix = load_mat(prev ? prev : cur);
    e.OBBtree_ptr = &t->tree;
    e.shouldCallback = cb;
    e.entity = ent;
    return e;
}

// Counting re-walk of OBBtree::IntersectOBBtreesRecursive (OBBtree.cpp:414-477)
// built only from the reference's public pieces; used to count SAT visits and
// cross-checked against the real combos by the tests.
void count_walk(const OBBtree::OBBtreeTraveler& a, const OBBtree::OBBtreeTraveler& b, const glm::mat4& m,
                uint64_t& visits, uint64_t& passes, uint64_t& combos, uint64_t& tri_tests, uint32_t depth, uint32_t& max_depth) {
    ++visits;
    if (depth > max_depth) max_depth = depth;
    Paralgram pa = a.GetOBB();
    Paralgram pb = m * b.GetOBB();
    if (!Paralgram::IntersectParalgramsBoolean(pa, pb)) return;
    ++passes;
    bool la = a.IsLeaf(), lb = b.IsLeaf();
    if (!la && !lb) {
        if (pa.GetSurface() >= pb.GetSurface()) {
            count_walk(a.GetLeftChildTraveler(), b, m, visits, passes, combos, tri_tests, depth + 1, max_depth);
            count_walk(a.GetRightChildTraveler(), b, m, visits, passes, combos, tri_tests, depth + 1, max_depth);
        } else {
            count_walk(a, b.GetLeftChildTraveler(), m, visits, passes, combos, tri_tests, depth + 1, max_depth);
            count_walk(a, b.GetRightChildTraveler(), m, visits, passes, combos, tri_tests, depth + 1, max_depth);
        }
    } else if (la && !lb) {
        count_walk(a, b.GetLeftChildTraveler(), m, visits, passes, combos, tri_tests, depth + 1, max_depth);
        count_walk(a, b.GetRightChildTraveler(), m, visits, passes, combos, tri_tests, depth + 1, max_depth);
    } else if (!la && lb) {
        count_walk(a.GetLeftChildTraveler(), b, m, visits, passes, combos, tri_tests, depth + 1, max_depth);
        count_walk(a.GetRightChildTraveler(), b, m, visits, passes, combos, tri_tests, depth + 1, max_depth);
    } else {
        ++combos;
        tri_tests += uint64_t(a.GetTrianglesCount()) * uint64_t(b.GetTrianglesCount());
    }
}

}  // namespace

extern "C" {

// ---- tree build ---------------------------------------------------------
// positions/normals: n_tri*9 floats (p0 p1 p2), vertex_ids: n_tri*3 u32.
void* imr_ref_tree_create(const float* positions, const float* normals, const uint32_t* vertex_ids, uint64_t n_tri) {
    std::vector<Triangle> tris;
    tris.reserve(n_tri);
    std::unordered_map<TriKey, std::vector<uint32_t>, TriKeyHash> by_key;
    for (uint64_t i = 0; i < n_tri; ++i) {
        const float* p = positions + 9 * i;
        TrianglePosition tp(glm::vec3(p[0], p[1], p[2]), glm::vec3(p[3], p[4], p[5]), glm::vec3(p[6], p[7], p[8]));
        TriangleNormal tn = normals
            ? TriangleNormal(glm::vec3(normals[9 * i + 0], normals[9 * i + 1], normals[9 * i + 2]),
                             glm::vec3(normals[9 * i + 3], normals[9 * i + 4], normals[9 * i + 5]),
                             glm::vec3(normals[9 * i + 6], normals[9 * i + 7], normals[9 * i + 8]))
            : TriangleNormal(tp.GetTriangleFaceNormal());   // Triangle.cpp:214-234 fallback
        TriangleIndices ti = vertex_ids ? TriangleIndices(vertex_ids[3 * i], vertex_ids[3 * i + 1], vertex_ids[3 * i + 2])
                                        : TriangleIndices(uint32_t(3 * i), uint32_t(3 * i + 1), uint32_t(3 * i + 2));
        tris.emplace_back(tp, tn, ti);
        by_key[make_key(tp, ti)].push_back(uint32_t(i));
    }
    RefTree* rt = new RefTree();
    rt->n_tri = n_tri;
    rt->tree = OBBtree(std::move(tris));
    // leaf order -> input order (duplicates consumed in input order)
    rt->orig_index.resize(n_tri);
    std::unordered_map<TriKey, size_t, TriKeyHash> cursor;
    for (uint64_t i = 0; i < n_tri; ++i) {
        TriKey k = make_key(rt->tree.GetTrianglePosition(i), rt->tree.GetTriangleIndices(i));
        size_t& c = cursor[k];
        rt->orig_index[i] = by_key[k][c++];
    }
    return rt;
}

void imr_ref_tree_destroy(void* t) { delete static_cast<RefTree*>(t); }

uint64_t imr_ref_tree_tri_count(void* t) { return static_cast<RefTree*>(t)->n_tri; }

// number of tree vertices (inner + leaves, root included), pre-order DFS
uint64_t imr_ref_tree_vertex_count(void* t) {
    Flattener f{};
    f.walk(static_cast<RefTree*>(t)->tree.GetRootTraveler(), false);
    return f.n;
}

// boxes: nv*12 floats (center,u,v,w); left/right: child vertex index or -1;
// tri_off/tri_cnt: leaf triangle range (leaf order) ; triangles in leaf order.
void imr_ref_tree_flatten(void* t, float* boxes, int32_t* left, int32_t* right, uint32_t* tri_off, uint32_t* tri_cnt,
                          float* tri_pos, float* tri_nrm, uint32_t* tri_vid, uint32_t* tri_orig) {
    RefTree* rt = static_cast<RefTree*>(t);
    Flattener f{boxes, left, right, tri_off, tri_cnt};
    f.walk(rt->tree.GetRootTraveler(), true);
    for (size_t i = 0; i < rt->n_tri; ++i) {
        TrianglePosition p = rt->tree.GetTrianglePosition(i);
        TriangleNormal nn = rt->tree.GetTriangleNormal(i);
        TriangleIndices ii = rt->tree.GetTriangleIndices(i);
        for (int v = 0; v < 3; ++v) {
            glm::vec3 q = p.GetP(v), m = nn.GetN(v);
            if (tri_pos) std::memcpy(tri_pos + 9 * i + 3 * v, &q, 12);
            if (tri_nrm) std::memcpy(tri_nrm + 9 * i + 3 * v, &m, 12);
            if (tri_vid) tri_vid[3 * i + v] = ii.GetI(v);
        }
        if (tri_orig) tri_orig[i] = rt->orig_index[i];
    }
}

// ---- single predicates (for unit-level differential tests) ---------------
// boxes: 12 floats each; returns IntersectParalgramsBoolean(lhs, m*rhs) (m may be NULL = no transform)
int imr_ref_sat(const float* lhs, const float* rhs, const float* m) {
    struct P : Paralgram { void set(const float* f) { std::memcpy(&center, f, 12); std::memcpy(&sideDirections, f + 3, 36);} };
    P a, b; a.set(lhs); b.set(rhs);
    Paralgram bb = m ? load_mat(m) * static_cast<Paralgram&>(b) : static_cast<Paralgram&>(b);
    return Paralgram::IntersectParalgramsBoolean(a, bb) ? 1 : 0;
}
float imr_ref_surface(const float* box, const float* m) {
    struct P : Paralgram { void set(const float* f) { std::memcpy(&center, f, 12); std::memcpy(&sideDirections, f + 3, 36);} };
    P a; a.set(box);
    Paralgram aa = m ? load_mat(m) * static_cast<Paralgram&>(a) : static_cast<Paralgram&>(a);
    return aa.GetSurface();
}
// out: 12 floats of m*box
void imr_ref_box_transform(const float* box, const float* m, float* out) {
    struct P : Paralgram { void set(const float* f) { std::memcpy(&center, f, 12); std::memcpy(&sideDirections, f + 3, 36);} };
    P a; a.set(box);
    Paralgram r = load_mat(m) * static_cast<Paralgram&>(a);
    glm::vec3 c = r.GetCenter(), u = r.GetSideDirectionU(), v = r.GetSideDirectionV(), w = r.GetSideDirectionW();
    std::memcpy(out, &c, 12); std::memcpy(out + 3, &u, 12); std::memcpy(out + 6, &v, 12); std::memcpy(out + 9, &w, 12);
}
// tri-tri on n pairs: a,b = n*9 floats; m (optional) applied to b as in CreateUncollideRays.cpp:84.
// flags bit0 = doIntersept, bit1 = areCoplanar; seg = n*6 floats (source,target), untouched garbage when undefined.
void imr_ref_tri_tri(const float* a, const float* b, const float* m, uint64_t n, uint8_t* flags, float* seg) {
    glm::mat4 M = m ? load_mat(m) : glm::mat4(1.f);
    for (uint64_t i = 0; i < n; ++i) {
        TrianglePosition ta(glm::vec3(a[9*i], a[9*i+1], a[9*i+2]), glm::vec3(a[9*i+3], a[9*i+4], a[9*i+5]), glm::vec3(a[9*i+6], a[9*i+7], a[9*i+8]));
        TrianglePosition tb(glm::vec3(b[9*i], b[9*i+1], b[9*i+2]), glm::vec3(b[9*i+3], b[9*i+4], b[9*i+5]), glm::vec3(b[9*i+6], b[9*i+7], b[9*i+8]));
        if (m) tb = M * tb;
        TrianglesIntersectionInfo info = Triangle::IntersectTriangles(ta, tb);
        flags[i] = uint8_t((info.doIntersept ? 1 : 0) | (info.areCoplanar ? 2 : 0));
        if (info.doIntersept && !info.areCoplanar) {
            std::memcpy(seg + 6 * i, &info.source, 12);
            std::memcpy(seg + 6 * i + 3, &info.target, 12);
        }
    }
}
// rel = inverse(a) * b   (OBBtreesCollision.cpp:15)
void imr_ref_pair_matrix(const float* a, const float* b, float* out) {
    glm::mat4 r = glm::inverse(load_mat(a)) * load_mat(b);
    std::memcpy(out, &r, 64);
}
// OBB fit of a point cloud (OBB.cpp:33-89); out 12 floats
void imr_ref_obb_from_points(const float* pts, uint64_t n, float* out) {
    std::vector<glm::vec3> p(n);
    std::memcpy(p.data(), pts, n * 12);
    OBB r = OBB::CreateOBBfromPoints(p);
    glm::vec3 c = r.GetCenter(), u = r.GetSideDirectionU(), v = r.GetSideDirectionV(), w = r.GetSideDirectionW();
    std::memcpy(out, &c, 12); std::memcpy(out + 3, &u, 12); std::memcpy(out + 6, &v, 12); std::memcpy(out + 9, &w, 12);
}
void imr_ref_sweep_axes(float* out9) {
    glm::vec3 U, V, W; sweep_axes(U, V, W);
    std::memcpy(out9, &U, 12); std::memcpy(out9 + 3, &V, 12); std::memcpy(out9 + 6, &W, 12);
}

// ---- broad phase ----------------------------------------------------------
// entries: n matrices (16 floats each), tree handle per entry, callback flag per entry. n <= 65534.
// Returns pair count; pairs written as (first,second) entry indices up to cap.
uint64_t imr_ref_broad(const float* mats, void* const* trees, const uint8_t* should_cb, uint64_t n,
                       uint32_t* pairs, uint64_t cap, double* seconds) {
    if (n > 65534) return uint64_t(-1);   // Entity is uint16_t (ECStypes.h:18, SweepAndPrune.cpp:26,59,73)
    std::vector<CollisionDetectionEntry> entries;
    entries.reserve(n);
    for (uint64_t i = 0; i < n; ++i)
        entries.push_back(make_entry(mats + 16 * i, nullptr, static_cast<RefTree*>(trees[i]), should_cb[i] != 0, uint16_t(i)));
    glm::vec3 U, V, W; sweep_axes(U, V, W);
    SweepAndPrune sap(U, V, W);
    auto t0 = std::chrono::steady_clock::now();
    auto res = sap.ExecuteSweepAndPrune(entries);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    uint64_t k = 0;
    for (auto& pr : res) {
        if (k < cap) { pairs[2 * k] = pr.first.entity; pairs[2 * k + 1] = pr.second.entity; }
        ++k;
    }
    return k;
}

// ---- mid phase -------------------------------------------------------------
// combos written as 4 u32 (offA,cntA,offB,cntB) in leaf order, up to cap; returns count.
uint64_t imr_ref_mid(void* tree_a, const float* mat_a, void* tree_b, const float* mat_b,
                     uint32_t* combos, uint64_t cap, double* seconds) {
    OBBtreesCollision mid;
    auto pr = std::make_pair(make_entry(mat_a, nullptr, static_cast<RefTree*>(tree_a), true, 1),
                             make_entry(mat_b, nullptr, static_cast<RefTree*>(tree_b), true, 2));
    auto t0 = std::chrono::steady_clock::now();
    CDentriesPairTrianglesPairs r = mid.ExecuteOBBtreesCollision(pr);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    uint64_t k = 0;
    for (auto& c : r.OBBtreesIntersectInfoObj.candidateTriangleRangeCombinations) {
        if (k < cap) {
            combos[4 * k] = uint32_t(c.first_obbtree_offset); combos[4 * k + 1] = uint32_t(c.first_obbtree_count);
            combos[4 * k + 2] = uint32_t(c.second_obbtree_offset); combos[4 * k + 3] = uint32_t(c.second_obbtree_count);
        }
        ++k;
    }
    return k;
}

// SAT visit statistics for one pair: out = {visits, passes, combos, tri_tests, max_depth}
void imr_ref_mid_stats(void* tree_a, const float* mat_a, void* tree_b, const float* mat_b, uint64_t* out5) {
    RefTree* a = static_cast<RefTree*>(tree_a); RefTree* b = static_cast<RefTree*>(tree_b);
    glm::mat4 rel = glm::inverse(load_mat(mat_a)) * load_mat(mat_b);
    uint64_t v = 0, p = 0, c = 0, tt = 0; uint32_t md = 0;
    count_walk(a->tree.GetRootTraveler(), b->tree.GetRootTraveler(), rel, v, p, c, tt, 0, md);
    out5[0] = v; out5[1] = p; out5[2] = c; out5[3] = tt; out5[4] = md;
}

// ---- narrow phase ------------------------------------------------------------
// Runs mid + CreateUncollideRays for one ordered pair.
//  hits: up to cap records of (triA_orig, triB_orig) u32 + 6 floats (source,target) + weight
//  summary[0]=n_combos [1]=n_tri_tests [2]=n_hits(non-coplanar) [3]=n_coplanar_hits
//         [4]=rays_first [5]=rays_second [6]=colliding(0/1)
//  avg: 6 floats (average_point_first_modelspace, average_point_second_modelspace)
//  seconds: [0]=mid [1]=narrow (CreateUncollideRays only)
void imr_ref_pair(void* tree_a, const float* mat_a, void* tree_b, const float* mat_b,
                  uint32_t* hit_ids, float* hit_seg, uint64_t cap, uint64_t* summary, float* avg, double* seconds) {
    RefTree* a = static_cast<RefTree*>(tree_a); RefTree* b = static_cast<RefTree*>(tree_b);
    OBBtreesCollision mid;
    CreateUncollideRays narrow;
    auto pr = std::make_pair(make_entry(mat_a, nullptr, a, true, 1), make_entry(mat_b, nullptr, b, true, 2));
    auto t0 = std::chrono::steady_clock::now();
    CDentriesPairTrianglesPairs m = mid.ExecuteOBBtreesCollision(pr);
    auto t1 = std::chrono::steady_clock::now();
    CDentriesUncollideRays rays;
    bool have = !m.OBBtreesIntersectInfoObj.candidateTriangleRangeCombinations.empty();   // CollisionDetection.cpp:51
    auto t2 = std::chrono::steady_clock::now();
    if (have) rays = narrow.ExecuteCreateUncollideRays(m);
    auto t3 = std::chrono::steady_clock::now();
    if (seconds) { seconds[0] = std::chrono::duration<double>(t1 - t0).count(); seconds[1] = std::chrono::duration<double>(t3 - t2).count(); }

    // list the individual hits with the reference's own predicate (CreateUncollideRays.cpp:62,74-88)
    const glm::mat4 rel = glm::inverse(pr.first.currentGlobalMatrix) * pr.second.currentGlobalMatrix;
    uint64_t n_tests = 0, n_hits = 0, n_copl = 0;
    for (auto& c : m.OBBtreesIntersectInfoObj.candidateTriangleRangeCombinations) {
        for (size_t i = 0; i != c.first_obbtree_count; ++i)
            for (size_t j = 0; j != c.second_obbtree_count; ++j) {
                TrianglePosition ta = a->tree.GetTrianglePosition(i + c.first_obbtree_offset);
                TrianglePosition tb = rel * b->tree.GetTrianglePosition(j + c.second_obbtree_offset);
                TrianglesIntersectionInfo info = Triangle::IntersectTriangles(ta, tb);
                ++n_tests;
                if (info.doIntersept && info.areCoplanar) ++n_copl;
                if (info.doIntersept && !info.areCoplanar) {
                    if (n_hits < cap) {
                        hit_ids[2 * n_hits] = a->orig_index[i + c.first_obbtree_offset];
                        hit_ids[2 * n_hits + 1] = b->orig_index[j + c.second_obbtree_offset];
                        std::memcpy(hit_seg + 7 * n_hits, &info.source, 12);
                        std::memcpy(hit_seg + 7 * n_hits + 3, &info.target, 12);
                        hit_seg[7 * n_hits + 6] = glm::length(info.source - info.target);   // CreateUncollideRays.cpp:93
                    }
                    ++n_hits;
                }
            }
    }
    summary[0] = m.OBBtreesIntersectInfoObj.candidateTriangleRangeCombinations.size();
    summary[1] = n_tests; summary[2] = n_hits; summary[3] = n_copl;
    summary[4] = have ? rays.rays_from_first_to_second.size() : 0;
    summary[5] = have ? rays.rays_from_second_to_first.size() : 0;
    summary[6] = (summary[4] || summary[5]) ? 1 : 0;                                        // CollisionDetection.cpp:63
    if (avg) {
        if (have) { std::memcpy(avg, &rays.average_point_first_modelspace, 12); std::memcpy(avg + 3, &rays.average_point_second_modelspace, 12); }
        else std::memset(avg, 0, 24);
    }
}

// ---- batch mid + narrow over a pair list (bench.py's --impl reference / cpu_baseline legs) ----------
// For each (first,second) entry pair: OBBtreesCollision::ExecuteOBBtreesCollision, then -- when there is at least
// one leaf combo (CollisionDetection.cpp:51) -- CreateUncollideRays::ExecuteCreateUncollideRays, exactly the loop of
// CollisionDetection.cpp:44-69.  Re-entrant (no shared state), so the caller may run several pair slices on
// several threads.  totals: [0] combos [1] tri-pair tests [2] colliding pairs [3] pairs with >= 1 combo.
// seconds: [0] mid [1] narrow.
void imr_ref_frame_pairs(const float* mats, void* const* trees, const uint32_t* pairs, uint64_t n_pairs,
                         uint64_t* totals, double* seconds) {
    OBBtreesCollision mid;
    CreateUncollideRays narrow;
    uint64_t combos = 0, tests = 0, colliding = 0, with_combos = 0;
    double s_mid = 0.0, s_narrow = 0.0;
    for (uint64_t k = 0; k < n_pairs; ++k) {
        const uint32_t ia = pairs[2 * k], ib = pairs[2 * k + 1];
        auto pr = std::make_pair(make_entry(mats + 16 * uint64_t(ia), nullptr, static_cast<RefTree*>(trees[ia]), true, 1),
                                 make_entry(mats + 16 * uint64_t(ib), nullptr, static_cast<RefTree*>(trees[ib]), true, 2));
        auto t0 = std::chrono::steady_clock::now();
        CDentriesPairTrianglesPairs m = mid.ExecuteOBBtreesCollision(pr);
        auto t1 = std::chrono::steady_clock::now();
        s_mid += std::chrono::duration<double>(t1 - t0).count();
        const auto& cc = m.OBBtreesIntersectInfoObj.candidateTriangleRangeCombinations;
        if (cc.empty()) continue;
        ++with_combos;
        combos += cc.size();
        for (auto& c : cc) tests += uint64_t(c.first_obbtree_count) * uint64_t(c.second_obbtree_count);
        auto t2 = std::chrono::steady_clock::now();
        CDentriesUncollideRays rays = narrow.ExecuteCreateUncollideRays(m);
        auto t3 = std::chrono::steady_clock::now();
        s_narrow += std::chrono::duration<double>(t3 - t2).count();
        if (rays.rays_from_first_to_second.size() || rays.rays_from_second_to_first.size()) ++colliding;   // CollisionDetection.cpp:63
    }
    totals[0] = combos; totals[1] = tests; totals[2] = colliding; totals[3] = with_combos;
    if (seconds) { seconds[0] = s_mid; seconds[1] = s_narrow; }
}

// The same loop with a per-pair verdict: out[5 * k ..] = { non-coplanar hits, coplanar hits, colliding (>= 1 ray, CollisionDetection.cpp:63),
// an order-free 64-bit fingerprint of the pair's hit set over ORIGINAL triangle indices (low word, high word) }.  Lets a test compare a whole
// full-size frame (every pair of the sweep) with the device pair by pair and ask for the hit lists (imr_ref_pair) only where they differ.
static inline uint64_t hit_mix(uint64_t a, uint64_t b) {
    uint64_t x = (a << 32) | b;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
void imr_ref_frame_pairs_detail(const float* mats, void* const* trees, const uint32_t* pairs, uint64_t n_pairs, uint32_t* out) {
    OBBtreesCollision mid;
    CreateUncollideRays narrow;
    for (uint64_t k = 0; k < n_pairs; ++k) {
        const uint32_t ia = pairs[2 * k], ib = pairs[2 * k + 1];
        RefTree* a = static_cast<RefTree*>(trees[ia]); RefTree* b = static_cast<RefTree*>(trees[ib]);
        auto pr = std::make_pair(make_entry(mats + 16 * uint64_t(ia), nullptr, a, true, 1), make_entry(mats + 16 * uint64_t(ib), nullptr, b, true, 2));
        uint32_t* o = out + 5 * k;
        o[0] = o[1] = o[2] = o[3] = o[4] = 0;
        CDentriesPairTrianglesPairs m = mid.ExecuteOBBtreesCollision(pr);
        const auto& cc = m.OBBtreesIntersectInfoObj.candidateTriangleRangeCombinations;
        if (cc.empty()) continue;
        CDentriesUncollideRays rays = narrow.ExecuteCreateUncollideRays(m);
        o[2] = (rays.rays_from_first_to_second.size() || rays.rays_from_second_to_first.size()) ? 1u : 0u;
        const glm::mat4 rel = glm::inverse(pr.first.currentGlobalMatrix) * pr.second.currentGlobalMatrix;      // CreateUncollideRays.cpp:62
        uint64_t fp = 0;
        for (auto& c : cc)
            for (size_t i = 0; i != c.first_obbtree_count; ++i)
                for (size_t j = 0; j != c.second_obbtree_count; ++j) {
                    TrianglePosition ta = a->tree.GetTrianglePosition(i + c.first_obbtree_offset);
                    TrianglePosition tb = rel * b->tree.GetTrianglePosition(j + c.second_obbtree_offset);
                    TrianglesIntersectionInfo info = Triangle::IntersectTriangles(ta, tb);
                    if (info.doIntersept && info.areCoplanar) ++o[1];
                    if (info.doIntersept && !info.areCoplanar) { ++o[0]; fp += hit_mix(a->orig_index[i + c.first_obbtree_offset], b->orig_index[j + c.second_obbtree_offset]); }
                }
        o[3] = uint32_t(fp); o[4] = uint32_t(fp >> 32);
    }
}

// ---- response rays --------------------------------------------------------------
// Ray::IntersectOBBtree (Ray.cpp:136-161) on one ray.  out3 = distanceFromOrigin, baryPosition; tri = leaf-order triangle index.
int imr_ref_ray_tree(void* tree, const float* m16, const float* origin3, const float* dir3, float* out3, uint32_t* tri, int* back) {
    RefTree* t = static_cast<RefTree*>(tree);
    Ray ray(glm::vec3(origin3[0], origin3[1], origin3[2]), glm::vec3(dir3[0], dir3[1], dir3[2]));
    RayOBBtreeIntersectInfo r = ray.IntersectOBBtree(t->tree, load_mat(m16));
    out3[0] = r.distanceFromOrigin; out3[1] = r.baryPosition.x; out3[2] = r.baryPosition.y;
    *tri = uint32_t(r.triangle_index); *back = r.itBackfaces ? 1 : 0;
    return r.doIntersect ? 1 : 0;
}

// One ordered pair through the engine's per-pair code, CollisionDetection.cpp:44-103 (the ECS fan-out after it needs an ECSwrapper and is
// exercised by the host adapter test instead).  delta6 = CollisionCallbackData.deltaVector of first, of second.  Returns colliding (0/1).
// PointMovementBetweenFrames is a private static of CollisionDetection (CollisionDetection.cpp:143-150); its six lines are repeated here.
int imr_ref_pair_delta(void* tree_a, const float* mat_a, const float* prev_a, void* tree_b, const float* mat_b, const float* prev_b,
                       float* delta6, uint64_t* n_rays2) {
    RefTree* a = static_cast<RefTree*>(tree_a); RefTree* b = static_cast<RefTree*>(tree_b);
    OBBtreesCollision mid;
    CreateUncollideRays narrow;
    ShootUncollideRays shoot(glm::radians(40.f), glm::radians(65.f), 1.01f);                 // CollisionDetection.cpp:22-24
    std::memset(delta6, 0, 24);
    auto pr = std::make_pair(make_entry(mat_a, prev_a, a, true, 1), make_entry(mat_b, prev_b, b, true, 2));
    CDentriesPairTrianglesPairs m = mid.ExecuteOBBtreesCollision(pr);
    if (m.OBBtreesIntersectInfoObj.candidateTriangleRangeCombinations.empty()) return 0;   // :51
    CDentriesUncollideRays rays = narrow.ExecuteCreateUncollideRays(m);
    if (n_rays2) { n_rays2[0] = rays.rays_from_first_to_second.size(); n_rays2[1] = rays.rays_from_second_to_first.size(); }
    if (!(rays.rays_from_first_to_second.size() || rays.rays_from_second_to_first.size())) return 0;   // :63
    if (rays.firstEntry.currentGlobalMatrix != rays.firstEntry.previousGlobalMatrix ||
        rays.secondEntry.currentGlobalMatrix != rays.secondEntry.previousGlobalMatrix) {     // :80-81
        glm::vec3 delta = shoot.ExecuteShootUncollideRays(rays);
        auto movement = [](const glm::vec3 point, const glm::mat4& m_first, const glm::mat4& m_second) {
            glm::vec3 p_first = glm::vec3(m_first * glm::vec4(point, 1.f));
            glm::vec3 p_second = glm::vec3(m_second * glm::vec4(point, 1.f));
            glm::vec3 v_diff = p_first - p_second;
            return glm::length(v_diff);
        };
        float first_movement = movement(rays.average_point_first_modelspace, rays.firstEntry.currentGlobalMatrix, rays.firstEntry.previousGlobalMatrix);
        float second_movement = movement(rays.average_point_second_modelspace, rays.secondEntry.currentGlobalMatrix, rays.secondEntry.previousGlobalMatrix);
        float total_movement = first_movement + second_movement;
        glm::vec3 d1 = - delta * (first_movement / total_movement);
        glm::vec3 d2 = + delta * (second_movement / total_movement);
        std::memcpy(delta6, &d1, 12); std::memcpy(delta6 + 3, &d2, 12);
    }
    return 1;
}


// ShootUncollideRays::ExecuteShootUncollideRays (ShootUncollideRays.cpp:14-93) on explicit ray lists (6 floats per ray: origin, direction,
// first's model space), so that a restatement can be compared on the same rays in the same order.
void imr_ref_shoot(void* tree_a, const float* mat_a, void* tree_b, const float* mat_b,
                   const float* rays_first, uint64_t n_first, const float* rays_second, uint64_t n_second, float* delta3) {
    RefTree* a = static_cast<RefTree*>(tree_a); RefTree* b = static_cast<RefTree*>(tree_b);
    ShootUncollideRays shoot(glm::radians(40.f), glm::radians(65.f), 1.01f);                 // CollisionDetection.cpp:22-24
    CDentriesUncollideRays in;
    in.firstEntry = make_entry(mat_a, nullptr, a, true, 1);
    in.secondEntry = make_entry(mat_b, nullptr, b, true, 2);
    for (uint64_t k = 0; k < n_first; ++k) in.rays_from_first_to_second.emplace_back(glm::vec3(rays_first[6 * k], rays_first[6 * k + 1], rays_first[6 * k + 2]), glm::vec3(rays_first[6 * k + 3], rays_first[6 * k + 4], rays_first[6 * k + 5]));
    for (uint64_t k = 0; k < n_second; ++k) in.rays_from_second_to_first.emplace_back(glm::vec3(rays_second[6 * k], rays_second[6 * k + 1], rays_second[6 * k + 2]), glm::vec3(rays_second[6 * k + 3], rays_second[6 * k + 4], rays_second[6 * k + 5]));
    glm::vec3 d = shoot.ExecuteShootUncollideRays(in);
    std::memcpy(delta3, &d, 12);
}

// The rays of one ordered pair in the reference's own order (CreateUncollideRays.cpp:180-184); returns the counts in n2.
void imr_ref_pair_rays(void* tree_a, const float* mat_a, void* tree_b, const float* mat_b, float* rays_first, float* rays_second, uint64_t cap, uint64_t* n2) {
    RefTree* a = static_cast<RefTree*>(tree_a); RefTree* b = static_cast<RefTree*>(tree_b);
    OBBtreesCollision mid;
    CreateUncollideRays narrow;
    auto pr = std::make_pair(make_entry(mat_a, nullptr, a, true, 1), make_entry(mat_b, nullptr, b, true, 2));
    CDentriesPairTrianglesPairs m = mid.ExecuteOBBtreesCollision(pr);
    n2[0] = n2[1] = 0;
    if (m.OBBtreesIntersectInfoObj.candidateTriangleRangeCombinations.empty()) return;
    CDentriesUncollideRays rays = narrow.ExecuteCreateUncollideRays(m);
    n2[0] = rays.rays_from_first_to_second.size(); n2[1] = rays.rays_from_second_to_first.size();
    auto put = [&](const std::vector<Ray>& v, float* out) {
        for (size_t k = 0; k < v.size() && k < cap; ++k) { glm::vec3 o = v[k].GetOrigin(), d = v[k].GetDirection(); std::memcpy(out + 6 * k, &o, 12); std::memcpy(out + 6 * k + 3, &d, 12); }
    };
    if (rays_first) put(rays.rays_from_first_to_second, rays_first);
    if (rays_second) put(rays.rays_from_second_to_first, rays_second);
}

// Triangle::CreateTriangleList (Triangle.cpp:259-280) on one primitive; returns the triangle count, fills the arrays when pos9 != nullptr.
uint64_t imr_ref_triangle_list(const float* points, uint64_t n_points, const float* normals, const uint32_t* indices, uint64_t n_indices, uint32_t mode,
                               float* pos9, float* nrm9, uint32_t* vid3) {
    std::vector<glm::vec3> pts(n_points), nrm;
    for (uint64_t i = 0; i < n_points; ++i) pts[i] = glm::vec3(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
    if (normals) { nrm.resize(n_points); for (uint64_t i = 0; i < n_points; ++i) nrm[i] = glm::vec3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]); }
    std::vector<uint32_t> idx(indices, indices + n_indices);
    std::vector<Triangle> tris = Triangle::CreateTriangleList(pts, nrm, idx, static_cast<glTFmode>(mode));
    if (pos9) {
        for (size_t i = 0; i < tris.size(); ++i)
            for (int q = 0; q < 3; ++q) {
                glm::vec3 p = tris[i].GetP(q), n = tris[i].GetN(q);
                std::memcpy(pos9 + 9 * i + 3 * q, &p, 12); std::memcpy(nrm9 + 9 * i + 3 * q, &n, 12);
                vid3[3 * i + q] = tris[i].GetI(q);
            }
    }
    return tris.size();
}

const char* imr_ref_build_info() { return "reference sources compiled in place: g++ -std=c++20 -O2 -ffp-contract=off"; }

}  // extern "C"
