// imrcd_rays.cu -- the response stage of a frame: "uncollide" rays shot through the two OBB trees
//   Ray::IntersectOBBtree / IntersectParalgram / IntersectTriangle        IMR/src/Geometry/Ray.cpp:13-236, glm fork gtx/intersect.inl:29-97
//   ShootUncollideRays::ExecuteShootUncollideRays / HermannPass / ...     IMR/src/CollisionDetection/ShootUncollideRays.cpp:14-179
//   deltaVector of both entities                                           IMR/src/CollisionDetection/CollisionDetection.cpp:80-103,143-150
// One thread per ray of a colliding pair whose entities moved since the last frame (the others get a zero deltaVector, :99-103):
// its Hermann pass, and when that succeeds the reflected ray's pass in the opposite direction (:73-89).  Tree descent is the reference's
// recursion made iterative with a per-thread stack: near child first, the far child re-checked against the best distance when it is popped,
// which is exactly when the recursion evaluates its condition (Ray.cpp:182-204), so the hit found is the same triangle bit for bit.
// Then one warp per colliding pair: FindResponse (:150-172) over the pair's responses and the split of the delta by movement.
#include "imrcd_internal.cuh"
#include <cfloat>
#include <cstdlib>
#include <algorithm>

#define FULL_MASK 0xffffffffu
#define RAY_STACK 128        // per-lane stack of the ray descent: one entry per level (the far child waits while the near one is walked).
                             // The Morton build's trees are at most ~70 deep (54 key bits + duplicates); imported trees deeper than this are
                             // refused at import (imrcd_mesh_import_tree), so a frame never meets a tree it cannot walk

struct RayHit { bool hit, back; float dist; uint32_t tri; float bx, by; };      // RayOBBtreeIntersectInfo, Ray.h:17-24

// glm fork, glm/gtx/intersect.inl:29-97 (intersectRayTriangle with itBackfaces); the ray's origin is (0,0,0) (Ray.cpp:138,153-158)
__device__ __forceinline__ bool ray_triangle0(V3 dir, V3 v0, V3 v1, V3 v2, float& bx, float& by, float& distance, bool& back) {
    const V3 edge1 = sub3(v1, v0), edge2 = sub3(v2, v0);
    const V3 p = cross3(dir, edge2);
    const float det = dot3(edge1, p);
    const V3 dist = sub3(mk3(0.f, 0.f, 0.f), v0);
    V3 perp;
    float x, y;
    if (det > FLT_EPSILON) {
        x = dot3(dist, p);
        if (x < 0.f || x > det) return false;
        perp = cross3(dist, edge1);
        y = dot3(dir, perp);
        if (y < 0.f || (x + y) > det) return false;
        back = false;
    } else if (det < -FLT_EPSILON) {
        x = dot3(dist, p);
        if (x > 0.f || x < det) return false;
        perp = cross3(dist, edge1);
        y = dot3(dir, perp);
        if (y > 0.f || (x + y) < det) return false;
        back = true;
    } else return false;
    const float inv_det = 1.f / det;
    distance = dot3(edge2, perp) * inv_det;
    bx = x * inv_det; by = y * inv_det;
    return true;
}

// one slab of Ray::IntersectParalgram, Ray.cpp:48-75 (the U, V and W blocks are the same code)
__device__ __forceinline__ bool ray_slab(V3 a, V3 b, V3 side, V3 ray_origin, V3 dir, float& mn, float& mx) {
    const V3 plane_dir = normalize3(cross3(a, b));
    const float d = -fabsf(dot3(plane_dir, side));
    const float c = dot3(plane_dir, ray_origin);
    const float v_n1 = +c + d;
    const float v_n2 = -c + d;
    const float vd = dot3(plane_dir, dir);
    if (fabsf(vd) >= FLT_EPSILON) {
        const float vd_inv = 1.f / vd;
        float t1 = -v_n1 * vd_inv;
        float t2 = +v_n2 * vd_inv;
        if (t1 > t2) { const float t = t1; t1 = t2; t2 = t; }
        mn = (t1 < mn) ? mn : t1;                  // std::max(t1, min_distance)
        mx = (mx < t2) ? mx : t2;                  // std::min(t2, max_distance)
        if (mn > mx || mx < 0.f) return false;
    } else if (v_n1 > 0 || v_n2 > 0) return false;
    return true;
}
// Ray::IntersectParalgram, Ray.cpp:38-134, origin (0,0,0)
__device__ __forceinline__ bool ray_box0(V3 dir, const Box& bx, float& tmin, float& tmax) {
    float mn = -INFINITY, mx = +INFINITY;
    const V3 ro = sub3(mk3(0.f, 0.f, 0.f), bx.c);
    if (!ray_slab(bx.v, bx.w, bx.u, ro, dir, mn, mx)) return false;       // U test
    if (!ray_slab(bx.w, bx.u, bx.v, ro, dir, mn, mx)) return false;       // V test
    if (!ray_slab(bx.u, bx.v, bx.w, ro, dir, mn, mx)) return false;       // W test
    tmin = mn; tmax = mx;
    return true;
}

// Ray::IntersectOBBtree + IntersectOBBtreeRecursive, Ray.cpp:136-236.  `m` maps the tree into the ray's space.
__device__ RayHit ray_tree(const TreeRec* __restrict__ recs, const TriRec* __restrict__ tris, Rel m, V3 origin, V3 dir, unsigned int* overflow) {
    RayHit best; best.hit = false; best.back = false; best.dist = INFINITY; best.tri = 0xffffffffu; best.bx = 0.f; best.by = 0.f;
    if (!(origin.x == 0.f && origin.y == 0.f && origin.z == 0.f)) {       // centered_matrix[3] -= vec4(origin, 0)  (:153-156)
        m.r0.w = m.r0.w - origin.x; m.r1.w = m.r1.w - origin.y; m.r2.w = m.r2.w - origin.z;
    }
    uint32_t st_node[RAY_STACK]; float st_min[RAY_STACK];
    int sp = 0;
    {
        const float4* rp = reinterpret_cast<const float4*>(recs);
        const Box root = box_transform(m, unpack_box(__ldg(rp), __ldg(rp + 1), __ldg(rp + 2)));
        float mn, mx;
        if (!(ray_box0(dir, root, mn, mx) && mx >= 0.f)) return best;    // :144-147
        st_node[0] = 0u; st_min[0] = -INFINITY; sp = 1;
    }
    while (sp > 0) {
        --sp;
        const uint32_t node = st_node[sp];
        if (!(st_min[sp] < best.dist)) continue;                          // the recursion's `min < best_so_far` at call time
        const float4 q3 = __ldg(reinterpret_cast<const float4*>(recs + node) + 3);
        const uint32_t child = __float_as_uint(q3.y);
        if (__float_as_uint(q3.w) == 0u) {                                // inner: children are records child (left), child + 1 (right)
            const float4* lp = reinterpret_cast<const float4*>(recs + child);
            const Box lb = box_transform(m, unpack_box(__ldg(lp), __ldg(lp + 1), __ldg(lp + 2)));
            const Box rb = box_transform(m, unpack_box(__ldg(lp + 4), __ldg(lp + 5), __ldg(lp + 6)));
            float lmin = 0.f, lmax = 0.f, rmin = 0.f, rmax = 0.f;
            const bool lh = ray_box0(dir, lb, lmin, lmax), rh = ray_box0(dir, rb, rmin, rmax);
            const bool lgo = lh && lmax >= 0.f, rgo = rh && rmax >= 0.f;
            // visit order (:182-204): both hit -> the one with the smaller entry distance first (left on lmin < rmin, else right)
            const bool left_first = !(lh && rh) || (lmin < rmin);
            if (sp + 2 > RAY_STACK) { atomicOr(overflow, (unsigned)OVF_RAYSTACK); return best; }
            if (left_first) {
                if (rgo) { st_node[sp] = child + 1u; st_min[sp] = rmin; ++sp; }
                if (lgo) { st_node[sp] = child; st_min[sp] = lmin; ++sp; }
            } else {
                if (lgo) { st_node[sp] = child; st_min[sp] = lmin; ++sp; }
                if (rgo) { st_node[sp] = child + 1u; st_min[sp] = rmin; ++sp; }
            }
        } else {
            const uint32_t cnt = __float_as_uint(q3.z);
            for (uint32_t i = 0; i < cnt; ++i) {                           // :221-234
                const float4* tp = reinterpret_cast<const float4*>(tris + child + i);
                const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                const V3 p0 = rel_mul(m, mk3(t0.x, t0.y, t0.z), 1.f), p1 = rel_mul(m, mk3(t1.x, t1.y, t1.z), 1.f), p2 = rel_mul(m, mk3(t2.x, t2.y, t2.z), 1.f);
                float bx = 0.f, by = 0.f, dist = INFINITY; bool back = false;
                if (ray_triangle0(dir, p0, p1, p2, bx, by, dist, back) && dist > 0.f && dist < best.dist) {
                    best.hit = true; best.back = back; best.dist = dist; best.tri = child + i; best.bx = bx; best.by = by;
                }
            }
        }
    }
    return best;
}

struct Hermann { bool ok; V3 response, normal; };

// ShootUncollideRays::HermannPass, ShootUncollideRays.cpp:116-148.  A = the ray's own object, B = the other one; matA / matB map them into the
// ray's space (first's model space), nmatB the normal matrix of B.
__device__ Hermann hermann_pass(const TreeRec* recsA, const TriRec* trisA, const Rel& matA, const TreeRec* recsB, const TriRec* trisB, const Rel& matB,
                                const M3& nmatB, const float* __restrict__ tri_nrm, uint32_t triB_base, V3 origin, V3 dir, unsigned int* overflow) {
    Hermann r; r.ok = false; r.response = mk3(0.f, 0.f, 0.f); r.normal = r.response;
    const RayHit p2 = ray_tree(recsB, trisB, matB, origin, dir, overflow);
    if (p2.hit && p2.back) {
        // Ray::MoveOriginEpsilonTowardsDirection(4.f), Ray.cpp:13-21
        float big = fabsf(origin.x); if (big < fabsf(origin.y)) big = fabsf(origin.y); if (big < fabsf(origin.z)) big = fabsf(origin.z);
        const float scaled = big * FLT_EPSILON;
        const V3 moved = add3(origin, scale3(dir, 4.f * scaled));
        const float eps_dist = length3(sub3(moved, origin));
        const RayHit p3 = ray_tree(recsA, trisA, matA, moved, dir, overflow);
        if (p2.dist <= p3.dist + eps_dist) {
            r.ok = true;
            r.response = scale3(dir, p2.dist);
            const float* nn = tri_nrm + 9ull * (triB_base + p2.tri);
            const V3 in = tri_interp_normal(mk3(nn[0], nn[1], nn[2]), mk3(nn[3], nn[4], nn[5]), mk3(nn[6], nn[7], nn[8]), p2.bx, p2.by);
            r.normal = normalize3(m3_mul(nmatB, in));                      // Triangle.cpp:197-204
        }
    }
    return r;
}

__device__ __forceinline__ Rel rel_identity() { Rel r; r.r0 = make_float4(1.f, 0.f, 0.f, 0.f); r.r1 = make_float4(0.f, 1.f, 0.f, 0.f); r.r2 = make_float4(0.f, 0.f, 1.f, 0.f); return r; }

// box `idx` of a tree, moved into the ray's space, against the ray from the origin (Ray.cpp:169-173); one copy of the slab code
__device__ __noinline__ bool ray_node_box(const TreeRec* __restrict__ recs, uint32_t idx, const Rel& m, V3 dir, float& mn, float& mx) {
    const float4* rp = reinterpret_cast<const float4*>(recs + idx);
    const Box b = box_transform(m, unpack_box(__ldg(rp), __ldg(rp + 1), __ldg(rp + 2)));
    return ray_box0(dir, b, mn, mx);
}

// One lane per kept ray: ExecuteShootUncollideRays' two lambdas (ShootUncollideRays.cpp:29-89) as a state machine.  A ray goes through up to
// four tree queries (step 0: the other object, then its own; step 1, the reflected ray: the same with the roles swapped).  A lane is always
// in one of four states - a box test due (the root of a new query, or the two children of an inner node), a leaf due, a query finished
// (Hermann bookkeeping due), or out of rays - and every iteration of the warp's loop runs the ONE kind of work that most lanes are waiting
// for, the others sitting the iteration out: each code path (boxes ~900 instructions, leaf ~600, bookkeeping ~200) executes with as many lanes
// as can take it instead of all three running back to back for a handful of lanes each.  The order of events inside one lane's query is the
// recursion's, so the hit is the reference's bit for bit.  (History: four inlined descents per ray, 12/32 lanes and 65 % of the issue slots
// lost to instruction-cache misses: 9.2 ms for 669 k rays on C3; one node per lane per iteration, all paths every iteration: 5.7 ms; majority
// path + rays handed out dynamically, since their costs differ by two orders of magnitude + 6 blocks per SM: 3.7 ms; C2: 28 -> 10 ms.)
enum { LANE_POP = 0, LANE_BOX = 1, LANE_LEAF = 2, LANE_ROOT = 3, LANE_DONE = 4 };

__global__ void __launch_bounds__(128, 6)
k_shoot(FrameCtl* ctl, unsigned long long cap_rays, const RayRec* __restrict__ rays, float4* __restrict__ resp, PairAcc* acc, const PairRec* __restrict__ pairrec,
        const TreeRec* __restrict__ recs, const TriRec* __restrict__ tris, const float* __restrict__ tri_nrm) {
    if (ctl->overflow) return;
    const unsigned long long n = ctl->n_rays_kept < cap_rays ? ctl->n_rays_kept : cap_rays;
    const uint32_t lane = threadIdx.x & 31u;
    // Lanes per warp that take rays: all 32 when the frame has a ray for every lane of the grid; fewer when it has not, so that a small
    // frame's rays spread over as many warps as there are (a warp runs its lanes' queries interleaved, one code path at a time: 32 rays in one
    // warp take ~32/14 times one ray's latency, and a game-sized frame of a few hundred rays used to sit in ten warps while 3500 idled).
    const unsigned long long total_warps = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    const unsigned long long per_warp = (n + total_warps - 1ull) / total_warps;
    const uint32_t quota = per_warp < 1ull ? 1u : (per_warp > 32ull ? 32u : (uint32_t)per_warp);
    bool exhausted = lane >= quota;                                         // the frame's rays have all been handed out (or: not this lane's to take)
    uint32_t st_node[RAY_STACK]; float st_min[RAY_STACK];
    int sp = 0;
    bool active = false;
    int kind = LANE_POP;
    uint32_t cur = 0, cur_cnt = 0;                                          // the node / leaf range the lane is about to work on
    // per-ray state
    unsigned long long k = 0; uint32_t p = 0, side = 0; int step = 0, phase = 0;
    V3 o = mk3(0, 0, 0), d = mk3(0, 0, 0), q_origin = mk3(0, 0, 0);
    Rel rel, m; rel.r0 = rel.r1 = rel.r2 = make_float4(0, 0, 0, 0); m = rel;
    uint32_t recA = 0, recB = 0, triA = 0, triB = 0;                        // first's / second's arena bases
    const TreeRec* q_recs = recs; const TriRec* q_tris = tris;             // the tree of the running query
    RayHit best, p2;
    best.hit = false; best.back = false; best.dist = INFINITY; best.tri = 0xffffffffu; best.bx = best.by = 0.f; p2 = best;
    float eps_dist = 0.f;
    double fx = 0.0, fy = 0.0, fz = 0.0; uint32_t n_ok = 0;
    float4 out0 = make_float4(0, 0, 0, 0), out1 = out0;

    for (;;) {
        // idle lanes take the next rays of the frame (one atomic per warp)
        const uint32_t m_idle = __ballot_sync(FULL_MASK, !active && !exhausted);
        if (m_idle) {
            const uint32_t leader = (uint32_t)__ffs(m_idle) - 1u;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(&ctl->ray_cursor, (unsigned long long)__popc(m_idle));
            base = __shfl_sync(FULL_MASK, base, leader);
            if (!active && !exhausted) { k = base + __popc(m_idle & ((1u << lane) - 1u)); if (k >= n) exhausted = true; }
        }
        if (!active && !exhausted) {
            const RayRec ray = rays[k];
            p = __float_as_uint(ray.o.w); side = __float_as_uint(ray.d.w);
            const PairRec pr = pairrec[p];
            rel.r0 = pr.r0; rel.r1 = pr.r1; rel.r2 = pr.r2; recA = pr.recA; recB = pr.recB; triA = pr.triA; triB = pr.triB;
            o = mk3(ray.o.x, ray.o.y, ray.o.z); d = mk3(ray.d.x, ray.d.y, ray.d.z);
            step = 0; phase = 0; fx = fy = fz = 0.0; n_ok = 0; out0 = out1 = make_float4(0, 0, 0, 0);
            active = true; kind = LANE_ROOT; q_origin = o;
        }
        // next event of the lane's query: drop pruned stack entries (the recursion's `min < best_so_far` at call time, Ray.cpp:182-216)
        if (active && kind == LANE_POP) {
            kind = LANE_DONE;
            while (sp > 0) {
                --sp;
                if (st_min[sp] < best.dist) {
                    const uint32_t node = st_node[sp];
                    const float4 q3 = __ldg(reinterpret_cast<const float4*>(q_recs + node) + 3);
                    cur = __float_as_uint(q3.y); cur_cnt = __float_as_uint(q3.z);
                    kind = __float_as_uint(q3.w) == 0u ? LANE_BOX : LANE_LEAF;
                    break;
                }
            }
        }
        const uint32_t m_box = __ballot_sync(FULL_MASK, active && (kind == LANE_BOX || kind == LANE_ROOT));
        const uint32_t m_leaf = __ballot_sync(FULL_MASK, active && kind == LANE_LEAF);
        const uint32_t m_done = __ballot_sync(FULL_MASK, active && kind == LANE_DONE);
        if ((m_box | m_leaf | m_done) == 0u) break;                         // no lane has a ray left
        const int n_box = __popc(m_box), n_leaf = __popc(m_leaf), n_done = __popc(m_done);
        const int path = (n_box >= n_leaf && n_box >= n_done) ? 0 : (n_leaf >= n_done ? 1 : 2);

        // f2s: first_to_second_ray_execute (objects: A = first, B = second), else second_to_first (A = second, B = first)  (:29-71)
        const bool f2s = (side == 0u) == (step == 0);
        if (path == 0) {
            if (active && kind == LANE_ROOT) {
                // start a query (Ray::IntersectOBBtree, Ray.cpp:136-161): phase 0 looks at B (point2), phase 1 at A (point3)
                const bool on_second = (phase == 0) == f2s;                 // which tree: second's (matrix rel) or first's (identity)
                q_recs = recs + (on_second ? recB : recA); q_tris = tris + (on_second ? triB : triA);
                m = on_second ? rel : rel_identity();
                if (!(q_origin.x == 0.f && q_origin.y == 0.f && q_origin.z == 0.f)) {      // centered_matrix[3] -= vec4(origin, 0)
                    m.r0.w = m.r0.w - q_origin.x; m.r1.w = m.r1.w - q_origin.y; m.r2.w = m.r2.w - q_origin.z;
                }
                best.hit = false; best.back = false; best.dist = INFINITY; best.tri = 0xffffffffu; best.bx = best.by = 0.f;
                sp = 0;
            }
            if (active && (kind == LANE_BOX || kind == LANE_ROOT)) {
                const bool root = kind == LANE_ROOT;
                float lmin = 0.f, lmax = 0.f, rmin = 0.f, rmax = 0.f;
                const bool lh = ray_node_box(q_recs, root ? 0u : cur, m, d, lmin, lmax);
                bool rh = false;
                if (!root) rh = ray_node_box(q_recs, cur + 1u, m, d, rmin, rmax);
                if (root) {
                    if (lh && lmax >= 0.f) { st_node[0] = 0u; st_min[0] = -INFINITY; sp = 1; }     // :144-147
                } else {
                    const bool lgo = lh && lmax >= 0.f, rgo = rh && rmax >= 0.f;
                    const bool left_first = !(lh && rh) || (lmin < rmin);   // :182-204
                    if (sp + 2 > RAY_STACK) { atomicOr(&ctl->overflow, (unsigned)OVF_RAYSTACK); sp = 0; }
                    else if (left_first) {
                        if (rgo) { st_node[sp] = cur + 1u; st_min[sp] = rmin; ++sp; }
                        if (lgo) { st_node[sp] = cur; st_min[sp] = lmin; ++sp; }
                    } else {
                        if (lgo) { st_node[sp] = cur; st_min[sp] = lmin; ++sp; }
                        if (rgo) { st_node[sp] = cur + 1u; st_min[sp] = rmin; ++sp; }
                    }
                }
                kind = LANE_POP;
            }
        } else if (path == 1) {
            if (active && kind == LANE_LEAF) {
                for (uint32_t i = 0; i < cur_cnt; ++i) {                    // :221-234
                    const float4* tp = reinterpret_cast<const float4*>(q_tris + cur + i);
                    const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                    const V3 p0 = rel_mul(m, mk3(t0.x, t0.y, t0.z), 1.f), p1 = rel_mul(m, mk3(t1.x, t1.y, t1.z), 1.f), p2v = rel_mul(m, mk3(t2.x, t2.y, t2.z), 1.f);
                    float bx = 0.f, by = 0.f, dist = INFINITY; bool back = false;
                    if (ray_triangle0(d, p0, p1, p2v, bx, by, dist, back) && dist > 0.f && dist < best.dist) {
                        best.hit = true; best.back = back; best.dist = dist; best.tri = cur + i; best.bx = bx; best.by = by;
                    }
                }
                kind = LANE_POP;
            }
        } else if (active && kind == LANE_DONE) {
            // the query is over: HermannPass (ShootUncollideRays.cpp:116-148)
            bool ray_done = false;
            if (phase == 0) {
                if (best.hit && best.back) {                                // point2: the other object, hit from inside
                    p2 = best;
                    // Ray::MoveOriginEpsilonTowardsDirection(4.f), Ray.cpp:13-21
                    float big = fabsf(o.x); if (big < fabsf(o.y)) big = fabsf(o.y); if (big < fabsf(o.z)) big = fabsf(o.z);
                    const float scaled = big * FLT_EPSILON;
                    q_origin = add3(o, scale3(d, 4.f * scaled));
                    eps_dist = length3(sub3(q_origin, o));
                    phase = 1; kind = LANE_ROOT;
                } else ray_done = true;
            } else {
                if (p2.dist <= best.dist + eps_dist) {                      // :134 (best = point3, the ray's own object)
                    const V3 response = scale3(d, p2.dist);
                    const uint32_t tri_base = f2s ? triB : triA;           // object B of this pass
                    const float* nn = tri_nrm + 9ull * (tri_base + p2.tri);
                    const V3 in = tri_interp_normal(mk3(nn[0], nn[1], nn[2]), mk3(nn[3], nn[4], nn[5]), mk3(nn[6], nn[7], nn[8]), p2.bx, p2.by);
                    const V3 normal = normalize3(f2s ? m3_mul(adjoint_transpose3(rel), in) : m3_mul(m3_identity(), in));     // Triangle.cpp:197-204
                    // CalcForceResponse (:104-114)
                    const V3 dirn = normalize3(response);
                    const float len = length3(response);
                    const float c = dot3(normal, dirn);
                    const V3 fr = scale3(dirn, (c * c) * len);
                    float4 rec;
                    if (f2s) { fx -= (double)fr.x; fy -= (double)fr.y; fz -= (double)fr.z; rec = make_float4(-response.x, -response.y, -response.z, 1.f); }
                    else { fx += (double)fr.x; fy += (double)fr.y; fz += (double)fr.z; rec = make_float4(response.x, response.y, response.z, 1.f); }
                    if (step == 0) out0 = rec; else out1 = rec;
                    ++n_ok;
                    if (step == 0) {                                        // ReflectHermannResult (:95-101), then the opposite direction (:73-89)
                        o = add3(o, response); d = mk3(-normal.x, -normal.y, -normal.z);
                        step = 1; phase = 0; kind = LANE_ROOT; q_origin = o;
                    } else ray_done = true;
                } else ray_done = true;
            }
            if (ray_done) {
                resp[2 * k] = out0; resp[2 * k + 1] = out1;
                if (n_ok) {
                    atomicAdd(&acc[p].force[0], fx); atomicAdd(&acc[p].force[1], fy); atomicAdd(&acc[p].force[2], fz);
                    atomicAdd(&acc[p].n_resp, n_ok);
                }
                active = false; kind = LANE_POP;
            }
        }
    }
}

// One warp per colliding entity pair: FindResponse (ShootUncollideRays.cpp:150-172) over the pair's responses, the world-space delta (:85-86)
// and its split between the two entities by how far each moved at its contact point (CollisionDetection.cpp:85-96,143-150).
__global__ void __launch_bounds__(256)
k_delta(FrameCtl* ctl, const uint32_t* __restrict__ epair_pair, imrcd_entity_pair* __restrict__ out, const PairAcc* __restrict__ acc,
        const float4* __restrict__ resp, const float* __restrict__ cur, const float* __restrict__ prev, float edge_a, float edge_b,
        const uint2* __restrict__ pairs) {
    if (ctl->overflow) return;
    const unsigned long long n = ctl->n_colliding;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long n_resp_total = 0;
    for (unsigned long long e = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5; e < n; e += ((unsigned long long)gridDim.x * blockDim.x) >> 5) {
        const uint32_t p = epair_pair[e];
        const PairAcc a = acc[p];
        if (!(a.flags & PAIR_MOVED)) continue;                              // deltaVector stays (0,0,0) (:99-103)
        const V3 force = mk3((float)a.force[0], (float)a.force[1], (float)a.force[2]);
        V3 local = mk3(0.f, 0.f, 0.f);
        if (!(force.x == 0.f && force.y == 0.f && force.z == 0.f)) {        // :81
            const V3 nf = normalize3(force);
            float max_len = 0.f;
            for (int part = 0; part < 2; ++part) {
                const uint32_t base = part ? a.ray_off_b : a.ray_off_a, cnt = part ? a.rays_b : a.rays_a;
                for (uint32_t k = lane; k < 2u * cnt; k += 32u) {
                    const float4 rr = resp[2ull * base + k];
                    if (rr.w == 0.f) continue;
                    const V3 rv = mk3(rr.x, rr.y, rr.z);
                    const V3 nr = normalize3(rv);
                    const float c = dot3(nr, nf);
                    const float len = length3(rv);
                    const float need = len / c;
                    const float tq = (c - edge_a) / (edge_b - edge_a);
                    const float tmp = tq < 0.f ? 0.f : (1.f < tq ? 1.f : tq);                // std::clamp
                    const float ss = tmp * tmp * tmp * (tmp * (tmp * 6 - 15) + 10);          // SmootherStep :174-178
                    const float cand = ss * need;
                    max_len = max_len < cand ? cand : max_len;
                }
            }
            for (int o = 16; o > 0; o >>= 1) { const float v = __shfl_xor_sync(FULL_MASK, max_len, o); max_len = max_len < v ? v : max_len; }
            local = scale3(scale3(nf, max_len), 1.01f);                     // ray_distance_bias_multiplier (CollisionDetection.cpp:24)
        }
        if (lane == 0) {
            imrcd_entity_pair o = out[e];
            const uint2 pr = pairs[p];                                      // this context's entry slots (o.entry_* are the caller's indices)
            const float* m1 = cur + 16 * (size_t)pr.x; const float* m2 = cur + 16 * (size_t)pr.y;
            const float* q1 = prev + 16 * (size_t)pr.x; const float* q2 = prev + 16 * (size_t)pr.y;
            const V3 delta = rel_mul(rel_from_mat(m1), local, 0.f);         // first.current * vec4(localspace_response, 0)
            const V3 pa = mk3(o.avg_first[0], o.avg_first[1], o.avg_first[2]), pb = mk3(o.avg_second[0], o.avg_second[1], o.avg_second[2]);
            const float mv1 = length3(sub3(rel_mul(rel_from_mat(m1), pa, 1.f), rel_mul(rel_from_mat(q1), pa, 1.f)));
            const float mv2 = length3(sub3(rel_mul(rel_from_mat(m2), pb, 1.f), rel_mul(rel_from_mat(q2), pb, 1.f)));
            const float total = mv1 + mv2;
            const float f1 = mv1 / total, f2 = mv2 / total;
            o.delta_first[0] = -delta.x * f1; o.delta_first[1] = -delta.y * f1; o.delta_first[2] = -delta.z * f1;       // :88
            o.delta_second[0] = delta.x * f2; o.delta_second[1] = delta.y * f2; o.delta_second[2] = delta.z * f2;       // :89
            out[e] = o;
            n_resp_total += a.n_resp;
        }
    }
    if (lane == 0 && n_resp_total) atomicAdd(&ctl->n_responses, n_resp_total);
}

int imr_frame_shoot_device(imrcd_ctx* ctx, FrameCtl* ctl, uint64_t* launches) {
    if (!ctx->prev_distinct) return IMRCD_OK;                              // nothing moved: every deltaVector is zero
    cudaStream_t s = ctx->stream;
    // glm::radians(40.f), glm::radians(65.f) -> cos (CollisionDetection.cpp:22-23, ShootUncollideRays.cpp:8-9): smoothstep runs from cos 65 to cos 40
    const float edge_b = cosf(0.01745329251994329576923690768489f * 40.f), edge_a = cosf(0.01745329251994329576923690768489f * 65.f);
    if (!ctx->shoot_blocks) {                                             // persistent grid: as many blocks as are co-resident
        int per_sm = 0;
        IMR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_shoot, 128, 0));
        ctx->shoot_blocks = ctx->sm_count * (per_sm < 1 ? 1 : per_sm);
    }
    k_shoot<<<ctx->shoot_blocks, 128, 0, s>>>(ctl, ctx->cap_rays, ctx->d_rays.as<RayRec>(), ctx->d_resp.as<float4>(), ctx->d_pairacc.as<PairAcc>(),
                                               ctx->d_pairrec.as<PairRec>(), ctx->d_recs.as<TreeRec>(), ctx->d_tris.as<TriRec>(), ctx->d_tri_nrm.as<float>());
    k_delta<<<ctx->sm_count * 4, 256, 0, s>>>(ctl, ctx->d_epair_pair.as<uint32_t>(), ctx->d_epairs.as<imrcd_entity_pair>() + 1, ctx->d_pairacc.as<PairAcc>(),
                                               ctx->d_resp.as<float4>(), ctx->d_cur.as<float>(), ctx->d_prev.as<float>(), edge_a, edge_b, ctx->d_pairs.as<uint2>());
    *launches += 2;
    return IMRCD_OK;
}

// ---- unit-level entry point: Ray::IntersectOBBtree on n rays against one mesh ----
__global__ void k_test_ray_tree(uint64_t n, const TreeRec* recs, const TriRec* tris, const float* mats, const float* origins, const float* dirs,
                                uint8_t* flags, float* out3, uint32_t* tri, unsigned int* overflow) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RayHit h = ray_tree(recs, tris, rel_from_mat(mats + 16 * i), mk3(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]),
                              mk3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]), overflow);
    flags[i] = (h.hit ? 1 : 0) | (h.back ? 2 : 0);
    out3[3 * i] = h.dist; out3[3 * i + 1] = h.bx; out3[3 * i + 2] = h.by;
    tri[i] = h.tri;
}

int imr_test_ray_tree_device(imrcd_ctx* ctx, uint32_t mesh_id, uint64_t n, const float* mats, const float* origins, const float* dirs,
                             uint8_t* flags, float* out3, uint32_t* tri) {
    const MeshDev md = ctx->meshes[mesh_id].dev;
    unsigned int* ovf = nullptr;
    IMR_CUDA(ctx, cudaMalloc(&ovf, 4));
    IMR_CUDA(ctx, cudaMemsetAsync(ovf, 0, 4, ctx->stream));
    k_test_ray_tree<<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(n, ctx->d_recs.as<TreeRec>() + md.rec_base, ctx->d_tris.as<TriRec>() + md.tri_base,
                                                                       mats, origins, dirs, flags, out3, tri, ovf);
    unsigned int h = 0;
    IMR_CUDA(ctx, cudaMemcpyAsync(&h, ovf, 4, cudaMemcpyDeviceToHost, ctx->stream));
    IMR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ovf);
    if (h) { ctx->err = "ray stack overflow (tree deeper than the per-thread stack)"; return IMRCD_E_CAPACITY; }
    return IMRCD_OK;
}
