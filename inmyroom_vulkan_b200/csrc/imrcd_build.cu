// imrcd_build.cu -- GPU OBB-tree construction: the replacement for OBBtree::OBBtree(std::vector<Triangle>&&)
// (IMR/src/Geometry/OBBtree.cpp:321-358 and the recursive OBBtreeSplitBuildNode ctor :8-108).
//
// IMRCD_BUILD_MORTON (default): bottom-up, Morton-ordered.
//   1. centroid bounds, 63-bit Morton keys, radix sort                         (k_bounds, k_morton, cub)
//   2. triangles gathered into sorted order = leaf order                       (k_gather_tris)
//   3. binary radix tree over the sorted keys (Karras 2012)                    (k_radix_tree)
//   4. subtrees of <= 4 triangles collapse into leaves (OBBtree.h:49); kept inner nodes get a
//      children pair slot by an exclusive scan -> sibling-adjacent 64-B records (k_flag_inner, cub scan, k_assign)
//   5. the fit of every box (imrcd_fit.cu): treelets of <= 128 triangles in shared memory (FP64 moments bottom-up, closed-form symmetric
//      3x3 eigen-solve, projections, boxes), then the few nodes above the treelets; boxes rounded to FP32 outward so that they contain
//      their triangles (the reference's +2*FLT_EPSILON pad, OBB.cpp:123, does not guarantee that)
// Nothing in the build waits for the device: the record count stays on the device until the end (the arena is reserved for the worst case
// and trimmed afterwards).
// The tree differs from the reference's top-down tree by construction; parity for this mode is asserted on
// tree-independent outputs (tests/test_gpu_build.py).
#include "imrcd_internal.cuh"
#include "imrcd_fit.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cfloat>
#include <cstring>

int imr_mesh_arena_alloc(imrcd_ctx* ctx, uint64_t n_rec, uint64_t n_tri, MeshHost* mh);
int imr_mesh_arena_reserve(imrcd_ctx* ctx, uint64_t n_rec, uint64_t n_tri);
int imr_build_mesh_reference(imrcd_ctx* ctx, const float* pos, const float* nrm, const uint32_t* vid, uint64_t n_tri, MeshHost* mh);

#define FULL_MASK 0xffffffffu

// ---- sortable encodings for atomic min/max ---------------------------------------------------
__device__ __forceinline__ unsigned long long f64_sortable(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_unsortable(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)u);
}
__device__ __forceinline__ uint32_t f32_sortable(float f) { uint32_t u = __float_as_uint(f); return (u >> 31) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float f32_unsortable(uint32_t u) { u = (u >> 31) ? (u & 0x7fffffffu) : ~u; return __uint_as_float(u); }

// ---- 1. bounds + Morton keys ---------------------------------------------------------------------
__global__ void k_bounds_init(uint32_t* b) { if (threadIdx.x < 3) b[threadIdx.x] = 0xffffffffu; else if (threadIdx.x < 6) b[threadIdx.x] = 0u; }

__global__ void k_bounds(uint32_t n, const float* __restrict__ pos, uint32_t* bounds) {
    float mn[3] = { INFINITY, INFINITY, INFINITY }, mx[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const float* p = pos + 9ull * t;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float c = (p[a] + p[3 + a] + p[6 + a]) * (1.0f / 3.0f);
            mn[a] = fminf(mn[a], c); mx[a] = fmaxf(mx[a], c);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int o = 16; o > 0; o >>= 1) { mn[a] = fminf(mn[a], __shfl_xor_sync(FULL_MASK, mn[a], o)); mx[a] = fmaxf(mx[a], __shfl_xor_sync(FULL_MASK, mx[a], o)); }
        if ((threadIdx.x & 31) == 0) { atomicMin(&bounds[a], f32_sortable(mn[a])); atomicMax(&bounds[3 + a], f32_sortable(mx[a])); }
    }
}

#define MORTON_BITS 18                   // per axis: 54-bit keys, seven 8-bit passes of the radix sort (21 bits would need eight)
__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {   // spread (up to) 21 bits to every third bit
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

__global__ void k_morton(uint32_t n, const float* __restrict__ pos, const uint32_t* __restrict__ bounds,
                         unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float* p = pos + 9ull * t;
    unsigned long long code = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float lo = f32_unsortable(bounds[a]), hi = f32_unsortable(bounds[3 + a]);
        float c = (p[a] + p[3 + a] + p[6 + a]) * (1.0f / 3.0f);
        double ext = (double)hi - (double)lo;
        double u = ext > 0.0 ? ((double)c - (double)lo) / ext : 0.0;
        u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
        unsigned long long q = (unsigned long long)(u * (double)((1u << MORTON_BITS) - 1u));
        code |= expand21(q) << (2 - a);
    }
    keys[t] = code; idx[t] = t;
}

// ---- 2. gather into sorted (= leaf) order ----------------------------------------------------------
__global__ void k_gather_tris(uint32_t n, const uint32_t* __restrict__ sorted_idx, const float* __restrict__ pos,
                              const float* __restrict__ nrm, const uint32_t* __restrict__ vid,
                              TriRec* __restrict__ tris, float* __restrict__ nrm_out, uint32_t* __restrict__ vid_out) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint32_t src = sorted_idx[t];
    const float* p = pos + 9ull * src;
    TriRec r;
    r.t0 = make_float4(p[0], p[1], p[2], __uint_as_float(src));
    r.t1 = make_float4(p[3], p[4], p[5], 0.f);
    r.t2 = make_float4(p[6], p[7], p[8], 0.f);
    { V3 N; float d; tt_plane(mk3(p[0], p[1], p[2]), mk3(p[3], p[4], p[5]), mk3(p[6], p[7], p[8]), N, d); r.t3 = make_float4(N.x, N.y, N.z, d); }
    tris[t] = r;
    float* no = nrm_out + 9ull * t;
    if (nrm) { const float* q = nrm + 9ull * src;
#pragma unroll
        for (int k = 0; k < 9; ++k) no[k] = q[k];
    } else {   // TriangleNormal fallback = face normal on all three corners (Triangle.cpp:141-147,214-234)
        V3 p0 = mk3(p[0], p[1], p[2]), p1 = mk3(p[3], p[4], p[5]), p2 = mk3(p[6], p[7], p[8]);
        V3 fn = normalize3(cross3(sub3(p1, p0), sub3(p2, p0)));
#pragma unroll
        for (int k = 0; k < 3; ++k) { no[3 * k] = fn.x; no[3 * k + 1] = fn.y; no[3 * k + 2] = fn.z; }
    }
    uint32_t* vo = vid_out + 3ull * t;
    if (vid) { vo[0] = vid[3ull * src]; vo[1] = vid[3ull * src + 1]; vo[2] = vid[3ull * src + 2]; }
    else { vo[0] = 3u * src; vo[1] = 3u * src + 1u; vo[2] = 3u * src + 2u; }
}

// ---- 3. binary radix tree (Karras, "Maximizing Parallelism in the Construction of BVHs...", 2012) ----
__device__ __forceinline__ int delta_fn(const unsigned long long* k, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    unsigned long long x = k[i] ^ k[j];
    if (x == 0ull) return 64 + __clz((unsigned)i ^ (unsigned)j);
    return __clzll((long long)x);
}
// child encoding: >= 0 internal node index ; < 0 : ~leaf(triangle) index
__global__ void k_radix_tree(int n, const unsigned long long* __restrict__ keys, int* __restrict__ left, int* __restrict__ right,
                             int* __restrict__ parent_int, int* __restrict__ parent_leaf, uint32_t* __restrict__ first, uint32_t* __restrict__ last) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta_fn(keys, n, i, i + 1) - delta_fn(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta_fn(keys, n, i, i - d);
    int lmax = 2;
    while (delta_fn(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1) if (delta_fn(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta_fn(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1; ; t = (t + 1) >> 1) {
        if (delta_fn(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lc = (lo == gamma) ? ~gamma : gamma;
    int rc = (hi == gamma + 1) ? ~(gamma + 1) : (gamma + 1);
    left[i] = lc; right[i] = rc;
    first[i] = (uint32_t)lo; last[i] = (uint32_t)hi;
    if (lc >= 0) parent_int[lc] = i; else parent_leaf[~lc] = i;
    if (rc >= 0) parent_int[rc] = i; else parent_leaf[~rc] = i;
    if (i == 0) parent_int[0] = -1;
}

// ---- 4. collapse <= 4-triangle subtrees, assign sibling-adjacent records ---------------------------
#define LEAF_MAX 4u     // OBBtreeSplitBuildNode::maxNumberOfTriangles, OBBtree.h:49

__global__ void k_flag_inner(int n_int, const uint32_t* __restrict__ first, const uint32_t* __restrict__ last, uint32_t* __restrict__ flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == n_int) flag[i] = 0u;                  // the scan runs over n_int + 1 elements: its last output is the number of kept nodes
    if (i >= n_int) return;
    flag[i] = (last[i] - first[i] + 1u > LEAF_MAX) ? 1u : 0u;
}

// The kept nodes become records (FitRec, imrcd_fit.cuh: arena indices): an inner node writes its own record and those of its children that
// are leaves (collapsed subtrees or single triangles); record 0 is the root, record 1 the padding beside it.  The same thread puts the
// node on the fit's lists: "upper" when its subtree is too large for one block (more than FIT_T triangles or FIT_R records), and each
// child whose subtree is small enough under such a node as a treelet root.  The inner nodes of a subtree over leaves [l, r] are consecutive in
// the radix tree's numbering - [l, r - 1] when its root is numbered l (a right child, or the root), [l + 1, r] when it is numbered r (a left
// child) - so the records of their children are consecutive in the arena: slots [slot[lo], slot[lo + r - l]).
// append `rec` (for the lanes with `yes`) to a list: one atomic per warp for the list position and one for the slots; converged code only
__device__ __forceinline__ void fit_list_append(bool yes, uint32_t rec, uint32_t* counter, uint32_t* n_slots, uint32_t* slot_of, uint2* troots, uint32_t* uppers) {
    const uint32_t m = __ballot_sync(0xffffffffu, yes);
    if (m == 0u) return;
    const uint32_t lane = threadIdx.x & 31u, leader = (uint32_t)__ffs(m) - 1u, k = (uint32_t)__popc(m);
    uint32_t base = 0, sbase = 0;
    if (lane == leader) { base = atomicAdd(counter, k); sbase = atomicAdd(n_slots, k); }
    base = __shfl_sync(0xffffffffu, base, leader); sbase = __shfl_sync(0xffffffffu, sbase, leader);
    if (yes) {
        const uint32_t off = (uint32_t)__popc(m & ((1u << lane) - 1u));
        slot_of[rec] = sbase + off;
        if (troots) troots[base + off] = make_uint2(rec, 0u); else uppers[base + off] = rec;
    }
}

__global__ void k_assign(int n_int, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ parent_int,
                         const uint32_t* __restrict__ first, const uint32_t* __restrict__ last, const uint32_t* __restrict__ flag,
                         const uint32_t* __restrict__ slot, FitRec* __restrict__ fit, uint32_t rec_base, uint32_t tri_base, FitLists L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { FitRec pad; pad.first = pad.last = pad.split = pad.child = 0u; pad.parent = 0xffffffffu; pad.kind = 2u; pad.desc = 0xffffffffu; pad.n_sub = 1u; fit[rec_base + 1u] = pad; }
    const bool kept_me = i < n_int && flag[i];
    bool me_upper = false, me_troot = false, c_troot[2] = { false, false };
    uint32_t myrec = 0, child_base = 0;
    if (kept_me) {
        // my own record: root -> 0, else the slot my parent reserved for its children
        uint32_t mypar = 0xffffffffu;
        if (i != 0) { const int p = parent_int[i]; myrec = 2u + 2u * slot[p] + (left[p] == i ? 0u : 1u); mypar = rec_base + (p == 0 ? 0u : 2u + 2u * slot[parent_int[p]] + (left[parent_int[p]] == p ? 0u : 1u)); }
        child_base = 2u + 2u * slot[i];
        const int lc = left[i], rc = right[i];
        const uint32_t my_tris = last[i] - first[i] + 1u;
        const uint32_t my_lo = ((uint32_t)i == first[i]) ? first[i] : first[i] + 1u, my_sub = 2u * (slot[my_lo + last[i] - first[i]] - slot[my_lo]) + 1u;
        FitRec me; me.first = tri_base + first[i]; me.last = tri_base + last[i]; me.child = rec_base + child_base; me.kind = 0u; me.parent = mypar;
        me.split = tri_base + ((lc >= 0) ? last[lc] : (uint32_t)(~lc));
        me.desc = rec_base + 2u + 2u * slot[my_lo]; me.n_sub = my_sub;
        fit[rec_base + myrec] = me;
        const bool me_small = FIT_SMALL(my_tris, my_sub);
        me_upper = !me_small; me_troot = me_small && i == 0;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const int c = side ? rc : lc;
            const bool kept = c >= 0 && flag[c];
            if (!kept) {                               // a leaf record (a kept inner node writes its own)
                FitRec d;
                if (c >= 0) { d.first = tri_base + first[c]; d.last = tri_base + last[c]; }
                else { d.first = d.last = tri_base + (uint32_t)(~c); }
                d.split = d.last; d.child = 0u; d.kind = 1u; d.parent = rec_base + myrec; d.desc = 0xffffffffu; d.n_sub = 1u;
                fit[rec_base + child_base + side] = d;
            }
            if (!me_small) {                           // a child with a small subtree under a large one: a treelet root
                bool c_small = true;
                if (kept) { const uint32_t c_lo = ((uint32_t)c == first[c]) ? first[c] : first[c] + 1u; c_small = FIT_SMALL(last[c] - first[c] + 1u, 2u * (slot[c_lo + last[c] - first[c]] - slot[c_lo]) + 1u); }
                c_troot[side] = c_small;
            }
        }
    }
    fit_list_append(me_upper, rec_base + myrec, &L.cnt->n_upper, &L.cnt->n_slots, L.slot_of, nullptr, L.uppers);
    fit_list_append(me_troot, rec_base + myrec, &L.cnt->n_troot, &L.cnt->n_slots, L.slot_of, L.troots, nullptr);
    fit_list_append(c_troot[0], rec_base + child_base, &L.cnt->n_troot, &L.cnt->n_slots, L.slot_of, L.troots, nullptr);
    fit_list_append(c_troot[1], rec_base + child_base + 1u, &L.cnt->n_troot, &L.cnt->n_slots, L.slot_of, L.troots, nullptr);
}

// single-leaf meshes (<= 4 triangles): OBBtree.cpp:346-356
__global__ void k_assign_single_leaf(FitRec* fit, uint32_t rec_base, uint32_t tri_base, uint32_t n, uint32_t* n_inner, FitLists L) {
    FitRec d; d.first = tri_base; d.last = tri_base + (n ? n - 1u : 0u); d.split = d.last; d.child = 0u; d.parent = 0xffffffffu; d.kind = 1u; d.desc = 0xffffffffu; d.n_sub = 1u;
    fit[rec_base] = d;
    d.first = d.last = d.split = 0u; d.kind = 2u; fit[rec_base + 1u] = d;
    *n_inner = 0u;
    L.slot_of[rec_base] = 0u; L.troots[0] = make_uint2(rec_base, 0u); L.cnt->n_slots = 1u; L.cnt->n_troot = 1u; L.cnt->n_upper = 0u;
}

// ---------------------------------------------------------------------------------------------------
// Refit (BASELINE config 5: re-posed meshes): new triangle positions, same topology.  Works on any tree in the arena
// (Morton-built, reference-built or imported) and on many meshes per call: the same fit as the build (imrcd_fit.cu), links untouched.
// The reference has no refit (SURVEY finding 4: skinned meshes never get collision trees); parity is defined against the
// reference REBUILDING its tree from the re-posed triangles (tests/test_gpu_refit.py).
// ---------------------------------------------------------------------------------------------------
__global__ void k_rf_scatter(uint32_t n, TriRec* tris, float* nrm_out, const float* __restrict__ pos, const float* __restrict__ nrm) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    TriRec r = tris[t];
    const uint32_t src = __float_as_uint(r.t0.w);
    const float* p = pos + 9ull * src;
    r.t0 = make_float4(p[0], p[1], p[2], r.t0.w); r.t1 = make_float4(p[3], p[4], p[5], 0.f); r.t2 = make_float4(p[6], p[7], p[8], 0.f);
    { V3 N; float d; tt_plane(mk3(p[0], p[1], p[2]), mk3(p[3], p[4], p[5]), mk3(p[6], p[7], p[8]), N, d); r.t3 = make_float4(N.x, N.y, N.z, d); }
    tris[t] = r;
    if (nrm) { float* no = nrm_out + 9ull * t; const float* q = nrm + 9ull * src;
#pragma unroll
        for (int k = 0; k < 9; ++k) no[k] = q[k]; }
}

// ---------------------------------------------------------------------------------------------------
static inline unsigned nb(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

static int build_morton(imrcd_ctx* ctx, const float* h_pos, const float* h_nrm, const uint32_t* h_vid, uint64_t n_tri, MeshHost* mh) {
    cudaStream_t s = ctx->stream;
    const uint32_t n = (uint32_t)n_tri;
    if (n_tri == 0) {   // empty mesh: one EmptyOBB-like leaf root (OBB.cpp:168-178)
        int rc = imr_mesh_arena_alloc(ctx, 2, 0, mh);
        if (rc) return rc;
        TreeRec r[2]; memset(r, 0, sizeof(r));
        r[0].q0 = make_float4(0, 0, 0, FLT_EPSILON); r[0].q1 = make_float4(0, 0, 0, FLT_EPSILON); r[0].q2 = make_float4(0, 0, 0, FLT_EPSILON);
        uint32_t one = 1u; float onef; memcpy(&onef, &one, 4);
        r[0].q3 = make_float4(0, 0, 0, onef); r[1] = r[0];
        IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_recs.as<TreeRec>() + mh->dev.rec_base, r, sizeof(r), cudaMemcpyHostToDevice, s));
        IMR_CUDA(ctx, cudaStreamSynchronize(s));
        memset(mh->root_box, 0, 48); mh->root_box[3] = mh->root_box[7] = mh->root_box[11] = FLT_EPSILON;
        return imr_mesh_finalize_records(ctx, mh->dev.rec_base, 2);
    }
    // ---- staging buffers (freed at the end; builds are load-time operations) ----
    DevBuf d_pos, d_nrm, d_vid, d_bounds, d_keys, d_keys2, d_idx, d_idx2, d_tmp, d_left, d_right, d_pint, d_pleaf, d_first, d_last, d_flag, d_slot;
    auto free_all = [&]() { DevBuf* all[] = { &d_pos, &d_nrm, &d_vid, &d_bounds, &d_keys, &d_keys2, &d_idx, &d_idx2, &d_tmp, &d_left, &d_right, &d_pint,
                                              &d_pleaf, &d_first, &d_last, &d_flag, &d_slot };
                            for (DevBuf* b : all) b->release(); };
#define BUILD_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e); free_all(); return IMRCD_E_CUDA; } } while (0)
    BUILD_CUDA(d_pos.reserve(36ull * n, 0, s));
    BUILD_CUDA(cudaMemcpyAsync(d_pos.p, h_pos, 36ull * n, cudaMemcpyDefault, s));
    if (h_nrm) { BUILD_CUDA(d_nrm.reserve(36ull * n, 0, s)); BUILD_CUDA(cudaMemcpyAsync(d_nrm.p, h_nrm, 36ull * n, cudaMemcpyDefault, s)); }
    if (h_vid) { BUILD_CUDA(d_vid.reserve(12ull * n, 0, s)); BUILD_CUDA(cudaMemcpyAsync(d_vid.p, h_vid, 12ull * n, cudaMemcpyDefault, s)); }
    BUILD_CUDA(d_bounds.reserve(64, 0, s));
    BUILD_CUDA(d_keys.reserve(8ull * n, 0, s)); BUILD_CUDA(d_keys2.reserve(8ull * n, 0, s));
    BUILD_CUDA(d_idx.reserve(4ull * n, 0, s)); BUILD_CUDA(d_idx2.reserve(4ull * n, 0, s));
    const uint32_t n_int = n > 1 ? n - 1 : 0;
    BUILD_CUDA(d_left.reserve(4ull * (n_int + 1), 0, s)); BUILD_CUDA(d_right.reserve(4ull * (n_int + 1), 0, s));
    BUILD_CUDA(d_pint.reserve(4ull * (n_int + 1), 0, s)); BUILD_CUDA(d_pleaf.reserve(4ull * n, 0, s));
    BUILD_CUDA(d_first.reserve(4ull * (n_int + 1), 0, s)); BUILD_CUDA(d_last.reserve(4ull * (n_int + 1), 0, s));
    BUILD_CUDA(d_flag.reserve(4ull * (n_int + 1), 0, s)); BUILD_CUDA(d_slot.reserve(4ull * (n_int + 2), 0, s));

    // every allocation happens before the timed region: sort / scan scratch, the arena for the worst-case record count (every inner node
    // kept; trimmed to the real count at the end), the fit's tables
    size_t tmp_bytes = 0, scan_bytes = 0;
    const int key_bits = 3 * MORTON_BITS;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys.as<unsigned long long>(), d_keys2.as<unsigned long long>(), d_idx.as<uint32_t>(), d_idx2.as<uint32_t>(), (int)n, 0, key_bits, s);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_flag.as<uint32_t>(), d_slot.as<uint32_t>(), (int)(n_int + 1), s);
    BUILD_CUDA(d_tmp.reserve(std::max(tmp_bytes, scan_bytes), 0, s));
    const uint64_t n_rec_max = n > LEAF_MAX ? 2ull + 2ull * n_int : 2ull;      // every inner node kept (a degenerate chain comes close)
    int rc = imr_mesh_arena_alloc(ctx, n_rec_max, n, mh);
    if (rc) { free_all(); return rc; }
    rc = imr_fit_reserve(ctx, ctx->n_rec_total);
    if (rc) { free_all(); return rc; }
    std::vector<FitSeg> segs(1);
    segs[0].rec_base = mh->dev.rec_base; segs[0].n_rec = (uint32_t)n_rec_max; segs[0].tri_base = mh->dev.tri_base; segs[0].n_tri = n;
    segs[0].origin[0] = segs[0].origin[1] = segs[0].origin[2] = 0.0;
    rc = imr_fit_prepare(ctx, segs, 0);
    if (rc) { free_all(); return rc; }
    const FitLists lists = imr_fit_lists(ctx);
    TriRec* tris = ctx->d_tris.as<TriRec>() + mh->dev.tri_base;
    TreeRec* recs = ctx->d_recs.as<TreeRec>() + mh->dev.rec_base;
    FitRec* fit = ctx->d_fit.as<FitRec>();
    uint32_t* n_inner_dev = d_slot.as<uint32_t>() + n_int;             // the scan's total: kept inner nodes

    cudaEvent_t e0 = ctx->ev[6], e1 = ctx->ev[7];
    BUILD_CUDA(cudaEventRecord(e0, s));
    // 1. keys + sort
    k_bounds_init<<<1, 32, 0, s>>>(d_bounds.as<uint32_t>());
    k_bounds<<<std::min<unsigned>(nb(n, 256), ctx->sm_count * 8), 256, 0, s>>>(n, d_pos.as<float>(), d_bounds.as<uint32_t>());
    k_morton<<<nb(n, 256), 256, 0, s>>>(n, d_pos.as<float>(), d_bounds.as<uint32_t>(), d_keys.as<unsigned long long>(), d_idx.as<uint32_t>());
    cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_keys.as<unsigned long long>(), d_keys2.as<unsigned long long>(), d_idx.as<uint32_t>(), d_idx2.as<uint32_t>(), (int)n, 0, key_bits, s);
    // 2. gather into leaf order
    k_gather_tris<<<nb(n, 256), 256, 0, s>>>(n, d_idx2.as<uint32_t>(), d_pos.as<float>(), h_nrm ? d_nrm.as<float>() : nullptr, h_vid ? d_vid.as<uint32_t>() : nullptr,
                                             tris, ctx->d_tri_nrm.as<float>() + 9ull * mh->dev.tri_base, ctx->d_tri_vid.as<uint32_t>() + 3ull * mh->dev.tri_base);
    if (n > LEAF_MAX) {
        // 3. radix tree, 4. kept nodes -> records
        k_radix_tree<<<nb(n_int, 128), 128, 0, s>>>((int)n, d_keys2.as<unsigned long long>(), d_left.as<int>(), d_right.as<int>(), d_pint.as<int>(),
                                                     d_pleaf.as<int>(), d_first.as<uint32_t>(), d_last.as<uint32_t>());
        k_flag_inner<<<nb(n_int + 1, 256), 256, 0, s>>>((int)n_int, d_first.as<uint32_t>(), d_last.as<uint32_t>(), d_flag.as<uint32_t>());
        cub::DeviceScan::ExclusiveSum(d_tmp.p, scan_bytes, d_flag.as<uint32_t>(), d_slot.as<uint32_t>(), (int)(n_int + 1), s);
        k_assign<<<nb(n_int, 128), 128, 0, s>>>((int)n_int, d_left.as<int>(), d_right.as<int>(), d_pint.as<int>(), d_first.as<uint32_t>(), d_last.as<uint32_t>(),
                                                d_flag.as<uint32_t>(), d_slot.as<uint32_t>(), fit, mh->dev.rec_base, mh->dev.tri_base, lists);
    } else k_assign_single_leaf<<<1, 1, 0, s>>>(fit, mh->dev.rec_base, mh->dev.tri_base, n, n_inner_dev, lists);
    // 5. the fit
    rc = imr_fit_launch(ctx, true, d_bounds.as<uint32_t>(), true);
    if (rc) { free_all(); return rc; }
    { TreeRec z; memset(&z, 0, sizeof(z)); uint32_t one = 1u; memcpy(&z.q3.w, &one, 4);      // the padding record beside the root
      BUILD_CUDA(cudaMemcpyAsync(recs + 1, &z, sizeof(z), cudaMemcpyHostToDevice, s)); }
    BUILD_CUDA(cudaEventRecord(e1, s));
    TreeRec root; uint32_t n_inner = 0;
    BUILD_CUDA(cudaMemcpyAsync(&root, recs, sizeof(TreeRec), cudaMemcpyDeviceToHost, s));
    BUILD_CUDA(cudaMemcpyAsync(&n_inner, n_inner_dev, 4, cudaMemcpyDeviceToHost, s));
    BUILD_CUDA(cudaStreamSynchronize(s));
    BUILD_CUDA(cudaGetLastError());
    cudaEventElapsedTime(&mh->build_ms, e0, e1);
    // trim the arena to the records the tree really has (this mesh is the last one in it)
    const uint64_t n_rec = 2ull + 2ull * n_inner;
    if (n_rec > n_rec_max) { ctx->err = "build: more records than reserved"; free_all(); return IMRCD_E_CAPACITY; }
    ctx->n_rec_total -= n_rec_max - n_rec; mh->dev.n_rec = (uint32_t)n_rec;
    mh->fit_plan = true;
    const float rb[12] = { root.q0.x, root.q0.y, root.q0.z, root.q0.w, root.q1.x, root.q1.y, root.q1.z, root.q1.w, root.q2.x, root.q2.y, root.q2.z, root.q2.w };
    memcpy(mh->root_box, rb, 48);
    free_all();
#undef BUILD_CUDA
    return IMRCD_OK;
}

int imr_build_mesh_device(imrcd_ctx* ctx, const float* pos, const float* nrm, const uint32_t* vid, uint64_t n_tri, uint32_t mode, MeshHost* out) {
    if (n_tri >= (1ull << 31)) { ctx->err = "imrcd_mesh_create: too many triangles"; return IMRCD_E_ARG; }
    if (mode == IMRCD_BUILD_MORTON) return build_morton(ctx, pos, nrm, vid, n_tri, out);
    if (mode == IMRCD_BUILD_REFERENCE) return imr_build_mesh_reference(ctx, pos, nrm, vid, n_tri, out);
    ctx->err = "imrcd_mesh_create: unknown build mode";
    return IMRCD_E_ARG;
}

// ---- refit: host side --------------------------------------------------------------------------------
int imr_mesh_update_positions_device(imrcd_ctx* ctx, uint32_t mesh_id, const float* h_pos, const float* h_nrm) {
    MeshHost& mh = ctx->meshes[mesh_id];
    const uint32_t n = mh.dev.n_tri;
    if (n == 0) return IMRCD_OK;
    cudaStream_t s = ctx->stream;
    IMR_CUDA(ctx, ctx->d_rf_stage.reserve(36ull * n * (h_nrm ? 2 : 1), 0, s));
    IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_rf_stage.p, h_pos, 36ull * n, cudaMemcpyHostToDevice, s));
    if (h_nrm) IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_rf_stage.as<char>() + 36ull * n, h_nrm, 36ull * n, cudaMemcpyHostToDevice, s));
    k_rf_scatter<<<nb(n, 256), 256, 0, s>>>(n, ctx->d_tris.as<TriRec>() + mh.dev.tri_base, ctx->d_tri_nrm.as<float>() + 9ull * mh.dev.tri_base,
                                             ctx->d_rf_stage.as<float>(), h_nrm ? ctx->d_rf_stage.as<float>() + 9ull * n : nullptr);
    IMR_CUDA(ctx, cudaGetLastError());
    IMR_CUDA(ctx, cudaStreamSynchronize(s));           // the staging buffer is reused by the next update
    mh.needs_refit = true;
    return IMRCD_OK;
}

int imr_meshes_refit_device(imrcd_ctx* ctx, const uint32_t* ids, uint64_t n_ids, float* ms_out) {
    std::vector<uint32_t> order(ids, ids + n_ids);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ctx->meshes[a].dev.rec_base < ctx->meshes[b].dev.rec_base; });
    order.erase(std::unique(order.begin(), order.end()), order.end());
    std::vector<FitSeg> segs;
    for (uint32_t id : order) {                          // segments in arena order
        MeshHost& mh = ctx->meshes[id];
        mh.needs_refit = false;
        if (mh.dev.n_tri == 0) continue;                // nothing to fit
        if (!mh.fit_plan) {                             // a tree that was not built by the Morton build: its FitRecs from its records, once
            const int rc = imr_fit_plan_from_records(ctx, mh.dev);
            if (rc) return rc;
            mh.fit_plan = true;
        }
        FitSeg sg; sg.rec_base = mh.dev.rec_base; sg.n_rec = mh.dev.n_rec; sg.tri_base = mh.dev.tri_base; sg.n_tri = mh.dev.n_tri;
        sg.origin[0] = sg.origin[1] = sg.origin[2] = 0.0;
        segs.push_back(sg);
    }
    if (segs.empty()) { if (ms_out) *ms_out = 0.f; return IMRCD_OK; }
    cudaStream_t s = ctx->stream;
    uint64_t key = 0xcbf29ce484222325ull;               // the same trees as the last call (the arena is append-only: base + size name a tree)
    for (const FitSeg& g : segs) { key = (key ^ g.rec_base) * 0x100000001b3ull; key = (key ^ g.n_rec) * 0x100000001b3ull; key = (key ^ g.n_tri) * 0x100000001b3ull; }
    int rc = imr_fit_prepare(ctx, segs, key | 1ull);
    if (rc) return rc;
    cudaEvent_t e0 = ctx->ev[6], e1 = ctx->ev[7];
    IMR_CUDA(ctx, cudaEventRecord(e0, s));
    rc = imr_fit_launch(ctx, false, nullptr, false);      // origins: the old root boxes' centres
    if (rc) return rc;
    IMR_CUDA(ctx, cudaEventRecord(e1, s));
    IMR_CUDA(ctx, cudaStreamSynchronize(s));
    IMR_CUDA(ctx, cudaGetLastError());
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out) *ms_out = ms;
    return IMRCD_OK;
}

// ---------------------------------------------------------------------------------------------------
// Triangle::CreateTriangleList on the device: the primitives of a mesh, as they lie in the glTF buffers, become the flat
// triangle arrays (positions, normals, vertex ids) the builds start from.  CreateIndicesTriplets, IMR/src/Geometry/Triangle.cpp:9-62:
//   points          (i, i, i)                          line strip      (i, i, i+1)
//   lines           (2i, 2i, 2i+1)                     triangles       (3i, 3i+1, 3i+2)
//   triangle strip  (i, i + (1 + i % 2), i + (2 - i % 2))              triangle fan   (i+1, i+2, 0)
// Normals are the vertex normals through the same triplets, or the triangle's face normal three times when the primitive has none
// (Triangle.cpp:141-147,223-232); vertex ids are the index values themselves (Triangle.cpp:242-250).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gltf_triplet(uint32_t mode, uint32_t i, uint32_t& a, uint32_t& b, uint32_t& c) {
    switch (mode) {
        case 0: a = i; b = i; c = i; break;
        case 1: a = 2u * i; b = 2u * i; c = 2u * i + 1u; break;
        case 3: a = i; b = i; c = i + 1u; break;
        case 4: a = 3u * i; b = 3u * i + 1u; c = 3u * i + 2u; break;
        case 5: a = i; b = i + (1u + i % 2u); c = i + (2u - i % 2u); break;
        default: a = i + 1u; b = i + 2u; c = 0u; break;      // triangle fan
    }
}

__global__ void k_assemble_primitive(uint32_t n_tri, uint32_t mode, uint32_t stride, const float* __restrict__ points, const float* __restrict__ normals,
                                     const uint32_t* __restrict__ indices, float* __restrict__ pos_out, float* __restrict__ nrm_out, uint32_t* __restrict__ vid_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tri) return;
    uint32_t k[3];
    gltf_triplet(mode, i, k[0], k[1], k[2]);
    uint32_t v[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) v[q] = indices ? indices[k[q]] : k[q];
    float p[9];
#pragma unroll
    for (int q = 0; q < 3; ++q) { const float* s = points + (size_t)stride * v[q]; p[3 * q] = s[0]; p[3 * q + 1] = s[1]; p[3 * q + 2] = s[2]; }
#pragma unroll
    for (int q = 0; q < 9; ++q) pos_out[9ull * i + q] = p[q];
    if (normals) {
#pragma unroll
        for (int q = 0; q < 3; ++q) { const float* s = normals + (size_t)stride * v[q]; nrm_out[9ull * i + 3 * q] = s[0]; nrm_out[9ull * i + 3 * q + 1] = s[1]; nrm_out[9ull * i + 3 * q + 2] = s[2]; }
    } else {
        const V3 n = normalize3(cross3(sub3(mk3(p[3], p[4], p[5]), mk3(p[0], p[1], p[2])), sub3(mk3(p[6], p[7], p[8]), mk3(p[0], p[1], p[2]))));   // Triangle.cpp:141-147
#pragma unroll
        for (int q = 0; q < 3; ++q) { nrm_out[9ull * i + 3 * q] = n.x; nrm_out[9ull * i + 3 * q + 1] = n.y; nrm_out[9ull * i + 3 * q + 2] = n.z; }
    }
    vid_out[3ull * i] = v[0]; vid_out[3ull * i + 1] = v[1]; vid_out[3ull * i + 2] = v[2];
}

int imr_mesh_assemble_device(imrcd_ctx* ctx, uint32_t build_mode, MeshHost* out) {
    cudaStream_t s = ctx->stream;
    uint64_t n_tri = 0;
    for (const auto& r : ctx->recording) n_tri += r.n_tri;
    if (n_tri >= (1ull << 32)) { ctx->err = "mesh exceeds 2^32 triangles"; return IMRCD_E_CAPACITY; }
    DevBuf d_pos, d_nrm, d_vid;
    auto free_all = [&]() { d_pos.release(); d_nrm.release(); d_vid.release(); };
    if (n_tri) {
        cudaError_t e;
        if ((e = d_pos.reserve(36ull * n_tri, 0, s)) != cudaSuccess || (e = d_nrm.reserve(36ull * n_tri, 0, s)) != cudaSuccess || (e = d_vid.reserve(12ull * n_tri, 0, s)) != cudaSuccess) {
            ctx->err = std::string("imr_mesh_assemble_device: ") + cudaGetErrorString(e); free_all(); return IMRCD_E_CUDA;
        }
        uint64_t at = 0;
        for (const auto& r : ctx->recording) {
            if (!r.n_tri) continue;
            k_assemble_primitive<<<nb(r.n_tri, 256), 256, 0, s>>>((uint32_t)r.n_tri, r.mode, r.stride, r.points.as<float>(), r.normals.p ? r.normals.as<float>() : nullptr,
                                                                   r.indices.p ? r.indices.as<uint32_t>() : nullptr, d_pos.as<float>() + 9ull * at, d_nrm.as<float>() + 9ull * at,
                                                                   d_vid.as<uint32_t>() + 3ull * at);
            at += r.n_tri;
        }
        if ((e = cudaGetLastError()) != cudaSuccess) { ctx->err = std::string("k_assemble_primitive: ") + cudaGetErrorString(e); free_all(); return IMRCD_E_CUDA; }
    }
    const int rc = imr_build_mesh_device(ctx, d_pos.as<float>(), n_tri ? d_nrm.as<float>() : nullptr, n_tri ? d_vid.as<uint32_t>() : nullptr, n_tri, build_mode, out);
    free_all();
    return rc;
}
