// imrcd_build_ref.cu -- IMRCD_BUILD_REFERENCE: the reference's OWN tree, built on the device.
//
// Reproduces OBBtree::OBBtree(std::vector<Triangle>&&) (IMR/src/Geometry/OBBtree.cpp:321-358) bit for bit: the recursive
// top-down OBBtreeSplitBuildNode (:8-108) becomes a level-synchronous loop (one warp per node per level); every node's box is
// OBB::CreateOBBfromPoints (IMR/src/Geometry/OBB.cpp:33-166) with the reference's SEQUENTIAL FP64 summation order
// (std::accumulate / std::inner_product, :52-73: each running sum is owned by one lane that walks the node's points in order;
// the nine sums run on nine lanes side by side), eig3 in FP64 (imrcd_eig3.cuh), the rows-of-V axes (:80-87), and the split rule
// of SplitOBBandCreateChildren (:43-108: axes by half-length, (min+max)/2 <= centre projection goes left, stable order, next axis
// when a side is empty, halves by index as the last resort, larger surface becomes the left child).
// It is the parity mode: slower than the Morton build (a node's sums are sequential by definition), identical to the reference.
#include "imrcd_internal.cuh"
#include "imrcd_eig3.cuh"
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cstring>
#include <vector>

int imr_mesh_arena_alloc(imrcd_ctx* ctx, uint64_t n_rec, uint64_t n_tri, MeshHost* mh);

#define FULL_MASK 0xffffffffu
#define REF_LEAF_MAX 4u              // OBBtreeSplitBuildNode::maxNumberOfTriangles, OBBtree.h:49

struct RefNodes {                    // structure of arrays, one entry per build node
    uint32_t* begin; uint32_t* count; uint32_t* left; uint32_t* right; uint32_t* parent; uint8_t* side; uint8_t* leaf; uint8_t* buf;
    float* box;                      // 12 floats per node
    float* surface;
    uint32_t* sub_tris; uint32_t* tri_off; uint32_t* inner_flag; uint32_t* inner_rank;
};

// points of a node in the reference's order: triangle by triangle, corner by corner (OBB.cpp:91-103)
struct TriPoints {
    const float* pos; const uint32_t* idx;
    __device__ __forceinline__ V3 operator()(uint64_t i) const {
        const uint32_t t = idx[i / 3u]; const uint32_t k = (uint32_t)(i % 3u);
        const float* p = pos + 9ull * t + 3u * k;
        return mk3(p[0], p[1], p[2]);
    }
};
struct RawPoints {
    const float* pts;
    __device__ __forceinline__ V3 operator()(uint64_t i) const { const float* p = pts + 3ull * i; return mk3(p[0], p[1], p[2]); }
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(FULL_MASK, v, src); }

// OBB::CreateOBBfromPoints by one warp.  The result is returned on every lane.
template <class Points>
__device__ Box ref_fit_warp(const Points& pt, uint64_t np, uint32_t lane) {
    Box out; out.c = mk3(0.f, 0.f, 0.f); out.u = mk3(FLT_EPSILON, 0.f, 0.f); out.v = mk3(0.f, FLT_EPSILON, 0.f); out.w = mk3(0.f, 0.f, FLT_EPSILON);   // EmptyOBB, OBB.cpp:168-178
    // unique-point probe (:35-41): stops once more than 3 distinct points were seen
    int nu = 0;
    if (lane == 0) {
        V3 uq[4];
        for (uint64_t i = 0; i < np && nu <= 3; ++i) {
            const V3 p = pt(i);
            bool seen = false;
            for (int k = 0; k < nu; ++k) if (uq[k].x == p.x && uq[k].y == p.y && uq[k].z == p.z) { seen = true; break; }
            if (!seen) { if (nu < 4) uq[nu] = p; ++nu; }
        }
    }
    nu = __shfl_sync(FULL_MASK, nu, 0);
    if (nu == 0) return out;
    if (nu == 1) { out.c = pt(0); return out; }

    const double dn = (double)np;
    // mean (:52-56): three sequential sums on lanes 0..2
    double s = 0.0;
    if (lane < 3) {
        for (uint64_t i = 0; i < np; ++i) { const V3 p = pt(i); const float c = lane == 0 ? p.x : (lane == 1 ? p.y : p.z); s = s + (double)c; }
        s = s / dn;
    }
    const double mx = shfl_d(s, 0), my = shfl_d(s, 1), mz = shfl_d(s, 2);
    // covariance (:58-73): six sequential sums on lanes 0..5, order xx yy zz xy xz yz
    double acc = 0.0;
    if (lane < 6) {
        const int a = lane < 3 ? (int)lane : (lane == 5 ? 1 : 0);
        const int b = lane < 3 ? (int)lane : (lane == 3 ? 1 : 2);
        const double ma = a == 0 ? mx : (a == 1 ? my : mz), mb = b == 0 ? mx : (b == 1 ? my : mz);
        for (uint64_t i = 0; i < np; ++i) {
            const V3 p = pt(i);
            const float pa = a == 0 ? p.x : (a == 1 ? p.y : p.z), pb = b == 0 ? p.x : (b == 1 ? p.y : p.z);
            acc = acc + ((double)pa - ma) * ((double)pb - mb);
        }
    }
    const double cxx = shfl_d(acc, 0), cyy = shfl_d(acc, 1), czz = shfl_d(acc, 2), cxy = shfl_d(acc, 3), cxz = shfl_d(acc, 4), cyz = shfl_d(acc, 5);
    double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (lane == 0) {
        double A[9] = { cxx, cxy, cxz, cxy, cyy, cyz, cxz, cyz, czz };
        for (int i = 0; i < 9; ++i) A[i] = A[i] / dn;                         // cov_mat /= double(points.size())
        double d[3];
        e3_eigen_decomposition(A, V, d);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) V[i] = shfl_d(V[i], 0);
    // CreateAABBfromPoints along the ROWS of V (OBB.cpp:80-87,105-166).  min / max do not depend on the order.
    double mn[3] = { INFINITY, INFINITY, INFINITY }, mxv[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (uint64_t i = lane; i < np; i += 32) {
        const V3 p = pt(i);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double tx = V[3 * a] * (double)p.x, ty = V[3 * a + 1] * (double)p.y, tz = V[3 * a + 2] * (double)p.z;
            const double proj = tx + ty + tz;
            mn[a] = (proj < mn[a]) ? proj : mn[a];
            mxv[a] = (mxv[a] < proj) ? proj : mxv[a];
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            const double om = __shfl_xor_sync(FULL_MASK, mn[a], o), ox = __shfl_xor_sync(FULL_MASK, mxv[a], o);
            mn[a] = (om < mn[a]) ? om : mn[a]; mxv[a] = (mxv[a] < ox) ? ox : mxv[a];
        }
    double center[3] = { 0.0, 0.0, 0.0 };
    V3 sides[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double delta = (mxv[a] - mn[a]) + 2.0 * (double)FLT_EPSILON;
        const double mid = (mxv[a] + mn[a]) / 2.0;
        center[0] += mid * V[3 * a]; center[1] += mid * V[3 * a + 1]; center[2] += mid * V[3 * a + 2];
        const double half = delta / 2.0;
        sides[a] = mk3((float)(half * V[3 * a]), (float)(half * V[3 * a + 1]), (float)(half * V[3 * a + 2]));
    }
    out.c = mk3((float)center[0], (float)center[1], (float)center[2]);
    out.u = sides[0]; out.v = sides[1]; out.w = sides[2];
    return out;
}

// Triangle::GetMinMaxProjectionToAxis (Triangle.cpp:113-139) and the split test of OBBtree.cpp:60-77
__device__ __forceinline__ bool ref_goes_left(const float* pos, uint32_t tri, V3 axis, float cproj) {
    float lo = INFINITY, hi = -INFINITY;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float* p = pos + 9ull * tri + 3 * k;
        const float pr = dot3(axis, mk3(p[0], p[1], p[2]));
        if (pr < lo) lo = pr;
        if (pr > hi) hi = pr;
    }
    const float mean = (lo + hi) / 2.f;
    return mean <= cproj;
}

// One level of the build: fit every node of the level, split the ones with more than 4 triangles.
__global__ void __launch_bounds__(128)
k_ref_level(const uint32_t* __restrict__ cur_list, uint32_t n_cur, const float* __restrict__ pos, const uint32_t* __restrict__ idx_in, uint32_t* __restrict__ idx_out,
            RefNodes nd, uint32_t* node_count, uint32_t* next_list, uint32_t* next_count, uint8_t level_buf) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_cur; w += warps_total) {
        const uint32_t v = cur_list[w];
        const uint32_t begin = nd.begin[v], count = nd.count[v];
        TriPoints pts; pts.pos = pos; pts.idx = idx_in + begin;
        const Box box = ref_fit_warp(pts, 3ull * count, lane);
        if (lane == 0) {
            float* b = nd.box + 12ull * v;
            b[0] = box.c.x; b[1] = box.c.y; b[2] = box.c.z; b[3] = box.u.x; b[4] = box.u.y; b[5] = box.u.z;
            b[6] = box.v.x; b[7] = box.v.y; b[8] = box.v.z; b[9] = box.w.x; b[10] = box.w.y; b[11] = box.w.z;
            nd.surface[v] = box_surface(box);
            nd.buf[v] = level_buf;
        }
        if (count <= REF_LEAF_MAX) { if (lane == 0) { nd.leaf[v] = 1; nd.left[v] = nd.right[v] = 0xffffffffu; } continue; }
        // ---- SplitOBBandCreateChildren (OBBtree.cpp:43-108) ----
        const V3 side[3] = { box.u, box.v, box.w };
        float len[3]; V3 ax[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { len[k] = length3(side[k]); ax[k] = normalize3(side[k]); }
        int ord[3] = { 0, 1, 2 };                       // std::sort by length, descending; 3 elements = insertion sort, ties keep order (:50-51)
        for (int i = 1; i < 3; ++i) { const int t = ord[i]; int j = i; while (j > 0 && len[t] > len[ord[j - 1]]) { ord[j] = ord[j - 1]; --j; } ord[j] = t; }
        uint32_t nl = 0, nr = 0; int chosen = -1;
        for (int attempt = 0; attempt < 3; ++attempt) {
            const V3 axis = ax[ord[attempt]];
            const float cproj = dot3(box.c, axis);      // GetCenterProjectionToAxis, Paralgram.cpp:192-196
            uint32_t l = 0;
            for (uint32_t i0 = 0; i0 < count; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool left = i < count && ref_goes_left(pos, idx_in[begin + i], axis, cproj);
                l += (uint32_t)__popc(__ballot_sync(FULL_MASK, left));
            }
            nl = l; nr = count - l;
            if (nl != 0 && nr != 0) { chosen = attempt; break; }
        }
        if (chosen >= 0) {                              // stable partition: left list first, both in the original order
            const V3 axis = ax[ord[chosen]];
            const float cproj = dot3(box.c, axis);
            uint32_t wl = 0, wr = 0;
            for (uint32_t i0 = 0; i0 < count; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool valid = i < count;
                const uint32_t t = valid ? idx_in[begin + i] : 0u;
                const bool left = valid && ref_goes_left(pos, t, axis, cproj);
                const uint32_t ml = __ballot_sync(FULL_MASK, left), mr = __ballot_sync(FULL_MASK, valid && !left);
                const uint32_t below = (1u << lane) - 1u;
                if (left) idx_out[begin + wl + __popc(ml & below)] = t;
                else if (valid) idx_out[begin + nl + wr + __popc(mr & below)] = t;
                wl += (uint32_t)__popc(ml); wr += (uint32_t)__popc(mr);
            }
        } else {                                        // halves by index (:83-95)
            nl = count / 2; nr = count - nl;
            for (uint32_t i = lane; i < count; i += 32) idx_out[begin + i] = idx_in[begin + i];
        }
        uint32_t base = 0, slot = 0;
        if (lane == 0) { base = atomicAdd(node_count, 2u); slot = atomicAdd(next_count, 2u); }
        base = __shfl_sync(FULL_MASK, base, 0); slot = __shfl_sync(FULL_MASK, slot, 0);
        if (lane < 2) {
            const uint32_t c = base + lane;
            nd.begin[c] = begin + (lane ? nl : 0u); nd.count[c] = lane ? nr : nl;
            nd.parent[c] = v; nd.side[c] = (uint8_t)lane; nd.leaf[c] = 0;
            next_list[slot + lane] = c;
        }
        if (lane == 0) { nd.leaf[v] = 0; nd.left[v] = base; nd.right[v] = base + 1u; }
    }
}

// larger surface becomes the left child (OBBtree.cpp:102-107)
__global__ void k_ref_swap(uint32_t n_nodes, RefNodes nd) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes || nd.leaf[v]) return;
    const uint32_t l = nd.left[v], r = nd.right[v];
    if (nd.surface[r] > nd.surface[l]) { nd.left[v] = r; nd.right[v] = l; nd.side[r] = 0; nd.side[l] = 1; }
}
__global__ void k_ref_sizes(uint32_t lo, uint32_t hi, RefNodes nd) {
    const uint32_t v = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= hi) return;
    nd.inner_flag[v] = nd.leaf[v] ? 0u : 1u;
    nd.sub_tris[v] = nd.leaf[v] ? nd.count[v] : nd.sub_tris[nd.left[v]] + nd.sub_tris[nd.right[v]];
}
// leaf triangles are appended in pre-order DFS, left first (OBBtree.cpp:207-214,385-394)
__global__ void k_ref_offsets(uint32_t lo, uint32_t hi, RefNodes nd) {
    const uint32_t v = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= hi || nd.leaf[v]) return;
    if (v == 0) nd.tri_off[0] = 0;
    const uint32_t l = nd.left[v], r = nd.right[v];
    nd.tri_off[l] = nd.tri_off[v]; nd.tri_off[r] = nd.tri_off[v] + nd.sub_tris[l];
}
__global__ void k_ref_emit_nodes(uint32_t n_nodes, RefNodes nd, TreeRec* __restrict__ recs) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    const uint32_t rec = v == 0 ? 0u : 2u + 2u * nd.inner_rank[nd.parent[v]] + nd.side[v];
    const float* b = nd.box + 12ull * v;
    TreeRec o;
    o.q0 = make_float4(b[0], b[1], b[2], b[3]); o.q1 = make_float4(b[4], b[5], b[6], b[7]); o.q2 = make_float4(b[8], b[9], b[10], b[11]);
    if (nd.leaf[v]) o.q3 = make_float4(nd.surface[v], __uint_as_float(v == 0 ? 0u : nd.tri_off[v]), __uint_as_float(nd.count[v]), __uint_as_float(1u));
    else o.q3 = make_float4(nd.surface[v], __uint_as_float(2u + 2u * nd.inner_rank[v]), __uint_as_float(0u), __uint_as_float(0u));
    recs[rec] = o;
    if (v == 0) { TreeRec pad; pad.q0 = pad.q1 = pad.q2 = make_float4(0.f, 0.f, 0.f, 0.f); pad.q3 = make_float4(0.f, 0.f, 0.f, __uint_as_float(1u)); recs[1] = pad; }
}
__global__ void k_ref_emit_tris(uint32_t n_nodes, RefNodes nd, const uint32_t* __restrict__ idx_a, const uint32_t* __restrict__ idx_b,
                                const float* __restrict__ pos, const float* __restrict__ nrm, const uint32_t* __restrict__ vid,
                                TriRec* __restrict__ tris, float* __restrict__ nrm_out, uint32_t* __restrict__ vid_out) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes || !nd.leaf[v]) return;
    const uint32_t* idx = (nd.buf[v] ? idx_b : idx_a) + nd.begin[v];
    const uint32_t off = v == 0 ? 0u : nd.tri_off[v];
    for (uint32_t i = 0; i < nd.count[v]; ++i) {
        const uint32_t src = idx[i], dst = off + i;
        const float* p = pos + 9ull * src;
        TriRec r;
        r.t0 = make_float4(p[0], p[1], p[2], __uint_as_float(src)); r.t1 = make_float4(p[3], p[4], p[5], 0.f); r.t2 = make_float4(p[6], p[7], p[8], 0.f);
        { V3 N; float d; tt_plane(mk3(p[0], p[1], p[2]), mk3(p[3], p[4], p[5]), mk3(p[6], p[7], p[8]), N, d); r.t3 = make_float4(N.x, N.y, N.z, d); }
        tris[dst] = r;
        float* no = nrm_out + 9ull * dst;
        if (nrm) { for (int k = 0; k < 9; ++k) no[k] = nrm[9ull * src + k]; }
        else {      // TriangleNormal fallback = face normal on all three corners (Triangle.cpp:141-147,214-234)
            const V3 fn = normalize3(cross3(sub3(mk3(p[3], p[4], p[5]), mk3(p[0], p[1], p[2])), sub3(mk3(p[6], p[7], p[8]), mk3(p[0], p[1], p[2]))));
            for (int k = 0; k < 3; ++k) { no[3 * k] = fn.x; no[3 * k + 1] = fn.y; no[3 * k + 2] = fn.z; }
        }
        for (int k = 0; k < 3; ++k) vid_out[3ull * dst + k] = vid ? vid[3ull * src + k] : 3u * src + (uint32_t)k;
    }
}
__global__ void k_ref_init(uint32_t n, uint32_t* idx, RefNodes nd, uint32_t* node_count, uint32_t* list, uint32_t* list_count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = i;
    if (i == 0) { nd.begin[0] = 0; nd.count[0] = n; nd.parent[0] = 0xffffffffu; nd.side[0] = 0; nd.leaf[0] = 0; nd.tri_off[0] = 0; *node_count = 1; list[0] = 0; *list_count = 1; }
}

static inline unsigned nb(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

int imr_build_mesh_reference(imrcd_ctx* ctx, const float* h_pos, const float* h_nrm, const uint32_t* h_vid, uint64_t n_tri, MeshHost* mh) {
    cudaStream_t s = ctx->stream;
    const uint32_t n = (uint32_t)n_tri;
    const uint64_t cap = 2ull * std::max<uint64_t>(n, 1) + 2;           // every split makes two nodes and leaves hold >= 1 triangle
    DevBuf d_pos, d_nrm, d_vid, d_idx_a, d_idx_b, d_list_a, d_list_b, d_cnt, d_u32, d_u8, d_box, d_tmp;
    auto free_all = [&]() { DevBuf* all[] = { &d_pos, &d_nrm, &d_vid, &d_idx_a, &d_idx_b, &d_list_a, &d_list_b, &d_cnt, &d_u32, &d_u8, &d_box, &d_tmp }; for (DevBuf* b : all) b->release(); };
#define REF_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e); free_all(); return IMRCD_E_CUDA; } } while (0)
    REF_CUDA(d_pos.reserve(36ull * std::max<uint32_t>(n, 1), 0, s));
    if (n) REF_CUDA(cudaMemcpyAsync(d_pos.p, h_pos, 36ull * n, cudaMemcpyDefault, s));
    if (h_nrm && n) { REF_CUDA(d_nrm.reserve(36ull * n, 0, s)); REF_CUDA(cudaMemcpyAsync(d_nrm.p, h_nrm, 36ull * n, cudaMemcpyDefault, s)); }
    if (h_vid && n) { REF_CUDA(d_vid.reserve(12ull * n, 0, s)); REF_CUDA(cudaMemcpyAsync(d_vid.p, h_vid, 12ull * n, cudaMemcpyDefault, s)); }
    REF_CUDA(d_idx_a.reserve(4ull * std::max<uint32_t>(n, 1), 0, s)); REF_CUDA(d_idx_b.reserve(4ull * std::max<uint32_t>(n, 1), 0, s));
    REF_CUDA(d_list_a.reserve(4ull * cap, 0, s)); REF_CUDA(d_list_b.reserve(4ull * cap, 0, s));
    REF_CUDA(d_cnt.reserve(64, 0, s));
    REF_CUDA(d_u32.reserve(4ull * cap * 9, 0, s)); REF_CUDA(d_u8.reserve(cap * 3, 0, s)); REF_CUDA(d_box.reserve(4ull * cap * 13, 0, s));
    RefNodes nd;
    uint32_t* u = d_u32.as<uint32_t>();
    nd.begin = u; nd.count = u + cap; nd.left = u + 2 * cap; nd.right = u + 3 * cap; nd.parent = u + 4 * cap; nd.sub_tris = u + 5 * cap;
    nd.tri_off = u + 6 * cap; nd.inner_flag = u + 7 * cap; nd.inner_rank = u + 8 * cap;
    nd.side = d_u8.as<uint8_t>(); nd.leaf = nd.side + cap; nd.buf = nd.side + 2 * cap;
    nd.box = d_box.as<float>(); nd.surface = nd.box + 12 * cap;
    uint32_t* node_count = d_cnt.as<uint32_t>(); uint32_t* cnt_a = node_count + 1; uint32_t* cnt_b = node_count + 2;

    cudaEvent_t e0 = ctx->ev[6], e1 = ctx->ev[7];
    REF_CUDA(cudaEventRecord(e0, s));
    k_ref_init<<<nb(std::max<uint32_t>(n, 1), 256), 256, 0, s>>>(n, d_idx_a.as<uint32_t>(), nd, node_count, d_list_a.as<uint32_t>(), cnt_a);
    std::vector<uint32_t> level_start;          // node ids are handed out level by level, so a level is a contiguous id range
    level_start.push_back(0);
    uint32_t n_cur = 1, n_nodes = 1;
    bool flip = false;
    while (n_cur) {
        uint32_t* cur = (flip ? d_list_b : d_list_a).as<uint32_t>(); uint32_t* nxt = (flip ? d_list_a : d_list_b).as<uint32_t>();
        uint32_t* nxt_cnt = flip ? cnt_a : cnt_b;
        REF_CUDA(cudaMemsetAsync(nxt_cnt, 0, 4, s));
        k_ref_level<<<std::min<unsigned>(nb(32ull * n_cur, 128), ctx->sm_count * 16), 128, 0, s>>>(cur, n_cur, d_pos.as<float>(), (flip ? d_idx_b : d_idx_a).as<uint32_t>(),
                (flip ? d_idx_a : d_idx_b).as<uint32_t>(), nd, node_count, nxt, nxt_cnt, (uint8_t)(flip ? 1 : 0));
        uint32_t h[3];
        REF_CUDA(cudaMemcpyAsync(h, node_count, 12, cudaMemcpyDeviceToHost, s));
        REF_CUDA(cudaStreamSynchronize(s));
        level_start.push_back(n_nodes);
        n_nodes = h[0];
        n_cur = flip ? h[1] : h[2];
        flip = !flip;
        if (level_start.size() > 4096) { ctx->err = "reference build: tree deeper than 4096 levels"; free_all(); return IMRCD_E_CAPACITY; }
    }
    level_start.push_back(n_nodes);
    // level L = ids [level_start[L+1]', ...): rebuild exact ranges: level 0 = [0,1), level k = [level_start[k], level_start[k+1])
    k_ref_swap<<<nb(n_nodes, 256), 256, 0, s>>>(n_nodes, nd);
    const size_t n_levels = level_start.size() - 1;
    for (size_t L = n_levels; L-- > 0;) {
        const uint32_t lo = L == 0 ? 0u : level_start[L], hi = L == 0 ? 1u : level_start[L + 1];
        if (hi > lo) k_ref_sizes<<<nb(hi - lo, 256), 256, 0, s>>>(lo, hi, nd);
    }
    for (size_t L = 0; L < n_levels; ++L) {
        const uint32_t lo = L == 0 ? 0u : level_start[L], hi = L == 0 ? 1u : level_start[L + 1];
        if (hi > lo) k_ref_offsets<<<nb(hi - lo, 256), 256, 0, s>>>(lo, hi, nd);
    }
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, nd.inner_flag, nd.inner_rank, (int)n_nodes, s);
    REF_CUDA(d_tmp.reserve(scan_bytes, 0, s));
    cub::DeviceScan::ExclusiveSum(d_tmp.p, scan_bytes, nd.inner_flag, nd.inner_rank, (int)n_nodes, s);
    uint32_t last_rank = 0, last_flag = 0;
    REF_CUDA(cudaMemcpyAsync(&last_rank, nd.inner_rank + (n_nodes - 1), 4, cudaMemcpyDeviceToHost, s));
    REF_CUDA(cudaMemcpyAsync(&last_flag, nd.inner_flag + (n_nodes - 1), 4, cudaMemcpyDeviceToHost, s));
    REF_CUDA(cudaStreamSynchronize(s));
    const uint32_t n_inner = last_rank + last_flag;
    const uint64_t n_rec = 2ull + 2ull * n_inner;
    int rc = imr_mesh_arena_alloc(ctx, n_rec, n, mh);
    if (rc) { free_all(); return rc; }
    TreeRec* recs = ctx->d_recs.as<TreeRec>() + mh->dev.rec_base;
    k_ref_emit_nodes<<<nb(n_nodes, 256), 256, 0, s>>>(n_nodes, nd, recs);
    if (n) k_ref_emit_tris<<<nb(n_nodes, 128), 128, 0, s>>>(n_nodes, nd, d_idx_a.as<uint32_t>(), d_idx_b.as<uint32_t>(), d_pos.as<float>(),
            h_nrm ? d_nrm.as<float>() : nullptr, h_vid ? d_vid.as<uint32_t>() : nullptr, ctx->d_tris.as<TriRec>() + mh->dev.tri_base,
            ctx->d_tri_nrm.as<float>() + 9ull * mh->dev.tri_base, ctx->d_tri_vid.as<uint32_t>() + 3ull * mh->dev.tri_base);
    REF_CUDA(cudaEventRecord(e1, s));
    TreeRec root;
    REF_CUDA(cudaMemcpyAsync(&root, recs, sizeof(TreeRec), cudaMemcpyDeviceToHost, s));
    REF_CUDA(cudaStreamSynchronize(s));
    REF_CUDA(cudaGetLastError());
    cudaEventElapsedTime(&mh->build_ms, e0, e1);
    const float rb[12] = { root.q0.x, root.q0.y, root.q0.z, root.q0.w, root.q1.x, root.q1.y, root.q1.z, root.q1.w, root.q2.x, root.q2.y, root.q2.z, root.q2.w };
    memcpy(mh->root_box, rb, 48);
    free_all();
#undef REF_CUDA
    return IMRCD_OK;
}

// ---- unit-level hook: OBB::CreateOBBfromPoints of a point cloud (OBB.cpp:33-89) -----------------------------
__global__ void k_ref_fit_points(const float* pts, uint64_t n, float* out12) {
    RawPoints rp; rp.pts = pts;
    const Box b = ref_fit_warp(rp, n, threadIdx.x & 31u);
    if (threadIdx.x == 0) {
        out12[0] = b.c.x; out12[1] = b.c.y; out12[2] = b.c.z; out12[3] = b.u.x; out12[4] = b.u.y; out12[5] = b.u.z;
        out12[6] = b.v.x; out12[7] = b.v.y; out12[8] = b.v.z; out12[9] = b.w.x; out12[10] = b.w.y; out12[11] = b.w.z;
    }
}

extern "C" int imrcd_test_obb_fit(imrcd_ctx* ctx, uint64_t n_points, const float* points, float* out12) {
    if (!ctx) return IMRCD_E_ARG;
    if (!out12 || (n_points && !points)) { ctx->err = "imrcd_test_obb_fit: bad argument"; return IMRCD_E_ARG; }
    cudaSetDevice(ctx->device);
    float* d_pts = nullptr; float* d_out = nullptr;
    int rc = IMRCD_OK;
    if (cudaMalloc(&d_pts, 12 * (n_points ? n_points : 1)) != cudaSuccess || cudaMalloc(&d_out, 48) != cudaSuccess) { ctx->err = "imrcd_test_obb_fit: alloc failed"; rc = IMRCD_E_CUDA; }
    if (rc == IMRCD_OK) {
        if (n_points) cudaMemcpyAsync(d_pts, points, 12 * n_points, cudaMemcpyHostToDevice, ctx->stream);
        k_ref_fit_points<<<1, 32, 0, ctx->stream>>>(d_pts, n_points, d_out);
        cudaMemcpyAsync(out12, d_out, 48, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) { ctx->err = "imrcd_test_obb_fit: kernel failed"; rc = IMRCD_E_CUDA; }
    }
    if (d_pts) cudaFree(d_pts);
    if (d_out) cudaFree(d_out);
    return rc;
}
