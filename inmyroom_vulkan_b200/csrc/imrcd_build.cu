// imrcd_build.cu -- GPU OBB-tree construction: the replacement for OBBtree::OBBtree(std::vector<Triangle>&&)
// (IMR/src/Geometry/OBBtree.cpp:321-358 and the recursive OBBtreeSplitBuildNode ctor :8-108).
//
// IMRCD_BUILD_MORTON (default): bottom-up, Morton-ordered.
//   1. centroid bounds, 63-bit Morton keys, radix sort                         (k_bounds, k_morton, cub)
//   2. triangles gathered into sorted order = leaf order                       (k_gather_tris)
//   3. binary radix tree over the sorted keys (Karras 2012)                    (k_radix_tree)
//   4. subtrees of <= 4 triangles collapse into leaves (OBBtree.h:49); kept inner nodes get a
//      children pair slot by an exclusive scan -> sibling-adjacent 64-B records (k_flag_inner, cub scan, k_assign)
//   5. raw second moments in FP64 reduced bottom-up with one atomic ticket per node (k_moments)
//   6. per node: covariance -> closed-form symmetric 3x3 eigen-solve -> box axes (k_axes)
//   7. every triangle walks root -> leaf once, projecting its points on each ancestor's axes with
//      warp-aggregated atomic min/max in FP64                                   (k_extents)
//   8. centre + half-extent vectors, rounded to FP32 and padded outward so that the FP32 box
//      contains its triangles (the reference's +2*FLT_EPSILON pad, OBB.cpp:123, does not guarantee that) (k_finalize)
// The tree differs from the reference's top-down tree by construction; parity for this mode is asserted on
// tree-independent outputs (tests/test_gpu_build.py).
#include "imrcd_internal.cuh"
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

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {   // spread 21 bits to every third bit
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
        unsigned long long q = (unsigned long long)(u * 2097151.0);
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
    if (i >= n_int) return;
    flag[i] = (last[i] - first[i] + 1u > LEAF_MAX) ? 1u : 0u;
}

// per-record build descriptor
struct RecDesc { uint32_t first, last, split, child; int src; uint32_t kind; };   // kind 0 inner, 1 leaf ; src >= 0 internal idx, < 0 ~tri

__global__ void k_assign(int n_int, const int* __restrict__ left, const int* __restrict__ right, const int* __restrict__ parent_int,
                         const uint32_t* __restrict__ first, const uint32_t* __restrict__ last, const uint32_t* __restrict__ flag,
                         const uint32_t* __restrict__ slot, RecDesc* __restrict__ desc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_int || !flag[i]) return;
    // my own record: root -> 0, else the slot my parent reserved for its children
    uint32_t myrec = 0;
    if (i != 0) { int p = parent_int[i]; myrec = 2u + 2u * slot[p] + (left[p] == i ? 0u : 1u); }
    const uint32_t child_base = 2u + 2u * slot[i];
    const int lc = left[i], rc = right[i];
    RecDesc me; me.first = first[i]; me.last = last[i]; me.child = child_base; me.src = i; me.kind = 0u;
    me.split = (lc >= 0) ? last[lc] : (uint32_t)(~lc);
    desc[myrec] = me;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const int c = side ? rc : lc;
        if (c >= 0 && flag[c]) continue;               // a kept inner node writes its own record
        RecDesc d;
        if (c >= 0) { d.first = first[c]; d.last = last[c]; d.src = c; }
        else { d.first = d.last = (uint32_t)(~c); d.src = c; }
        d.split = d.last; d.child = d.first; d.kind = 1u;
        desc[child_base + side] = d;
    }
}

// ---- 5. FP64 raw moments, bottom-up ---------------------------------------------------------------
// m[0..2] = sum(p - o), m[3..8] = sum of (xx, yy, zz, xy, xz, yz) of (p - o), m[9] = point count
__device__ __forceinline__ void tri_moments(const TriRec& t, const double o[3], double m[10]) {
    const float px[3] = { t.t0.x, t.t1.x, t.t2.x }, py[3] = { t.t0.y, t.t1.y, t.t2.y }, pz[3] = { t.t0.z, t.t1.z, t.t2.z };
#pragma unroll
    for (int k = 0; k < 10; ++k) m[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double x = (double)px[k] - o[0], y = (double)py[k] - o[1], z = (double)pz[k] - o[2];
        m[0] += x; m[1] += y; m[2] += z;
        m[3] += x * x; m[4] += y * y; m[5] += z * z; m[6] += x * y; m[7] += x * z; m[8] += y * z;
    }
    m[9] = 3.0;
}

__global__ void k_moments(int n, const TriRec* __restrict__ tris, const int* __restrict__ left, const int* __restrict__ right,
                          const int* __restrict__ parent_int, const int* __restrict__ parent_leaf,
                          const uint32_t* __restrict__ bounds, double* mom, int* ticket) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double o[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) o[a] = 0.5 * ((double)f32_unsortable(bounds[a]) + (double)f32_unsortable(bounds[3 + a]));
    double m[10];
    tri_moments(tris[t], o, m);
    int me = ~t;                       // child encoding of the node whose total `m` holds
    int cur = parent_leaf[t];
    while (cur >= 0) {
        __threadfence();               // my subtree total (stored below) is visible before I take the ticket
        if (atomicAdd(&ticket[cur], 1) == 0) return;      // first child to arrive leaves; the second one finishes the node
        const int l = left[cur], r = right[cur];
        const int sib = (l == me) ? r : l;
        if (sib >= 0) {
#pragma unroll
            for (int k = 0; k < 10; ++k) m[k] += __ldcg(mom + 10ull * sib + k);
        } else {
            double sm[10];
            tri_moments(tris[~sib], o, sm);
#pragma unroll
            for (int k = 0; k < 10; ++k) m[k] += sm[k];
        }
        double* slot = mom + 10ull * cur;
#pragma unroll
        for (int k = 0; k < 10; ++k) __stcg(slot + k, m[k]);
        me = cur;
        cur = parent_int[cur];
    }
}

// ---- 6. axes: covariance -> closed-form symmetric eigen-solve ---------------------------------------
// Non-iterative symmetric 3x3 eigenvectors (trigonometric eigenvalues; eigenvector of the best separated
// eigenvalue from cross products of rows, the other two from the 2x2 problem in its orthogonal complement).
__device__ __forceinline__ void cross_d(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot_d(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

__device__ void sym_eig3_axes(double a00, double a11, double a22, double a01, double a02, double a12, double ax[9]) {
    // default: coordinate axes
    ax[0] = 1; ax[1] = 0; ax[2] = 0; ax[3] = 0; ax[4] = 1; ax[5] = 0; ax[6] = 0; ax[7] = 0; ax[8] = 1;
    const double mxabs = fmax(fmax(fabs(a00), fabs(a11)), fmax(fabs(a22), fmax(fabs(a01), fmax(fabs(a02), fabs(a12)))));
    if (!(mxabs > 0.0) || !isfinite(mxabs)) return;
    const double inv = 1.0 / mxabs;                  // scale to [-1,1] for robustness
    a00 *= inv; a11 *= inv; a22 *= inv; a01 *= inv; a02 *= inv; a12 *= inv;
    const double p1 = a01 * a01 + a02 * a02 + a12 * a12;
    if (p1 < 1e-30) return;                          // already diagonal
    const double q = (a00 + a11 + a22) / 3.0;
    const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
    const double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1) / 6.0);
    const double ip = 1.0 / p;
    const double c00 = b00 * ip, c11 = b11 * ip, c22 = b22 * ip, c01 = a01 * ip, c02 = a02 * ip, c12 = a12 * ip;
    double hd = 0.5 * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02));
    hd = fmin(1.0, fmax(-1.0, hd));
    const double ang = acos(hd) / 3.0;
    const double e_hi = q + 2.0 * p * cos(ang);
    const double e_lo = q + 2.0 * p * cos(ang + 2.0943951023931954923);
    const double e_mid = 3.0 * q - e_hi - e_lo;
    // eigenvector of the best separated eigenvalue
    const bool use_hi = (e_hi - e_mid) >= (e_mid - e_lo);
    const double ev = use_hi ? e_hi : e_lo;
    const double r0[3] = { a00 - ev, a01, a02 }, r1[3] = { a01, a11 - ev, a12 }, r2[3] = { a02, a12, a22 - ev };
    double c0[3], c1[3], c2[3];
    cross_d(r0, r1, c0); cross_d(r0, r2, c1); cross_d(r1, r2, c2);
    const double d0 = dot_d(c0, c0), d1 = dot_d(c1, c1), d2 = dot_d(c2, c2);
    double w[3]; double dmax = d0; w[0] = c0[0]; w[1] = c0[1]; w[2] = c0[2];
    if (d1 > dmax) { dmax = d1; w[0] = c1[0]; w[1] = c1[1]; w[2] = c1[2]; }
    if (d2 > dmax) { dmax = d2; w[0] = c2[0]; w[1] = c2[1]; w[2] = c2[2]; }
    if (!(dmax > 1e-60)) return;
    const double iw = 1.0 / sqrt(dmax);
    w[0] *= iw; w[1] *= iw; w[2] *= iw;
    // orthonormal complement (u, v) of w
    double u[3], v[3];
    if (fabs(w[0]) > fabs(w[1])) { const double il = 1.0 / sqrt(w[0] * w[0] + w[2] * w[2]); u[0] = -w[2] * il; u[1] = 0.0; u[2] = w[0] * il; }
    else { const double il = 1.0 / sqrt(w[1] * w[1] + w[2] * w[2]); u[0] = 0.0; u[1] = w[2] * il; u[2] = -w[1] * il; }
    cross_d(w, u, v);
    // 2x2 problem of A restricted to span(u, v)
    const double Au[3] = { a00 * u[0] + a01 * u[1] + a02 * u[2], a01 * u[0] + a11 * u[1] + a12 * u[2], a02 * u[0] + a12 * u[1] + a22 * u[2] };
    const double Av[3] = { a00 * v[0] + a01 * v[1] + a02 * v[2], a01 * v[0] + a11 * v[1] + a12 * v[2], a02 * v[0] + a12 * v[1] + a22 * v[2] };
    const double m00 = dot_d(u, Au), m01 = dot_d(u, Av), m11 = dot_d(v, Av);
    const double th = 0.5 * atan2(2.0 * m01, m00 - m11);
    const double cs = cos(th), sn = sin(th);
    double e1[3] = { cs * u[0] + sn * v[0], cs * u[1] + sn * v[1], cs * u[2] + sn * v[2] };
    double e2[3];
    cross_d(w, e1, e2);
    ax[0] = w[0]; ax[1] = w[1]; ax[2] = w[2]; ax[3] = e1[0]; ax[4] = e1[1]; ax[5] = e1[2]; ax[6] = e2[0]; ax[7] = e2[1]; ax[8] = e2[2];
#pragma unroll
    for (int k = 0; k < 9; ++k) if (!isfinite(ax[k])) { ax[0] = 1; ax[1] = 0; ax[2] = 0; ax[3] = 0; ax[4] = 1; ax[5] = 0; ax[6] = 0; ax[7] = 0; ax[8] = 1; break; }
}

__global__ void k_axes(uint32_t n_rec, const RecDesc* __restrict__ desc, const TriRec* __restrict__ tris, const double* __restrict__ mom,
                       const uint32_t* __restrict__ bounds, double* __restrict__ axes, unsigned long long* __restrict__ ext) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    unsigned long long* e = ext + 6ull * r;
#pragma unroll
    for (int k = 0; k < 3; ++k) { e[2 * k] = 0xffffffffffffffffull; e[2 * k + 1] = 0ull; }
    double* ax = axes + 9ull * r;
    if (r == 1) { for (int k = 0; k < 9; ++k) ax[k] = (k % 4 == 0) ? 1.0 : 0.0; return; }   // padding record
    const RecDesc d = desc[r];
    double m[10];
    if (d.src >= 0) {
#pragma unroll
        for (int k = 0; k < 10; ++k) m[k] = mom[10ull * d.src + k];
    } else {
        double o[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) o[a] = 0.5 * ((double)f32_unsortable(bounds[a]) + (double)f32_unsortable(bounds[3 + a]));
        tri_moments(tris[~d.src], o, m);
    }
    const double in = 1.0 / m[9];
    const double mx = m[0] * in, my = m[1] * in, mz = m[2] * in;
    sym_eig3_axes(m[3] * in - mx * mx, m[4] * in - my * my, m[5] * in - mz * mz, m[6] * in - mx * my, m[7] * in - mx * mz, m[8] * in - my * mz, ax);
}

// ---- 7. extents: each triangle walks root -> leaf --------------------------------------------------
// Triangles are in leaf (= sorted) order and every node covers a contiguous range of them, so the lanes of a warp that sit in the same node are
// consecutive: their minima / maxima are folded by a segmented scan first, and a run that fills whole warps is folded across the block in
// shared memory, before one atomic per run and box face.  (First version: plain per-lane atomics unless the whole warp was in one node; the 6
// words of the top nodes then took one atomic per warp of the mesh each: 6.5 of the 11 ms of kernel time of a 10 M-triangle build.)
#define EXT_BLOCK 512
__global__ void __launch_bounds__(EXT_BLOCK)
k_extents(uint32_t n, const TriRec* __restrict__ tris, const RecDesc* __restrict__ desc, const double* __restrict__ axes,
          unsigned long long* ext) {
    __shared__ double s_v[EXT_BLOCK / 32][6];
    __shared__ uint32_t s_node[EXT_BLOCK / 32];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool valid = t < n;
    double px[3] = { 0, 0, 0 }, py[3] = { 0, 0, 0 }, pz[3] = { 0, 0, 0 };
    if (valid) {
        const TriRec tr = tris[t];
        px[0] = tr.t0.x; py[0] = tr.t0.y; pz[0] = tr.t0.z; px[1] = tr.t1.x; py[1] = tr.t1.y; pz[1] = tr.t1.z; px[2] = tr.t2.x; py[2] = tr.t2.y; pz[2] = tr.t2.z;
    }
    uint32_t node = 0;
    bool active = valid;
    for (int depth = 0; depth < 4096 && __syncthreads_or(active); ++depth) {
        double v[6] = { INFINITY, -INFINITY, INFINITY, -INFINITY, INFINITY, -INFINITY };      // min, max per axis
        RecDesc d; d.kind = 1u; d.child = 0; d.split = 0;
        if (active) {
            d = desc[node];
            const double* ax = axes + 9ull * node;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const double ux = ax[3 * a], uy = ax[3 * a + 1], uz = ax[3 * a + 2];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const double pr = ux * px[k] + uy * py[k] + uz * pz[k];
                    v[2 * a] = fmin(v[2 * a], pr); v[2 * a + 1] = fmax(v[2 * a + 1], pr);
                }
            }
        }
        // segmented scan over runs of equal node in consecutive lanes (inactive lanes: node id ~0)
        const uint32_t key = active ? node : 0xffffffffu;
        const uint32_t prev = __shfl_up_sync(FULL_MASK, key, 1);
        const uint32_t heads = __ballot_sync(FULL_MASK, lane == 0u || prev != key);
        const uint32_t head = 31u - (uint32_t)__clz(heads & (0xffffffffu >> (31u - lane)));
        const uint32_t after = heads & ~(0xffffffffu >> (31u - lane));
        const uint32_t tail = after ? (uint32_t)__ffs(after) - 2u : 31u;
#pragma unroll
        for (uint32_t o = 1; o < 32u; o <<= 1) {
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const double w = __shfl_up_sync(FULL_MASK, v[q], o);
                if (lane >= head + o) v[q] = (q & 1) ? fmax(v[q], w) : fmin(v[q], w);
            }
        }
        const bool whole = heads == 1u && active;                       // the warp is one run (uniform in lane 31's view: key != ~0)
        // runs that fill whole warps go through shared memory: the last warp of a run of equal nodes adds up the run
        if (lane == 31u) {
            s_node[warp] = whole ? node : 0xffffffffu;
            if (whole) {
#pragma unroll
                for (int q = 0; q < 6; ++q) s_v[warp][q] = v[q];
            }
        }
        __syncthreads();
        if (lane == tail && active) {
            bool publish = true;
            if (whole) {
                if (warp + 1u < EXT_BLOCK / 32 && s_node[warp + 1u] == node) publish = false;      // a later warp of the block carries the run on
                else {
                    for (int w = (int)warp - 1; w >= 0 && s_node[w] == node; --w) {
#pragma unroll
                        for (int q = 0; q < 6; ++q) v[q] = (q & 1) ? fmax(v[q], s_v[w][q]) : fmin(v[q], s_v[w][q]);
                    }
                }
            }
            if (publish) {
#pragma unroll
                for (int a = 0; a < 3; ++a) { atomicMin(&ext[6ull * node + 2 * a], f64_sortable(v[2 * a])); atomicMax(&ext[6ull * node + 2 * a + 1], f64_sortable(v[2 * a + 1])); }
            }
        }
        if (active) {
            if (d.kind == 1u) active = false;
            else node = d.child + (t > d.split ? 1u : 0u);
        }
    }
}

// ---- 8. boxes -------------------------------------------------------------------------------------
// Outward padding of a box whose centre has components up to cmax and whose largest half extent is hmax.  It has to cover (a) the FP32
// rounding of the centre and of the side vectors and (b) the rounding of the 15-axis SAT itself (Paralgram.cpp:17-173 runs in FP32 on
// coordinates of this size): two tight boxes around triangles that cross with a penetration depth of a few ulp would otherwise be called
// separated and the hit lost (measured on BASELINE config 2: ~1e-4 of the hits with 4 ulp; the reference's own pad is 2 * FLT_EPSILON
// absolute, OBB.cpp:123, but its boxes are loose).  32 ulp of the box's scale, plus the reference's absolute pad.
__device__ __forceinline__ double box_pad(double cmax, double hmax) { return 32.0 * 5.9604644775390625e-8 * (cmax + hmax) + (double)FLT_EPSILON; }

__global__ void k_finalize_boxes(uint32_t n_rec, const RecDesc* __restrict__ desc, const double* __restrict__ axes,
                                 const unsigned long long* __restrict__ ext, TreeRec* __restrict__ recs) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    TreeRec out;
    if (r == 1) { out.q0 = out.q1 = out.q2 = out.q3 = make_float4(0.f, 0.f, 0.f, 0.f); out.q3.w = __uint_as_float(1u); recs[r] = out; return; }
    const RecDesc d = desc[r];
    const double* ax = axes + 9ull * r;
    const unsigned long long* e = ext + 6ull * r;
    double c[3] = { 0, 0, 0 }, half[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double mn = f64_unsortable(e[2 * a]), mx = f64_unsortable(e[2 * a + 1]);
        const double mid = 0.5 * (mx + mn);
        half[a] = 0.5 * (mx - mn);
        c[0] += mid * ax[3 * a]; c[1] += mid * ax[3 * a + 1]; c[2] += mid * ax[3 * a + 2];
    }
    const float cf[3] = { (float)c[0], (float)c[1], (float)c[2] };
    // outward padding: covers the FP32 rounding of the centre and of the side vectors, plus the reference's own pad
    const double cmax = fmax(fabs(c[0]), fmax(fabs(c[1]), fabs(c[2])));
    const double hmax = fmax(half[0], fmax(half[1], half[2]));
    const double pad = box_pad(cmax, hmax);
    float s[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double h = half[a] * (1.0 + 2.384185791015625e-7) + pad;
        s[3 * a] = (float)(h * ax[3 * a]); s[3 * a + 1] = (float)(h * ax[3 * a + 1]); s[3 * a + 2] = (float)(h * ax[3 * a + 2]);
    }
    out.q0 = make_float4(cf[0], cf[1], cf[2], s[0]);
    out.q1 = make_float4(s[1], s[2], s[3], s[4]);
    out.q2 = make_float4(s[5], s[6], s[7], s[8]);
    Box b; b.c = mk3(cf[0], cf[1], cf[2]); b.u = mk3(s[0], s[1], s[2]); b.v = mk3(s[3], s[4], s[5]); b.w = mk3(s[6], s[7], s[8]);
    const float surf = box_surface(b);
    if (d.kind == 1u) out.q3 = make_float4(surf, __uint_as_float(d.first), __uint_as_float(d.last - d.first + 1u), __uint_as_float(1u));
    else out.q3 = make_float4(surf, __uint_as_float(d.child), __uint_as_float(0u), __uint_as_float(0u));
    recs[r] = out;
}

// single-leaf meshes (<= 4 triangles): OBBtree.cpp:346-356
__global__ void k_desc_single_leaf(RecDesc* desc, uint32_t n) {
    RecDesc d; d.first = 0; d.last = n ? n - 1 : 0; d.split = d.last; d.child = 0; d.src = 0; d.kind = 1u;
    desc[0] = d; desc[1] = d;
}
// moments of a <= 4-triangle mesh into mom[0]
__global__ void k_moments_small(uint32_t n, const TriRec* __restrict__ tris, const uint32_t* __restrict__ bounds, double* mom) {
    double o[3];
    for (int a = 0; a < 3; ++a) o[a] = 0.5 * ((double)f32_unsortable(bounds[a]) + (double)f32_unsortable(bounds[3 + a]));
    double acc[10]; for (int k = 0; k < 10; ++k) acc[k] = 0.0;
    for (uint32_t t = 0; t < n; ++t) { double m[10]; tri_moments(tris[t], o, m); for (int k = 0; k < 10; ++k) acc[k] += m[k]; }
    for (int k = 0; k < 10; ++k) mom[k] = acc[k];
}

// ---------------------------------------------------------------------------------------------------
// Refit (BASELINE config 5: re-posed meshes): new triangle positions, same topology.  Works on any tree in the arena
// (Morton-built, reference-built or imported) and on many meshes per launch: the work of all meshes being refitted is
// concatenated and every thread finds its mesh by binary search in a small segment table.
//   k_rf_parents   parent link of every record                                     (top-down, trivial)
//   k_rf_up        leaf records: triangle range + FP64 moments, then climb to the root with one ticket per record
//   k_rf_axes / k_rf_extents / k_rf_boxes   the same fit as the Morton build (covariance -> closed-form eigen-solve ->
//                  every triangle walks root -> leaf -> outward-rounded FP32 boxes)
// The reference has no refit (SURVEY finding 4: skinned meshes never get collision trees); parity is defined against the
// reference REBUILDING its tree from the re-posed triangles (tests/test_gpu_refit.py).
// ---------------------------------------------------------------------------------------------------
struct RefitSeg { uint32_t rec_base, tri_base, n_rec, n_tri, rec_prefix, tri_prefix; };

__device__ __forceinline__ uint32_t rf_locate(const uint32_t* __restrict__ prefix, uint32_t n_seg, uint32_t g) {   // last s with prefix[s] <= g
    uint32_t lo = 0, hi = n_seg;
    while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (prefix[mid] <= g) lo = mid; else hi = mid; }
    return lo;
}
__device__ __forceinline__ void rf_origin(const TreeRec* recs, const RefitSeg& sg, double o[3]) {
    const float4 q = recs[sg.rec_base].q0;            // centre of the old root box: a well-conditioned origin for the raw moments
    o[0] = (double)q.x; o[1] = (double)q.y; o[2] = (double)q.z;
}

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

__global__ void k_rf_parents(uint32_t total_rec, const RefitSeg* __restrict__ segs, const uint32_t* __restrict__ rec_prefix, uint32_t n_seg,
                             const TreeRec* __restrict__ recs, uint32_t* __restrict__ parent, int* __restrict__ ticket) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_rec) return;
    const RefitSeg sg = segs[rf_locate(rec_prefix, n_seg, g)];
    const uint32_t r = g - sg.rec_prefix;
    ticket[g] = 0;
    if (r == 0) parent[g] = 0xffffffffu;
    if (r == 1) { parent[g] = 0xfffffffeu; return; }                       // padding record
    const float4 q3 = recs[sg.rec_base + r].q3;
    if (__float_as_uint(q3.w) == 0u) { const uint32_t c = __float_as_uint(q3.y); parent[sg.rec_prefix + c] = r; parent[sg.rec_prefix + c + 1u] = r; }
}

__global__ void k_rf_up(uint32_t total_rec, const RefitSeg* __restrict__ segs, const uint32_t* __restrict__ rec_prefix, uint32_t n_seg,
                        const TreeRec* __restrict__ recs, const TriRec* __restrict__ tris, const uint32_t* __restrict__ parent,
                        RecDesc* desc, double* mom, int* ticket) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_rec) return;
    const RefitSeg sg = segs[rf_locate(rec_prefix, n_seg, g)];
    uint32_t r = g - sg.rec_prefix;
    if (r == 1) { RecDesc d; d.first = d.last = d.split = d.child = 0; d.src = 1; d.kind = 2u; desc[g] = d; return; }      // kind 2 = padding
    const float4 q3 = recs[sg.rec_base + r].q3;
    if (__float_as_uint(q3.w) == 0u) return;                               // inner records are finished by their second child
    double o[3]; rf_origin(recs, sg, o);
    RecDesc d;
    d.first = __float_as_uint(q3.y); const uint32_t cnt = __float_as_uint(q3.z);
    d.last = cnt ? d.first + cnt - 1u : d.first; d.split = d.last; d.child = d.first; d.src = (int)g; d.kind = 1u;
    double m[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) m[k] = 0.0;
    for (uint32_t i = 0; i < cnt; ++i) { double tm[10]; tri_moments(tris[sg.tri_base + d.first + i], o, tm);
#pragma unroll
        for (int k = 0; k < 10; ++k) m[k] += tm[k]; }
    desc[g] = d;
#pragma unroll
    for (int k = 0; k < 10; ++k) __stcg(mom + 10ull * g + k, m[k]);
    uint32_t first = d.first, last = d.last;
    for (;;) {
        const uint32_t p = parent[sg.rec_prefix + r];
        if (p >= 0xfffffffeu) return;                                      // reached the root
        __threadfence();
        if (atomicAdd(&ticket[sg.rec_prefix + p], 1) == 0) return;        // the sibling's subtree is not finished yet: it will continue
        const uint32_t child = __float_as_uint(recs[sg.rec_base + p].q3.y);
        const uint32_t sib = (r == child) ? child + 1u : child;
        const RecDesc sd = desc[sg.rec_prefix + sib];
#pragma unroll
        for (int k = 0; k < 10; ++k) m[k] += __ldcg(mom + 10ull * (sg.rec_prefix + sib) + k);
        const RecDesc ld = (sib == child) ? sd : RecDesc{first, last, 0, 0, 0, 0};
        first = min(first, sd.first); last = max(last, sd.last);
        RecDesc pd; pd.first = first; pd.last = last; pd.split = ld.last; pd.child = child; pd.src = (int)(sg.rec_prefix + p); pd.kind = 0u;
        desc[sg.rec_prefix + p] = pd;
#pragma unroll
        for (int k = 0; k < 10; ++k) __stcg(mom + 10ull * (sg.rec_prefix + p) + k, m[k]);
        r = p;
    }
}

__global__ void k_rf_axes(uint32_t total_rec, const RecDesc* __restrict__ desc, const double* __restrict__ mom, double* __restrict__ axes, unsigned long long* __restrict__ ext) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_rec) return;
    unsigned long long* e = ext + 6ull * g;
#pragma unroll
    for (int k = 0; k < 3; ++k) { e[2 * k] = 0xffffffffffffffffull; e[2 * k + 1] = 0ull; }
    double* ax = axes + 9ull * g;
    const RecDesc d = desc[g];
    if (d.kind == 2u) { for (int k = 0; k < 9; ++k) ax[k] = (k % 4 == 0) ? 1.0 : 0.0; return; }   // padding record
    double m[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) m[k] = mom[10ull * g + k];
    if (!(m[9] > 0.0)) { for (int k = 0; k < 9; ++k) ax[k] = (k % 4 == 0) ? 1.0 : 0.0; return; }
    const double in = 1.0 / m[9];
    const double mx = m[0] * in, my = m[1] * in, mz = m[2] * in;
    sym_eig3_axes(m[3] * in - mx * mx, m[4] * in - my * my, m[5] * in - mz * mz, m[6] * in - mx * my, m[7] * in - mx * mz, m[8] * in - my * mz, ax);
}

__global__ void __launch_bounds__(EXT_BLOCK)
k_rf_extents(uint32_t total_tri, const RefitSeg* __restrict__ segs, const uint32_t* __restrict__ tri_prefix, uint32_t n_seg,
             const TriRec* __restrict__ tris, const RecDesc* __restrict__ desc, const double* __restrict__ axes, unsigned long long* ext) {
    // same folding as k_extents: runs of equal record in consecutive lanes, whole-warp runs across the block, one atomic per run and face
    __shared__ double s_v[EXT_BLOCK / 32][6];
    __shared__ uint32_t s_node[EXT_BLOCK / 32];
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    bool active = g < total_tri;
    RefitSeg sg; sg.rec_base = sg.tri_base = sg.n_rec = sg.n_tri = sg.rec_prefix = sg.tri_prefix = 0u;
    uint32_t t = 0;
    double px[3] = { 0, 0, 0 }, py[3] = { 0, 0, 0 }, pz[3] = { 0, 0, 0 };
    if (active) {
        sg = segs[rf_locate(tri_prefix, n_seg, g)];
        t = g - sg.tri_prefix;
        const TriRec tr = tris[sg.tri_base + t];
        px[0] = tr.t0.x; px[1] = tr.t1.x; px[2] = tr.t2.x; py[0] = tr.t0.y; py[1] = tr.t1.y; py[2] = tr.t2.y; pz[0] = tr.t0.z; pz[1] = tr.t1.z; pz[2] = tr.t2.z;
    }
    uint32_t node = 0;
    for (int depth = 0; depth < 4096 && __syncthreads_or(active); ++depth) {
        double v[6] = { INFINITY, -INFINITY, INFINITY, -INFINITY, INFINITY, -INFINITY };
        RecDesc d; d.kind = 1u; d.child = 0; d.split = 0;
        const uint32_t rec = sg.rec_prefix + node;
        if (active) {
            d = desc[rec];
            const double* ax = axes + 9ull * rec;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
#pragma unroll
                for (int k = 0; k < 3; ++k) { const double pr = ax[3 * a] * px[k] + ax[3 * a + 1] * py[k] + ax[3 * a + 2] * pz[k]; v[2 * a] = fmin(v[2 * a], pr); v[2 * a + 1] = fmax(v[2 * a + 1], pr); }
            }
        }
        const uint32_t key = active ? rec : 0xffffffffu;
        const uint32_t prev = __shfl_up_sync(FULL_MASK, key, 1);
        const uint32_t heads = __ballot_sync(FULL_MASK, lane == 0u || prev != key);
        const uint32_t head = 31u - (uint32_t)__clz(heads & (0xffffffffu >> (31u - lane)));
        const uint32_t after = heads & ~(0xffffffffu >> (31u - lane));
        const uint32_t tail = after ? (uint32_t)__ffs(after) - 2u : 31u;
#pragma unroll
        for (uint32_t o = 1; o < 32u; o <<= 1) {
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const double w = __shfl_up_sync(FULL_MASK, v[q], o);
                if (lane >= head + o) v[q] = (q & 1) ? fmax(v[q], w) : fmin(v[q], w);
            }
        }
        const bool whole = heads == 1u && active;
        if (lane == 31u) {
            s_node[warp] = whole ? rec : 0xffffffffu;
            if (whole) {
#pragma unroll
                for (int q = 0; q < 6; ++q) s_v[warp][q] = v[q];
            }
        }
        __syncthreads();
        if (lane == tail && active) {
            bool publish = true;
            if (whole) {
                if (warp + 1u < EXT_BLOCK / 32 && s_node[warp + 1u] == rec) publish = false;
                else {
                    for (int w = (int)warp - 1; w >= 0 && s_node[w] == rec; --w) {
#pragma unroll
                        for (int q = 0; q < 6; ++q) v[q] = (q & 1) ? fmax(v[q], s_v[w][q]) : fmin(v[q], s_v[w][q]);
                    }
                }
            }
            if (publish) {
                unsigned long long* e = ext + 6ull * rec;
#pragma unroll
                for (int a = 0; a < 3; ++a) { atomicMin(&e[2 * a], f64_sortable(v[2 * a])); atomicMax(&e[2 * a + 1], f64_sortable(v[2 * a + 1])); }
            }
        }
        if (active) {
            if (d.kind == 1u) active = false;
            else node = d.child + (t > d.split ? 1u : 0u);
        }
    }
}

__global__ void k_rf_boxes(uint32_t total_rec, const RefitSeg* __restrict__ segs, const uint32_t* __restrict__ rec_prefix, uint32_t n_seg,
                           const RecDesc* __restrict__ desc, const double* __restrict__ axes, const unsigned long long* __restrict__ ext, TreeRec* __restrict__ recs) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_rec) return;
    const RefitSeg sg = segs[rf_locate(rec_prefix, n_seg, g)];
    const uint32_t r = g - sg.rec_prefix;
    if (r == 1) return;                                                   // padding record stays as it is
    const RecDesc d = desc[g];
    const double* ax = axes + 9ull * g;
    const unsigned long long* e = ext + 6ull * g;
    double c[3] = { 0, 0, 0 }, half[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double mn = f64_unsortable(e[2 * a]), mx = f64_unsortable(e[2 * a + 1]);
        const double mid = 0.5 * (mx + mn);
        half[a] = 0.5 * (mx - mn);
        c[0] += mid * ax[3 * a]; c[1] += mid * ax[3 * a + 1]; c[2] += mid * ax[3 * a + 2];
    }
    const float cf[3] = { (float)c[0], (float)c[1], (float)c[2] };
    const double cmax = fmax(fabs(c[0]), fmax(fabs(c[1]), fabs(c[2])));
    const double hmax = fmax(half[0], fmax(half[1], half[2]));
    const double pad = box_pad(cmax, hmax);     // same outward rounding as k_finalize_boxes
    float s[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double h = half[a] * (1.0 + 2.384185791015625e-7) + pad;
        s[3 * a] = (float)(h * ax[3 * a]); s[3 * a + 1] = (float)(h * ax[3 * a + 1]); s[3 * a + 2] = (float)(h * ax[3 * a + 2]);
    }
    TreeRec out = recs[sg.rec_base + r];                                   // q3 (links, leaf range) is topology: unchanged
    out.q0 = make_float4(cf[0], cf[1], cf[2], s[0]); out.q1 = make_float4(s[1], s[2], s[3], s[4]); out.q2 = make_float4(s[5], s[6], s[7], s[8]);
    Box b; b.c = mk3(cf[0], cf[1], cf[2]); b.u = mk3(s[0], s[1], s[2]); b.v = mk3(s[3], s[4], s[5]); b.w = mk3(s[6], s[7], s[8]);
    out.q3.x = box_surface(b);
    recs[sg.rec_base + r] = out;
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
    DevBuf d_pos, d_nrm, d_vid, d_bounds, d_keys, d_keys2, d_idx, d_idx2, d_tmp, d_left, d_right, d_pint, d_pleaf, d_first, d_last,
           d_flag, d_slot, d_desc, d_mom, d_ticket, d_axes, d_ext;
    auto free_all = [&]() { DevBuf* all[] = { &d_pos, &d_nrm, &d_vid, &d_bounds, &d_keys, &d_keys2, &d_idx, &d_idx2, &d_tmp, &d_left, &d_right, &d_pint,
                                              &d_pleaf, &d_first, &d_last, &d_flag, &d_slot, &d_desc, &d_mom, &d_ticket, &d_axes, &d_ext };
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
    BUILD_CUDA(d_flag.reserve(4ull * (n_int + 1), 0, s)); BUILD_CUDA(d_slot.reserve(4ull * (n_int + 1), 0, s));
    BUILD_CUDA(d_mom.reserve(80ull * (n_int + 1), 0, s)); BUILD_CUDA(d_ticket.reserve(4ull * (n_int + 1), 0, s));

    // every allocation happens before the timed region: sort / scan scratch, the worst-case record count (every inner node kept), the arena
    size_t tmp_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys.as<unsigned long long>(), d_keys2.as<unsigned long long>(), d_idx.as<uint32_t>(), d_idx2.as<uint32_t>(), (int)n, 0, 63, s);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_flag.as<uint32_t>(), d_slot.as<uint32_t>(), (int)(n_int + 1), s);
    BUILD_CUDA(d_tmp.reserve(std::max(tmp_bytes, scan_bytes), 0, s));
    {
        const uint64_t n_rec_max = 2ull + 2ull * n_int;
        BUILD_CUDA(d_desc.reserve(sizeof(RecDesc) * n_rec_max, 0, s));
        BUILD_CUDA(d_axes.reserve(72ull * n_rec_max, 0, s));
        BUILD_CUDA(d_ext.reserve(48ull * n_rec_max, 0, s));
        int rrc = imr_mesh_arena_reserve(ctx, n_rec_max, n);
        if (rrc) { free_all(); return rrc; }
    }
    cudaEvent_t e0 = ctx->ev[6], e1 = ctx->ev[7];
    BUILD_CUDA(cudaEventRecord(e0, s));
    // 1. keys + sort
    k_bounds_init<<<1, 32, 0, s>>>(d_bounds.as<uint32_t>());
    k_bounds<<<std::min<unsigned>(nb(n, 256), ctx->sm_count * 8), 256, 0, s>>>(n, d_pos.as<float>(), d_bounds.as<uint32_t>());
    k_morton<<<nb(n, 256), 256, 0, s>>>(n, d_pos.as<float>(), d_bounds.as<uint32_t>(), d_keys.as<unsigned long long>(), d_idx.as<uint32_t>());
    cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_keys.as<unsigned long long>(), d_keys2.as<unsigned long long>(), d_idx.as<uint32_t>(), d_idx2.as<uint32_t>(), (int)n, 0, 63, s);

    // how many records?  known only after the scan -> reserve the worst case (every inner node kept) in the arena first
    uint32_t n_inner = 0;
    if (n > LEAF_MAX) {
        // 3. radix tree + 4. flags/scan
        k_radix_tree<<<nb(n_int, 128), 128, 0, s>>>((int)n, d_keys2.as<unsigned long long>(), d_left.as<int>(), d_right.as<int>(), d_pint.as<int>(),
                                                     d_pleaf.as<int>(), d_first.as<uint32_t>(), d_last.as<uint32_t>());
        BUILD_CUDA(cudaMemsetAsync(d_flag.p, 0, 4ull * (n_int + 1), s));
        k_flag_inner<<<nb(n_int, 256), 256, 0, s>>>((int)n_int, d_first.as<uint32_t>(), d_last.as<uint32_t>(), d_flag.as<uint32_t>());
        cub::DeviceScan::ExclusiveSum(d_tmp.p, scan_bytes, d_flag.as<uint32_t>(), d_slot.as<uint32_t>(), (int)(n_int + 1), s);
        BUILD_CUDA(cudaMemcpyAsync(&n_inner, d_slot.as<uint32_t>() + n_int, 4, cudaMemcpyDeviceToHost, s));
        BUILD_CUDA(cudaStreamSynchronize(s));
    }
    const uint64_t n_rec = 2ull + 2ull * n_inner;
    int rc = imr_mesh_arena_alloc(ctx, n_rec, n, mh);
    if (rc) { free_all(); return rc; }
    TriRec* tris = ctx->d_tris.as<TriRec>() + mh->dev.tri_base;
    TreeRec* recs = ctx->d_recs.as<TreeRec>() + mh->dev.rec_base;
    BUILD_CUDA(d_desc.reserve(sizeof(RecDesc) * n_rec, 0, s));
    BUILD_CUDA(d_axes.reserve(72ull * n_rec, 0, s));
    BUILD_CUDA(d_ext.reserve(48ull * n_rec, 0, s));
    // 2. gather
    k_gather_tris<<<nb(n, 256), 256, 0, s>>>(n, d_idx2.as<uint32_t>(), d_pos.as<float>(), h_nrm ? d_nrm.as<float>() : nullptr, h_vid ? d_vid.as<uint32_t>() : nullptr,
                                             tris, ctx->d_tri_nrm.as<float>() + 9ull * mh->dev.tri_base, ctx->d_tri_vid.as<uint32_t>() + 3ull * mh->dev.tri_base);
    if (n > LEAF_MAX) {
        BUILD_CUDA(cudaMemsetAsync(d_desc.p, 0, sizeof(RecDesc) * n_rec, s));
        k_assign<<<nb(n_int, 128), 128, 0, s>>>((int)n_int, d_left.as<int>(), d_right.as<int>(), d_pint.as<int>(), d_first.as<uint32_t>(), d_last.as<uint32_t>(),
                                                d_flag.as<uint32_t>(), d_slot.as<uint32_t>(), d_desc.as<RecDesc>());
        BUILD_CUDA(cudaMemsetAsync(d_ticket.p, 0, 4ull * (n_int + 1), s));
        k_moments<<<nb(n, 128), 128, 0, s>>>((int)n, tris, d_left.as<int>(), d_right.as<int>(), d_pint.as<int>(), d_pleaf.as<int>(), d_bounds.as<uint32_t>(), d_mom.as<double>(), d_ticket.as<int>());
    } else {
        k_desc_single_leaf<<<1, 1, 0, s>>>(d_desc.as<RecDesc>(), n);
        k_moments_small<<<1, 1, 0, s>>>(n, tris, d_bounds.as<uint32_t>(), d_mom.as<double>());
    }
    k_axes<<<nb(n_rec, 128), 128, 0, s>>>((uint32_t)n_rec, d_desc.as<RecDesc>(), tris, d_mom.as<double>(), d_bounds.as<uint32_t>(), d_axes.as<double>(),
                                          d_ext.as<unsigned long long>());
    k_extents<<<nb(n, EXT_BLOCK), EXT_BLOCK, 0, s>>>(n, tris, d_desc.as<RecDesc>(), d_axes.as<double>(), d_ext.as<unsigned long long>());
    k_finalize_boxes<<<nb(n_rec, 128), 128, 0, s>>>((uint32_t)n_rec, d_desc.as<RecDesc>(), d_axes.as<double>(), d_ext.as<unsigned long long>(), recs);
    BUILD_CUDA(cudaEventRecord(e1, s));
    TreeRec root;
    BUILD_CUDA(cudaMemcpyAsync(&root, recs, sizeof(TreeRec), cudaMemcpyDeviceToHost, s));
    BUILD_CUDA(cudaStreamSynchronize(s));
    BUILD_CUDA(cudaGetLastError());
    cudaEventElapsedTime(&mh->build_ms, e0, e1);
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
    std::vector<RefitSeg> segs; std::vector<uint32_t> recp, trip;
    uint64_t tot_rec = 0, tot_tri = 0;
    for (uint64_t k = 0; k < n_ids; ++k) {
        MeshHost& mh = ctx->meshes[ids[k]];
        RefitSeg sg; sg.rec_base = mh.dev.rec_base; sg.tri_base = mh.dev.tri_base; sg.n_rec = mh.dev.n_rec; sg.n_tri = mh.dev.n_tri;
        sg.rec_prefix = (uint32_t)tot_rec; sg.tri_prefix = (uint32_t)tot_tri;
        if (sg.n_tri == 0) continue;                    // nothing to fit
        segs.push_back(sg); recp.push_back(sg.rec_prefix); trip.push_back(sg.tri_prefix);
        tot_rec += sg.n_rec; tot_tri += sg.n_tri;
        mh.needs_refit = false;
    }
    if (segs.empty()) { if (ms_out) *ms_out = 0.f; return IMRCD_OK; }
    if (tot_rec >= (1ull << 32) || tot_tri >= (1ull << 32)) { ctx->err = "refit: too many records in one call"; return IMRCD_E_CAPACITY; }
    cudaStream_t s = ctx->stream;
    const uint32_t ns = (uint32_t)segs.size();
    IMR_CUDA(ctx, ctx->d_rf_segs.reserve(sizeof(RefitSeg) * ns + 8ull * ns, 0, s));
    RefitSeg* d_segs = ctx->d_rf_segs.as<RefitSeg>();
    uint32_t* d_recp = reinterpret_cast<uint32_t*>(d_segs + ns); uint32_t* d_trip = d_recp + ns;
    IMR_CUDA(ctx, cudaMemcpyAsync(d_segs, segs.data(), sizeof(RefitSeg) * ns, cudaMemcpyHostToDevice, s));
    IMR_CUDA(ctx, cudaMemcpyAsync(d_recp, recp.data(), 4ull * ns, cudaMemcpyHostToDevice, s));
    IMR_CUDA(ctx, cudaMemcpyAsync(d_trip, trip.data(), 4ull * ns, cudaMemcpyHostToDevice, s));
    IMR_CUDA(ctx, ctx->d_rf_scratch.reserve((sizeof(RecDesc) + 80 + 72 + 48 + 4 + 4) * tot_rec, 0, s));
    char* base = ctx->d_rf_scratch.as<char>();
    double* mom = reinterpret_cast<double*>(base); double* axes = mom + 10ull * tot_rec;
    unsigned long long* ext = reinterpret_cast<unsigned long long*>(axes + 9ull * tot_rec);
    RecDesc* desc = reinterpret_cast<RecDesc*>(ext + 6ull * tot_rec);
    uint32_t* parent = reinterpret_cast<uint32_t*>(desc + tot_rec); int* ticket = reinterpret_cast<int*>(parent + tot_rec);
    const TreeRec* recs = ctx->d_recs.as<TreeRec>(); const TriRec* tris = ctx->d_tris.as<TriRec>();
    cudaEvent_t e0 = ctx->ev[6], e1 = ctx->ev[7];
    IMR_CUDA(ctx, cudaEventRecord(e0, s));
    k_rf_parents<<<nb(tot_rec, 256), 256, 0, s>>>((uint32_t)tot_rec, d_segs, d_recp, ns, recs, parent, ticket);
    k_rf_up<<<nb(tot_rec, 128), 128, 0, s>>>((uint32_t)tot_rec, d_segs, d_recp, ns, recs, tris, parent, desc, mom, ticket);
    k_rf_axes<<<nb(tot_rec, 128), 128, 0, s>>>((uint32_t)tot_rec, desc, mom, axes, ext);
    k_rf_extents<<<nb(tot_tri, EXT_BLOCK), EXT_BLOCK, 0, s>>>((uint32_t)tot_tri, d_segs, d_trip, ns, tris, desc, axes, ext);
    k_rf_boxes<<<nb(tot_rec, 128), 128, 0, s>>>((uint32_t)tot_rec, d_segs, d_recp, ns, desc, axes, ext, ctx->d_recs.as<TreeRec>());
    IMR_CUDA(ctx, cudaEventRecord(e1, s));
    IMR_CUDA(ctx, cudaStreamSynchronize(s));
    IMR_CUDA(ctx, cudaGetLastError());
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    if (ms_out) *ms_out = ms;
    // root boxes (host copies used nowhere on the hot path, kept coherent for imrcd_mesh_info-style queries)
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
