// imrcd_fit.cu -- the OBB fit of a whole tree (or of many trees at once), shared by the Morton build and by the refit of re-posed meshes:
// the replacement for OBB::CreateOBBfromPoints / CreateAABBfromPoints (IMR/src/Geometry/OBB.cpp:33-166) applied to every node of a tree
// whose topology is given (OBBtree.cpp:8-108 builds boxes and topology together, top-down, re-walking a node's points ~11 times per level).
//
// A node's box needs (a) the covariance of its points -> axes (OBB.cpp:52-87) and (b) the extents of its points along those axes
// (OBB.cpp:105-166).  (a) is a bottom-up sum of raw moments; (b) is not: every node projects all of its triangles on its OWN axes, so a
// triangle is projected once per ancestor.  The tree is cut into TREELETS - maximal subtrees of at most FIT_T triangles, a contiguous
// slice of the leaf-ordered triangle array:
//   k_fit_treelets  one block per treelet: its triangles go to shared memory once; FP64 moments of every node bottom-up (leaf sums, then
//                   child + child), covariance -> closed-form symmetric 3x3 eigen-solve -> axes, every triangle walks treelet root -> leaf
//                   projecting on each node's axes (shared-memory min / max), boxes written; the treelet root's moments go to HBM
//   k_fit_climb     the nodes ABOVE the treelets (1 in ~FIT_T/2 of all nodes): moments by child + child with one ticket per node
//   k_fit_axes      ... their axes
//   k_fit_upper     one block per treelet again: its triangles against every ancestor above it, one warp-reduced min / max per block
//                   and ancestor, merged by atomics that are issued only when they would change the value
//   k_fit_boxes     ... their boxes
// Triangles are read twice (48 of their 64 bytes), node records written once: the pass is bandwidth-shaped, not atomic-shaped (round 1
// walked every triangle root -> leaf through FP64 atomics: 0.03 of the HBM roofline).
// Covariance sums are FP64 with explicit FMAs; projections are FP32 (explicit FMAs) of (point - node mean) on the FP32-rounded axes, and
// the box is built from those same axes in FP64 and padded outward (box_pad) so that the FP32 box contains its triangles - unlike the
// reference's, whose 2 * FLT_EPSILON pad (OBB.cpp:123) does not guarantee that.  Boxes use true PCA axes (the reference uses the ROWS of
// eig3's V, SURVEY finding 3), so this is not the reference's tree: parity for it is asserted on tree-independent outputs.
#include "imrcd_internal.cuh"
#include "imrcd_fit.cuh"
#include <algorithm>
#include <cfloat>

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ uint32_t f32_ord(float f) { uint32_t u = __float_as_uint(f); return (u >> 31) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float f32_unord(uint32_t u) { u = (u >> 31) ? (u & 0x7fffffffu) : ~u; return __uint_as_float(u); }

// ---- closed-form symmetric 3x3 eigenvectors (trigonometric eigenvalues; eigenvector of the best separated eigenvalue from cross products of
//      rows, the other two from the 2x2 problem in its orthogonal complement).  Replaces eig3's iterative tred2 / tql2 (eig3.cpp:21-254).
//      The covariance comes in FP64 and is scaled to [-1, 1]; the solve itself runs in FP32: any orthonormal frame gives a valid box (the
//      extents are measured on whatever axes come out), the axes only have to be close to the principal ones for the box to be tight. ----
__device__ __forceinline__ void cross_f(const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ float dot_f(const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

__device__ void sym_eig3_axes(double d00, double d11, double d22, double d01, double d02, double d12, float ax[9]) {
    ax[0] = 1; ax[1] = 0; ax[2] = 0; ax[3] = 0; ax[4] = 1; ax[5] = 0; ax[6] = 0; ax[7] = 0; ax[8] = 1;          // default: coordinate axes
    const double mxabs = fmax(fmax(fabs(d00), fabs(d11)), fmax(fabs(d22), fmax(fabs(d01), fmax(fabs(d02), fabs(d12)))));
    if (!(mxabs > 0.0) || !isfinite(mxabs)) return;
    const double inv = 1.0 / mxabs;
    const float a00 = (float)(d00 * inv), a11 = (float)(d11 * inv), a22 = (float)(d22 * inv), a01 = (float)(d01 * inv), a02 = (float)(d02 * inv), a12 = (float)(d12 * inv);
    const float p1 = a01 * a01 + a02 * a02 + a12 * a12;
    if (p1 < 1e-12f) return;                         // already diagonal (to FP32)
    const float q = (a00 + a11 + a22) * (1.f / 3.f);
    const float b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
    const float p = sqrtf((b00 * b00 + b11 * b11 + b22 * b22 + 2.f * p1) * (1.f / 6.f));
    const float ip = 1.f / p;
    const float c00 = b00 * ip, c11 = b11 * ip, c22 = b22 * ip, c01 = a01 * ip, c02 = a02 * ip, c12 = a12 * ip;
    float hd = 0.5f * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) + c02 * (c01 * c12 - c11 * c02));
    hd = fminf(1.f, fmaxf(-1.f, hd));
    const float ang = acosf(hd) * (1.f / 3.f);
    const float e_hi = q + 2.f * p * cosf(ang);
    const float e_lo = q + 2.f * p * cosf(ang + 2.0943951023931954923f);
    const float e_mid = 3.f * q - e_hi - e_lo;
    const bool use_hi = (e_hi - e_mid) >= (e_mid - e_lo);      // eigenvector of the best separated eigenvalue
    const float ev = use_hi ? e_hi : e_lo;
    const float r0[3] = { a00 - ev, a01, a02 }, r1[3] = { a01, a11 - ev, a12 }, r2[3] = { a02, a12, a22 - ev };
    float c0[3], c1[3], c2[3];
    cross_f(r0, r1, c0); cross_f(r0, r2, c1); cross_f(r1, r2, c2);
    const float d0 = dot_f(c0, c0), d1 = dot_f(c1, c1), d2 = dot_f(c2, c2);
    float w[3]; float dmax = d0; w[0] = c0[0]; w[1] = c0[1]; w[2] = c0[2];
    if (d1 > dmax) { dmax = d1; w[0] = c1[0]; w[1] = c1[1]; w[2] = c1[2]; }
    if (d2 > dmax) { dmax = d2; w[0] = c2[0]; w[1] = c2[1]; w[2] = c2[2]; }
    if (!(dmax > 1e-30f)) return;
    const float iw = rsqrtf(dmax);
    w[0] *= iw; w[1] *= iw; w[2] *= iw;
    float u[3], v[3];                                // orthonormal complement (u, v) of w
    if (fabsf(w[0]) > fabsf(w[1])) { const float il = rsqrtf(w[0] * w[0] + w[2] * w[2]); u[0] = -w[2] * il; u[1] = 0.f; u[2] = w[0] * il; }
    else { const float il = rsqrtf(w[1] * w[1] + w[2] * w[2]); u[0] = 0.f; u[1] = w[2] * il; u[2] = -w[1] * il; }
    cross_f(w, u, v);
    const float Au[3] = { a00 * u[0] + a01 * u[1] + a02 * u[2], a01 * u[0] + a11 * u[1] + a12 * u[2], a02 * u[0] + a12 * u[1] + a22 * u[2] };
    const float Av[3] = { a00 * v[0] + a01 * v[1] + a02 * v[2], a01 * v[0] + a11 * v[1] + a12 * v[2], a02 * v[0] + a12 * v[1] + a22 * v[2] };
    const float m00 = dot_f(u, Au), m01 = dot_f(u, Av), m11 = dot_f(v, Av);      // 2x2 problem of A restricted to span(u, v)
    const float th = 0.5f * atan2f(2.f * m01, m00 - m11);
    float sn, cs; sincosf(th, &sn, &cs);
    float e1[3] = { cs * u[0] + sn * v[0], cs * u[1] + sn * v[1], cs * u[2] + sn * v[2] };
    const float i1 = rsqrtf(dot_f(e1, e1));
    e1[0] *= i1; e1[1] *= i1; e1[2] *= i1;
    float e2[3];
    cross_f(w, e1, e2);
    const float i2 = rsqrtf(dot_f(e2, e2));
    ax[0] = w[0]; ax[1] = w[1]; ax[2] = w[2]; ax[3] = e1[0]; ax[4] = e1[1]; ax[5] = e1[2]; ax[6] = e2[0] * i2; ax[7] = e2[1] * i2; ax[8] = e2[2] * i2;
#pragma unroll
    for (int k = 0; k < 9; ++k) if (!isfinite(ax[k])) { ax[0] = 1; ax[1] = 0; ax[2] = 0; ax[3] = 0; ax[4] = 1; ax[5] = 0; ax[6] = 0; ax[7] = 0; ax[8] = 1; break; }
}

// ---- shared pieces --------------------------------------------------------------------------------------------------------------------
// raw moments of a node relative to the call's origin: m[0..2] = sum(p - o), m[3..8] = sum of (xx, yy, zz, xy, xz, yz), m[9] = points
__device__ __forceinline__ void mom_add_point(double m[10], double x, double y, double z) {
    m[0] += x; m[1] += y; m[2] += z;
    m[3] = __fma_rn(x, x, m[3]); m[4] = __fma_rn(y, y, m[4]); m[5] = __fma_rn(z, z, m[5]);
    m[6] = __fma_rn(x, y, m[6]); m[7] = __fma_rn(x, z, m[7]); m[8] = __fma_rn(y, z, m[8]);
    m[9] += 1.0;
}

// moments -> the node's frame: 12 floats = mean (the origin of the projections) + three unit axes, all rounded to FP32 once; every later
// step (projections, the box itself) uses exactly these values
__device__ __forceinline__ void frame_from_moments(const double m[10], const double o[3], float fr[12]) {
    float ax[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    double mean[3] = { o[0], o[1], o[2] };
    if (m[9] > 0.0) {
        const double in = 1.0 / m[9];
        const double mx = m[0] * in, my = m[1] * in, mz = m[2] * in;
        mean[0] += mx; mean[1] += my; mean[2] += mz;
        sym_eig3_axes(m[3] * in - mx * mx, m[4] * in - my * my, m[5] * in - mz * mz, m[6] * in - mx * my, m[7] * in - mx * mz, m[8] * in - my * mz, ax);
    }
    fr[0] = (float)mean[0]; fr[1] = (float)mean[1]; fr[2] = (float)mean[2];
#pragma unroll
    for (int k = 0; k < 9; ++k) fr[3 + k] = ax[k];
}

// a triangle's three points on the three axes of a frame: min / max per axis as orderable integers (for integer min / max reductions)
struct Ext6 { uint32_t v[6]; };
__device__ __forceinline__ Ext6 project_tri(const float fr[12], const float p[9]) {
    Ext6 e;
    float mn[3] = { INFINITY, INFINITY, INFINITY }, mx[3] = { -INFINITY, -INFINITY, -INFINITY };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float x = p[3 * k] - fr[0], y = p[3 * k + 1] - fr[1], z = p[3 * k + 2] - fr[2];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float pr = __fmaf_rn(fr[3 + 3 * a + 2], z, __fmaf_rn(fr[3 + 3 * a + 1], y, fr[3 + 3 * a] * x));
            mn[a] = fminf(mn[a], pr); mx[a] = fmaxf(mx[a], pr);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { e.v[2 * a] = f32_ord(mn[a]); e.v[2 * a + 1] = f32_ord(mx[a]); }
    return e;
}

// Outward padding of a box whose centre has components up to cmax and whose largest half extent is hmax.  It covers the FP32 rounding of
// the projections, of the centre and of the side vectors, and the rounding of the 15-axis SAT itself (Paralgram.cpp:17-173 runs in FP32 on
// coordinates of this size).  32 ulp of the box's scale, plus the reference's own absolute pad (OBB.cpp:123).
__device__ __forceinline__ double box_pad(double cmax, double hmax) { return 32.0 * 5.9604644775390625e-8 * (cmax + hmax) + (double)FLT_EPSILON; }

// frame + extents -> the box (centre + three half-extent vectors, Paralgram.h:29-35) rounded to FP32 outward, with its surface
__device__ __forceinline__ void box_from_frame(const float fr[12], const uint32_t ext[6], float4& q0, float4& q1, float4& q2, float& surface) {
    double c[3] = { (double)fr[0], (double)fr[1], (double)fr[2] }, half[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double mn = (double)f32_unord(ext[2 * a]), mx = (double)f32_unord(ext[2 * a + 1]);
        const double mid = 0.5 * (mx + mn);
        half[a] = 0.5 * (mx - mn);
        c[0] = __fma_rn(mid, (double)fr[3 + 3 * a], c[0]); c[1] = __fma_rn(mid, (double)fr[3 + 3 * a + 1], c[1]); c[2] = __fma_rn(mid, (double)fr[3 + 3 * a + 2], c[2]);
    }
    const float cf[3] = { (float)c[0], (float)c[1], (float)c[2] };
    const double cmax = fmax(fabs(c[0]), fmax(fabs(c[1]), fabs(c[2])));
    const double hmax = fmax(half[0], fmax(half[1], half[2]));
    const double pad = box_pad(cmax, hmax);
    float s[9];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double h = half[a] * (1.0 + 2.384185791015625e-7) + pad;
        s[3 * a] = (float)(h * (double)fr[3 + 3 * a]); s[3 * a + 1] = (float)(h * (double)fr[3 + 3 * a + 1]); s[3 * a + 2] = (float)(h * (double)fr[3 + 3 * a + 2]);
    }
    q0 = make_float4(cf[0], cf[1], cf[2], s[0]); q1 = make_float4(s[1], s[2], s[3], s[4]); q2 = make_float4(s[5], s[6], s[7], s[8]);
    Box b; b.c = mk3(cf[0], cf[1], cf[2]); b.u = mk3(s[0], s[1], s[2]); b.v = mk3(s[3], s[4], s[5]); b.w = mk3(s[6], s[7], s[8]);
    surface = box_surface(b);
}

__device__ __forceinline__ void store_rec(TreeRec* recs, uint32_t rec, const FitRec& f, const FitSeg& sg, const float4& q0, const float4& q1, const float4& q2, float surface, bool write_links) {
    TreeRec* out = recs + rec;
    out->q0 = q0; out->q1 = q1; out->q2 = q2;
    if (!write_links) { out->q3.x = surface; return; }                  // refit: links and leaf ranges are topology, unchanged
    if (f.kind == 1u) out->q3 = make_float4(surface, __uint_as_float(f.first - sg.tri_base), __uint_as_float(f.last - f.first + 1u), __uint_as_float(1u));
    else out->q3 = make_float4(surface, __uint_as_float(f.child - sg.rec_base), __uint_as_float(0u), __uint_as_float(0u));
}

// ---- classify the records of the call: above the treelets / treelet root / inside a treelet ------------------------------------------
__device__ __forceinline__ uint32_t seg_locate(const uint32_t* __restrict__ prefix, uint32_t n_seg, uint32_t g) {   // last s with prefix[s] <= g
    uint32_t lo = 0, hi = n_seg;
    while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (prefix[mid] <= g) lo = mid; else hi = mid; }
    return lo;
}

__global__ void k_fit_collect(uint32_t total_rec, const FitSeg* __restrict__ segs, const uint32_t* __restrict__ rec_prefix, uint32_t n_seg,
                              const FitRec* __restrict__ fit, uint32_t* __restrict__ slot_of, uint2* __restrict__ troots, uint32_t* __restrict__ uppers,
                              FitCounters* cnt, TreeRec* recs, int write_links) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_rec) return;
    const uint32_t s = seg_locate(rec_prefix, n_seg, g);
    const uint32_t rec = segs[s].rec_base + (g - rec_prefix[s]);
    const FitRec f = fit[rec];
    if (f.kind == 2u) {                                                  // the padding record beside the root
        if (write_links) { TreeRec z; z.q0 = z.q1 = z.q2 = z.q3 = make_float4(0.f, 0.f, 0.f, 0.f); z.q3.w = __uint_as_float(1u); recs[rec] = z; }
        return;
    }
    const uint32_t n = f.last - f.first + 1u;
    if (!FIT_SMALL(n, f.n_sub)) { const uint32_t sl = atomicAdd(&cnt->n_slots, 1u); slot_of[rec] = sl; uppers[atomicAdd(&cnt->n_upper, 1u)] = rec; return; }
    bool root = f.parent == 0xffffffffu;
    if (!root) { const FitRec p = fit[f.parent]; root = !FIT_SMALL(p.last - p.first + 1u, p.n_sub); }
    if (root) { const uint32_t sl = atomicAdd(&cnt->n_slots, 1u); slot_of[rec] = sl; troots[atomicAdd(&cnt->n_troot, 1u)] = make_uint2(rec, s); }
}

// ---- the treelets ---------------------------------------------------------------------------------------------------------------------
// One block per treelet, five barriers: [triangles + the subtree's records -> shared memory] [moments of every TRIANGLE relative to the
// treelet's first vertex, inclusive prefix over the treelet's triangles] [every record at once: moments = prefix difference over its
// contiguous range, frame] [every triangle walks treelet root -> leaf, min / max per record] [boxes].  Prefix differences of FP64 sums
// over <= 128 nearby triangles lose two of sixteen digits: no bottom-up order is needed, so no level loop and no idle threads at barriers.
struct FitSmem {
    float p[FIT_T][9];                      // the treelet's triangles (stride 9: conflict-free for consecutive threads)
    double pre[FIT_T + 1][10];              // pre[t] = moments of triangles [0, t)
    float fr[FIT_R][12];
    uint32_t ext[FIT_R][6];
    FitRec rec[FIT_R];                      // the subtree's records, [0] = the treelet root; parent / child rewritten to LOCAL indices
    uint32_t arena[FIT_R];                  // their arena indices
    double wtot[FIT_T / 32][10];
    uint32_t n_loc;
};

__global__ void __launch_bounds__(FIT_T)
k_fit_treelets(const FitCounters* __restrict__ cnt, const uint2* __restrict__ troots, const FitSeg* __restrict__ segs, const FitRec* __restrict__ fit,
               const TriRec* __restrict__ tris, const uint32_t* __restrict__ slot_of, double* __restrict__ mom_out, TreeRec* __restrict__ recs, int write_links,
               uint32_t* cursor) {
    extern __shared__ __align__(16) unsigned char fit_smem[];
    FitSmem& sm = *reinterpret_cast<FitSmem*>(fit_smem);
    __shared__ uint32_t s_next;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (;;) {                                                           // treelets are handed out through a counter: 1 to 128 triangles each
        __syncthreads();
        if (tid == 0) s_next = atomicAdd(cursor, 1u);
        __syncthreads();
        const uint32_t b = s_next;
        if (b >= cnt->n_troot) break;
        const uint2 tr = troots[b];
        const FitSeg sg = segs[tr.y];
        const FitRec root = fit[tr.x];
        const uint32_t n = root.last - root.first + 1u;
        __syncthreads();                                                 // the previous treelet's tables are no longer read
        float p[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        if (tid < n) {
            const float4* tp = reinterpret_cast<const float4*>(tris + root.first + tid);
            const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
            p[0] = t0.x; p[1] = t0.y; p[2] = t0.z; p[3] = t1.x; p[4] = t1.y; p[5] = t1.z; p[6] = t2.x; p[7] = t2.y; p[8] = t2.z;
#pragma unroll
            for (int k = 0; k < 9; ++k) sm.p[tid][k] = p[k];
        }
        // -- the subtree's records --
        uint32_t n_loc = root.n_sub;
        if (root.desc != 0xffffffffu) {                                   // contiguous in the arena (Morton builds): one coalesced load
            const uint32_t* src = reinterpret_cast<const uint32_t*>(fit + root.desc);
            uint32_t* dst = reinterpret_cast<uint32_t*>(sm.rec + 1);
            for (uint32_t k = tid; k < (n_loc - 1u) * 8u; k += FIT_T) dst[k] = __ldg(src + k);
            if (tid == 0) { sm.rec[0] = root; sm.arena[0] = tr.x; }
            for (uint32_t k = 1u + tid; k < n_loc; k += FIT_T) sm.arena[k] = root.desc + k - 1u;
            __syncthreads();
            for (uint32_t k = tid; k < n_loc; k += FIT_T) {              // links -> local indices
                FitRec& f = sm.rec[k];
                if (f.kind == 0u) f.child = f.child - root.desc + 1u;
                f.parent = k == 0u ? 0xffffffffu : (f.parent == tr.x ? 0u : f.parent - root.desc + 1u);
            }
        } else {                                                          // any other tree: level by level from the root
            if (tid == 0) { sm.rec[0] = root; sm.rec[0].parent = 0xffffffffu; sm.arena[0] = tr.x; sm.n_loc = 1u; }
            __syncthreads();
            uint32_t begin = 0, end = 1;
            while (begin < end) {
                for (uint32_t i = begin + tid; i < end; i += FIT_T) {
                    FitRec& f = sm.rec[i];
                    if (f.kind == 0u) {
                        const uint32_t lc = atomicAdd(&sm.n_loc, 2u), c = f.child;
                        sm.rec[lc] = fit[c]; sm.rec[lc + 1u] = fit[c + 1u]; sm.arena[lc] = c; sm.arena[lc + 1u] = c + 1u;
                        sm.rec[lc].parent = i; sm.rec[lc + 1u].parent = i;
                        f.child = lc;
                    }
                }
                __syncthreads();
                begin = end; end = sm.n_loc;
                __syncthreads();
            }
            n_loc = end;
        }
        // -- per-triangle moments relative to the treelet's first vertex, inclusive prefix over the treelet --
        const double o0[3] = { (double)__shfl_sync(FULL_MASK, p[0], 0), (double)__shfl_sync(FULL_MASK, p[1], 0), (double)__shfl_sync(FULL_MASK, p[2], 0) };
        __shared__ double s_o[3];
        if (tid == 0) { s_o[0] = o0[0]; s_o[1] = o0[1]; s_o[2] = o0[2]; }
        double m[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) m[k] = 0.0;
        __syncthreads();
        const double ox = s_o[0], oy = s_o[1], oz = s_o[2];
        if (tid < n) {
#pragma unroll
            for (int k = 0; k < 3; ++k) mom_add_point(m, (double)p[3 * k] - ox, (double)p[3 * k + 1] - oy, (double)p[3 * k + 2] - oz);
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int k = 0; k < 10; ++k) { const double v = __shfl_up_sync(FULL_MASK, m[k], o); if (lane >= (uint32_t)o) m[k] += v; }
        }
        if (lane == 31u) {
#pragma unroll
            for (int k = 0; k < 10; ++k) sm.wtot[warp][k] = m[k];
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < (int)(FIT_T / 32) - 1; ++w) if ((int)warp > w) {
#pragma unroll
            for (int k = 0; k < 10; ++k) m[k] += sm.wtot[w][k];
        }
#pragma unroll
        for (int k = 0; k < 10; ++k) sm.pre[tid + 1u][k] = m[k];
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < 10; ++k) sm.pre[0][k] = 0.0;
        }
        __syncthreads();
        // -- every record: moments of its range, frame; the root's moments (moved to the call's origin) travel up --
        const double local_o[3] = { ox, oy, oz };
        for (uint32_t i = tid; i < n_loc; i += FIT_T) {
            const FitRec f = sm.rec[i];
            const uint32_t t0 = f.first - root.first, t1 = f.last - root.first + 1u;
            double mm[10];
#pragma unroll
            for (int k = 0; k < 10; ++k) mm[k] = sm.pre[t1][k] - sm.pre[t0][k];
            float fr[12];
            frame_from_moments(mm, local_o, fr);
#pragma unroll
            for (int k = 0; k < 12; ++k) sm.fr[i][k] = fr[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) { sm.ext[i][2 * k] = 0xffffffffu; sm.ext[i][2 * k + 1] = 0u; }
            if (i == 0u) {
                const double dx = ox - sg.origin[0], dy = oy - sg.origin[1], dz = oz - sg.origin[2], cntp = mm[9];
                double* mo = mom_out + 10ull * slot_of[tr.x];
                mo[0] = mm[0] + cntp * dx; mo[1] = mm[1] + cntp * dy; mo[2] = mm[2] + cntp * dz;
                mo[3] = mm[3] + 2.0 * dx * mm[0] + cntp * dx * dx; mo[4] = mm[4] + 2.0 * dy * mm[1] + cntp * dy * dy; mo[5] = mm[5] + 2.0 * dz * mm[2] + cntp * dz * dz;
                mo[6] = mm[6] + dx * mm[1] + dy * mm[0] + cntp * dx * dy; mo[7] = mm[7] + dx * mm[2] + dz * mm[0] + cntp * dx * dz; mo[8] = mm[8] + dy * mm[2] + dz * mm[1] + cntp * dy * dz;
                mo[9] = cntp;
            }
        }
        __syncthreads();
        // -- extents: every triangle walks treelet root -> leaf; lanes of a warp that sit in the same node reduce among themselves first --
        {
            bool active = tid < n;
            uint32_t node = 0;
            const uint32_t my_tri = root.first + tid;
            while (__any_sync(FULL_MASK, active)) {
                Ext6 e;
#pragma unroll
                for (int q = 0; q < 6; ++q) e.v[q] = (q & 1) ? 0u : 0xffffffffu;
                if (active) e = project_tri(sm.fr[node], p);
                const uint32_t lead_node = __shfl_sync(FULL_MASK, node, __ffs(__ballot_sync(FULL_MASK, active)) - 1);
                const bool uniform = __all_sync(FULL_MASK, !active || node == lead_node);
                if (uniform) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) e.v[q] = (q & 1) ? __reduce_max_sync(FULL_MASK, e.v[q]) : __reduce_min_sync(FULL_MASK, e.v[q]);
                    if (lane == 0) {
#pragma unroll
                        for (int q = 0; q < 6; ++q) { if (q & 1) atomicMax(&sm.ext[lead_node][q], e.v[q]); else atomicMin(&sm.ext[lead_node][q], e.v[q]); }
                    }
                } else if (active) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) { if (q & 1) atomicMax(&sm.ext[node][q], e.v[q]); else atomicMin(&sm.ext[node][q], e.v[q]); }
                }
                if (active) {
                    const FitRec& f = sm.rec[node];
                    if (f.kind == 1u) active = false;
                    else node = f.child + (my_tri > f.split ? 1u : 0u);
                }
            }
        }
        __syncthreads();
        // -- boxes --
        for (uint32_t i = tid; i < n_loc; i += FIT_T) {
            float4 q0, q1, q2; float surface;
            box_from_frame(sm.fr[i], sm.ext[i], q0, q1, q2, surface);
            FitRec f = sm.rec[i];
            if (f.kind == 0u) f.child = sm.arena[f.child];                // back to the arena index for the record's link
            store_rec(recs, sm.arena[i], f, sg, q0, q1, q2, surface, write_links != 0);
        }
    }
}

// ---- above the treelets -----------------------------------------------------------------------------------------------------------------
// one thread per treelet root climbs: the second child to arrive at a node adds its sibling's moments and goes on
__global__ void k_fit_climb(const FitCounters* __restrict__ cnt, const uint2* __restrict__ troots, const FitRec* __restrict__ fit,
                            const uint32_t* __restrict__ slot_of, double* mom, uint32_t* ticket) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt->n_troot) return;
    uint32_t cur = troots[t].x;
    double m[10];
    { const double* mo = mom + 10ull * slot_of[cur];
#pragma unroll
      for (int k = 0; k < 10; ++k) m[k] = __ldcg(mo + k); }
    for (;;) {
        const uint32_t par = fit[cur].parent;
        if (par == 0xffffffffu) return;
        const uint32_t ps = slot_of[par];
        __threadfence();                                                 // my subtree's total is visible before I take the ticket
        if (atomicAdd(&ticket[ps], 1u) == 0u) return;                    // the sibling's subtree is not finished: it will go on from here
        const uint32_t child = fit[par].child;
        const uint32_t sib = (cur == child) ? child + 1u : child;
        const double* so = mom + 10ull * slot_of[sib];
#pragma unroll
        for (int k = 0; k < 10; ++k) m[k] += __ldcg(so + k);
        double* po = mom + 10ull * ps;
#pragma unroll
        for (int k = 0; k < 10; ++k) __stcg(po + k, m[k]);
        cur = par;
    }
}

__global__ void k_fit_axes(const FitCounters* __restrict__ cnt, const uint32_t* __restrict__ uppers, const FitSeg* __restrict__ segs, const uint32_t* __restrict__ seg_of_upper,
                           const uint32_t* __restrict__ slot_of, const double* __restrict__ mom, float* __restrict__ frames, uint32_t* __restrict__ ext) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt->n_upper) return;
    const uint32_t rec = uppers[t], sl = slot_of[rec];
    double m[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) m[k] = mom[10ull * sl + k];
    float fr[12];
    frame_from_moments(m, segs[seg_of_upper[t]].origin, fr);
#pragma unroll
    for (int k = 0; k < 12; ++k) frames[12ull * sl + k] = fr[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ext[6ull * sl + 2 * k] = 0xffffffffu; ext[6ull * sl + 2 * k + 1] = 0u; }
}

// the ancestors above each treelet, once per topology: one thread per treelet chases the parent pointers (a hundred thousand chains in
// flight hide the latency that one block per treelet would wait for, link by link) and leaves their slots in the treelet's chain row
__global__ void k_fit_chains(const FitCounters* __restrict__ cnt, const uint2* __restrict__ troots, const FitRec* __restrict__ fit, const uint32_t* __restrict__ slot_of,
                             uint32_t* __restrict__ chain, uint32_t* __restrict__ chain_len) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt->n_troot) return;
    uint32_t n = 0;
    for (uint32_t a = fit[troots[t].x].parent; a != 0xffffffffu; a = fit[a].parent) {
        if (n < FIT_CHAIN) chain[(size_t)t * FIT_CHAIN + n] = slot_of[a];
        ++n;
    }
    chain_len[t] = n;
}

// one block per treelet: its triangles against every ancestor above the treelet.  The ancestors' frames are fetched together (their
// slots are in the chain row), every warp projects its triangles on each of them and reduces (hardware integer min / max on orderable
// floats), the warps' partial results meet in shared memory, and an atomic goes out only where it would change the stored value: the
// top of the tree hears from every treelet, and almost none of them moves its extents.
__global__ void __launch_bounds__(FIT_T)
k_fit_upper(const FitCounters* __restrict__ cnt, const uint2* __restrict__ troots, const FitRec* __restrict__ fit, const TriRec* __restrict__ tris,
            const uint32_t* __restrict__ slot_of, const uint32_t* __restrict__ chain, const uint32_t* __restrict__ chain_len, const float* __restrict__ frames, uint32_t* ext,
            const TreeRec* __restrict__ recs, uint32_t* cursor) {
    __shared__ uint32_t s_part[FIT_CHAIN][FIT_T / 32][6];
    __shared__ float s_fr[FIT_CHAIN][12];
    __shared__ uint32_t s_slot[FIT_CHAIN];
    __shared__ uint32_t s_need[FIT_CHAIN];
    __shared__ uint32_t s_next;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (;;) {                                                           // handed out through a counter: how many ancestors a treelet still moves differs
        __syncthreads();
        if (tid == 0) s_next = atomicAdd(cursor, 1u);
        __syncthreads();
        const uint32_t b = s_next;
        if (b >= cnt->n_troot) break;
        const FitRec root = fit[troots[b].x];
        const uint32_t n = root.last - root.first + 1u;
        const uint32_t n_all = chain_len[b], n_anc = n_all < FIT_CHAIN ? n_all : FIT_CHAIN;
        __syncthreads();                                                 // the previous treelet's tables are no longer read
        if (tid < n_anc) s_slot[tid] = chain[(size_t)b * FIT_CHAIN + tid];
        float p[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
        if (tid < n) {
            const float4* tp = reinterpret_cast<const float4*>(tris + root.first + tid);
            const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
            p[0] = t0.x; p[1] = t0.y; p[2] = t0.z; p[3] = t1.x; p[4] = t1.y; p[5] = t1.z; p[6] = t2.x; p[7] = t2.y; p[8] = t2.z;
        }
        __syncthreads();
        for (uint32_t k = tid; k < n_anc * 12u; k += FIT_T) s_fr[k / 12u][k % 12u] = frames[12ull * s_slot[k / 12u] + k % 12u];
        __syncthreads();
        // Which ancestors can this treelet still move?  Its own box (written by k_fit_treelets) holds all of its triangles, so the box's
        // projection on an ancestor's axes bounds theirs: where that already lies inside the ancestor's extents so far, there is nothing
        // to add - the case for almost every treelet at the upper levels once the first blocks have been through.
        if (tid < n_anc) {
            const TreeRec* tb = recs + troots[b].x;
            const float4 q0 = __ldg(&tb->q0), q1 = __ldg(&tb->q1), q2 = __ldg(&tb->q2);
            const float* fr = s_fr[tid];
            const uint32_t* ex = ext + 6ull * s_slot[tid];
            bool need = false;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float ax = fr[3 + 3 * k], ay = fr[3 + 3 * k + 1], az = fr[3 + 3 * k + 2];
                const float cp = ax * (q0.x - fr[0]) + ay * (q0.y - fr[1]) + az * (q0.z - fr[2]);
                const float r = fabsf(ax * q0.w + ay * q1.x + az * q1.y) + fabsf(ax * q1.z + ay * q1.w + az * q2.x) + fabsf(ax * q2.y + ay * q2.z + az * q2.w);
                const float slack = 9.5367431640625e-7f * (fabsf(cp) + r);      // 16 ulp: the rounding of this bound and of the triangles' own projections
                const uint32_t lo = f32_ord(cp - r - slack), hi = f32_ord(cp + r + slack);
                need |= lo < *((volatile const uint32_t*)(ex + 2 * k)) || hi > *((volatile const uint32_t*)(ex + 2 * k + 1));
            }
            s_need[tid] = need ? 1u : 0u;
        }
        __syncthreads();
        for (uint32_t a = 0; a < n_anc; ++a) {
            if (!s_need[a]) continue;
            Ext6 e;
#pragma unroll
            for (int q = 0; q < 6; ++q) e.v[q] = (q & 1) ? 0u : 0xffffffffu;
            if (tid < n) e = project_tri(s_fr[a], p);
#pragma unroll
            for (int q = 0; q < 6; ++q) e.v[q] = (q & 1) ? __reduce_max_sync(FULL_MASK, e.v[q]) : __reduce_min_sync(FULL_MASK, e.v[q]);
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 6; ++q) s_part[a][warp][q] = e.v[q];
            }
        }
        __syncthreads();
        for (uint32_t k = tid; k < n_anc * 6u; k += FIT_T) {
            const uint32_t a = k / 6u, q = k % 6u;
            if (!s_need[a]) continue;
            uint32_t v = s_part[a][0][q];
#pragma unroll
            for (int w = 1; w < FIT_T / 32; ++w) v = (q & 1u) ? max(v, s_part[a][w][q]) : min(v, s_part[a][w][q]);
            uint32_t* dst = ext + 6ull * s_slot[a] + q;
            const uint32_t seen = *((volatile uint32_t*)dst);           // a stale value only costs an atomic that changes nothing
            if (q & 1u) { if (v > seen) atomicMax(dst, v); } else { if (v < seen) atomicMin(dst, v); }
        }
        if (n_all > FIT_CHAIN) {                                         // a chain longer than its row (a degenerate tree): the rest link by link
            uint32_t a = fit[troots[b].x].parent;
            for (uint32_t k = 0; k < FIT_CHAIN; ++k) a = fit[a].parent;
            for (; a != 0xffffffffu; a = fit[a].parent) {
                const uint32_t sl = slot_of[a];
                __syncthreads();
                if (tid < 12) s_fr[0][tid] = frames[12ull * sl + tid];
                __syncthreads();
                Ext6 e;
#pragma unroll
                for (int q = 0; q < 6; ++q) e.v[q] = (q & 1) ? 0u : 0xffffffffu;
                if (tid < n) e = project_tri(s_fr[0], p);
#pragma unroll
                for (int q = 0; q < 6; ++q) e.v[q] = (q & 1) ? __reduce_max_sync(FULL_MASK, e.v[q]) : __reduce_min_sync(FULL_MASK, e.v[q]);
                if (lane == 0) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) { if (q & 1) atomicMax(ext + 6ull * sl + q, e.v[q]); else atomicMin(ext + 6ull * sl + q, e.v[q]); }
                }
            }
        }
    }
}

__global__ void k_fit_boxes(const FitCounters* __restrict__ cnt, const uint32_t* __restrict__ uppers, const FitSeg* __restrict__ segs, const uint32_t* __restrict__ seg_of_upper,
                            const FitRec* __restrict__ fit, const uint32_t* __restrict__ slot_of, const float* __restrict__ frames, const uint32_t* __restrict__ ext,
                            TreeRec* __restrict__ recs, int write_links) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt->n_upper) return;
    const uint32_t rec = uppers[t], sl = slot_of[rec];
    float fr[12]; uint32_t ex[6];
#pragma unroll
    for (int k = 0; k < 12; ++k) fr[k] = frames[12ull * sl + k];
#pragma unroll
    for (int k = 0; k < 6; ++k) ex[k] = ext[6ull * sl + k];
    float4 q0, q1, q2; float surface;
    box_from_frame(fr, ex, q0, q1, q2, surface);
    store_rec(recs, rec, fit[rec], segs[seg_of_upper[t]], q0, q1, q2, surface, write_links != 0);
}

// uppers were collected without their segment: look it up once (the list is short)
__global__ void k_fit_upper_segs(const FitCounters* __restrict__ cnt, const uint32_t* __restrict__ uppers, const FitSeg* __restrict__ segs, uint32_t n_seg, uint32_t* __restrict__ seg_of_upper) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt->n_upper) return;
    const uint32_t rec = uppers[t];
    uint32_t lo = 0, hi = n_seg;                                          // segments are sorted by rec_base (arena order)
    while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (segs[mid].rec_base <= rec) lo = mid; else hi = mid; }
    seg_of_upper[t] = lo;
}

// ---- FitRec of a tree that was not built here (imported, reference mode): from the records' links ---------------------------------------
__global__ void k_plan_links(uint32_t n_rec, uint32_t rec_base, uint32_t tri_base, const TreeRec* __restrict__ recs, FitRec* fit, uint32_t* ticket) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    ticket[r] = 0u;
    FitRec& f = fit[rec_base + r];
    if (r == 0u) f.parent = 0xffffffffu;
    f.desc = 0xffffffffu; f.n_sub = 1u;
    if (r == 1u) { f.first = f.last = f.split = f.child = 0u; f.parent = 0xffffffffu; f.kind = 2u; return; }
    const float4 q3 = recs[rec_base + r].q3;
    if (__float_as_uint(q3.w) == 0u) {
        const uint32_t c = rec_base + __float_as_uint(q3.y);
        f.child = c; f.kind = 0u;
        fit[c].parent = rec_base + r; fit[c + 1u].parent = rec_base + r;
    } else {
        const uint32_t cntt = __float_as_uint(q3.z);
        f.first = tri_base + __float_as_uint(q3.y); f.last = cntt ? f.first + cntt - 1u : f.first; f.split = f.last; f.child = 0u; f.kind = 1u;
    }
}
__global__ void k_plan_ranges(uint32_t n_rec, uint32_t rec_base, FitRec* fit, uint32_t* ticket) {
    const uint32_t r0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (r0 >= n_rec || r0 == 1u) return;
    uint32_t cur = rec_base + r0;
    if (fit[cur].kind != 1u) return;                                     // leaves start; the second child to arrive finishes the parent
    for (;;) {
        const uint32_t par = fit[cur].parent;
        if (par == 0xffffffffu) return;
        __threadfence();
        if (atomicAdd(&ticket[par - rec_base], 1u) == 0u) return;
        const uint32_t child = fit[par].child;
        const FitRec l = fit[child], r = fit[child + 1u];
        volatile FitRec* p = fit + par;
        p->first = min(l.first, r.first); p->last = max(l.last, r.last); p->split = l.last; p->n_sub = l.n_sub + r.n_sub + 1u;
        cur = par;
    }
}

// ---- host -------------------------------------------------------------------------------------------------------------------------------
static inline unsigned nb(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

int imr_fit_reserve(imrcd_ctx* ctx, uint64_t n_rec_total) {
    IMR_CUDA(ctx, ctx->d_fit.reserve(sizeof(FitRec) * n_rec_total, std::min<uint64_t>(ctx->d_fit.cap, sizeof(FitRec) * n_rec_total), ctx->stream));
    return IMRCD_OK;
}

int imr_fit_plan_from_records(imrcd_ctx* ctx, const MeshDev& md) {
    cudaStream_t s = ctx->stream;
    int rc = imr_fit_reserve(ctx, ctx->n_rec_total); if (rc) return rc;
    IMR_CUDA(ctx, ctx->d_fit_ticket.reserve(4ull * md.n_rec, 0, s));
    k_plan_links<<<nb(md.n_rec, 256), 256, 0, s>>>(md.n_rec, md.rec_base, md.tri_base, ctx->d_recs.as<TreeRec>(), ctx->d_fit.as<FitRec>(), ctx->d_fit_ticket.as<uint32_t>());
    k_plan_ranges<<<nb(md.n_rec, 256), 256, 0, s>>>(md.n_rec, md.rec_base, ctx->d_fit.as<FitRec>(), ctx->d_fit_ticket.as<uint32_t>());
    IMR_CUDA(ctx, cudaGetLastError());
    return IMRCD_OK;
}

// origin of the raw moments of a segment: the centre of the mesh's bounds (build) or of its old root box (refit), taken on the device
__global__ void k_fit_origin_bounds(FitSeg* seg, const uint32_t* __restrict__ bounds) {
    for (int a = 0; a < 3; ++a) seg->origin[a] = 0.5 * ((double)f32_unord(bounds[a]) + (double)f32_unord(bounds[3 + a]));
}
__global__ void k_fit_origin_roots(uint32_t n_seg, FitSeg* segs, const TreeRec* __restrict__ recs) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    const float4 q = recs[segs[s].rec_base].q0;
    segs[s].origin[0] = (double)q.x; segs[s].origin[1] = (double)q.y; segs[s].origin[2] = (double)q.z;
}

struct FitLayout { FitCounters* cnt; FitSeg* segs; uint32_t* prefix; double* mom; uint2* troots; float* frames; uint32_t* ext; uint32_t* uppers; uint32_t* seg_of_upper;
                   uint32_t* ticket; uint32_t* chain; uint32_t* chain_len; uint32_t* cursor; };
static FitLayout fit_layout(imrcd_ctx* ctx) {
    FitLayout L;
    char* sb = ctx->d_fit_segs.as<char>();
    L.cnt = reinterpret_cast<FitCounters*>(sb);
    L.segs = reinterpret_cast<FitSeg*>(sb + 64);
    L.prefix = reinterpret_cast<uint32_t*>(L.segs + ctx->fit_ns);
    char* b = ctx->d_fit_scratch.as<char>();
    L.mom = reinterpret_cast<double*>(b); b += 80 * ctx->fit_max_slots;
    L.troots = reinterpret_cast<uint2*>(b); b += 8 * ctx->fit_max_troot;
    L.frames = reinterpret_cast<float*>(b); b += 48 * ctx->fit_max_slots;
    L.ext = reinterpret_cast<uint32_t*>(b); b += 24 * ctx->fit_max_slots;
    L.uppers = reinterpret_cast<uint32_t*>(b); b += 4 * ctx->fit_max_slots;
    L.seg_of_upper = reinterpret_cast<uint32_t*>(b); b += 4 * ctx->fit_max_slots;
    L.ticket = reinterpret_cast<uint32_t*>(b); b += 4 * ctx->fit_max_slots;
    L.chain_len = reinterpret_cast<uint32_t*>(b); b += 4 * ctx->fit_max_troot;
    L.chain = reinterpret_cast<uint32_t*>(b); b += 4ull * FIT_CHAIN * ctx->fit_max_troot;
    L.cursor = reinterpret_cast<uint32_t*>(b);                      // [0] k_fit_treelets, [1] k_fit_upper: next treelet to hand out (inside the 256 spare bytes)
    return L;
}
FitLists imr_fit_lists(imrcd_ctx* ctx) {
    const FitLayout L = fit_layout(ctx);
    FitLists r; r.cnt = L.cnt; r.troots = L.troots; r.uppers = L.uppers; r.slot_of = ctx->d_fit_slot.as<uint32_t>();
    return r;
}
int imr_fit_reset_counters(imrcd_ctx* ctx) {
    IMR_CUDA(ctx, cudaMemsetAsync(fit_layout(ctx).cnt, 0, sizeof(FitCounters), ctx->stream));
    return IMRCD_OK;
}

// First half of a fit: buffers and the segment table (segments in arena order; their origins are filled in on the device).  Waits for
// the stream once (the pinned staging block may still feed the previous call's copy): call it before a timed region.  Returns with
// ctx->fit_lists_valid set when the call is for the same trees as the last one and their topology has not changed: the treelet / upper
// lists, slots and chain rows of that call are still in place and are used again (a character is re-posed every frame).
int imr_fit_prepare(imrcd_ctx* ctx, const std::vector<FitSeg>& segs, uint64_t topology_key) {
    cudaStream_t s = ctx->stream;
    const uint32_t ns = (uint32_t)segs.size();
    if (topology_key != 0 && topology_key == ctx->fit_key && ns == ctx->fit_ns) { ctx->fit_lists_valid = true; return IMRCD_OK; }
    ctx->fit_lists_valid = false; ctx->fit_key = topology_key;
    std::vector<uint32_t> prefix(ns);
    uint64_t tot_rec = 0, tot_tri = 0;
    for (uint32_t k = 0; k < ns; ++k) { prefix[k] = (uint32_t)tot_rec; tot_rec += segs[k].n_rec; tot_tri += segs[k].n_tri; }
    if (tot_rec >= (1ull << 32)) { ctx->err = "fit: too many records in one call"; return IMRCD_E_CAPACITY; }
    // a treelet root's parent holds more than FIT_T triangles or FIT_R records, and the treelets are disjoint; the nodes above the treelets
    // are fewer than the treelets
    ctx->fit_ns = ns; ctx->fit_tot_rec = tot_rec;
    ctx->fit_max_troot = 2 * (tot_tri / (FIT_T / 2 + 1)) + 2 * (tot_rec / (FIT_R / 2 + 1)) + ns + 16; ctx->fit_max_slots = 2 * ctx->fit_max_troot + 16;
    IMR_CUDA(ctx, ctx->d_fit_segs.reserve(sizeof(FitSeg) * ns + 4ull * ns + 64, 0, s));
    IMR_CUDA(ctx, ctx->p_fit_segs.reserve(sizeof(FitSeg) * ns + 4ull * ns, 0, s));
    IMR_CUDA(ctx, ctx->d_fit_slot.reserve(4ull * (ctx->d_recs.cap / sizeof(TreeRec)) + 64, 0, s));
    IMR_CUDA(ctx, ctx->d_fit_scratch.reserve((8ull + 4 + 4ull * FIT_CHAIN) * ctx->fit_max_troot + (80 + 48 + 24 + 4 + 4 + 4) * ctx->fit_max_slots + 256, 0, s));
    IMR_CUDA(ctx, cudaStreamSynchronize(s));
    memcpy(ctx->p_fit_segs.p, segs.data(), sizeof(FitSeg) * ns);
    memcpy(ctx->p_fit_segs.as<char>() + sizeof(FitSeg) * ns, prefix.data(), 4ull * ns);
    const FitLayout L = fit_layout(ctx);
    IMR_CUDA(ctx, cudaMemcpyAsync(L.segs, ctx->p_fit_segs.p, sizeof(FitSeg) * ns + 4ull * ns, cudaMemcpyHostToDevice, s));
    IMR_CUDA(ctx, cudaMemsetAsync(L.cnt, 0, sizeof(FitCounters), s));
    if (!ctx->fit_attr_set) {
        IMR_CUDA(ctx, cudaFuncSetAttribute(k_fit_treelets, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FitSmem)));
        int per_sm = 0;
        IMR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fit_treelets, FIT_T, sizeof(FitSmem)));
        ctx->fit_blocks = ctx->sm_count * std::max(per_sm, 1);
        ctx->fit_attr_set = true;
    }
    return IMRCD_OK;
}

// Second half: every kernel of the fit, enqueued on the context's stream; nothing waits.  `bounds` (build: the mesh's centroid bounds, still on
// the device) or the old root boxes (refit) give the origins.  `classified`: the builder has put the records on the treelet / upper lists
// itself (k_assign); otherwise k_fit_collect does, unless the lists of the last call still stand (imr_fit_prepare).
int imr_fit_launch(imrcd_ctx* ctx, bool write_links, const uint32_t* bounds, bool classified) {
    cudaStream_t s = ctx->stream;
    const FitLayout L = fit_layout(ctx);
    const uint32_t ns = ctx->fit_ns;
    const uint64_t max_troot = ctx->fit_max_troot, max_slots = ctx->fit_max_slots;
    IMR_CUDA(ctx, cudaMemsetAsync(L.ticket, 0, 4 * max_slots, s));
    IMR_CUDA(ctx, cudaMemsetAsync(L.cursor, 0, 8, s));
    if (bounds) k_fit_origin_bounds<<<1, 1, 0, s>>>(L.segs, bounds);
    else k_fit_origin_roots<<<nb(ns, 128), 128, 0, s>>>(ns, L.segs, ctx->d_recs.as<TreeRec>());
    const FitRec* fit = ctx->d_fit.as<FitRec>();
    uint32_t* slot_of = ctx->d_fit_slot.as<uint32_t>();
    TreeRec* recs = ctx->d_recs.as<TreeRec>();
    const TriRec* tris = ctx->d_tris.as<TriRec>();
    const int wl = write_links ? 1 : 0;
    const unsigned g_troot = (unsigned)std::min<uint64_t>(max_troot, 1u << 30);
    if (!ctx->fit_lists_valid) {
        if (!classified) k_fit_collect<<<nb(ctx->fit_tot_rec, 256), 256, 0, s>>>((uint32_t)ctx->fit_tot_rec, L.segs, L.prefix, ns, fit, slot_of, L.troots, L.uppers, L.cnt, recs, wl);
        k_fit_upper_segs<<<nb(max_slots, 256), 256, 0, s>>>(L.cnt, L.uppers, L.segs, ns, L.seg_of_upper);
        k_fit_chains<<<nb(max_troot, 128), 128, 0, s>>>(L.cnt, L.troots, fit, slot_of, L.chain, L.chain_len);
    }
    k_fit_treelets<<<std::min<unsigned>(g_troot, ctx->fit_blocks), FIT_T, sizeof(FitSmem), s>>>(L.cnt, L.troots, L.segs, fit, tris, slot_of, L.mom, recs, wl, L.cursor);
    k_fit_climb<<<nb(max_troot, 128), 128, 0, s>>>(L.cnt, L.troots, fit, slot_of, L.mom, L.ticket);
    k_fit_axes<<<nb(max_slots, 128), 128, 0, s>>>(L.cnt, L.uppers, L.segs, L.seg_of_upper, slot_of, L.mom, L.frames, L.ext);
    k_fit_upper<<<std::min<unsigned>(g_troot, ctx->sm_count * 12), FIT_T, 0, s>>>(L.cnt, L.troots, fit, tris, slot_of, L.chain, L.chain_len, L.frames, L.ext, recs, L.cursor + 1);
    k_fit_boxes<<<nb(max_slots, 128), 128, 0, s>>>(L.cnt, L.uppers, L.segs, L.seg_of_upper, fit, slot_of, L.frames, L.ext, recs, wl);
    IMR_CUDA(ctx, cudaGetLastError());
    return IMRCD_OK;
}
