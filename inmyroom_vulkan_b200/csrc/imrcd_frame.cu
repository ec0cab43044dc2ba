// imrcd_frame.cu -- the per-frame collision pipeline on the device:
//   broad  : k_entry_prep -> radix sort on U-min -> k_sweep           (SweepAndPrune.cpp:15-88)
//   setup  : k_pair_setup  rel = inverse(first.M) * second.M           (OBBtreesCollision.cpp:15)
//   mid    : imrcd_traverse.cu                                         (OBBtree.cpp:396-477, Paralgram.cpp:17-173)
//   narrow : k_tritri      leaf x leaf triangle tests                  (CreateUncollideRays.cpp:74-115, Triangle.cpp:866-1002)
//   reduce : k_finalize    colliding entity pairs                      (CollisionDetection.cpp:60-67)
#include "imrcd_internal.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cstring>
#include <cstdlib>

#define FULL_MASK 0xffffffffu

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) { return *((const volatile unsigned long long*)p); }
__device__ __forceinline__ long long ld_volatile_s64(const long long* p) { return *((const volatile long long*)p); }
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// ------------------------------------------------------------------------------------------
// broad phase
// ------------------------------------------------------------------------------------------
// One thread per entry: world-space root box (SweepAndPrune.cpp:23), its extents on the three fixed
// sweep axes (Paralgram.cpp:175-190), the sort key, and inverse(M) for the pair stage.
// With `frecs` (the few-flagged-entries path, see k_pairs_few_flagged) the entries WITH shouldCallback are also appended, in no particular
// order, to a dense list of sweep records whose last word carries the caller's entry index.
__global__ void k_entry_prep(uint32_t n, const float* __restrict__ cur, const uint32_t* __restrict__ mesh_id,
                             const MeshDev* __restrict__ meshes, const TreeRec* __restrict__ recs,
                             float* __restrict__ inv_out, float* __restrict__ ext_out,
                             uint32_t* __restrict__ keys, uint32_t* __restrict__ idx,
                             const uint8_t* __restrict__ cb, const uint32_t* __restrict__ gidx, SweepRec* __restrict__ frecs, FrameCtl* ctl) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float m[16];
    const float4* mp = reinterpret_cast<const float4*>(cur + 16 * (size_t)e);
#pragma unroll
    for (int k = 0; k < 4; ++k) { float4 v = mp[k]; m[4 * k] = v.x; m[4 * k + 1] = v.y; m[4 * k + 2] = v.z; m[4 * k + 3] = v.w; }
    const TreeRec& root = recs[meshes[mesh_id[e]].rec_base];
    Box b = box_transform(rel_from_mat(m), unpack_box(root.q0, root.q1, root.q2));
    V3 U, V, W; sweep_axes(U, V, W);
    float mn, mx;
    float* eo = ext_out + 6 * (size_t)e;
    SweepRec sr;
    box_minmax(b, U, mn, mx); eo[0] = mn; eo[1] = mx; sr.umin = mn; sr.umax = mx;
    if (keys) { keys[e] = float_orderable(mn + 0.0f);   // -0 -> +0 so equal floats get equal keys; ties then keep entry order (stable sort)
                idx[e] = e; }
    box_minmax(b, V, mn, mx); eo[2] = mn; eo[3] = mx; sr.vmin = mn; sr.vmax = mx;
    box_minmax(b, W, mn, mx); eo[4] = mn; eo[5] = mx; sr.wmin = mn; sr.wmax = mx;
    if (frecs && cb[e]) { sr.idx = e; sr.cb = gidx ? gidx[e] : e; frecs[atomicAdd(&ctl->n_flagged, 1ull)] = sr; }
    float inv[16];
    mat4_inverse(m, inv);
    float4* ip = reinterpret_cast<float4*>(inv_out + 16 * (size_t)e);
#pragma unroll
    for (int k = 0; k < 4; ++k) ip[k] = make_float4(inv[4 * k], inv[4 * k + 1], inv[4 * k + 2], inv[4 * k + 3]);
}

__global__ void k_gather_sorted(uint32_t n, const uint32_t* __restrict__ sorted_idx, const float* __restrict__ ext,
                                const uint8_t* __restrict__ cb, SweepRec* __restrict__ out, uint32_t* __restrict__ cb_flag) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) cb_flag[n] = 0u;          // the scans run over n + 1 elements so that element n of the output is the total
    if (p >= n) return;
    uint32_t e = sorted_idx[p];
    const float* eo = ext + 6 * (size_t)e;
    SweepRec r;
    r.umin = eo[0]; r.umax = eo[1]; r.vmin = eo[2]; r.vmax = eo[3]; r.wmin = eo[4]; r.wmax = eo[5];
    r.idx = e; r.cb = cb[e];
    out[p] = r;
    cb_flag[p] = r.cb ? 1u : 0u;
}

// The sweep (SweepAndPrune.cpp:50-85) reports (a,e), a before e in U-min order, iff e.umin <= a.umax on U (:58), the same
// on V and W, and a.shouldCallback || e.shouldCallback (:60).  So an entry WITH the flag must look at every later entry
// of its U window, an entry WITHOUT it only at the later entries that have it.  Two sorted lists make both windows
// contiguous: S_all (everything) and S_C (the flagged entries, same order; cpos[p] = flagged entries before position p).
__global__ void k_compact_flagged(uint32_t n, const SweepRec* __restrict__ sorted, const uint32_t* __restrict__ cpos, SweepRec* __restrict__ sorted_c) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const SweepRec r = sorted[p];
    if (r.cb) sorted_c[cpos[p]] = r;
}

#define SWEEP_CHUNK 512u     // candidates per work item of k_sweep

// Window of every entry by binary search (the lists are sorted by umin): candidates [start, start + len) of its list.
__global__ void k_window(uint32_t n, const SweepRec* __restrict__ sorted, const SweepRec* __restrict__ sorted_c, const uint32_t* __restrict__ cpos,
                         uint32_t* __restrict__ wlen, uint32_t* __restrict__ chunks) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) chunks[n] = 0u;
    if (p >= n) return;
    const SweepRec a = sorted[p];
    const SweepRec* list = a.cb ? sorted : sorted_c;
    const uint32_t start = a.cb ? p + 1u : cpos[p];
    uint32_t lo = start, hi = a.cb ? n : cpos[n];
    while (lo < hi) {                                   // first index whose umin > a.umax  (expiry is strict, :58)
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (!(a.umax < list[mid].umin)) lo = mid + 1u; else hi = mid;
    }
    const uint32_t len = lo - start;
    wlen[p] = len;
    chunks[p] = (len + SWEEP_CHUNK - 1u) / SWEEP_CHUNK;
}

// Largest p in [0, n) with arr[p] <= c, found by the whole warp 32 probes at a time (arr is non-decreasing, arr[0] = 0).
__device__ __forceinline__ uint32_t warp_find_owner(const uint32_t* __restrict__ arr, uint32_t n, uint32_t c, uint32_t lane) {
    uint32_t lo = 0, hi = n;                            // arr[lo] <= c ; hi == n or arr[hi] > c
    while (hi - lo > 1u) {
        const uint32_t span = hi - lo - 1u;             // candidates lo+1 .. hi-1
        uint32_t idx;
        if (span <= 32u) idx = lo + 1u + lane;
        else idx = lo + 1u + (uint32_t)(((unsigned long long)span * lane) >> 5);
        const bool ok = (idx < hi) && (__ldg(arr + idx) <= c);
        const uint32_t m = __ballot_sync(FULL_MASK, ok);        // monotone: a prefix of ones
        const uint32_t cnt = (uint32_t)__popc(m);
        const uint32_t new_lo = cnt ? __shfl_sync(FULL_MASK, idx, cnt - 1u) : lo;
        const uint32_t nxt = __shfl_sync(FULL_MASK, idx, cnt < 32u ? cnt : 31u);
        const uint32_t new_hi = (cnt < 32u && nxt < hi) ? nxt : hi;
        if (span <= 32u) { lo = new_lo; break; }
        lo = new_lo; hi = new_hi;
    }
    return lo;
}

// Load-balanced sweep: one warp per chunk of SWEEP_CHUNK candidates, so that the few entries with very long windows
// (a floor spanning the whole scene) are spread over the machine.  Orientation = U order (:63).
// Multi-GPU (n_ranks > 1): the lists hold this rank's share of the frame (every flagged entry + the unflagged entries it owns, see
// imrcd_frame_add_entries), so a pair with an unflagged entity exists on exactly one rank and is always kept; a pair of two flagged
// entities is seen by every rank and kept by rank (gidx_a + gidx_e) % n_ranks.  When NO entry of the frame is unflagged the lists are the
// same on all ranks and the sweep itself is dealt out instead: rank r looks only at the chunks c with c % n_ranks == r.
__global__ void __launch_bounds__(256)
k_sweep(uint32_t n, const SweepRec* __restrict__ sorted, const SweepRec* __restrict__ sorted_c, const uint32_t* __restrict__ cpos,
        const uint32_t* __restrict__ wlen, const uint32_t* __restrict__ chunk_off, uint2* __restrict__ pairs, unsigned long long cap,
        FrameCtl* ctl, uint32_t rank, uint32_t n_ranks, const uint32_t* __restrict__ gidx, uint32_t deal_chunks) {
    const uint32_t lane = lane_id();
    const uint32_t warps_total = (gridDim.x * blockDim.x) >> 5;
    const uint32_t total = chunk_off[n];
    const uint32_t c_first = deal_chunks ? rank : 0u, c_step = deal_chunks ? n_ranks : 1u;
    const bool by_pair = n_ranks > 1u && !deal_chunks;
    for (uint32_t c = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * c_step + c_first; c < total; c += warps_total * c_step) {
        const uint32_t p = warp_find_owner(chunk_off, n, c, lane);
        const SweepRec a = sorted[p];
        const SweepRec* list = a.cb ? sorted : sorted_c;
        const uint32_t start = a.cb ? p + 1u : cpos[p];
        const uint32_t k0 = (c - chunk_off[p]) * SWEEP_CHUNK;
        const uint32_t len = wlen[p];
        const uint32_t k1 = (k0 + SWEEP_CHUNK < len) ? k0 + SWEEP_CHUNK : len;
        for (uint32_t kb = k0; kb < k1; kb += 32u) {
            const uint32_t k = kb + lane;
            bool emit = false;
            SweepRec e;
            if (k < k1) {
                e = list[start + k];
                if (a.cb | e.cb) {
                    const bool v_ok = (a.vmin <= e.vmin) ? !(a.vmax < e.vmin) : !(e.vmax < a.vmin);
                    const bool w_ok = (a.wmin <= e.wmin) ? !(a.wmax < e.wmin) : !(e.wmax < a.wmin);
                    emit = v_ok && w_ok;
                    if (emit && by_pair && a.cb && e.cb) emit = (gidx[a.idx] + gidx[e.idx]) % n_ranks == rank;
                }
            }
            const uint32_t m = __ballot_sync(FULL_MASK, emit);
            if (m) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&ctl->n_pairs, (unsigned long long)__popc(m));
                base = __shfl_sync(FULL_MASK, base, 0);
                if (emit) {
                    const unsigned long long slot = base + __popc(m & ((1u << lane) - 1));
                    if (slot < cap) pairs[slot] = make_uint2(a.idx, e.idx);
                    else atomicOr(&ctl->overflow, (unsigned)OVF_PAIRS);
                }
            }
        }
    }
}

// Frames with few flagged entries (a static scene of some hundred parts against any number of bodies that only collide with it, BASELINE
// config 3; or a game-sized frame): no sort at all.  A pair needs shouldCallback on one side (SweepAndPrune.cpp:60), so every pair has a
// flagged member: each entry walks the dense list of the flagged ones (k_entry_prep) from shared memory and applies the sweep's own
// predicate - closed overlap on U, V and W, orientation by U-min with the lower entry index first on ties (:42-63) - to every candidate.
// An unflagged entry reports all its pairs, a flagged one those with flagged entries of lower index (each pair once); with n_ranks > 1 a
// pair of two flagged entries belongs to rank (gidx_a + gidx_b) % n_ranks, every other pair to the one rank that holds the unflagged entry.
#define FEW_FLAGGED_MAX 2048u
__global__ void __launch_bounds__(256)
k_pairs_few_flagged(uint32_t n, const float* __restrict__ ext, const uint8_t* __restrict__ cb, const uint32_t* __restrict__ gidx,
                    const SweepRec* __restrict__ frecs, uint2* __restrict__ pairs, unsigned long long cap, FrameCtl* ctl, uint32_t rank, uint32_t n_ranks) {
    __shared__ SweepRec tile[256];
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x, lane = lane_id();
    const bool valid = e < n;
    SweepRec me; me.umin = me.umax = me.vmin = me.vmax = me.wmin = me.wmax = 0.f; me.idx = e; me.cb = 0u;
    uint32_t ge = e;
    if (valid) {
        const float* eo = ext + 6 * (size_t)e;
        me.umin = eo[0]; me.umax = eo[1]; me.vmin = eo[2]; me.vmax = eo[3]; me.wmin = eo[4]; me.wmax = eo[5]; me.cb = cb[e];
        ge = gidx ? gidx[e] : e;
    }
    const uint32_t nf = (uint32_t)ctl->n_flagged;
    for (uint32_t t0 = 0; t0 < nf; t0 += 256u) {
        __syncthreads();
        if (t0 + threadIdx.x < nf) tile[threadIdx.x] = frecs[t0 + threadIdx.x];
        __syncthreads();
        const uint32_t tn = nf - t0 < 256u ? nf - t0 : 256u;
        for (uint32_t k = 0; k < tn; ++k) {
            const SweepRec f = tile[k];                       // every lane reads the same record: a broadcast
            const uint32_t gf = f.cb;                         // the flagged entry's caller index (k_entry_prep)
            bool emit = valid && (me.cb ? gf < ge : true);
            bool me_first = false;
            if (emit) {
                me_first = (me.umin < f.umin) || (me.umin == f.umin && ge < gf);
                const SweepRec& a = me_first ? me : f; const SweepRec& b = me_first ? f : me;      // a before b on U
                const bool u_ok = !(a.umax < b.umin);                                               // expiry is strict (:58)
                const bool v_ok = (a.vmin <= b.vmin) ? !(a.vmax < b.vmin) : !(b.vmax < a.vmin);
                const bool w_ok = (a.wmin <= b.wmin) ? !(a.wmax < b.wmin) : !(b.wmax < a.wmin);
                emit = u_ok && v_ok && w_ok;
                if (emit && me.cb && n_ranks > 1u) emit = (ge + gf) % n_ranks == rank;
            }
            const uint32_t m = __ballot_sync(FULL_MASK, emit);
            if (m) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&ctl->n_pairs, (unsigned long long)__popc(m));
                base = __shfl_sync(FULL_MASK, base, 0);
                if (emit) {
                    const unsigned long long slot = base + __popc(m & ((1u << lane) - 1));
                    if (slot < cap) pairs[slot] = me_first ? make_uint2(e, f.idx) : make_uint2(f.idx, e);
                    else atomicOr(&ctl->overflow, (unsigned)OVF_PAIRS);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// pair setup
// ------------------------------------------------------------------------------------------
// Plane.cpp:5-21 through TrianglePosition::GetTrianglePlane: normal = normalize(cross(p1-p0, p2-p0)), d = -dot(p0, normal),
// then the Plane ctor divides both by length(normal).
struct PlaneN { V3 n; float d; };
IMR_D PlaneN plane_from_tri(V3 p0, V3 p1, V3 p2) {
    const V3 nrm = normalize3(cross3(sub3(p1, p0), sub3(p2, p0)));
    const float d = -dot3(p0, nrm);
    const float len = length3(nrm);
    PlaneN pl; pl.n = mk3(nrm.x / len, nrm.y / len, nrm.z / len); pl.d = d / len;
    return pl;
}
IMR_D bool plane_outside(const PlaneN& pl, V3 p) { return dot3(p, pl.n) + pl.d > 0.f; }     // Plane.cpp:23-29



// One thread per pair: rel = glm::inverse(first.M) * second.M (OBBtreesCollision.cpp:15), the bases of the
// two meshes, the pair's accumulators, and the root work item (root_obb vs root_obb, OBBtree.cpp:396-411).
__global__ void k_pair_setup(const FrameCtl* ctl, unsigned long long cap_pairs, const uint2* __restrict__ pairs,
                             const float* __restrict__ cur, const float* __restrict__ prev /* or NULL: nothing moved */, const float* __restrict__ inv,
                             const uint32_t* __restrict__ mesh_id, const MeshDev* __restrict__ meshes, PairRec* __restrict__ pairrec,
                             PairAcc* __restrict__ acc, WorkItem* __restrict__ queue, unsigned long long cap_queue) {
    unsigned long long n = ctl->n_pairs < cap_pairs ? ctl->n_pairs : cap_pairs;
    for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < n; p += (unsigned long long)gridDim.x * blockDim.x) {
        uint2 pr = pairs[p];
        float a[16], b[16], r[16];
        const float4* ap = reinterpret_cast<const float4*>(inv + 16 * (size_t)pr.x);
        const float4* bp = reinterpret_cast<const float4*>(cur + 16 * (size_t)pr.y);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4 v = ap[k]; a[4 * k] = v.x; a[4 * k + 1] = v.y; a[4 * k + 2] = v.z; a[4 * k + 3] = v.w;
            float4 w = bp[k]; b[4 * k] = w.x; b[4 * k + 1] = w.y; b[4 * k + 2] = w.z; b[4 * k + 3] = w.w;
        }
        mat4_mul(a, b, r);
        MeshDev ma = meshes[mesh_id[pr.x]], mb = meshes[mesh_id[pr.y]];
        PairRec o;
        o.r0 = make_float4(r[0], r[4], r[8], r[12]);
        o.r1 = make_float4(r[1], r[5], r[9], r[13]);
        o.r2 = make_float4(r[2], r[6], r[10], r[14]);
        o.recA = ma.rec_base; o.recB = mb.rec_base; o.triA = ma.tri_base; o.triB = mb.tri_base;
        pairrec[p] = o;
        PairAcc z; memset(&z, 0, sizeof(z));
        if (prev) {                                     // either matrix changed since the last frame (glm mat4 !=, CollisionDetection.cpp:80-81)
            bool moved = false;
            const float* cx = cur + 16 * (size_t)pr.x; const float* px = prev + 16 * (size_t)pr.x;
            const float* py = prev + 16 * (size_t)pr.y;
#pragma unroll
            for (int k = 0; k < 16; ++k) moved |= (cx[k] != px[k]) | (b[k] != py[k]);
            if (moved) z.flags = PAIR_MOVED;
        }
        acc[p] = z;
        if (p < cap_queue) queue[p] = make_uint4((uint32_t)p, ma.rec_base, mb.rec_base, 1u);      // arena indices of the two roots
    }
}


// ------------------------------------------------------------------------------------------
// narrow phase: the loops of CreateUncollideRays.cpp:74-115 over the leaf combos.
//
// A warp takes a tile of 32 leaf combos (<= 4 x 4 triangle pairs each) and runs three dense passes over it:
//   pass 0  every triangle of the second entity's leaves is moved to first's model space ONCE per combo
//           (seconds_triangle = rel * tri, CreateUncollideRays.cpp:84) together with its plane (Triangle.cpp:905-910);
//           the result lives in shared memory, structure-of-arrays, for the other two passes;
//   pass 1  all triangle pairs of the tile, one per lane with no idle (i,j) slots: the two plane-side rejection tests of
//           tri_tri_intersect_with_isectline (Triangle.cpp:884-926), 12 dot products against hoisted planes.  Survivors
//           (a few percent) are compacted into a shared-memory list;
//   pass 2  the survivors, again one per lane: interval / segment computation (Triangle.cpp:928-1001), hit records
//           appended with one atomic per warp, per-pair accumulators with one atomic per distinct pair.
// The split keeps lanes busy: the monolithic one-lane-per-(i,j)-slot kernel ran at 10 of 32 active lanes (ncu).
// ------------------------------------------------------------------------------------------
#define NT_WARPS 8
#define NT_TILE 32u                       // combos per warp tile
#define NT_SLOTS (NT_TILE * 4u)           // transformed second-entity triangles per tile
#define NT_TESTS (NT_TILE * 16u)          // triangle pairs per tile (upper bound)

struct NarrowWarp {
    float ux[3][NT_SLOTS], uy[3][NT_SLOTS], uz[3][NT_SLOTS];   // [vertex][slot], slot = 4 * combo_in_tile + j
    float nx[NT_SLOTS], ny[NT_SLOTS], nz[NT_SLOTS], nd[NT_SLOTS];   // plane of the transformed triangle
    uint32_t absB[NT_SLOTS];              // arena index of the second entity's triangle
    uint4 cmb[NT_TILE];                   // (pair, absolute index of first's leaf triangle 0, cntA | cntB << 16, unused)
    uint16_t test[NT_TESTS];              // dense list of the tile's pairs: combo | i << 5 | j << 7; pass 1 compacts the
                                          // survivors of the rejection tests into its front (in place: writes trail reads)
};

__global__ void __launch_bounds__(NT_WARPS * 32, 3)
k_tritri(FrameCtl* ctl, const Combo* __restrict__ combos, unsigned long long cap_combos, const PairRec* __restrict__ pairrec,
         const TriRec* __restrict__ tris, imrcd_tri_hit* __restrict__ hits, unsigned long long cap_hits, PairAcc* acc, HitAux* __restrict__ aux) {
    extern __shared__ __align__(16) unsigned char nt_smem[];
    NarrowWarp& sm = reinterpret_cast<NarrowWarp*>(nt_smem)[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    const uint32_t lt_mask = (1u << lane) - 1u;
    const unsigned long long n = ctl->n_combos < cap_combos ? ctl->n_combos : cap_combos;
    const unsigned long long n_tiles = (n + NT_TILE - 1) / NT_TILE;
    unsigned long long my_cop = 0;

    // tiles are handed out through a counter: their costs differ (4 to 512 triangle pairs, a few percent of them going the whole way), and a
    // fixed stride left the last warps of the grid working alone (0.453 -> 0.423 ms on C3)
    for (;;) {
        unsigned long long tile = 0;
        if (lane == 0) tile = atomicAdd(&ctl->tile_cursor, 1ull);
        tile = __shfl_sync(FULL_MASK, tile, 0);
        if (tile >= n_tiles) break;
        // ---- tile setup: lane = combo ----
        const unsigned long long ci = tile * NT_TILE + lane;
        uint32_t cntA = 0, cntB = 0, pair = 0, triB0 = 0;
        if (ci < n) {
            const Combo cb = __ldg(combos + ci);
            pair = cb.x; cntA = cb.w & 0xffffu; cntB = cb.w >> 16;
            const uint4 bases = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float4*>(pairrec + pair) + 3));
            sm.cmb[lane] = make_uint4(pair, bases.z + cb.y, cb.w, 0u);
            triB0 = bases.w + cb.z;
        }
        const uint32_t nt = cntA * cntB;
        uint32_t off = nt;                                   // exclusive prefix sum of nt over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL_MASK, off, o); if (lane >= (uint32_t)o) off += v; }
        const uint32_t total = __shfl_sync(FULL_MASK, off, 31);
        off -= nt;
        for (uint32_t i = 0; i < cntA; ++i)
            for (uint32_t j = 0; j < cntB; ++j) sm.test[off + i * cntB + j] = (uint16_t)(lane | (i << 5) | (j << 7));

        // ---- pass 0: second entity's triangles -> first's model space, with their planes ----
#pragma unroll
        for (uint32_t q = 0; q < 4; ++q) {
            const uint32_t slot = q * 32u + lane, c = slot >> 2, j = slot & 3u;
            const uint32_t c_cntB = __shfl_sync(FULL_MASK, cntB, c), c_pair = __shfl_sync(FULL_MASK, pair, c), c_triB0 = __shfl_sync(FULL_MASK, triB0, c);
            if (j < c_cntB) {
                const float4* pp = reinterpret_cast<const float4*>(pairrec + c_pair);
                Rel rel; rel.r0 = __ldg(pp); rel.r1 = __ldg(pp + 1); rel.r2 = __ldg(pp + 2);
                const float4* tb = reinterpret_cast<const float4*>(tris + c_triB0 + j);
                const float4 b0 = __ldg(tb), b1 = __ldg(tb + 1), b2 = __ldg(tb + 2);
                const V3 U0 = rel_mul(rel, mk3(b0.x, b0.y, b0.z), 1.f);       // Triangle.cpp:69-78
                const V3 U1 = rel_mul(rel, mk3(b1.x, b1.y, b1.z), 1.f);
                const V3 U2 = rel_mul(rel, mk3(b2.x, b2.y, b2.z), 1.f);
                V3 N2; float d2;
                tt_plane(U0, U1, U2, N2, d2);
                sm.ux[0][slot] = U0.x; sm.uy[0][slot] = U0.y; sm.uz[0][slot] = U0.z;
                sm.ux[1][slot] = U1.x; sm.uy[1][slot] = U1.y; sm.uz[1][slot] = U1.z;
                sm.ux[2][slot] = U2.x; sm.uy[2][slot] = U2.y; sm.uz[2][slot] = U2.z;
                sm.nx[slot] = N2.x; sm.ny[slot] = N2.y; sm.nz[slot] = N2.z; sm.nd[slot] = d2;
                sm.absB[slot] = c_triB0 + j;
            }
        }
        __syncwarp();

        // ---- pass 1: plane-side rejection for every pair of the tile ----
        uint32_t n_surv = 0;
        for (uint32_t t0 = 0; t0 < total; t0 += 32u) {
            const uint32_t t = t0 + lane;
            bool keep = false;
            uint32_t code = 0;
            if (t < total) {
                code = sm.test[t];
                const uint32_t c = code & 31u, i = (code >> 5) & 3u, slot = 4u * c + (code >> 7);
                const float4* ta = reinterpret_cast<const float4*>(tris + sm.cmb[c].y + i);
                const float4 a3 = __ldg(ta + 3);
                const V3 N1 = mk3(a3.x, a3.y, a3.z);
                const V3 U0 = mk3(sm.ux[0][slot], sm.uy[0][slot], sm.uz[0][slot]);
                const V3 U1 = mk3(sm.ux[1][slot], sm.uy[1][slot], sm.uz[1][slot]);
                const V3 U2 = mk3(sm.ux[2][slot], sm.uy[2][slot], sm.uz[2][slot]);
                float s0, s1, s2, s01, s02;
                if (!tt_side(N1, a3.w, U0, U1, U2, s0, s1, s2, s01, s02)) {
                    const float4 a0 = __ldg(ta), a1 = __ldg(ta + 1), a2 = __ldg(ta + 2);
                    const V3 N2 = mk3(sm.nx[slot], sm.ny[slot], sm.nz[slot]);
                    keep = !tt_side(N2, sm.nd[slot], mk3(a0.x, a0.y, a0.z), mk3(a1.x, a1.y, a1.z), mk3(a2.x, a2.y, a2.z), s0, s1, s2, s01, s02);
                }
            }
            const uint32_t km = __ballot_sync(FULL_MASK, keep);
            __syncwarp();                                                         // every lane has read its test[t]: the writes below trail the reads
            if (keep) sm.test[n_surv + __popc(km & lt_mask)] = (uint16_t)code;
            n_surv += (uint32_t)__popc(km);
        }
        __syncwarp();

        // ---- pass 2: segment computation for the survivors, hit records, contact candidates ----
        for (uint32_t t0 = 0; t0 < n_surv; t0 += 32u) {
            const uint32_t t = t0 + lane;
            bool hit = false;
            V3 src = mk3(0, 0, 0), tgt = mk3(0, 0, 0);
            uint32_t hpair = 0xffffffffu, triA = 0, triB = 0, origA = 0, bits_a = 7u, bits_b = 7u, code = 0;
            if (t < n_surv) {
                code = sm.test[t];
                const uint32_t c = code & 31u, i = (code >> 5) & 3u, slot = 4u * c + (code >> 7);
                const uint4 cm = sm.cmb[c];
                const float4* ta = reinterpret_cast<const float4*>(tris + cm.y + i);
                const float4 a0 = __ldg(ta), a1 = __ldg(ta + 1), a2 = __ldg(ta + 2), a3 = __ldg(ta + 3);
                const V3 V0 = mk3(a0.x, a0.y, a0.z), V1 = mk3(a1.x, a1.y, a1.z), V2 = mk3(a2.x, a2.y, a2.z);
                const V3 U0 = mk3(sm.ux[0][slot], sm.uy[0][slot], sm.uz[0][slot]);
                const V3 U1 = mk3(sm.ux[1][slot], sm.uy[1][slot], sm.uz[1][slot]);
                const V3 U2 = mk3(sm.ux[2][slot], sm.uy[2][slot], sm.uz[2][slot]);
                const V3 N1 = mk3(a3.x, a3.y, a3.z), N2 = mk3(sm.nx[slot], sm.ny[slot], sm.nz[slot]);
                float du0, du1, du2, du0du1, du0du2, dv0, dv1, dv2, dv0dv1, dv0dv2;
                tt_side(N1, a3.w, U0, U1, U2, du0, du1, du2, du0du1, du0du2);          // same inputs, same bits as in pass 1
                tt_side(N2, sm.nd[slot], V0, V1, V2, dv0, dv1, dv2, dv0dv1, dv0dv2);
                const int f = tt_segment(V0, V1, V2, U0, U1, U2, N1, N2, du0, du1, du2, du0du1, du0du2, dv0, dv1, dv2, dv0dv1, dv0dv2, src, tgt);   // :86
                hit = (f == 1);                                                        // doIntersept && !areCoplanar (:88)
                if (f == 3) ++my_cop;
            }
            const uint32_t hm = __ballot_sync(FULL_MASK, hit);
            if (hm == 0u) continue;
            // the slots of this iteration's hits: the atomic goes out first, the per-hit work below runs while it is on its way through L2
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&ctl->n_hits, (unsigned long long)__popc(hm));
            float weight = 0.f;
            bool zero_w = false;
            if (hit) {
                const uint32_t c = code & 31u, i = (code >> 5) & 3u, slot = 4u * c + (code >> 7);
                const uint4 cm = sm.cmb[c];
                const float4* ta = reinterpret_cast<const float4*>(tris + cm.y + i);
                const float4 a0 = __ldg(ta), a1 = __ldg(ta + 1), a2 = __ldg(ta + 2);
                const V3 V0 = mk3(a0.x, a0.y, a0.z), V1 = mk3(a1.x, a1.y, a1.z), V2 = mk3(a2.x, a2.y, a2.z);
                const V3 U0 = mk3(sm.ux[0][slot], sm.uy[0][slot], sm.uz[0][slot]);
                const V3 U1 = mk3(sm.ux[1][slot], sm.uy[1][slot], sm.uz[1][slot]);
                const V3 U2 = mk3(sm.ux[2][slot], sm.uy[2][slot], sm.uz[2][slot]);
                hpair = cm.x; triA = cm.y + i; triB = sm.absB[slot]; origA = __float_as_uint(a0.w);
                // each vertex against the other triangle's plane (:102-112)
                const PlaneN pa = plane_from_tri(V0, V1, V2), pb = plane_from_tri(U0, U1, U2);
                if (plane_outside(pb, V0)) bits_a &= ~1u; if (plane_outside(pb, V1)) bits_a &= ~2u; if (plane_outside(pb, V2)) bits_a &= ~4u;
                if (plane_outside(pa, U0)) bits_b &= ~1u; if (plane_outside(pa, U1)) bits_b &= ~2u; if (plane_outside(pa, U2)) bits_b &= ~4u;
                weight = length3(sub3(src, tgt));                                      // :93
            }
            base = __shfl_sync(FULL_MASK, base, 0);
            if (hit) {
                const unsigned long long slot = base + __popc(hm & lt_mask);
                if (slot < cap_hits) {
                    imrcd_tri_hit h;
                    h.pair = hpair; h.tri_first = origA; h.tri_second = __float_as_uint(__ldg(reinterpret_cast<const float4*>(tris + triB)).w);
                    h.source[0] = src.x; h.source[1] = src.y; h.source[2] = src.z;
                    h.target[0] = tgt.x; h.target[1] = tgt.y; h.target[2] = tgt.z;
                    h.weight = weight;
                    hits[slot] = h;
                    HitAux x; x.triA = triA; x.triB = triB; x.flags = bits_a | (bits_b << 3) | (((code >> 5) & 3u) << 6) | ((code >> 7) << 8);
                    aux[slot] = x;
                } else atomicOr(&ctl->overflow, (unsigned)OVF_HITS);
                zero_w = (weight == 0.0f);
            }
            // per-pair accumulators: one atomic per distinct pair among the hitting lanes.
            // A candidate survives IsNull() iff its accumulated weight != 0 (CreateUncollideRays.cpp:22-25,117-127);
            // weights are >= 0 (or NaN), so that is "some hit of the triangle has weight != 0".
            const uint32_t peers = __match_any_sync(FULL_MASK, hpair);
            if (hit) {
                const uint32_t nz = __ballot_sync(peers, !zero_w) & peers;
                if (lane == (uint32_t)(__ffs(peers) - 1)) {
                    atomicAdd(&acc[hpair].n_hits, (uint32_t)__popc(peers));
                    if (nz) atomicOr(&acc[hpair].flags, 1u);
                }
            }
        }
        __syncwarp();
    }
    for (int o = 16; o > 0; o >>= 1) my_cop += __shfl_down_sync(FULL_MASK, my_cop, o);
    if (lane == 0 && my_cop) atomicAdd(&ctl->n_coplanar, my_cop);
}

// ------------------------------------------------------------------------------------------
// reduce: ray origins per pair and side (find_rays_lambda, CreateUncollideRays.cpp:131-178), then the colliding entity
// pairs with their contact points (CreateUncollideRays.cpp:185-198, CollisionDetection.cpp:60-78)
// ------------------------------------------------------------------------------------------
// The hits of a frame come out in no particular order; the reduction of CreateUncollideRays.cpp:117-178 is per entity pair.
// So: (1) k_hit_lists gives every pair with hits a slice of the grouping array (power-of-two sized, one atomic per warp of pairs) and
// puts it on the list of its size class; (2) k_group_hits drops each hit index into its pair's slice; (3) one block
// per pair reduces the slice (k_pair_contacts_hash below).  Nothing depends on the order in which the hits were produced.
// size classes with their tables in shared memory: <= 256 hits (128 threads), <= 512 (512 threads), <= 1024 (1024 threads), and
// the large pairs (more hits), which go through the grid-wide passes k_large_* with their tables in a global scratch
#define PC_S_MAX 256u
#define PC_M1_MAX 512u
#define PC_M_MAX 1024u
#define PC_CLASSES 4

__global__ void k_hit_lists(FrameCtl* ctl, unsigned long long cap_pairs, PairAcc* acc,
                            uint32_t* __restrict__ lists /* PC_CLASSES x cap_pairs */, uint32_t large_min) {
    const unsigned long long n = ctl->n_pairs < cap_pairs ? ctl->n_pairs : cap_pairs;
    const uint32_t lane = lane_id();
    for (unsigned long long p0 = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) & ~31ull; p0 < n; p0 += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long p = p0 + lane;
        uint32_t h = 0;
        if (p < n) h = acc[p].n_hits;
        // the pair's slice of the grouping array: power-of-two sized, handed out by one atomic per warp (no scan over all pairs)
        uint32_t m = 0;
        if (h) { m = 1u; while (m < h) m <<= 1; }
        uint32_t incl = m;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= (uint32_t)o) incl += v; }
        const uint32_t warp_total = __shfl_sync(FULL_MASK, incl, 31);
        if (warp_total) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&ctl->grouped_used, (unsigned long long)warp_total);
            base = __shfl_sync(FULL_MASK, base, 0);
            if (h) acc[p].off = (uint32_t)(base + incl - m);
        }
        const int cls = h == 0 ? -1 : (h > large_min ? 3 : (h <= PC_S_MAX ? 0 : (h <= PC_M1_MAX ? 1 : 2)));
#pragma unroll
        for (int c = 0; c < PC_CLASSES; ++c) {
            const uint32_t mm = __ballot_sync(FULL_MASK, cls == c);
            if (mm) {
                unsigned long long b = 0;
                if (lane == 0) b = atomicAdd(&ctl->n_class[c * 16], (unsigned long long)__popc(mm));
                b = __shfl_sync(FULL_MASK, b, 0);
                uint32_t* list = lists + (unsigned long long)c * cap_pairs;
                if (cls == c) list[b + __popc(mm & ((1u << lane) - 1u))] = (uint32_t)p;
            }
        }
    }
}

__global__ void k_group_hits(const FrameCtl* ctl, unsigned long long cap_hits, const imrcd_tri_hit* __restrict__ hits, PairAcc* acc,
                             uint32_t* __restrict__ grouped) {
    if (ctl->overflow & (OVF_PAIRS | OVF_QUEUE | OVF_COMBOS | OVF_HITS)) return;      // slices would not fit: the frame is re-run
    const unsigned long long n = ctl->n_hits < cap_hits ? ctl->n_hits : cap_hits;
    const uint32_t lane = lane_id();
    for (unsigned long long h0 = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) & ~31ull; h0 < n; h0 += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long h = h0 + lane;
        const uint32_t pair = h < n ? hits[h].pair : 0xffffffffu;
        const uint32_t peers = __match_any_sync(FULL_MASK, pair);             // neighbouring hits mostly share the pair
        uint32_t base = 0;
        const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
        if (h < n && lane == leader) base = atomicAdd(&acc[pair].cursor, (uint32_t)__popc(peers));
        base = __shfl_sync(FULL_MASK, base, leader);
        if (h < n) grouped[acc[pair].off + base + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)h;
    }
}

struct SideSum { double x, y, z; uint32_t rays; };

// ---- one block per pair, no sort -----------------------------------------------------------------------------------------
// What a side needs of a pair's hits is a keyed reduction: per own triangle, AND of the not-outside bits and the sums of weight and
// weight * midpoint over the hits of every combo whose weight is not 0 (TriangleCandidateRays::Merge, :51-57; IsNull, :22-25).  Weights are
// >= 0, so a combo is dropped iff every one of its hits has weight 0: a hit with weight != 0 always contributes, and a hit with weight 0
// contributes (its bits only) iff the same (own triangle, other LEAF) has another hit with weight != 0 - checked by a scan over the pair's
// staged hits for those few.  The reduction runs in a hash table, one thread per hit: AND and FP64 adds are order-free, so the FP32
// result does not depend on which thread comes first.  Then one thread per occupied slot turns the candidate into rays (:139-166), vertex
// rays going through the `emplaced` set (:139-141), and one thread per ray adds its origin and - for pairs whose entities moved since the
// last frame, the only ones the response stage looks at (CollisionDetection.cpp:80-81) - writes the ray (origin, -normal) to the frame's ray array.
// The three per-pair size classes keep their tables in shared memory; pairs with more than 1024 hits go through the grid-wide passes (k_large_*).
struct PcSlot { double w, cx, cy, cz; };


// Find or claim the slot of `tri`; `claimed` tells the caller to append the slot to the dense list of occupied slots (done
// afterwards in converged code with one ballot, so that the later passes run over occupied slots only).
__device__ __forceinline__ uint32_t pc_slot_of(uint32_t* keys, uint32_t mask, uint32_t tri, bool& claimed) {
    uint32_t slot = ((tri * 2654435761u) >> 7) & mask;
    for (;;) {
        const uint32_t old = atomicCAS(&keys[slot], 0xffffffffu, tri);
        if (old == 0xffffffffu) { claimed = true; return slot; }
        if (old == tri) return slot;
        slot = (slot + 1u) & mask;
    }
}

// vertex set: entry = vid << 32 | smallest (triangle * 4 + corner) that has it; returns the slot when this call claimed it, else ~0
__device__ __forceinline__ uint32_t vset_insert_m(unsigned long long* vset, uint32_t mask, unsigned long long ent) {
    const uint32_t vid = (uint32_t)(ent >> 32);
    uint32_t slot = ((vid * 2654435761u) >> 9) & mask;
    for (;;) {
        const unsigned long long old = atomicCAS(&vset[slot], ~0ull, ent);
        if (old == ~0ull) return slot;
        if ((uint32_t)(old >> 32) == vid) { atomicMin(&vset[slot], ent); return 0xffffffffu; }
        slot = (slot + 1u) & mask;
    }
}

// Set of (own triangle, other LEAF) keys that have a hit with weight != 0: answers "is this combo's weight for this triangle 0?" (:117-127) for
// the hits whose own weight is 0 in O(1).  (A scan over the pair's hits per such hit looked harmless - they are 0.07 % of the hits on C3 - until
// the instance-vs-instance scene C2 turned out to have 3 % of them: one scan of a 700-hit pair is 34 us of one warp with the block waiting.)
__device__ __forceinline__ uint32_t pc_combo_hash(unsigned long long key) { return (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 40); }
__device__ __forceinline__ void pc_combo_insert(unsigned long long* set, uint32_t mask, unsigned long long key) {
    uint32_t slot = pc_combo_hash(key) & mask;
    for (;;) {
        const unsigned long long old = atomicCAS(&set[slot], ~0ull, key);
        if (old == ~0ull || old == key) return;
        slot = (slot + 1u) & mask;
    }
}
template <bool G>
__device__ __forceinline__ bool pc_combo_contains(const unsigned long long* set, uint32_t mask, unsigned long long key) {
    uint32_t slot = pc_combo_hash(key) & mask;
    for (;;) {
        const unsigned long long v = G ? __ldcg(set + slot) : set[slot];
        if (v == key) return true;
        if (v == ~0ull) return false;
        slot = (slot + 1u) & mask;
    }
}

// warp-aggregated append of `v` (for the lanes with `yes`) to list[*count ...]; converged code only
template <class LT>
__device__ __forceinline__ void pc_append(bool yes, uint32_t v, LT* list, uint32_t* count, uint32_t lane) {
    const uint32_t m = __ballot_sync(FULL_MASK, yes);
    if (m == 0u) return;
    const uint32_t leader = (uint32_t)__ffs(m) - 1u;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(FULL_MASK, base, leader);
    if (yes) list[base + __popc(m & ((1u << lane) - 1u))] = (LT)v;
}


template <int T, uint32_t M_MAX>
__global__ void __launch_bounds__(T)
k_pair_contacts_hash(FrameCtl* ctl, const uint32_t* __restrict__ list, int cls, PairAcc* acc, const uint32_t* __restrict__ grouped,
                     const imrcd_tri_hit* __restrict__ hits, const HitAux* __restrict__ aux, const PairRec* __restrict__ pairrec,
                     const TriRec* __restrict__ tris, const uint32_t* __restrict__ tri_vid, const float* __restrict__ tri_nrm,
                     RayRec* __restrict__ rays, unsigned long long cap_rays) {
    typedef uint16_t LT;                               // slot indices (< 4 M_MAX <= 4096)
    extern __shared__ __align__(16) unsigned char pc_smem[];
    __shared__ double s_red[3][T / 32];
    __shared__ uint32_t s_redc[T / 32];
    __shared__ uint32_t s_ncand, s_nvert, s_navg;
    __shared__ __align__(16) uint32_t s_raybase;      // on its own 16 bytes: never part of a vector load of the counters
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    if (ctl->overflow & (OVF_PAIRS | OVF_QUEUE | OVF_COMBOS | OVF_HITS)) return;      // the frame is re-run with larger buffers
    const unsigned long long n_list = ctl->n_class[cls * 16];
    __shared__ unsigned long long s_next;
    // pairs are handed out through a counter (the word behind the class's count): a pair of 16 hits and one of 250 do not cost the same
    for (;;) {
        __syncthreads();                                             // the last pair's use of the shared tables (and of s_next) is over
        if (tid == 0) s_next = atomicAdd(&ctl->n_class[cls * 16 + 1], 1ull);
        __syncthreads();
        const unsigned long long b = s_next;
        if (b >= n_list) break;
        const uint32_t p = list[b];
        const uint32_t n = acc[p].n_hits;
        const bool keep_rays = (acc[p].flags & PAIR_MOVED) != 0u;
        uint32_t m = 16u; while (m < n) m <<= 1;                     // tables sized by the pair: 2m candidate slots (<= n distinct triangles), 4m vertex slots (<= 3n)
        const uint32_t slots = 2u * m, vslots = 4u * m;
        const uint32_t cap = M_MAX;                                  // array stride
        unsigned char* base = pc_smem;
        PcSlot* s_sum = reinterpret_cast<PcSlot*>(base);                                             // 2 cap
        unsigned long long* s_vset = reinterpret_cast<unsigned long long*>(s_sum + 2u * cap);        // 4 cap
        uint32_t* s_key = reinterpret_cast<uint32_t*>(s_vset + 4u * cap);                            // 2 cap
        uint32_t* s_bits = s_key + 2u * cap;                                                         // 2 cap
        uint32_t* s_ta = s_bits + 2u * cap;                                                          // cap: the pair's hits, staged once for both sides
        uint32_t* s_tb = s_ta + cap;                                                                 // cap
        uint32_t* s_fl = s_tb + cap;                                                                 // cap: HitAux.flags | (weight != 0) << 31
        LT* s_cand = reinterpret_cast<LT*>(s_fl + cap);                                              // cap: claimed candidate slots
        LT* s_vert = s_cand + cap;                                                                   // 3 cap: claimed vertex slots
        LT* s_avgl = s_vert + 3u * cap;                                                              // cap: candidates that fall back to the average point
        const uint32_t* grp = grouped + acc[p].off;
        const float4* pp = reinterpret_cast<const float4*>(pairrec + p);
        Rel rel; rel.r0 = __ldg(pp); rel.r1 = __ldg(pp + 1); rel.r2 = __ldg(pp + 2);
        const M3 nmat = adjoint_transpose3(rel);                                                     // CreateUncollideRays.cpp:65
        bool zero_w = false;
        for (uint32_t k = tid; k < n; k += T) {
            const uint32_t h = grp[k];
            const HitAux x = aux[h];
            const bool nz = !(hits[h].weight == 0.f);
            zero_w |= !nz;
            s_ta[k] = x.triA; s_tb[k] = x.triB; s_fl[k] = x.flags | (nz ? 0x80000000u : 0u);
        }
        const bool any_zero_w = __syncthreads_or(zero_w) != 0;          // block-uniform: does the pair have hits of weight 0 at all?
        for (uint32_t side = 0; side < 2; ++side) {
            for (uint32_t k = tid; k < slots; k += T) { s_key[k] = 0xffffffffu; s_bits[k] = 7u; s_sum[k].w = 0.0; s_sum[k].cx = 0.0; s_sum[k].cy = 0.0; s_sum[k].cz = 0.0; }
            for (uint32_t k = tid; k < vslots; k += T) s_vset[k] = ~0ull;
            if (tid == 0) { s_ncand = 0u; s_nvert = 0u; s_navg = 0u; }
            __syncthreads();
            if (any_zero_w) {
                // the vertex table is idle until the candidates are walked: it first holds the set of (own triangle, other leaf) with weight
                for (uint32_t k = tid; k < n; k += T) {
                    const uint32_t fl = s_fl[k];
                    if (fl >> 31) {
                        const uint32_t own = side ? s_tb[k] : s_ta[k];
                        const uint32_t leaf = side ? s_ta[k] - ((fl >> 6) & 3u) : s_tb[k] - ((fl >> 8) & 3u);
                        pc_combo_insert(s_vset, vslots - 1u, ((unsigned long long)own << 32) | leaf);
                    }
                }
                __syncthreads();
            }
            // ---- one thread per hit: merge into the own triangle's candidate ----
            for (uint32_t k0 = 0; k0 < n; k0 += T) {
                const uint32_t k = k0 + tid;
                uint32_t slot = 0xffffffffu, bits = 7u;
                double w = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
                bool claimed = false;
                if (k < n) {
                    const uint32_t fl = s_fl[k];
                    const uint32_t own = side ? s_tb[k] : s_ta[k];
                    bool contributes = (fl >> 31) != 0u;
                    if (!contributes) {                                                  // is the combo's weight for this triangle 0? (:117-127)
                        const uint32_t leaf = side ? s_ta[k] - ((fl >> 6) & 3u) : s_tb[k] - ((fl >> 8) & 3u);
                        contributes = pc_combo_contains<false>(s_vset, vslots - 1u, ((unsigned long long)own << 32) | leaf);
                    }
                    if (contributes) {
                        const imrcd_tri_hit hh = hits[grp[k]];
                        const V3 sum = add3(mk3(hh.source[0], hh.source[1], hh.source[2]), mk3(hh.target[0], hh.target[1], hh.target[2]));
                        slot = pc_slot_of(s_key, slots - 1u, own, claimed);
                        bits = side ? ((fl >> 3) & 7u) : (fl & 7u);
                        w = (double)hh.weight;
                        cx = (double)((hh.weight * sum.x) / 2.f); cy = (double)((hh.weight * sum.y) / 2.f); cz = (double)((hh.weight * sum.z) / 2.f);   // :94-100
                    }
                }
                pc_append<LT>(claimed, slot, s_cand, &s_ncand, lane);
                // Hits come out of the narrow phase combo by combo, so one triangle's hits mostly sit in consecutive lanes (and a large triangle
                // collects many): add up each run of equal slots inside the warp first (segmented scan), then one update per run.
                const uint32_t prev = __shfl_up_sync(FULL_MASK, slot, 1);
                const uint32_t heads = __ballot_sync(FULL_MASK, lane == 0u || prev != slot);
                const uint32_t head = 31u - (uint32_t)__clz(heads & (0xffffffffu >> (31u - lane)));           // first lane of this lane's run
                const uint32_t after = heads & ~(0xffffffffu >> (31u - lane));                                // run heads above this lane
                const uint32_t tail = after ? (uint32_t)__ffs(after) - 2u : 31u;                              // last lane of the run
#pragma unroll
                for (uint32_t d = 1; d < 32u; d <<= 1) {
                    const double vw = __shfl_up_sync(FULL_MASK, w, d), vx = __shfl_up_sync(FULL_MASK, cx, d), vy = __shfl_up_sync(FULL_MASK, cy, d), vz = __shfl_up_sync(FULL_MASK, cz, d);
                    const uint32_t vb = __shfl_up_sync(FULL_MASK, bits, d);
                    if (lane >= head + d) { w += vw; cx += vx; cy += vy; cz += vz; bits &= vb; }
                }
                if (lane == tail && slot != 0xffffffffu) {
                    atomicAnd(&s_bits[slot], bits);
                    atomicAdd(&s_sum[slot].w, w); atomicAdd(&s_sum[slot].cx, cx); atomicAdd(&s_sum[slot].cy, cy); atomicAdd(&s_sum[slot].cz, cz);
                }
            }
            __syncthreads();
            if (any_zero_w) {                                                            // the vertex table back to empty
                for (uint32_t k = tid; k < vslots; k += T) s_vset[k] = ~0ull;
                __syncthreads();
            }
            // ---- one thread per candidate: vertex rays into the `emplaced` set, average-point rays onto their list (:139-166) ----
            const uint32_t n_cand = s_ncand;
            for (uint32_t c0 = 0; c0 < n_cand; c0 += T) {
                const uint32_t c = c0 + tid;
                uint32_t q0 = 0xffffffffu, q1 = 0xffffffffu, q2 = 0xffffffffu;           // vertex slots claimed by this lane
                uint32_t fallback = 0xffffffffu;
                if (c < n_cand) {
                    const uint32_t k = s_cand[c];
                    const uint32_t tri = s_key[k], bits = s_bits[k];
                    if (bits == 0u) fallback = k;                                        // ShouldFallbackToAvgPoint (:27-30)
                    else {
                        const uint32_t v0 = tri_vid[3ull * tri], v1 = tri_vid[3ull * tri + 1], v2 = tri_vid[3ull * tri + 2];
                        if (bits & 1u) q0 = vset_insert_m(s_vset, vslots - 1u, ((unsigned long long)v0 << 32) | (unsigned long long)(tri * 4u));
                        if (bits & 2u) q1 = vset_insert_m(s_vset, vslots - 1u, ((unsigned long long)v1 << 32) | (unsigned long long)(tri * 4u + 1u));
                        if (bits & 4u) q2 = vset_insert_m(s_vset, vslots - 1u, ((unsigned long long)v2 << 32) | (unsigned long long)(tri * 4u + 2u));
                    }
                }
                pc_append<LT>(q0 != 0xffffffffu, q0, s_vert, &s_nvert, lane);
                pc_append<LT>(q1 != 0xffffffffu, q1, s_vert, &s_nvert, lane);
                pc_append<LT>(q2 != 0xffffffffu, q2, s_vert, &s_nvert, lane);
                pc_append<LT>(fallback != 0xffffffffu, fallback, s_avgl, &s_navg, lane);
            }
            __syncthreads();
            const uint32_t n_vert = s_nvert, n_avg = s_navg;
            // the pair's slice of the frame's ray array (pairs that moved only)
            if (tid == 0) s_raybase = keep_rays ? (uint32_t)atomicAdd(&ctl->n_rays_kept, (unsigned long long)(n_vert + n_avg)) : 0u;
            __syncthreads();
            const uint32_t ray_base = s_raybase;
            const bool emit = keep_rays && (unsigned long long)ray_base + n_vert + n_avg <= cap_rays;
            if (keep_rays && !emit && tid == 0) atomicOr(&ctl->overflow, (unsigned)OVF_RAYS);
            SideSum r; r.x = r.y = r.z = 0.0; r.rays = 0u;
            // ---- rays at the weighted average point (:34-37,157-166) ----
            for (uint32_t c = tid; c < n_avg; c += T) {
                const uint32_t k = s_avgl[c];
                // the quotient is taken in FP64 and rounded once: rounding the two sums first would turn the last-bit noise of the
                // order-free FP64 sums into FP32 differences whenever a sum sits on a rounding tie (two equal-exponent addends do)
                const double w = s_sum[k].w;
                const V3 pos = mk3((float)(s_sum[k].cx / w), (float)(s_sum[k].cy / w), (float)(s_sum[k].cz / w));
                r.x += (double)pos.x; r.y += (double)pos.y; r.z += (double)pos.z; r.rays += 1u;
                if (emit) {
                    const uint32_t tri = s_key[k];
                    const float4* tp = reinterpret_cast<const float4*>(tris + tri);
                    const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                    V3 p0 = mk3(t0.x, t0.y, t0.z), p1 = mk3(t1.x, t1.y, t1.z), p2 = mk3(t2.x, t2.y, t2.z);
                    if (side) { p0 = rel_mul(rel, p0, 1.f); p1 = rel_mul(rel, p1, 1.f); p2 = rel_mul(rel, p2, 1.f); }     // :134
                    float bx, by;
                    tri_barycentric(p0, p1, p2, pos, bx, by);                            // :160
                    const float* nn = tri_nrm + 9ull * tri;
                    V3 nrm = tri_interp_normal(mk3(nn[0], nn[1], nn[2]), mk3(nn[3], nn[4], nn[5]), mk3(nn[6], nn[7], nn[8]), bx, by);
                    nrm = normalize3(side ? m3_mul(nmat, nrm) : nrm);                    // :163-164, Triangle.cpp:197-212
                    RayRec o; o.o = make_float4(pos.x, pos.y, pos.z, __uint_as_float(p)); o.d = make_float4(-nrm.x, -nrm.y, -nrm.z, __uint_as_float(side));
                    rays[ray_base + c] = o;
                }
            }
            // ---- vertex rays: one per distinct vertex id (the `emplaced` set, :139-155) ----
            for (uint32_t c = tid; c < n_vert; c += T) {
                const uint32_t ref = (uint32_t)s_vset[s_vert[c]];
                const float4 q = __ldg(reinterpret_cast<const float4*>(tris + (ref >> 2)) + (ref & 3u));
                V3 pos = mk3(q.x, q.y, q.z);
                if (side) pos = rel_mul(rel, pos, 1.f);                                  // second's triangles live in first's space (:84,:134)
                r.x += (double)pos.x; r.y += (double)pos.y; r.z += (double)pos.z; r.rays += 1u;
                if (emit) {
                    const float* nn = tri_nrm + 9ull * (ref >> 2) + 3u * (ref & 3u);
                    V3 nrm = mk3(nn[0], nn[1], nn[2]);
                    nrm = normalize3(side ? m3_mul(nmat, nrm) : nrm);                    // :150-151, Triangle.cpp:180-195
                    RayRec o; o.o = make_float4(pos.x, pos.y, pos.z, __uint_as_float(p)); o.d = make_float4(-nrm.x, -nrm.y, -nrm.z, __uint_as_float(side));
                    rays[ray_base + n_avg + c] = o;
                }
            }
            for (int o = 16; o > 0; o >>= 1) {
                r.x += __shfl_down_sync(FULL_MASK, r.x, o); r.y += __shfl_down_sync(FULL_MASK, r.y, o); r.z += __shfl_down_sync(FULL_MASK, r.z, o);
                r.rays += __shfl_down_sync(FULL_MASK, r.rays, o);
            }
            if (lane == 0u) { s_red[0][tid >> 5] = r.x; s_red[1][tid >> 5] = r.y; s_red[2][tid >> 5] = r.z; s_redc[tid >> 5] = r.rays; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < T / 32; ++w) { r.x += s_red[0][w]; r.y += s_red[1][w]; r.z += s_red[2][w]; r.rays += s_redc[w]; }
                PairAcc* pa = acc + p;
                double* sum = side ? pa->sum_b : pa->sum_a;
                sum[0] = r.x; sum[1] = r.y; sum[2] = r.z;
                if (side) { pa->rays_b = r.rays; pa->ray_off_b = ray_base; } else { pa->rays_a = r.rays; pa->ray_off_a = ray_base; }
                if (!emit) pa->flags &= ~(uint32_t)PAIR_MOVED;                           // no rays kept: the response stage skips the pair
            }
            __syncthreads();
        }
    }
}

// ---- size class L (> 1024 hits): all large pairs together, the whole machine on them ---------------------------------------
// A block per pair is the wrong shape for a pair with tens of thousands of hits (two deeply interpenetrating meshes: C2, C5): the same keyed
// reduction runs here as a short sequence of grid-wide passes over ALL large pairs, tables in a global scratch (2m candidate slots and 4m vertex
// slots per side for a pair padded to m hits), every atomic a native L2 RED (AND, FP64 add, 64-bit min).  `pref` = exclusive prefix of the
// pairs' padded sizes: thread t of a pass serves unit t - pref[i] of pair i (found by binary search).
struct LargeSide { uint32_t n_avg, n_vert, ray_base, cursor; };
#define PCL_BYTES_PER_UNIT 224ull      // per padded hit, both sides: 2 x (2 keys + 2 bits + 2 PcSlot + 4 vertex entries)

struct LargeTables { uint32_t* key; uint32_t* bits; PcSlot* sum; unsigned long long* vset; };
__device__ __forceinline__ LargeTables pcl_tables(unsigned char* scratch, unsigned long long unit0, uint32_t m, uint32_t side) {
    unsigned char* b = scratch + unit0 * PCL_BYTES_PER_UNIT + (unsigned long long)side * 112ull * m;
    LargeTables t;
    t.sum = reinterpret_cast<PcSlot*>(b);                                    // 2m x 32 B
    t.vset = reinterpret_cast<unsigned long long*>(b + 64ull * m);           // 4m x 8 B
    t.key = reinterpret_cast<uint32_t*>(b + 96ull * m);                      // 2m x 4 B
    t.bits = reinterpret_cast<uint32_t*>(b + 104ull * m);                    // 2m x 4 B
    return t;
}
__device__ __forceinline__ uint32_t pcl_padded(uint32_t n) { uint32_t m = 16u; while (m < n) m <<= 1; return m; }

// one block: exclusive prefix of the padded sizes of the large pairs, scratch budget check, per-pair-side counters
__global__ void __launch_bounds__(1024)
k_large_layout(FrameCtl* ctl, const uint32_t* __restrict__ list, const PairAcc* __restrict__ acc, unsigned long long* __restrict__ pref, LargeSide* __restrict__ sides,
               unsigned long long cap_scratch) {
    __shared__ unsigned long long s_part[1024];
    const unsigned long long n_list = (ctl->overflow & (OVF_PAIRS | OVF_QUEUE | OVF_COMBOS | OVF_HITS)) ? 0ull : ctl->n_class[3 * 16];
    const uint32_t tid = threadIdx.x;
    const unsigned long long per = (n_list + 1023ull) / 1024ull, lo = tid * per, hi = lo + per < n_list ? lo + per : n_list;
    unsigned long long sum = 0;
    for (unsigned long long i = lo; i < hi; ++i) sum += pcl_padded(acc[list[i]].n_hits);
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) { unsigned long long run = 0; for (int k = 0; k < 1024; ++k) { const unsigned long long v = s_part[k]; s_part[k] = run; run += v; } pref[n_list] = run;
                    ctl->scratch_used = run * PCL_BYTES_PER_UNIT; if (run * PCL_BYTES_PER_UNIT > cap_scratch) { atomicOr(&ctl->overflow, (unsigned)OVF_SCRATCH); pref[n_list] = 0; } }
    __syncthreads();
    unsigned long long run = s_part[tid];
    for (unsigned long long i = lo; i < hi; ++i) {
        pref[i] = run; run += pcl_padded(acc[list[i]].n_hits);
        LargeSide z; z.n_avg = z.n_vert = z.ray_base = z.cursor = 0u; sides[2 * i] = z; sides[2 * i + 1] = z;
    }
}

// largest i in [0, n) with pref[i] <= t  (pref is non-decreasing, pref[0] = 0)
__device__ __forceinline__ uint32_t pcl_find(const unsigned long long* __restrict__ pref, uint32_t n, unsigned long long t) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1u) { const uint32_t mid = (lo + hi) >> 1; if (pref[mid] <= t) lo = mid; else hi = mid; }
    return lo;
}

__global__ void k_large_init(const FrameCtl* ctl, const uint32_t* __restrict__ list, const PairAcc* __restrict__ acc, const unsigned long long* __restrict__ pref, unsigned char* scratch) {
    const uint32_t n_list = (uint32_t)ctl->n_class[3 * 16];
    if (ctl->overflow) return;
    const unsigned long long total = pref[n_list];
    for (unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t i = pcl_find(pref, n_list, t);
        const uint32_t m = pcl_padded(acc[list[i]].n_hits), u = (uint32_t)(t - pref[i]);
#pragma unroll
        for (uint32_t side = 0; side < 2; ++side) {
            const LargeTables tb = pcl_tables(scratch, pref[i], m, side);
#pragma unroll
            for (uint32_t q = 0; q < 2; ++q) { const uint32_t k = 2u * u + q; tb.key[k] = 0xffffffffu; tb.bits[k] = 7u; tb.sum[k].w = 0.0; tb.sum[k].cx = 0.0; tb.sum[k].cy = 0.0; tb.sum[k].cz = 0.0; }
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) tb.vset[4u * u + q] = ~0ull;
        }
    }
}

// the (own triangle, other leaf) keys with weight, both sides, into the (still idle) vertex tables; k_large_unmark empties them again
__global__ void __launch_bounds__(256)
k_large_mark(const FrameCtl* ctl, const uint32_t* __restrict__ list, const PairAcc* __restrict__ acc, const unsigned long long* __restrict__ pref, unsigned char* scratch,
             const uint32_t* __restrict__ grouped, const imrcd_tri_hit* __restrict__ hits, const HitAux* __restrict__ aux) {
    const uint32_t n_list = (uint32_t)ctl->n_class[3 * 16];
    if (ctl->overflow) return;
    const unsigned long long total = pref[n_list];
    for (unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t i = pcl_find(pref, n_list, t);
        const PairAcc& pa = acc[list[i]];
        const uint32_t n = pa.n_hits, m = pcl_padded(n), k = (uint32_t)(t - pref[i]);
        if (k >= n) continue;
        const uint32_t h = grouped[pa.off + k];
        if (hits[h].weight == 0.f) continue;
        const HitAux x = aux[h];
        pc_combo_insert(pcl_tables(scratch, pref[i], m, 0).vset, 4u * m - 1u, ((unsigned long long)x.triA << 32) | (x.triB - ((x.flags >> 8) & 3u)));
        pc_combo_insert(pcl_tables(scratch, pref[i], m, 1).vset, 4u * m - 1u, ((unsigned long long)x.triB << 32) | (x.triA - ((x.flags >> 6) & 3u)));
    }
}
__global__ void k_large_unmark(const FrameCtl* ctl, const uint32_t* __restrict__ list, const PairAcc* __restrict__ acc, const unsigned long long* __restrict__ pref, unsigned char* scratch) {
    const uint32_t n_list = (uint32_t)ctl->n_class[3 * 16];
    if (ctl->overflow) return;
    const unsigned long long total = pref[n_list];
    for (unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t i = pcl_find(pref, n_list, t);
        const uint32_t m = pcl_padded(acc[list[i]].n_hits), u = (uint32_t)(t - pref[i]);
#pragma unroll
        for (uint32_t side = 0; side < 2; ++side) {
            unsigned long long* vs = pcl_tables(scratch, pref[i], m, side).vset;
#pragma unroll
            for (uint32_t q = 0; q < 4; ++q) vs[4u * u + q] = ~0ull;
        }
    }
}

// one thread per hit of a large pair, both sides
__global__ void __launch_bounds__(256)
k_large_hits(const FrameCtl* ctl, const uint32_t* __restrict__ list, const PairAcc* __restrict__ acc, const unsigned long long* __restrict__ pref, unsigned char* scratch,
             const uint32_t* __restrict__ grouped, const imrcd_tri_hit* __restrict__ hits, const HitAux* __restrict__ aux) {
    const uint32_t n_list = (uint32_t)ctl->n_class[3 * 16];
    if (ctl->overflow) return;
    const unsigned long long total = pref[n_list];
    const uint32_t lane = lane_id();
    for (unsigned long long t0 = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) & ~31ull; t0 < total; t0 += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long t = t0 + lane;
        bool have = false;
        uint32_t i = 0, m = 0, n = 0, k = 0;
        const uint32_t* grp = nullptr;
        HitAux x; x.triA = x.triB = x.flags = 0u;
        imrcd_tri_hit hh; hh.weight = 0.f; hh.source[0] = hh.source[1] = hh.source[2] = hh.target[0] = hh.target[1] = hh.target[2] = 0.f;
        if (t < total) {
            i = pcl_find(pref, n_list, t);
            const PairAcc& pa = acc[list[i]];
            n = pa.n_hits; m = pcl_padded(n); k = (uint32_t)(t - pref[i]); grp = grouped + pa.off;
            if (k < n) { have = true; const uint32_t h = grp[k]; x = aux[h]; hh = hits[h]; }
        }
#pragma unroll
        for (uint32_t side = 0; side < 2; ++side) {
            uint32_t slot = 0xffffffffu, bits = 7u, tab = 0xffffffffu;
            double w = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
            LargeTables tb; tb.key = nullptr; tb.bits = nullptr; tb.sum = nullptr; tb.vset = nullptr;
            if (have) {
                const uint32_t own = side ? x.triB : x.triA;
                bool contributes = !(hh.weight == 0.f);
                tb = pcl_tables(scratch, pref[i], m, side);
                if (!contributes) {                                                      // is the combo's weight for this triangle 0? (:117-127)
                    const uint32_t leaf = side ? x.triA - ((x.flags >> 6) & 3u) : x.triB - ((x.flags >> 8) & 3u);
                    contributes = pc_combo_contains<true>(tb.vset, 4u * m - 1u, ((unsigned long long)own << 32) | leaf);
                }
                if (contributes) {
                    bool claimed = false;
                    slot = pc_slot_of(tb.key, 2u * m - 1u, own, claimed);
                    tab = i;
                    const V3 sum = add3(mk3(hh.source[0], hh.source[1], hh.source[2]), mk3(hh.target[0], hh.target[1], hh.target[2]));
                    bits = side ? ((x.flags >> 3) & 7u) : (x.flags & 7u);
                    w = (double)hh.weight;
                    cx = (double)((hh.weight * sum.x) / 2.f); cy = (double)((hh.weight * sum.y) / 2.f); cz = (double)((hh.weight * sum.z) / 2.f);   // :94-100
                }
            }
            // runs of equal (pair, slot) in consecutive lanes are added up inside the warp first (see k_pair_contacts_hash)
            const uint32_t prev_s = __shfl_up_sync(FULL_MASK, slot, 1), prev_t = __shfl_up_sync(FULL_MASK, tab, 1);
            const uint32_t heads = __ballot_sync(FULL_MASK, lane == 0u || prev_s != slot || prev_t != tab);
            const uint32_t head = 31u - (uint32_t)__clz(heads & (0xffffffffu >> (31u - lane)));
            const uint32_t after = heads & ~(0xffffffffu >> (31u - lane));
            const uint32_t tail = after ? (uint32_t)__ffs(after) - 2u : 31u;
#pragma unroll
            for (uint32_t d = 1; d < 32u; d <<= 1) {
                const double vw = __shfl_up_sync(FULL_MASK, w, d), vx = __shfl_up_sync(FULL_MASK, cx, d), vy = __shfl_up_sync(FULL_MASK, cy, d), vz = __shfl_up_sync(FULL_MASK, cz, d);
                const uint32_t vb = __shfl_up_sync(FULL_MASK, bits, d);
                if (lane >= head + d) { w += vw; cx += vx; cy += vy; cz += vz; bits &= vb; }
            }
            if (lane == tail && slot != 0xffffffffu) {
                atomicAnd(&tb.bits[slot], bits);
                atomicAdd(&tb.sum[slot].w, w); atomicAdd(&tb.sum[slot].cx, cx); atomicAdd(&tb.sum[slot].cy, cy); atomicAdd(&tb.sum[slot].cz, cz);
            }
        }
    }
}

// one thread per candidate slot (2m per side): vertex rays into the `emplaced` set, average-point rays counted (:139-166)
__global__ void __launch_bounds__(256)
k_large_candidates(const FrameCtl* ctl, const uint32_t* __restrict__ list, const PairAcc* __restrict__ acc, const unsigned long long* __restrict__ pref, unsigned char* scratch,
                   LargeSide* __restrict__ sides, const uint32_t* __restrict__ tri_vid) {
    const uint32_t n_list = (uint32_t)ctl->n_class[3 * 16];
    if (ctl->overflow) return;
    const unsigned long long total = pref[n_list] * 4ull;                       // 2 sides x 2m slots per padded hit
    for (unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t i = pcl_find(pref, n_list, t >> 2);
        const uint32_t m = pcl_padded(acc[list[i]].n_hits);
        const unsigned long long r = t - pref[i] * 4ull;                          // [0, 4m)
        const uint32_t side = (uint32_t)(r / (2ull * m)), k = (uint32_t)(r % (2ull * m));
        const LargeTables tb = pcl_tables(scratch, pref[i], m, side);
        const uint32_t tri = __ldcg(tb.key + k);
        if (tri == 0xffffffffu) continue;
        const uint32_t bits = __ldcg(tb.bits + k);
        if (bits == 0u) atomicAdd(&sides[2 * i + side].n_avg, 1u);                // ShouldFallbackToAvgPoint (:27-30)
        else {
            uint32_t claimed = 0;
#pragma unroll
            for (uint32_t pi = 0; pi < 3; ++pi)
                if ((bits >> pi) & 1u)
                    claimed += vset_insert_m(tb.vset, 4u * m - 1u, ((unsigned long long)tri_vid[3ull * tri + pi] << 32) | (unsigned long long)(tri * 4u + pi)) != 0xffffffffu;
            if (claimed) atomicAdd(&sides[2 * i + side].n_vert, claimed);
        }
    }
}

// one thread per pair side: ray counts, the pair's slice of the ray array
__global__ void k_large_alloc(FrameCtl* ctl, const uint32_t* __restrict__ list, PairAcc* acc, LargeSide* __restrict__ sides, unsigned long long cap_rays) {
    const uint32_t n_list = (uint32_t)ctl->n_class[3 * 16];
    if (ctl->overflow & ~(unsigned)OVF_RAYS) return;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < 2u * n_list; t += gridDim.x * blockDim.x) {
        const uint32_t i = t >> 1, side = t & 1u;
        PairAcc* pa = acc + list[i];
        LargeSide& sd = sides[t];
        const uint32_t cnt = sd.n_avg + sd.n_vert;
        if (side) pa->rays_b = cnt; else pa->rays_a = cnt;
        if (pa->flags & PAIR_MOVED) {
            const unsigned long long base = atomicAdd(&ctl->n_rays_kept, (unsigned long long)cnt);
            if (base + cnt > cap_rays) { atomicOr(&ctl->overflow, (unsigned)OVF_RAYS); sd.ray_base = 0xffffffffu; }
            else { sd.ray_base = (uint32_t)base; if (side) pa->ray_off_b = (uint32_t)base; else pa->ray_off_a = (uint32_t)base; }
        } else sd.ray_base = 0xffffffffu;
    }
}

// one thread per candidate slot and per vertex slot: the ray's origin into the pair's FP64 sum, the ray into the pair's slice
__global__ void __launch_bounds__(256)
k_large_rays(const FrameCtl* ctl, const uint32_t* __restrict__ list, PairAcc* acc, const unsigned long long* __restrict__ pref, unsigned char* scratch,
             LargeSide* __restrict__ sides, const PairRec* __restrict__ pairrec, const TriRec* __restrict__ tris, const float* __restrict__ tri_nrm, RayRec* __restrict__ rays) {
    const uint32_t n_list = (uint32_t)ctl->n_class[3 * 16];
    if (ctl->overflow) return;
    const unsigned long long total = pref[n_list] * 12ull;                      // per padded hit: 2 sides x (2 candidate slots + 4 vertex slots)
    for (unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t i = pcl_find(pref, n_list, t / 12ull);
        const uint32_t p = list[i];
        const uint32_t m = pcl_padded(acc[p].n_hits);
        const unsigned long long r = t - pref[i] * 12ull;                         // [0, 12m)
        const uint32_t side = (uint32_t)(r / (6ull * m)), k = (uint32_t)(r % (6ull * m));
        const LargeTables tb = pcl_tables(scratch, pref[i], m, side);
        LargeSide& sd = sides[2 * i + side];
        V3 pos, nrm;
        if (k < 2u * m) {                                                      // candidate slot: only the average-point rays (:157-166)
            const uint32_t tri = __ldcg(tb.key + k);
            if (tri == 0xffffffffu || __ldcg(tb.bits + k) != 0u) continue;
            const double w = __ldcg(&tb.sum[k].w);
            pos = mk3((float)(__ldcg(&tb.sum[k].cx) / w), (float)(__ldcg(&tb.sum[k].cy) / w), (float)(__ldcg(&tb.sum[k].cz) / w));     // FP64 quotient, rounded once
            if (sd.ray_base != 0xffffffffu) {
                const float4* pp = reinterpret_cast<const float4*>(pairrec + p);
                Rel rel; rel.r0 = __ldg(pp); rel.r1 = __ldg(pp + 1); rel.r2 = __ldg(pp + 2);
                const float4* tp = reinterpret_cast<const float4*>(tris + tri);
                const float4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);
                V3 p0 = mk3(t0.x, t0.y, t0.z), p1 = mk3(t1.x, t1.y, t1.z), p2 = mk3(t2.x, t2.y, t2.z);
                if (side) { p0 = rel_mul(rel, p0, 1.f); p1 = rel_mul(rel, p1, 1.f); p2 = rel_mul(rel, p2, 1.f); }
                float bx, by;
                tri_barycentric(p0, p1, p2, pos, bx, by);
                const float* nn = tri_nrm + 9ull * tri;
                nrm = tri_interp_normal(mk3(nn[0], nn[1], nn[2]), mk3(nn[3], nn[4], nn[5]), mk3(nn[6], nn[7], nn[8]), bx, by);
                nrm = normalize3(side ? m3_mul(adjoint_transpose3(rel), nrm) : nrm);
            }
        } else {                                                               // vertex slot (:139-155)
            const unsigned long long ent = __ldcg(tb.vset + (k - 2u * m));
            if (ent == ~0ull) continue;
            const uint32_t ref = (uint32_t)ent;
            const float4 q = __ldg(reinterpret_cast<const float4*>(tris + (ref >> 2)) + (ref & 3u));
            pos = mk3(q.x, q.y, q.z);
            const float4* pp = reinterpret_cast<const float4*>(pairrec + p);
            Rel rel; rel.r0 = __ldg(pp); rel.r1 = __ldg(pp + 1); rel.r2 = __ldg(pp + 2);
            if (side) pos = rel_mul(rel, pos, 1.f);
            if (sd.ray_base != 0xffffffffu) {
                const float* nn = tri_nrm + 9ull * (ref >> 2) + 3u * (ref & 3u);
                nrm = mk3(nn[0], nn[1], nn[2]);
                nrm = normalize3(side ? m3_mul(adjoint_transpose3(rel), nrm) : nrm);
            }
        }
        double* sum = side ? acc[p].sum_b : acc[p].sum_a;
        atomicAdd(sum, (double)pos.x); atomicAdd(sum + 1, (double)pos.y); atomicAdd(sum + 2, (double)pos.z);
        if (sd.ray_base != 0xffffffffu) {
            RayRec o; o.o = make_float4(pos.x, pos.y, pos.z, __uint_as_float(p)); o.d = make_float4(-nrm.x, -nrm.y, -nrm.z, __uint_as_float(side));
            rays[sd.ray_base + atomicAdd(&sd.cursor, 1u)] = o;
        }
    }
}

// row 0 of the result block: the number of records that follow and the frame's overflow bits (what a fixed-capacity all-gather of the
// block needs to carry: every rank learns from the gathered headers whether any rank has to run its frame again)
__global__ void k_epairs_header(const FrameCtl* ctl, imrcd_entity_pair* block) {
    imrcd_entity_pair h; memset(&h, 0, sizeof(h));
    const unsigned long long n = ctl->n_colliding;
    h.entry_first = (uint32_t)n; h.entry_second = (uint32_t)(n >> 32); h.entity_first = ctl->overflow;
    block[0] = h;
}

__global__ void k_finalize(FrameCtl* ctl, unsigned long long cap_pairs, const uint2* __restrict__ pairs, const PairAcc* __restrict__ acc,
                           const uint32_t* __restrict__ entity, const float* __restrict__ cur, const float* __restrict__ inv,
                           imrcd_entity_pair* __restrict__ out, uint32_t* __restrict__ out_pair, const uint32_t* __restrict__ gidx) {
    unsigned long long n = ctl->n_pairs < cap_pairs ? ctl->n_pairs : cap_pairs;
    for (unsigned long long p = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; p < n; p += (unsigned long long)gridDim.x * blockDim.x) {
        if (!(acc[p].flags & 1u)) continue;
        const PairAcc a = acc[p];
        unsigned long long slot = atomicAdd(&ctl->n_colliding, 1ull);
        uint2 pr = pairs[p];
        imrcd_entity_pair o;
        memset(&o, 0, sizeof(o));
        o.entry_first = gidx ? gidx[pr.x] : pr.x; o.entry_second = gidx ? gidx[pr.y] : pr.y;      // the caller's entry indices
        o.entity_first = entity[pr.x]; o.entity_second = entity[pr.y];
        o.n_hits = a.n_hits; o.flags = 1u;
        o.n_rays_first = a.rays_a; o.n_rays_second = a.rays_b;
        atomicAdd(&ctl->n_rays, (unsigned long long)(a.rays_a + a.rays_b));
        // average_point_first_modelspace = sum / count (:185-189); second: back to its own model space through
        // inverse(second_to_first_space_matrix) (:191-198).  0 rays on a side gives NaN exactly like the reference.
        const double fa = (double)a.rays_a, fb = (double)a.rays_b;            // FP64 quotient, rounded once (see k_pair_contacts_hash)
        o.avg_first[0] = (float)(a.sum_a[0] / fa); o.avg_first[1] = (float)(a.sum_a[1] / fa); o.avg_first[2] = (float)(a.sum_a[2] / fa);
        const V3 sb = mk3((float)(a.sum_b[0] / fb), (float)(a.sum_b[1] / fb), (float)(a.sum_b[2] / fb));
        float rel[16], rinv[16];
        mat4_mul(inv + 16 * (size_t)pr.x, cur + 16 * (size_t)pr.y, rel);            // the same rel as k_pair_setup (OBBtreesCollision.cpp:15)
        mat4_inverse(rel, rinv);
        const V3 back = rel_mul(rel_from_mat(rinv), sb, 1.f);
        o.avg_second[0] = back.x; o.avg_second[1] = back.y; o.avg_second[2] = back.z;
        out[slot] = o;
        out_pair[slot] = (uint32_t)p;
    }
}

// ------------------------------------------------------------------------------------------
// host orchestration
// ------------------------------------------------------------------------------------------
static inline unsigned blocks_for(unsigned long long n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

int imr_comm_reserve(imrcd_ctx* ctx);
int imr_comm_allgather(imrcd_ctx* ctx);
int imr_comm_after_gather(imrcd_ctx* ctx, uint64_t spec_rows);
int imr_comm_decide(imrcd_ctx* ctx, uint64_t spec_rows, bool* retry, bool* fatal);

// rows of the result that travel to the host speculatively, right behind the frame's kernels (the count is not known on the host yet):
// what the last frame had plus a margin.  A frame with more records pays one extra copy after the wait.
uint64_t imr_frame_spec_rows(const imrcd_ctx* ctx) {
    const uint64_t want = std::max<uint64_t>(256, ctx->spec_hint + ctx->spec_hint / 4 + 64);
    if (want >= 2048) return (want + 1023) & ~1023ull;      // coarse steps: the copy's size is part of what a captured frame is made of
    uint64_t r = 256;
    while (r < want) r <<= 1;
    return r;
}

// Second half of a frame: wait for the stream, read the control block back, grow whatever overflowed (the caller re-runs the frame) or
// fill in the statistics.  Split from the enqueue half so that a caller can put more work on the stream before the host looks at the frame
// (imrcd_frame_run_async / imrcd_frame_finish).  With a communicator the re-run decision is taken from the GATHERED headers, so every rank
// takes the same one and issues the same number of collectives.
int imr_frame_complete(imrcd_ctx* ctx, bool* retry) {
    cudaStream_t s = ctx->stream;
    *retry = false;
    IMR_CUDA(ctx, cudaStreamSynchronize(s));
    IMR_CUDA(ctx, cudaGetLastError());
    ctx->ctl_host = *ctx->p_ctl.as<FrameCtl>();
    const FrameCtl& c = ctx->ctl_host;
    ctx->queue_dirty = std::min<uint64_t>(std::max(c.q_tail, c.q_head), ctx->cap_queue);
    bool fatal = (c.overflow & OVF_RAYSTACK) != 0;
    if (ctx->comm) { const int rc = imr_comm_decide(ctx, ctx->spec_rows_sent, retry, &fatal); if (rc) return rc; }
    else *retry = c.overflow != 0;
    if (fatal) { ctx->err = "ray stack overflow (tree deeper than the per-thread stack of the response stage)"; return IMRCD_E_CAPACITY; }

    if (c.overflow) {
        if (c.overflow & OVF_PAIRS) ctx->cap_pairs = std::max<uint64_t>(c.n_pairs + c.n_pairs / 8, ctx->cap_pairs * 2);
        if (c.overflow & OVF_QUEUE) ctx->cap_queue = std::max<uint64_t>(ctx->cap_queue * 2, ctx->cap_pairs + (1ull << 22));
        if (ctx->cap_queue < ctx->cap_pairs) ctx->cap_queue = ctx->cap_pairs + (1ull << 22);
        if (c.overflow & OVF_COMBOS) ctx->cap_combos = std::max<uint64_t>(c.n_combos + c.n_combos / 8, ctx->cap_combos * 2);
        if (c.overflow & OVF_HITS) ctx->cap_hits = std::max<uint64_t>(c.n_hits + c.n_hits / 8, ctx->cap_hits * 2);
        if (c.overflow & OVF_RAYS) ctx->cap_rays = std::max<uint64_t>(c.n_rays_kept + c.n_rays_kept / 8, ctx->cap_rays * 2);
        if (c.overflow & OVF_SCRATCH) ctx->cap_lscratch = std::max<uint64_t>(c.scratch_used + c.scratch_used / 8, ctx->cap_lscratch * 2);
    }
    if (*retry) return IMRCD_OK;   // re-run the frame (with the larger buffers)

    ctx->spec_hint = ctx->comm ? ctx->n_merged : c.n_colliding;
    imrcd_frame_stats& st = ctx->stats;
    st.n_pairs = c.n_pairs; st.n_sat_tests = c.n_sat; st.n_combos = c.n_combos; st.n_tri_tests = c.n_tri_tests;
    st.n_hits = c.n_hits; st.n_coplanar_hits = c.n_coplanar; st.n_colliding = c.n_colliding;
    st.n_contact_pairs = c.n_class[0] + c.n_class[16] + c.n_class[32] + c.n_class[48]; st.n_rays = c.n_rays;
    st.traverse_launches = 1; st.total_launches = ctx->pending_launches; st.n_queue_items = c.n_donated; st.n_warp_iterations = c.n_iterations; st.trav_busy_cycles = c.busy_cycles; st.trav_idle_polls = c.idle_polls;
    cudaEventElapsedTime(&st.ms_total, ctx->ev[0], ctx->ev[5]);
    cudaEventElapsedTime(&st.ms_broad, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&st.ms_pair_setup, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&st.ms_traverse, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&st.ms_narrow, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&st.ms_reduce, ctx->ev[4], ctx->ev[5]);
    cudaEventElapsedTime(&st.ms_response, ctx->ev[6], ctx->ev[5]);
    st.n_rays_shot = c.n_rays_kept; st.n_responses = c.n_responses;
    st.n_merged = ctx->comm ? ctx->n_merged : c.n_colliding;
    return IMRCD_OK;
}

// once per context: kernel attributes, persistent grid sizes, initial capacities (all persist across frames; capacities grow on overflow)
static int frame_prepare(imrcd_ctx* ctx) {
    if (ctx->cap_pairs == 0) ctx->cap_pairs = std::max<uint64_t>(1u << 16, 4ull * ctx->n_entries);      // grown (and the frame re-run) when a frame has more pairs
    if (ctx->cap_queue == 0) ctx->cap_queue = ctx->cap_pairs + (1ull << 22);
    if (ctx->cap_combos == 0) ctx->cap_combos = 1ull << 22;
    if (ctx->cap_hits == 0) ctx->cap_hits = 1ull << 20;
    if (ctx->cap_rays == 0) ctx->cap_rays = 1ull << 18;
    if (ctx->cap_lscratch == 0) {
        ctx->cap_lscratch = 16ull << 20;
        const char* ev = getenv("IMRCD_PC_LARGE_MIN");                 // pairs with more hits than this take the grid-wide passes (tuning knob)
        if (ev) ctx->pc_large_min = std::min<uint32_t>((uint32_t)atoi(ev), PC_M_MAX);
    }
    if (ctx->few_flagged_max == 0xffffffffu) {
        const char* ev = getenv("IMRCD_FEW_FLAGGED_MAX");              // 0 = always the sort-and-sweep broad phase (tests run both)
        ctx->few_flagged_max = ev ? std::min<uint32_t>((uint32_t)atoi(ev), FEW_FLAGGED_MAX) : FEW_FLAGGED_MAX;
    }
    if (ctx->narrow_blocks == 0) {
        const int smem = (int)(NT_WARPS * sizeof(NarrowWarp));
        IMR_CUDA(ctx, cudaFuncSetAttribute(k_tritri, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int per_sm = 0;
        IMR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tritri, NT_WARPS * 32, smem));
        if (per_sm < 1) per_sm = 1;
        ctx->narrow_blocks = per_sm * ctx->sm_count;
    }
    { const int rc = imr_traverse_prepare(ctx); if (rc) return rc; }
    if (!ctx->pc_attr_set) {
        const size_t per_hit = 2 * (sizeof(PcSlot) + 8) + 4 * 8 + 12 + 5 * sizeof(uint16_t);
        IMR_CUDA(ctx, cudaFuncSetAttribute(k_pair_contacts_hash<512, PC_M1_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PC_M1_MAX * per_hit)));
        IMR_CUDA(ctx, cudaFuncSetAttribute(k_pair_contacts_hash<1024, PC_M_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PC_M_MAX * per_hit)));
        ctx->pc_attr_set = true;
    }
    return IMRCD_OK;
}

// Every buffer a frame of the current size and capacities needs.  Separate from the enqueue half so that the latter makes no allocation
// and can run under stream capture (imr_frame_enqueue_all replays the captured frame while nothing that shapes it has changed).
int imr_frame_reserve(imrcd_ctx* ctx) {
    cudaStream_t s = ctx->stream;
    const uint32_t n = (uint32_t)ctx->n_entries;
    { const int rc = frame_prepare(ctx); if (rc) return rc; }
    size_t cub_bytes = 0;
    if (n >= 2) {
        cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, ctx->d_keys.as<uint32_t>(), ctx->d_keys2.as<uint32_t>(),
                                        ctx->d_idx.as<uint32_t>(), ctx->d_idx2.as<uint32_t>(), (int)n, 0, 32, s);
        size_t scan_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, ctx->d_flag.as<uint32_t>(), ctx->d_cpos.as<uint32_t>(), (int)(n + 1), s);
        cub_bytes = std::max(cub_bytes, scan_bytes);
    }
    IMR_CUDA(ctx, ctx->d_ctl.reserve(sizeof(FrameCtl), 0, s));
    IMR_CUDA(ctx, ctx->p_ctl.reserve(sizeof(FrameCtl)));
    IMR_CUDA(ctx, ctx->d_epairs.reserve(sizeof(imrcd_entity_pair) * (std::max<uint64_t>(ctx->cap_pairs, ctx->gcap) + 1), 0, s));      // row 0 = header (record count), see imrcd_frame_results_block
    IMR_CUDA(ctx, ctx->d_inv.reserve(64ull * n, 0, s));
    IMR_CUDA(ctx, ctx->d_ext.reserve(24ull * n, 0, s));
    IMR_CUDA(ctx, ctx->d_keys.reserve(4ull * n, 0, s));
    IMR_CUDA(ctx, ctx->d_keys2.reserve(4ull * n, 0, s));
    IMR_CUDA(ctx, ctx->d_idx.reserve(4ull * n, 0, s));
    IMR_CUDA(ctx, ctx->d_idx2.reserve(4ull * n, 0, s));
    IMR_CUDA(ctx, ctx->d_sorted.reserve(sizeof(SweepRec) * (size_t)n, 0, s));
    IMR_CUDA(ctx, ctx->d_sorted_c.reserve(sizeof(SweepRec) * (size_t)std::max<uint32_t>(n, FEW_FLAGGED_MAX), 0, s));
    IMR_CUDA(ctx, ctx->d_flag.reserve(4ull * (n + 1), 0, s));
    IMR_CUDA(ctx, ctx->d_cpos.reserve(4ull * (n + 1), 0, s));
    IMR_CUDA(ctx, ctx->d_wlen.reserve(4ull * (n + 1), 0, s));
    IMR_CUDA(ctx, ctx->d_chunks.reserve(4ull * (n + 1), 0, s));
    IMR_CUDA(ctx, ctx->d_chunkoff.reserve(4ull * (n + 1), 0, s));
    IMR_CUDA(ctx, ctx->d_cubtmp.reserve(cub_bytes, 0, s));
    IMR_CUDA(ctx, ctx->d_pairs.reserve(8ull * ctx->cap_pairs, 0, s));
    IMR_CUDA(ctx, ctx->d_pairrec.reserve(sizeof(PairRec) * ctx->cap_pairs, 0, s));
    IMR_CUDA(ctx, ctx->d_pairacc.reserve(sizeof(PairAcc) * ctx->cap_pairs, 0, s));
    if (16ull * ctx->cap_queue > ctx->d_queue.cap) { IMR_CUDA(ctx, ctx->d_queue.reserve(16ull * ctx->cap_queue, 0, s)); ctx->queue_dirty = ctx->cap_queue; }
    IMR_CUDA(ctx, ctx->d_combos.reserve(16ull * ctx->cap_combos, 0, s));
    IMR_CUDA(ctx, ctx->d_hits.reserve(sizeof(imrcd_tri_hit) * ctx->cap_hits, 0, s));
    IMR_CUDA(ctx, ctx->d_aux.reserve(sizeof(HitAux) * ctx->cap_hits, 0, s));
    IMR_CUDA(ctx, ctx->d_grouped.reserve(4ull * 2 * ctx->cap_hits, 0, s));
    IMR_CUDA(ctx, ctx->d_lscratch.reserve(ctx->cap_lscratch, 0, s));
    IMR_CUDA(ctx, ctx->d_lpref.reserve(8ull * (ctx->cap_pairs + 1), 0, s));
    IMR_CUDA(ctx, ctx->d_lsides.reserve(sizeof(LargeSide) * 2ull * ctx->cap_pairs, 0, s));
    IMR_CUDA(ctx, ctx->d_epair_pair.reserve(4ull * ctx->cap_pairs, 0, s));
    if (ctx->prev_distinct) IMR_CUDA(ctx, ctx->d_rays.reserve(sizeof(RayRec) * ctx->cap_rays, 0, s));
    if (ctx->prev_distinct) IMR_CUDA(ctx, ctx->d_resp.reserve(32ull * ctx->cap_rays, 0, s));
    IMR_CUDA(ctx, ctx->d_lsmall.reserve(4ull * PC_CLASSES * ctx->cap_pairs, 0, s));      // the size-class lists, cap_pairs entries each
    IMR_CUDA(ctx, ctx->p_epairs.reserve(sizeof(imrcd_entity_pair) * std::min<uint64_t>(imr_frame_spec_rows(ctx), ctx->cap_pairs), 0, s));
    if (ctx->comm) { const int rc = imr_comm_reserve(ctx); if (rc) return rc; }
    return IMRCD_OK;
}

static inline uint64_t queue_clear_slots(const imrcd_ctx* ctx) {
    uint64_t r = 1ull << 16;
    while (r < ctx->queue_dirty) r <<= 1;
    return std::min<uint64_t>(r, ctx->cap_queue);
}

// a stage-boundary event: inside a stream capture it has to be recorded as an external node to stay usable for cudaEventElapsedTime
static inline cudaError_t frame_event(imrcd_ctx* ctx, cudaEvent_t ev, cudaStream_t s) {
    return ctx->capturing ? cudaEventRecordWithFlags(ev, s, cudaEventRecordExternal) : cudaEventRecord(ev, s);
}

// First half of a frame: every kernel of it, the control block's way back to the host and a speculative copy of the result rows, all on the
// context's stream; nothing here waits for the device.
int imr_frame_enqueue(imrcd_ctx* ctx) {
    cudaStream_t s = ctx->stream;
    const uint32_t n = (uint32_t)ctx->n_entries;
    { const int rc = frame_prepare(ctx); if (rc) return rc; }
    FrameCtl* ctl = ctx->d_ctl.as<FrameCtl>();
    uint64_t launches = 0;
    IMR_CUDA(ctx, frame_event(ctx, ctx->ev[0], s));
    IMR_CUDA(ctx, cudaMemsetAsync(ctl, 0, sizeof(FrameCtl), s));
    if (n < 2) {
        // a shard can be left with fewer than two entries of a frame that has more: it has no pairs, but it still answers the collective
        for (int k = 1; k <= 6; ++k) IMR_CUDA(ctx, frame_event(ctx, ctx->ev[k], s));
        IMR_CUDA(ctx, cudaMemsetAsync(ctx->d_epairs.p, 0, sizeof(imrcd_entity_pair), s));
        IMR_CUDA(ctx, cudaMemcpyAsync(ctx->p_ctl.p, ctl, sizeof(FrameCtl), cudaMemcpyDeviceToHost, s));
        ctx->pending_launches = 0; ctx->spec_rows_sent = 0;
        return IMRCD_OK;
    }
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, ctx->d_keys.as<uint32_t>(), ctx->d_keys2.as<uint32_t>(),
                                    ctx->d_idx.as<uint32_t>(), ctx->d_idx2.as<uint32_t>(), (int)n, 0, 32, s);
    {
        size_t scan_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, ctx->d_flag.as<uint32_t>(), ctx->d_cpos.as<uint32_t>(), (int)(n + 1), s);
        cub_bytes = std::max(cub_bytes, scan_bytes);
    }
    {
        // contact reduction scratch: per-pair slices are power-of-two padded, so at most 2 x hits slots in total
        if (ctx->prev_distinct) {
        }
        if (ctx->queue_dirty) {                       // clear publication flags left by the previous frame (a power-of-two many: see frame_key)
            IMR_CUDA(ctx, cudaMemsetAsync(ctx->d_queue.p, 0, 16ull * queue_clear_slots(ctx), s));
        }
        // ---- broad ----
        const uint32_t* a_gidx = ctx->shard_n > 1 ? ctx->d_gidx.as<uint32_t>() : nullptr;
        if (ctx->n_flagged_global <= ctx->few_flagged_max) {
            // few flagged entries: every entry against the dense list of the flagged ones, no sort (k_pairs_few_flagged)
            k_entry_prep<<<blocks_for(n, 128), 128, 0, s>>>(n, ctx->d_cur.as<float>(), ctx->d_mesh.as<uint32_t>(), ctx->d_meshes.as<MeshDev>(),
                                                             ctx->d_recs.as<TreeRec>(), ctx->d_inv.as<float>(), ctx->d_ext.as<float>(), nullptr, nullptr,
                                                             ctx->d_cb.as<uint8_t>(), a_gidx, ctx->d_sorted_c.as<SweepRec>(), ctl);
            k_pairs_few_flagged<<<blocks_for(n, 256), 256, 0, s>>>(n, ctx->d_ext.as<float>(), ctx->d_cb.as<uint8_t>(), a_gidx, ctx->d_sorted_c.as<SweepRec>(),
                                                                    ctx->d_pairs.as<uint2>(), ctx->cap_pairs, ctl, ctx->shard_rank, ctx->shard_n);
            launches += 2;
        } else {
        k_entry_prep<<<blocks_for(n, 128), 128, 0, s>>>(n, ctx->d_cur.as<float>(), ctx->d_mesh.as<uint32_t>(), ctx->d_meshes.as<MeshDev>(),
                                                         ctx->d_recs.as<TreeRec>(), ctx->d_inv.as<float>(), ctx->d_ext.as<float>(),
                                                         ctx->d_keys.as<uint32_t>(), ctx->d_idx.as<uint32_t>(), nullptr, nullptr, nullptr, ctl);
        cub::DeviceRadixSort::SortPairs(ctx->d_cubtmp.p, cub_bytes, ctx->d_keys.as<uint32_t>(), ctx->d_keys2.as<uint32_t>(),
                                        ctx->d_idx.as<uint32_t>(), ctx->d_idx2.as<uint32_t>(), (int)n, 0, 32, s);
        k_gather_sorted<<<blocks_for(n, 256), 256, 0, s>>>(n, ctx->d_idx2.as<uint32_t>(), ctx->d_ext.as<float>(), ctx->d_cb.as<uint8_t>(),
                                                            ctx->d_sorted.as<SweepRec>(), ctx->d_flag.as<uint32_t>());
        cub::DeviceScan::ExclusiveSum(ctx->d_cubtmp.p, cub_bytes, ctx->d_flag.as<uint32_t>(), ctx->d_cpos.as<uint32_t>(), (int)(n + 1), s);
        k_compact_flagged<<<blocks_for(n, 256), 256, 0, s>>>(n, ctx->d_sorted.as<SweepRec>(), ctx->d_cpos.as<uint32_t>(), ctx->d_sorted_c.as<SweepRec>());
        k_window<<<blocks_for(n, 256), 256, 0, s>>>(n, ctx->d_sorted.as<SweepRec>(), ctx->d_sorted_c.as<SweepRec>(), ctx->d_cpos.as<uint32_t>(),
                                                     ctx->d_wlen.as<uint32_t>(), ctx->d_chunks.as<uint32_t>());
        cub::DeviceScan::ExclusiveSum(ctx->d_cubtmp.p, cub_bytes, ctx->d_chunks.as<uint32_t>(), ctx->d_chunkoff.as<uint32_t>(), (int)(n + 1), s);
        k_sweep<<<ctx->sm_count * 8, 256, 0, s>>>(n, ctx->d_sorted.as<SweepRec>(), ctx->d_sorted_c.as<SweepRec>(), ctx->d_cpos.as<uint32_t>(),
                                                   ctx->d_wlen.as<uint32_t>(), ctx->d_chunkoff.as<uint32_t>(), ctx->d_pairs.as<uint2>(), ctx->cap_pairs,
                                                   ctl, ctx->shard_rank, ctx->shard_n, a_gidx,
                                                   (ctx->shard_n > 1 && ctx->n_flagged_global == ctx->n_entries_global) ? 1u : 0u);
        launches += 6 + 4 + 2 * 2;   // + radix sort (histogram + onesweep passes, counted as 4) + two decoupled-look-back scans (init + scan)
        }
        IMR_CUDA(ctx, frame_event(ctx, ctx->ev[1], s));
        // ---- pair setup ----
        { const int rc = imr_traverse_queue_init(ctx, ctl); if (rc) return rc; }
        k_pair_setup<<<ctx->sm_count * 8, 128, 0, s>>>(ctl, ctx->cap_pairs, ctx->d_pairs.as<uint2>(), ctx->d_cur.as<float>(), ctx->prev_distinct ? ctx->d_prev.as<float>() : nullptr, ctx->d_inv.as<float>(),
                                                        ctx->d_mesh.as<uint32_t>(), ctx->d_meshes.as<MeshDev>(), ctx->d_pairrec.as<PairRec>(),
                                                        ctx->d_pairacc.as<PairAcc>(), ctx->d_queue.as<WorkItem>(), ctx->cap_queue);
        launches += 2;
        IMR_CUDA(ctx, frame_event(ctx, ctx->ev[2], s));
        // ---- mid ----
        { const int rc = imr_traverse_launch(ctx, ctl); if (rc) return rc; }
        launches += 1;
        IMR_CUDA(ctx, frame_event(ctx, ctx->ev[3], s));
        // ---- narrow ----
        k_tritri<<<ctx->narrow_blocks, NT_WARPS * 32, NT_WARPS * sizeof(NarrowWarp), s>>>(ctl, ctx->d_combos.as<Combo>(), ctx->cap_combos, ctx->d_pairrec.as<PairRec>(),
                                                    ctx->d_tris.as<TriRec>(), ctx->d_hits.as<imrcd_tri_hit>(), ctx->cap_hits,
                                                    ctx->d_pairacc.as<PairAcc>(), ctx->d_aux.as<HitAux>());
        launches += 1;
        IMR_CUDA(ctx, frame_event(ctx, ctx->ev[4], s));
        // ---- reduce ----
        k_hit_lists<<<ctx->sm_count * 4, 256, 0, s>>>(ctl, ctx->cap_pairs, ctx->d_pairacc.as<PairAcc>(),
                                                       ctx->d_lsmall.as<uint32_t>(), ctx->pc_large_min);
        k_group_hits<<<ctx->sm_count * 8, 256, 0, s>>>(ctl, ctx->cap_hits, ctx->d_hits.as<imrcd_tri_hit>(), ctx->d_pairacc.as<PairAcc>(), ctx->d_grouped.as<uint32_t>());
        {
            const size_t per_hit = 2 * (sizeof(PcSlot) + 8) + 4 * 8 + 12 + 5 * sizeof(uint16_t);
            const size_t smem_s = PC_S_MAX * per_hit, smem_m1 = PC_M1_MAX * per_hit, smem_m = PC_M_MAX * per_hit;
            PairAcc* a_acc = ctx->d_pairacc.as<PairAcc>(); const uint32_t* a_grp = ctx->d_grouped.as<uint32_t>();
            const imrcd_tri_hit* a_hits = ctx->d_hits.as<imrcd_tri_hit>(); const HitAux* a_aux = ctx->d_aux.as<HitAux>();
            const PairRec* a_pr = ctx->d_pairrec.as<PairRec>(); const TriRec* a_tris = ctx->d_tris.as<TriRec>(); const uint32_t* a_vid = ctx->d_tri_vid.as<uint32_t>();
            const float* a_nrm = ctx->d_tri_nrm.as<float>(); RayRec* a_rays = ctx->d_rays.as<RayRec>(); unsigned char* a_scr = ctx->d_lscratch.as<unsigned char>();
            const uint32_t* l0 = ctx->d_lsmall.as<uint32_t>(); const uint32_t* l1 = l0 + ctx->cap_pairs; const uint32_t* l2 = l1 + ctx->cap_pairs; const uint32_t* l3 = l2 + ctx->cap_pairs;
            // the size classes are independent: the rarer ones run beside the common one on a second stream, the ones with the largest
            // shared-memory footprint first (a 137-KB block would otherwise wait for the small-class blocks to drain)
            IMR_CUDA(ctx, cudaEventRecord(ctx->ev_fork, s));
            IMR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
            IMR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream3, ctx->ev_fork, 0));
            IMR_CUDA(ctx, cudaStreamWaitEvent(ctx->stream4, ctx->ev_fork, 0));
            k_pair_contacts_hash<1024, PC_M_MAX><<<ctx->sm_count, 1024, smem_m, ctx->stream2>>>(ctl, l2, 2, a_acc, a_grp, a_hits, a_aux, a_pr, a_tris, a_vid, a_nrm, a_rays, ctx->cap_rays);
            {   // large pairs: grid-wide passes (k_large_*), on their own side stream
                cudaStream_t s2 = ctx->stream4;
                unsigned long long* a_pref = ctx->d_lpref.as<unsigned long long>(); LargeSide* a_sides = ctx->d_lsides.as<LargeSide>();
                const unsigned gl = ctx->sm_count * 4;
                k_large_layout<<<1, 1024, 0, s2>>>(ctl, l3, a_acc, a_pref, a_sides, ctx->cap_lscratch);
                k_large_init<<<gl, 256, 0, s2>>>(ctl, l3, a_acc, a_pref, a_scr);
                k_large_mark<<<gl, 256, 0, s2>>>(ctl, l3, a_acc, a_pref, a_scr, a_grp, a_hits, a_aux);
                k_large_hits<<<gl, 256, 0, s2>>>(ctl, l3, a_acc, a_pref, a_scr, a_grp, a_hits, a_aux);
                k_large_unmark<<<gl, 256, 0, s2>>>(ctl, l3, a_acc, a_pref, a_scr);
                k_large_candidates<<<gl, 256, 0, s2>>>(ctl, l3, a_acc, a_pref, a_scr, a_sides, a_vid);
                k_large_alloc<<<ctx->sm_count, 256, 0, s2>>>(ctl, l3, a_acc, a_sides, ctx->cap_rays);
                k_large_rays<<<gl, 256, 0, s2>>>(ctl, l3, a_acc, a_pref, a_scr, a_sides, a_pr, a_tris, a_nrm, a_rays);
            }
            k_pair_contacts_hash<512, PC_M1_MAX><<<ctx->sm_count * 3, 512, smem_m1, ctx->stream3>>>(ctl, l1, 1, a_acc, a_grp, a_hits, a_aux, a_pr, a_tris, a_vid, a_nrm, a_rays, ctx->cap_rays);
            k_pair_contacts_hash<128, PC_S_MAX><<<ctx->sm_count * 8, 128, smem_s, s>>>(ctl, l0, 0, a_acc, a_grp, a_hits, a_aux, a_pr, a_tris, a_vid, a_nrm, a_rays, ctx->cap_rays);
            IMR_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->stream2));
            IMR_CUDA(ctx, cudaEventRecord(ctx->ev_join3, ctx->stream3));
            IMR_CUDA(ctx, cudaEventRecord(ctx->ev_join4, ctx->stream4));
            IMR_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_join, 0));
            IMR_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_join3, 0));
            IMR_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_join4, 0));
        }
        launches += 13;     // lists, group, three per-pair size classes, eight passes over the large pairs
        k_finalize<<<ctx->sm_count * 4, 256, 0, s>>>(ctl, ctx->cap_pairs, ctx->d_pairs.as<uint2>(), ctx->d_pairacc.as<PairAcc>(),
                                                      ctx->d_entity.as<uint32_t>(), ctx->d_cur.as<float>(), ctx->d_inv.as<float>(),
                                                      ctx->d_epairs.as<imrcd_entity_pair>() + 1, ctx->d_epair_pair.as<uint32_t>(),
                                                      ctx->shard_n > 1 ? ctx->d_gidx.as<uint32_t>() : nullptr);
        launches += 2;
        IMR_CUDA(ctx, frame_event(ctx, ctx->ev[6], s));
        { int rc = imr_frame_shoot_device(ctx, ctl, &launches); if (rc != IMRCD_OK) return rc; }
        k_epairs_header<<<1, 1, 0, s>>>(ctl, ctx->d_epairs.as<imrcd_entity_pair>());      // after the last kernel that can raise an overflow bit
        IMR_CUDA(ctx, frame_event(ctx, ctx->ev[5], s));
        IMR_CUDA(ctx, cudaMemcpyAsync(ctx->p_ctl.p, ctl, sizeof(FrameCtl), cudaMemcpyDeviceToHost, s));
        ctx->pending_launches = launches;
        ctx->spec_rows_sent = 0;
        if (!ctx->comm) {            // the records themselves, speculatively (imrcd_frame_fetch copies the rest when the frame has more)
            const uint64_t rows = std::min<uint64_t>(imr_frame_spec_rows(ctx), ctx->cap_pairs);
            IMR_CUDA(ctx, cudaMemcpyAsync(ctx->p_epairs.p, ctx->d_epairs.as<imrcd_entity_pair>() + 1, sizeof(imrcd_entity_pair) * rows, cudaMemcpyDeviceToHost, s));
            ctx->spec_rows_sent = rows;
        }
    }
    return IMRCD_OK;
}

// everything that shapes the frame's launches: sizes, capacities, the buffers' addresses
static uint64_t frame_key(imrcd_ctx* ctx) {
    uint64_t h = 0xcbf29ce484222325ull;
    auto mix = [&](uint64_t v) { h = (h ^ v) * 0x100000001b3ull; };
    mix(ctx->n_entries); mix(ctx->n_entries_global); mix(ctx->n_flagged_global <= ctx->few_flagged_max); mix(ctx->n_flagged_global == ctx->n_entries_global);
    mix(ctx->prev_distinct); mix(ctx->shard_rank); mix(ctx->shard_n); mix(ctx->cap_pairs); mix(ctx->cap_queue); mix(ctx->cap_combos); mix(ctx->cap_hits);
    mix(ctx->cap_rays); mix(ctx->cap_lscratch); mix(ctx->pc_large_min); mix(ctx->queue_dirty ? queue_clear_slots(ctx) : 0); mix(imr_frame_spec_rows(ctx));
    mix(ctx->gcap); mix((uint64_t)ctx->comm); mix(ctx->meshes.size()); mix((uint64_t)ctx->p2p_buf); mix((uint64_t)(ctx->p2p_state + 1)); mix(ctx->p2p_gcap);
    const DevBuf* bufs[] = { &ctx->d_recs, &ctx->d_tris, &ctx->d_tri_nrm, &ctx->d_tri_vid, &ctx->d_meshes, &ctx->d_cur, &ctx->d_prev, &ctx->d_mesh, &ctx->d_cb, &ctx->d_entity, &ctx->d_gidx,
                             &ctx->d_inv, &ctx->d_ext, &ctx->d_keys, &ctx->d_keys2, &ctx->d_idx, &ctx->d_idx2, &ctx->d_sorted, &ctx->d_sorted_c, &ctx->d_flag, &ctx->d_cpos, &ctx->d_wlen,
                             &ctx->d_chunks, &ctx->d_chunkoff, &ctx->d_cubtmp, &ctx->d_pairs, &ctx->d_pairrec, &ctx->d_pairacc, &ctx->d_queue, &ctx->d_combos, &ctx->d_hits, &ctx->d_epairs,
                             &ctx->d_ctl, &ctx->d_aux, &ctx->d_grouped, &ctx->d_lscratch, &ctx->d_lpref, &ctx->d_lsides, &ctx->d_lsmall, &ctx->d_rays, &ctx->d_resp, &ctx->d_epair_pair,
                             &ctx->d_gather };
    for (const DevBuf* b : bufs) mix((uint64_t)b->p);
    mix((uint64_t)ctx->p_ctl.p); mix((uint64_t)ctx->p_epairs.p); mix((uint64_t)ctx->p_gather.p);
    return h | 1ull;
}

static int frame_enqueue_body(imrcd_ctx* ctx) {
    int rc = imr_frame_enqueue(ctx);
    if (rc) return rc;
    if (ctx->comm) {
        rc = imr_comm_allgather(ctx); if (rc) return rc;
        ctx->spec_rows_sent = imr_frame_spec_rows(ctx);
        rc = imr_comm_after_gather(ctx, ctx->spec_rows_sent); if (rc) return rc;
    }
    return IMRCD_OK;
}

// The whole frame (and, with a communicator, its end-of-frame merge) goes onto the stream.  The ~25 launches, memsets and copies of a frame
// are recorded once as a CUDA graph and replayed while nothing that shapes them changes (entry count, capacities, buffer addresses): on a
// frame of a millisecond the host's launch calls are a tenth of the end-to-end time, and with eight ranks a fifth.  IMRCD_GRAPH=0 turns
// it off; the traversal's diagnostic trace runs eagerly.
static int frame_enqueue_all(imrcd_ctx* ctx) {
    int rc = imr_frame_reserve(ctx);
    if (rc) return rc;
    if (ctx->use_graph < 0) { const char* ev = getenv("IMRCD_GRAPH"); ctx->use_graph = (ev && atoi(ev) == 0) ? 0 : 1; if (getenv("IMRCD_TRAV_TRACE")) ctx->use_graph = 0; }
    if (!ctx->use_graph) return frame_enqueue_body(ctx);
    cudaStream_t s = ctx->stream;
    const uint64_t key = frame_key(ctx);
    if (ctx->graph_exec && key == ctx->graph_key) {
        IMR_CUDA(ctx, cudaGraphLaunch(ctx->graph_exec, s));
        ctx->pending_launches = ctx->graph_launches; ctx->spec_rows_sent = ctx->graph_spec_rows;
        return IMRCD_OK;
    }
    // a frame shape is captured the second time in a row it is seen: one-off frames are not worth a capture, and the first collective on a
    // communicator must run eagerly (NCCL sets its connections up inside it, which a capture cannot hold)
    if (key != ctx->graph_seen_key) { ctx->graph_seen_key = key; return frame_enqueue_body(ctx); }
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; ctx->graph_key = 0; }
    IMR_CUDA(ctx, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    rc = frame_enqueue_body(ctx);
    ctx->capturing = false;
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
    if (e != cudaSuccess || !graph) { ctx->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e); cudaGetLastError(); return IMRCD_E_CUDA; }
    const cudaError_t ei = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) { ctx->graph_exec = nullptr; ctx->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ei); return IMRCD_E_CUDA; }
    ctx->graph_key = key; ctx->graph_launches = ctx->pending_launches; ctx->graph_spec_rows = ctx->spec_rows_sent;
    IMR_CUDA(ctx, cudaGraphLaunch(ctx->graph_exec, s));
    return IMRCD_OK;
}

void imr_frame_begin(imrcd_ctx* ctx) {
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    memset(&ctx->ctl_host, 0, sizeof(ctx->ctl_host));
    ctx->stats.n_entries = ctx->n_entries_global; ctx->stats.n_entries_local = ctx->n_entries;
    ctx->hits_fetched = false; ctx->merged_valid = false; ctx->n_merged = 0; ctx->spec_rows_sent = 0;
}

int imr_frame_run_device(imrcd_ctx* ctx) {
    imr_frame_begin(ctx);
    if (ctx->n_entries_global < 2) {                              // CollisionDetection.cpp:40 (every rank of a sharded frame sees the same count)
        if (ctx->d_epairs.p) IMR_CUDA(ctx, cudaMemsetAsync(ctx->d_epairs.p, 0, sizeof(imrcd_entity_pair), ctx->stream));      // the result block says "no records"
        ctx->merged_valid = true;
        return IMRCD_OK;
    }
    for (int attempt = 0; attempt < 10; ++attempt) {
        int rc = frame_enqueue_all(ctx);
        if (rc) return rc;
        if (ctx->enqueue_only) return IMRCD_OK;                   // imrcd_frame_run_async: imrcd_frame_finish does the rest
        bool retry = false;
        rc = imr_frame_complete(ctx, &retry);
        if (rc) return rc;
        if (!retry) return IMRCD_OK;
    }
    ctx->err = "frame buffers could not be grown enough (10 attempts)";
    return IMRCD_E_CAPACITY;
}

// imrcd_frame_finish: 0 = the enqueued frame stands, 1 = a buffer overflowed (on this rank or, with a communicator, on any) and the frame
// was run again, synchronously
int imr_frame_finish_device(imrcd_ctx* ctx) {
    if (ctx->n_entries_global < 2) return IMRCD_OK;
    bool retry = false;
    int rc = imr_frame_complete(ctx, &retry);
    if (rc != IMRCD_OK) return rc;
    if (!retry) return IMRCD_OK;
    ctx->enqueue_only = false;
    rc = imr_frame_run_device(ctx);
    return rc != IMRCD_OK ? rc : 1;
}
