// imrcd_frame.cu -- the per-frame collision pipeline on the device:
//   broad  : k_entry_prep -> radix sort on U-min -> k_sweep           (SweepAndPrune.cpp:15-88)
//   setup  : k_pair_setup  rel = inverse(first.M) * second.M           (OBBtreesCollision.cpp:15)
//   mid    : imrcd_traverse.cu                                         (OBBtree.cpp:396-477, Paralgram.cpp:17-173)
//   narrow : imrcd_narrow.cu    leaf x leaf triangle tests             (CreateUncollideRays.cpp:74-115, Triangle.cpp:866-1002)
//   reduce : imrcd_contacts.cu  contact reduction                      (CreateUncollideRays.cpp:13-58,117-198)
//            k_finalize        colliding entity pairs                  (CollisionDetection.cpp:60-67)
#include "imrcd_frame.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cstring>
#include <cstdio>
#include <cstdlib>

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
    // High-water mark, with headroom: how many slots the next frame clears is part of what a captured frame is made of (frame_key), and
    // q_head depends on how many tickets idle warps happened to take - a per-frame value made the key flip between two powers of two
    // from frame to frame (N = 4 on C3: 200-300 k slots around 262,144), and a frame whose key changes is not replayed but re-captured.
    {
        const uint64_t used = std::min<uint64_t>(std::max(c.q_tail, c.q_head), ctx->cap_queue);
        ctx->queue_dirty = std::max<uint64_t>(ctx->queue_dirty, std::min<uint64_t>(used + used / 2, ctx->cap_queue));
    }
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
    if (ctx->comm && ctx->ev_merge[0] && getenv("IMRCD_MERGE_DEBUG")) {
        float a = 0.f, b = 0.f, t = 0.f;
        cudaEventElapsedTime(&a, ctx->ev[5], ctx->ev_merge[0]); cudaEventElapsedTime(&b, ctx->ev_merge[0], ctx->ev_merge[2]); cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev_merge[2]);
        fprintf(stderr, "[imrcd rank %u] frame %.3f ms | header -> push done %.3f | wait + compact + D2H %.3f | whole %.3f\n", ctx->comm_rank, st.ms_total, a, b, t);
    }
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
    { const int rc = imr_narrow_prepare(ctx); if (rc) return rc; }
    { const int rc = imr_traverse_prepare(ctx); if (rc) return rc; }
    { const int rc = imr_contacts_prepare(ctx); if (rc) return rc; }
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
        { const int rc = imr_narrow_launch(ctx, ctl); if (rc) return rc; }
        launches += 1;
        IMR_CUDA(ctx, frame_event(ctx, ctx->ev[4], s));
        // ---- reduce ----
        { const int rc = imr_contacts_enqueue(ctx, ctl); if (rc) return rc; }
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
        static const bool dbg = getenv("IMRCD_MERGE_DEBUG") != nullptr;      // stage times of the merge, printed by imr_frame_complete
        if (dbg && !ctx->ev_merge[0]) for (auto& e : ctx->ev_merge) IMR_CUDA(ctx, cudaEventCreate(&e));
        rc = imr_comm_allgather(ctx); if (rc) return rc;
        if (dbg) IMR_CUDA(ctx, frame_event(ctx, ctx->ev_merge[0], ctx->stream));
        ctx->spec_rows_sent = imr_frame_spec_rows(ctx);
        rc = imr_comm_after_gather(ctx, ctx->spec_rows_sent); if (rc) return rc;
        if (dbg) IMR_CUDA(ctx, frame_event(ctx, ctx->ev_merge[2], ctx->stream));
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
