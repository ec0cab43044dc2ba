// imrcd_traverse.cu -- the mid phase of a frame: the dual-tree descent of OBBtree::IntersectOBBtreesRecursive (IMR/src/Geometry/OBBtree.cpp:396-477)
// with the 15-axis SAT on parallelepipeds (IMR/src/Geometry/Paralgram.cpp:17-173) as a persistent work-queue kernel.
#include "imrcd_internal.cuh"
#include <algorithm>
#include <cstdlib>

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) { return *((const volatile unsigned long long*)p); }
__device__ __forceinline__ long long ld_volatile_s64(const long long* p) { return *((const volatile long long*)p); }
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// The roots are dealt evenly: warp w of the traversal starts on slots [w * k, w * k + k), k = ceil(roots / warps) <= 32 (its first "ticket"
// needs no atomic); later tickets come from q_head, which therefore starts behind the fixed ones.
__device__ __forceinline__ uint32_t trav_first_ticket(unsigned long long n_roots, uint32_t n_warps) {
    const unsigned long long k = (n_roots + n_warps - 1) / n_warps;
    return k > 32ull ? 32u : (k < 1ull ? 1u : (uint32_t)k);
}
__global__ void k_queue_init(FrameCtl* ctl, unsigned long long cap_pairs, unsigned long long cap_queue, uint32_t trav_warps) {
    unsigned long long n = ctl->n_pairs < cap_pairs ? ctl->n_pairs : cap_pairs;
    if (n > cap_queue) { n = cap_queue; atomicOr(&ctl->overflow, (unsigned)OVF_QUEUE); }
    ctl->q_head = (unsigned long long)trav_first_ticket(n, trav_warps) * trav_warps; ctl->q_tail = n; ctl->pending = (long long)n; ctl->n_roots = n;
}

// ------------------------------------------------------------------------------------------
// mid phase: persistent work-queue dual-tree traversal
//
// One persistent grid (as many CTAs as are co-resident).  Every warp owns a deque of work items in shared
// memory and runs the descent depth-first on it, 32 SAT visits per iteration (one per lane).  Load balancing
// goes through ONE global linear queue used as a ticket rendezvous:
//   * a warp whose deque is empty takes a ticket for 32 consecutive slots (one atomicAdd on q_head, no CAS
//     retry loop) and then polls only the publication flags of ITS OWN slots - no shared hot word;
//   * a warp with more than 32 items (more than it can start on next iteration) looks at q_head > q_tail
//     ("somebody is waiting on unfilled slots") and, if so, moves its oldest items - the ones closest to the
//     roots, i.e. the largest subtrees - to the slots at q_tail (one atomicAdd, then payload, fence, flag);
//   * `pending` counts alive items (queue + deques); warps add their net production lazily (when they go idle
//     or publish), and everybody leaves when it reads 0.
// Producers never wait, consumers wait only on slots that a producer has already reserved or will never
// fill once pending == 0, so the scheme cannot deadlock.
// ------------------------------------------------------------------------------------------
#define TRAV_WARPS 4
#define STK_CAP 256u          // per-warp deque capacity (power of two)
#define STK_MASK (STK_CAP - 1u)

struct TravStack { uint32_t pair[TRAV_WARPS][STK_CAP]; uint32_t a[TRAV_WARPS][STK_CAP]; uint32_t b[TRAV_WARPS][STK_CAP]; };

// Move k (<= 32) oldest items of this warp's deque to the global queue.  The caller has flushed `delta`
// so that `pending` already counts them.
__device__ __forceinline__ void trav_donate(TravStack& st, uint32_t warp, uint32_t lane, uint32_t& bot, uint32_t k,
                                            FrameCtl* ctl, WorkItem* queue, unsigned long long cap_queue) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&ctl->q_tail, (unsigned long long)k);
    base = __shfl_sync(FULL_MASK, base, 0);
    bool dropped = false;
    if (lane < k) {
        uint32_t s = (bot + lane) & STK_MASK;
        unsigned long long slot = base + lane;
        if (slot < cap_queue) {
            uint4 it = make_uint4(st.pair[warp][s], st.a[warp][s], st.b[warp][s], 0u);
            __stcg(&queue[slot], it);
            __threadfence();
            *((volatile uint32_t*)&queue[slot].w) = 1u;       // publish
        } else dropped = true;
    }
    uint32_t dm = __ballot_sync(FULL_MASK, dropped);
    if (lane == 0) {
        if (dm) {                                              // queue full: the frame will be re-run with a larger queue
            atomicOr(&ctl->overflow, (unsigned)OVF_QUEUE);
            atomicAdd((unsigned long long*)&ctl->pending, (unsigned long long)(-(long long)__popc(dm)));
        }
        atomicAdd(&ctl->n_donated, (unsigned long long)k);
    }
    bot += k;
    __syncwarp();
}

template <int STRAIGHT, int MIN_BLOCKS>
__global__ void __launch_bounds__(TRAV_WARPS * 32, MIN_BLOCKS)
k_traverse(FrameCtl* ctl, const PairRec* __restrict__ pairrec, const TreeRec* __restrict__ recs,
           WorkItem* queue, unsigned long long cap_queue, Combo* __restrict__ combos, unsigned long long cap_combos,
           uint32_t keep_items, uint32_t backoff_max, uint4* __restrict__ trace, uint32_t trace_cap, uint32_t coop) {
    __shared__ TravStack st;
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t top = 0, bot = 0;          // deque: items live in [bot, top); a and b are ARENA record indices (no base to wait for)
    int delta = 0;                      // alive-item change not yet added to ctl->pending
    // ticket: this warp consumes global slots own_base + lane for set bits of own_mask.  The first one is fixed by the warp's id.
    const uint32_t k_first = trav_first_ticket(ctl->n_roots, n_warps);
    unsigned long long own_base = (unsigned long long)gwarp * k_first;
    uint32_t own_mask = k_first >= 32u ? FULL_MASK : ((1u << k_first) - 1u);
    unsigned long long my_sat = 0, my_tri = 0, my_iter = 0, my_busy = 0, my_polls = 0;
    bool finished = false;

    while (!finished) {
        uint32_t cnt = top - bot;
        if (cnt == 0) {
            // ---- refill: wait on this warp's own global slots; what arrives goes on the deque ----
            if (lane == 0 && delta != 0) atomicAdd((unsigned long long*)&ctl->pending, (unsigned long long)(long long)delta);
            delta = 0;
            uint32_t backoff = 32;
            for (;;) {
                if (own_mask == 0u) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(&ctl->q_head, 32ull);
                    own_base = __shfl_sync(FULL_MASK, base, 0);
                    own_mask = FULL_MASK;
                }
                const unsigned long long slot = own_base + lane;
                uint32_t f = 0;
                if (((own_mask >> lane) & 1u) && slot < cap_queue) f = *((volatile uint32_t*)&queue[slot].w);
                const uint32_t ready = __ballot_sync(FULL_MASK, f != 0u);
                if (ready) {
                    if (f) {
                        __threadfence();
                        const uint4 it = __ldcg(&queue[slot]);
                        const uint32_t s = (top + (uint32_t)__popc(ready & lt_mask)) & STK_MASK;
                        st.pair[warp][s] = it.x; st.a[warp][s] = it.y; st.b[warp][s] = it.z;
                    }
                    top += (uint32_t)__popc(ready);
                    own_mask &= ~ready;
                    __syncwarp();
                    break;
                }
                int done = 0;
                if (lane == 0) done = (ld_volatile_s64(&ctl->pending) == 0) ? 1 : 0;
                done = __shfl_sync(FULL_MASK, done, 0);
                if (done) { finished = true; break; }
                ++my_polls;
                __nanosleep(backoff);
                if (backoff < backoff_max) backoff <<= 1;
            }
            if (finished) break;
            cnt = top - bot;
        }
        // ---- this iteration's node pairs: the newest ones (depth first).  A warp with fewer pairs than lanes deals the 15 axes of each
        //      to a group of g lanes (box_sat_part): the visit's latency, which is what a ramp or a tail of the frame waits for, drops ----
        const uint32_t take = cnt < 32u ? cnt : 32u;
        const uint32_t g_log = (coop && take <= 4u) ? (take > 2u ? 3u : 4u) : 0u;       // measured: groups of 2 or 4 lanes do not pay (the indexed axis costs more than it saves)
        const uint32_t item = lane >> g_log, sub = lane & ((1u << g_log) - 1u);
        const bool have = item < take, leader = have && sub == 0u;
        uint32_t ip = 0, ia = 0, ib = 0;
        if (have) {
            const uint32_t s = (top - 1u - item) & STK_MASK;
            ip = st.pair[warp][s]; ia = st.a[warp][s]; ib = st.b[warp][s];
        }
        top -= take;
        __syncwarp();
        const long long t_begin = clock64();
        uint32_t trace_t0 = 0;
        if (trace) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); trace_t0 = (uint32_t)t; }
        const uint32_t trace_slot = (uint32_t)my_iter;
        ++my_iter;

        // ---- one SAT visit per lane group (IntersectOBBtreesRecursive, OBBtree.cpp:414-477) ----
        bool push = false, emit = false, ok = false, leafA = false, leafB = false;
        uint32_t c0a = 0, c0b = 0, c1a = 0, c1b = 0, childA = 0, childB = 0, cnt_ab = 0;
        float surf_a = 0.f, surf_b = 0.f;
        if (have) {
            // the three records are independent loads (the item carries arena indices): one memory latency per visit, not two
            const float4* pp = reinterpret_cast<const float4*>(pairrec + ip);
            const float4* ra = reinterpret_cast<const float4*>(recs + ia);
            const float4* rb = reinterpret_cast<const float4*>(recs + ib);
            Rel rel; rel.r0 = __ldg(pp); rel.r1 = __ldg(pp + 1); rel.r2 = __ldg(pp + 2);
            const float4 a0 = __ldg(ra), a1 = __ldg(ra + 1), a2 = __ldg(ra + 2), a3 = __ldg(ra + 3);
            const float4 b0 = __ldg(rb), b1 = __ldg(rb + 1), b2 = __ldg(rb + 2), b3 = __ldg(rb + 3);
            const uint2 bases = __ldg(reinterpret_cast<const uint2*>(pp + 3));
            leafA = __float_as_uint(a3.w) != 0u; leafB = __float_as_uint(b3.w) != 0u;
            childA = __float_as_uint(a3.y); childB = __float_as_uint(b3.y);
            cnt_ab = __float_as_uint(a3.z) | (__float_as_uint(b3.z) << 16);
            surf_a = a3.x;
            const Box first = unpack_box(a0, a1, a2);
            const Box second = box_transform(rel, unpack_box(b0, b1, b2));        // :420
            if (g_log == 0u) ok = box_sat_t<STRAIGHT>(first, second);             // :422
            else ok = box_sat_part(first, second, (int)sub, 1 << g_log);
            if (!leafA && !leafB) surf_b = box_surface(second);                   // for :426
            if (!leafA) childA += bases.x;                                        // arena indices of the children
            if (!leafB) childB += bases.y;
        }
        if (g_log != 0u) {                                                        // the verdict of a group: every lane's axes overlap
            const uint32_t okm = __ballot_sync(FULL_MASK, ok);
            const uint32_t gm = ((1u << (1u << g_log)) - 1u) << (item << g_log);
            ok = have && (okm & gm) == gm;
        }
        if (leader) {
            ++my_sat;
            if (ok) {
                if (leafA && leafB) {
                    emit = true;                                                   // :473
                    my_tri += (unsigned long long)(cnt_ab & 0xffffu) * (cnt_ab >> 16);
                } else {
                    bool descend_first;
                    if (!leafA && !leafB) descend_first = surf_a >= surf_b;        // :426 (TreeRec caches first.GetSurface())
                    else descend_first = !leafA;
                    push = true;
                    if (descend_first) { c0a = childA; c1a = childA + 1u; c0b = ib; c1b = ib; }
                    else { c0a = ia; c1a = ia; c0b = childB; c1b = childB + 1u; }
                }
            }
        }
        const uint32_t have_m = __ballot_sync(FULL_MASK, leader);
        const uint32_t push_m = __ballot_sync(FULL_MASK, push);
        const uint32_t emit_m = __ballot_sync(FULL_MASK, emit);
        const uint32_t total = 2u * (uint32_t)__popc(push_m);
        delta += (int)total - (int)__popc(have_m);

        // ---- leaf combos: warp-aggregated append; the atomic goes out now, its result is used after the deque work ----
        unsigned long long combo_base = 0;
        if (emit_m && lane == 0) combo_base = atomicAdd(&ctl->n_combos, (unsigned long long)__popc(emit_m));

        // ---- make room, then push the children on the warp's deque ----
        cnt = top - bot;
        if (cnt + total > STK_CAP) {
            if (lane == 0 && delta != 0) atomicAdd((unsigned long long*)&ctl->pending, (unsigned long long)(long long)delta);
            delta = 0;
            trav_donate(st, warp, lane, bot, 32u, ctl, queue, cap_queue);
            trav_donate(st, warp, lane, bot, 32u, ctl, queue, cap_queue);
        }
        if (push) {
            const uint32_t s0 = (top + 2u * (uint32_t)__popc(push_m & lt_mask)) & STK_MASK, s1 = (s0 + 1u) & STK_MASK;
            st.pair[warp][s0] = ip; st.a[warp][s0] = c0a; st.b[warp][s0] = c0b;
            st.pair[warp][s1] = ip; st.a[warp][s1] = c1a; st.b[warp][s1] = c1b;
        }
        top += total;
        __syncwarp();

        // ---- feed waiting warps with what this warp cannot start on in its next iteration ----
        cnt = top - bot;
        uint32_t give = 0;
        if (cnt > keep_items && lane == 0) {           // only a warp with a surplus looks at the shared control words
            const unsigned long long qh = ld_volatile_u64(&ctl->q_head), qt = ld_volatile_u64(&ctl->q_tail);
            if (qh > qt) {
                const unsigned long long want = qh - qt;
                give = cnt - keep_items;
                if (give > 64u) give = 64u;
                if ((unsigned long long)give > want) give = (uint32_t)want;
            }
        }
        give = __shfl_sync(FULL_MASK, give, 0);
        if (emit_m) {
            combo_base = __shfl_sync(FULL_MASK, combo_base, 0);
            if (emit) {
                const unsigned long long slot = combo_base + __popc(emit_m & lt_mask);
                if (slot < cap_combos) combos[slot] = make_uint4(ip, childA, childB, cnt_ab);
                else atomicOr(&ctl->overflow, (unsigned)OVF_COMBOS);
            }
        }
        my_busy += (unsigned long long)(clock64() - t_begin);
        if (trace && lane == 0 && trace_slot < trace_cap) {            // diagnostic timeline (IMRCD_TRAV_TRACE): begin, end (ns), lanes with an item, deque size after
            unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            trace[(size_t)gwarp * trace_cap + trace_slot] = make_uint4(trace_t0, (uint32_t)t1, (uint32_t)__popc(have_m), cnt);
        }
        if (give) {
            if (lane == 0 && delta != 0) atomicAdd((unsigned long long*)&ctl->pending, (unsigned long long)(long long)delta);
            delta = 0;
            while (give) {
                const uint32_t k = give < 32u ? give : 32u;
                trav_donate(st, warp, lane, bot, k, ctl, queue, cap_queue);
                give -= k;
            }
        }
    }

    // ---- statistics ----
    for (int o = 16; o > 0; o >>= 1) {
        my_sat += __shfl_down_sync(FULL_MASK, my_sat, o);
        my_tri += __shfl_down_sync(FULL_MASK, my_tri, o);
    }
    if (lane == 0) {
        atomicAdd(&ctl->n_iterations, my_iter); atomicAdd(&ctl->busy_cycles, my_busy); atomicAdd(&ctl->idle_polls, my_polls);
        if (my_sat) atomicAdd(&ctl->n_sat, my_sat);
        if (my_tri) atomicAdd(&ctl->n_tri_tests, my_tri);
    }
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
// IMRCD_TRAV_VARIANT picks the SAT form (straight-line / per-axis exits / one exit after the face normals) and the occupancy target; the
// default was chosen on C3 and C2 (0.68 / 1.63 ms).  A two-phase form of the kernel (face-normal axes for 32 node pairs, survivors parked
// in shared memory, edge-edge axes for 32 survivors at once) was measured in round 2 and dropped: a warp's wavefront is ~32 node pairs wide
// (more is donated to starving warps), so the second phase ran half empty exactly like the divergent lanes it was meant to remove, and
// letting warps hoard 64+ pairs to fill it starved the others (C3: 0.74-0.80 ms against 0.68; profiles/r2_traverse_experiments.txt).
int imr_traverse_prepare(imrcd_ctx* ctx) {
    if (ctx->trav_blocks != 0) return IMRCD_OK;
    int per_sm = 0;
    const char* ev = getenv("IMRCD_TRAV_VARIANT");
    ctx->trav_variant = ev ? atoi(ev) : 0;
    switch (ctx->trav_variant) {
        case 1: ctx->trav_fn = (const void*)k_traverse<1, 6>; break;
        case 2: ctx->trav_fn = (const void*)k_traverse<1, 8>; break;
        case 3: ctx->trav_fn = (const void*)k_traverse<0, 8>; break;
        case 4: ctx->trav_fn = (const void*)k_traverse<1, 4>; break;
        case 5: ctx->trav_fn = (const void*)k_traverse<2, 6>; break;
        case 6: ctx->trav_fn = (const void*)k_traverse<0, 6>; break;
        case 7: ctx->trav_fn = (const void*)k_traverse<2, 8>; break;
        default: ctx->trav_fn = (const void*)k_traverse<2, 5>; break;      // face axes straight-line, one exit, edge axes straight-line; 5 blocks/SM (C3 0.74, C2 1.62 ms against 0.76 / 1.73 for <0, 6>)
    }
    IMR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ctx->trav_fn, TRAV_WARPS * 32, 0));
    if (per_sm < 1) per_sm = 1;
    if (getenv("IMRCD_TRAV_BLOCKS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(getenv("IMRCD_TRAV_BLOCKS_PER_SM"))));      // probe: a smaller persistent grid
    ctx->trav_blocks = per_sm * ctx->sm_count;
    // a quarter of a frame or less (4+ ranks): the frontier is too thin for the whole grid, and fewer warps polling for it finish sooner
    // (4 instead of 5 blocks per SM: 0.195 -> 0.185 ms on an eighth of C3, +3 % on the whole frame; profiles/r2_traverse_experiments.txt)
    ctx->trav_blocks_shard = std::min(per_sm, 4) * ctx->sm_count;
    return IMRCD_OK;
}
static inline int trav_grid(const imrcd_ctx* ctx) { return ctx->shard_n >= 4 ? ctx->trav_blocks_shard : ctx->trav_blocks; }

int imr_traverse_queue_init(imrcd_ctx* ctx, FrameCtl* ctl) {
    k_queue_init<<<1, 1, 0, ctx->stream>>>(ctl, ctx->cap_pairs, ctx->cap_queue, (uint32_t)(trav_grid(ctx) * TRAV_WARPS));
    return IMRCD_OK;
}

int imr_traverse_launch(imrcd_ctx* ctx, FrameCtl* ctl) {
    cudaStream_t s = ctx->stream;
    const PairRec* a_pairrec = ctx->d_pairrec.as<PairRec>(); const TreeRec* a_recs = ctx->d_recs.as<TreeRec>();
    WorkItem* a_queue = ctx->d_queue.as<WorkItem>(); Combo* a_combos = ctx->d_combos.as<Combo>();
    static uint32_t keep_items = getenv("IMRCD_TRAV_KEEP") ? (uint32_t)atoi(getenv("IMRCD_TRAV_KEEP")) : 32u;
    static uint32_t backoff_max = getenv("IMRCD_TRAV_BACKOFF") ? (uint32_t)atoi(getenv("IMRCD_TRAV_BACKOFF")) : 1024u;
    static uint32_t coop = getenv("IMRCD_TRAV_COOP") ? (uint32_t)atoi(getenv("IMRCD_TRAV_COOP")) : 1u;
    uint4* a_trace = nullptr; uint32_t trace_cap = 0;
    if (getenv("IMRCD_TRAV_TRACE")) {                            // diagnostic: per-warp iteration timeline, dumped by imrcd_debug_trav_trace
        trace_cap = 256;
        IMR_CUDA(ctx, ctx->d_trace.reserve(16ull * trace_cap * ctx->trav_blocks * TRAV_WARPS, 0, s));
        IMR_CUDA(ctx, cudaMemsetAsync(ctx->d_trace.p, 0, 16ull * trace_cap * ctx->trav_blocks * TRAV_WARPS, s));
        a_trace = ctx->d_trace.as<uint4>();
    }
    void* targs[] = { &ctl, &a_pairrec, &a_recs, &a_queue, &ctx->cap_queue, &a_combos, &ctx->cap_combos, &keep_items, &backoff_max, &a_trace, &trace_cap, &coop };
    IMR_CUDA(ctx, cudaLaunchKernel(ctx->trav_fn, dim3(trav_grid(ctx)), dim3(TRAV_WARPS * 32), targs, 0, s));
    return IMRCD_OK;
}

// diagnostic (IMRCD_TRAV_TRACE=1): the traversal's per-warp iteration timeline of the last frame, 256 x uint4 per warp
extern "C" int imrcd_debug_trav_trace(imrcd_ctx* ctx, uint32_t* out, uint64_t cap_u32, uint32_t* n_warps) {
    if (!ctx || !n_warps) return IMRCD_E_ARG;
    *n_warps = (uint32_t)(ctx->trav_blocks * TRAV_WARPS);
    const uint64_t need = 4ull * 256 * *n_warps;
    if (!out || cap_u32 < need || !ctx->d_trace.p) return IMRCD_OK;
    cudaSetDevice(ctx->device);
    IMR_CUDA(ctx, cudaMemcpy(out, ctx->d_trace.p, 4 * need, cudaMemcpyDeviceToHost));
    return IMRCD_OK;
}
