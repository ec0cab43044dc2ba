// imrcd_contacts.cu -- the contact reduction of a frame (CreateUncollideRays.cpp:13-58,117-198): the hits grouped per entity pair, folded into
// TriangleCandidateRays per triangle and turned into rays and contact points.
#include "imrcd_frame.cuh"
#include <algorithm>

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

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

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
int imr_contacts_prepare(imrcd_ctx* ctx) {
    if (ctx->pc_attr_set) return IMRCD_OK;
    const size_t per_hit = 2 * (sizeof(PcSlot) + 8) + 4 * 8 + 12 + 5 * sizeof(uint16_t);
    IMR_CUDA(ctx, cudaFuncSetAttribute(k_pair_contacts_hash<512, PC_M1_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PC_M1_MAX * per_hit)));
    IMR_CUDA(ctx, cudaFuncSetAttribute(k_pair_contacts_hash<1024, PC_M_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PC_M_MAX * per_hit)));
    ctx->pc_attr_set = true;
    return IMRCD_OK;
}

// 13 launches: lists, group, three per-pair size classes, eight passes over the large pairs
int imr_contacts_enqueue(imrcd_ctx* ctx, FrameCtl* ctl) {
    cudaStream_t s = ctx->stream;
    {
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
    }
    IMR_CUDA(ctx, cudaGetLastError());
    return IMRCD_OK;
}
