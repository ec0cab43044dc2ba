// imrcd_narrow.cu -- the narrow phase of a frame: the loops of CreateUncollideRays.cpp:74-115 over the leaf combos with
// tri_tri_intersect_with_isectline (IMR/src/Geometry/Triangle.cpp:866-1002) cut into three dense passes per warp tile.
#include "imrcd_frame.cuh"
#include <algorithm>

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

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

int imr_narrow_prepare(imrcd_ctx* ctx) {
    if (ctx->narrow_blocks != 0) return IMRCD_OK;
    const int smem = (int)(NT_WARPS * sizeof(NarrowWarp));
    IMR_CUDA(ctx, cudaFuncSetAttribute(k_tritri, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 0;
    IMR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tritri, NT_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    ctx->narrow_blocks = per_sm * ctx->sm_count;
    return IMRCD_OK;
}

int imr_narrow_launch(imrcd_ctx* ctx, FrameCtl* ctl) {
    k_tritri<<<ctx->narrow_blocks, NT_WARPS * 32, NT_WARPS * sizeof(NarrowWarp), ctx->stream>>>(ctl, ctx->d_combos.as<Combo>(), ctx->cap_combos, ctx->d_pairrec.as<PairRec>(),
                                                ctx->d_tris.as<TriRec>(), ctx->d_hits.as<imrcd_tri_hit>(), ctx->cap_hits,
                                                ctx->d_pairacc.as<PairAcc>(), ctx->d_aux.as<HitAux>());
    IMR_CUDA(ctx, cudaGetLastError());
    return IMRCD_OK;
}
