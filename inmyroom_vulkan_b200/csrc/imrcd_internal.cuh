// imrcd_internal.cuh -- shared declarations of libimrcd.so (context, HBM layouts, helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <cstring>
#include <vector>
#include "../../include/imrcd.h"
#include "imrcd_math.cuh"

// ------------------------------------------------------------------------------------------
// HBM layouts
// ------------------------------------------------------------------------------------------
// Tree vertex record, 64 B = 4 x float4, siblings adjacent (children of one node share a 128-B line):
//   q0 = (c.x c.y c.z u.x)  q1 = (u.y u.z v.x v.y)  q2 = (v.z w.x w.y w.z)
//   q3 = (surface, child_or_tri_off [u32 bits], tri_cnt [u32 bits], kind [u32 bits: 0 inner, 1 leaf])
// Record 0 of a mesh is its root (OBBtree::root_obb, OBBtree.h:129); record 1 is padding;
// an inner record's children are records `child` (left) and `child+1` (right).
// `surface` caches Paralgram::GetSurface() of the untransformed box (Paralgram.cpp:203-210).
struct __align__(16) TreeRec { float4 q0, q1, q2, q3; };

// Triangle, 64 B = 4 x float4 in LEAF order (OBBtree.cpp:207-214):
//   t0 = (p0.xyz, orig_index bits)  t1 = (p1.xyz, 0)  t2 = (p2.xyz, 0)
//   t3 = (N.xyz, d): the triangle's plane exactly as tri_tri_intersect_with_isectline computes it for its FIRST argument
//        (N = (p1-p0) x (p2-p0), d = -N.p0, Triangle.cpp:877-882) -- a pure function of the model-space triangle, hoisted
//        out of the per-pair test (the first entity's triangles are never transformed, CreateUncollideRays.cpp:82-86).
struct __align__(16) TriRec { float4 t0, t1, t2, t3; };

struct MeshDev { uint32_t rec_base, tri_base, n_rec, n_tri; };

// Per broad-phase pair, 64 B: rel = inverse(first.M) * second.M stored by rows (imrcd_math.cuh Rel)
// plus the two meshes' bases so the hot kernels do not chase entry -> mesh -> base.
struct __align__(16) PairRec { float4 r0, r1, r2; uint32_t recA, recB, triA, triB; };

// Work item of the traversal queue: test record a of first's tree against record b of second's tree.
// .w is the publication flag of a global-queue slot (0 = not yet written).
typedef uint4 WorkItem;   // (pair, a, b, ready)

// Leaf x leaf candidate (OBBtreesIntersectInfo::CandidateTriangleRangeCombination, OBBtree.h:9-18)
typedef uint4 Combo;      // (pair, offA, offB, cntA | cntB << 16), offsets relative to the mesh

// Per broad-phase pair accumulators, 112 B.  sum_* / rays_* are the ray origins of CreateUncollideRays.cpp:131-178 summed per side
// (first's model space) and force is average_force_responses of ShootUncollideRays.cpp:26,42,60, all in FP64 so that the order of the
// atomic adds does not show in the FP32 result.  ray_off_*: the pair's slices of the frame's ray array (only pairs that moved keep rays).
struct __align__(16) PairAcc { uint32_t n_hits, flags, rays_a, rays_b, cursor, off, ray_off_a, ray_off_b; double sum_a[3], sum_b[3], force[3]; uint32_t n_resp, pad; };
enum { PAIR_COLLIDING = 1u, PAIR_MOVED = 2u };

// Side record of one hit for the contact reduction (the 40-B imrcd_tri_hit keeps source, target and weight), 12 B:
// arena indices of the two triangles and flags = bitsA | bitsB << 3 | i << 6 | j << 8, where bitsX has bit k set iff vertex k of
// that triangle is NOT outside the other triangle's plane (CreateUncollideRays.cpp:102-112) and (i, j) are the positions of the
// triangles inside their leaves (so tri - i / tri - j name the leaf, i.e. the combo the hit came from).
struct HitAux { uint32_t triA, triB, flags; };

// One "uncollide" ray (CreateUncollideRays.cpp:153,165), first's model space, 32 B: o = (origin, pair index bits), d = (direction, side bits:
// 0 = from first to second, 1 = from second to first).  Its up to two Hermann responses (ShootUncollideRays.cpp:73-89) are float4
// (response with the sign it enters ray_responses with, valid ? 1 : 0) at resp[2 * ray + step].
struct __align__(16) RayRec { float4 o, d; };

// sorted sweep record, 32 B
struct __align__(16) SweepRec { float umin, umax, vmin, vmax, wmin, wmax; uint32_t idx, cb; };

// Device control block of one frame (read back once per frame).  Every hot word sits on its own 128-B line so
// that polling warps and the atomics of different kernels do not serialise on one L2 line.
struct __align__(128) FrameCtl {
    unsigned long long q_head;      unsigned long long _p0[15];   // global traversal queue: next slot to pop
    unsigned long long q_tail;      unsigned long long _p1[15];   //                          next slot to reserve
    long long pending;              unsigned long long _p2[15];   // alive work items (queue + stacks): termination detector
    unsigned int idle_warps;        unsigned int _p3[31];         // traversal warps currently starving (drives donation)
    unsigned long long n_pairs;     unsigned long long _p4[15];   // broad-phase pairs emitted (may exceed capacity -> overflow)
    unsigned long long n_combos;    unsigned long long _p5[15];
    unsigned long long n_hits;      unsigned long long _p6[15];
    unsigned long long n_colliding; unsigned long long _p7[15];
    unsigned long long tile_cursor; unsigned long long _p8[15];   // narrow phase: next tile of 32 leaf combos to hand out (k_tritri)
    unsigned long long n_class[64];                               // pairs with hits per size class of the contact reduction ([0], [16], [32], [48])
    unsigned long long n_coplanar;
    unsigned long long n_sat;
    unsigned long long n_tri_tests;
    unsigned long long n_donated;          // items that went through the global queue after the roots
    unsigned long long n_iterations;       // traversal warp-iterations (diagnostic)
    unsigned long long busy_cycles, idle_polls;
    unsigned long long n_rays;             // rays of all colliding pairs, both sides
    unsigned long long n_rays_kept;        // rays written to the ray array (pairs that moved): bump allocator of the contact reduction
    unsigned long long scratch_used;       // bytes of the large-pair scratch handed out (bump allocator)
    unsigned long long n_responses;        // successful Hermann passes of the frame
    unsigned long long grouped_used;       // slots of the hit-grouping array handed out (contact reduction)
    unsigned long long n_roots;            // root items of the traversal (k_queue_init)
    unsigned long long n_flagged;          // entries with shouldCallback appended by k_entry_prep (few-flagged broad phase)
    unsigned long long ray_cursor;         // next ray to hand out (k_shoot fetches dynamically: ray costs differ by two orders of magnitude)
    unsigned int overflow;                 // bit0 pairs, bit1 queue, bit2 combos, bit3 hits, bit4 rays, bit5 large-pair scratch, bit6 ray stack
    unsigned int pad;
};
enum { OVF_PAIRS = 1, OVF_QUEUE = 2, OVF_COMBOS = 4, OVF_HITS = 8, OVF_RAYS = 16, OVF_SCRATCH = 32, OVF_RAYSTACK = 64 };

// ---- device helpers shared by the .cu files ----
__device__ __forceinline__ Box unpack_box(const float4& q0, const float4& q1, const float4& q2) {
    Box b;
    b.c = mk3(q0.x, q0.y, q0.z); b.u = mk3(q0.w, q1.x, q1.y); b.v = mk3(q1.z, q1.w, q2.x); b.w = mk3(q2.y, q2.z, q2.w);
    return b;
}
__device__ __forceinline__ Rel rel_from_mat(const float* m) {   // rows of a column-major mat4
    Rel r;
    r.r0 = make_float4(m[0], m[4], m[8], m[12]);
    r.r1 = make_float4(m[1], m[5], m[9], m[13]);
    r.r2 = make_float4(m[2], m[6], m[10], m[14]);
    return r;
}

// ------------------------------------------------------------------------------------------
// host-side helpers
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    // grow to at least `bytes`; keep the first `keep` bytes
    cudaError_t reserve(size_t bytes, size_t keep, cudaStream_t s) {
        if (bytes <= cap) return cudaSuccess;
        size_t ncap = cap ? cap : 4096;
        while (ncap < bytes) ncap += ncap / 2 + 4096;
        void* np = nullptr;
        cudaError_t e = cudaMalloc(&np, ncap);
        if (e != cudaSuccess) return e;
        if (keep && p) { e = cudaMemcpyAsync(np, p, keep, cudaMemcpyDeviceToDevice, s); if (e != cudaSuccess) return e; e = cudaStreamSynchronize(s); if (e != cudaSuccess) return e; }
        if (p) cudaFree(p);
        p = np; cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    // grow to at least `bytes`, keeping the first `keep` bytes; `s` is drained before the old block is freed
    // (asynchronous copies may still be reading it)
    cudaError_t reserve(size_t bytes, size_t keep = 0, cudaStream_t s = nullptr) {
        if (bytes <= cap) return cudaSuccess;
        size_t ncap = cap ? cap : 4096;
        while (ncap < bytes) ncap += ncap / 2 + 4096;
        void* np = nullptr;
        cudaError_t e = cudaMallocHost(&np, ncap);
        if (e != cudaSuccess) return e;
        if (p) {
            if (keep) memcpy(np, p, keep);
            if (s) cudaStreamSynchronize(s);
            cudaFreeHost(p);
        }
        p = np; cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct MeshHost {
    MeshDev dev;
    float root_box[12];
    float build_ms = 0.f;
    bool needs_refit = false;            // positions were updated since the last refit
    int skin = -1;                       // imrcd_mesh_bind_skin
    bool fit_plan = false;               // d_fit holds this mesh's FitRecs (Morton builds leave them; other trees get them at their first refit)
};

struct imrcd_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;

    // mesh arena
    std::vector<MeshHost> meshes;
    DevBuf d_recs, d_tris, d_tri_nrm, d_tri_vid, d_meshes;
    void* copy_pool = nullptr;           // host threads for the staging copy of large submissions (imrcd_api.cu)
    void* skins = nullptr; std::vector<uint32_t> skin_max_joint;                       // re-posing (imrcd_repose.cu)
    PinBuf p_repose; DevBuf d_repose_in, d_repose_prod, d_repose_vtx, d_scalar; bool repose_pending = false; float last_repose_ms = 0.f;
    DevBuf d_rf_stage;                               // re-posed positions on their way into the arena (imrcd_build.cu)
    DevBuf d_fit, d_fit_slot, d_fit_segs, d_fit_scratch, d_fit_ticket; PinBuf p_fit_segs;      // the tree fit (imrcd_fit.cu): FitRec per arena record, per-call tables
    uint32_t fit_ns = 0; uint64_t fit_tot_rec = 0, fit_max_troot = 0, fit_max_slots = 0; bool fit_attr_set = false; int fit_blocks = 0;
    uint64_t fit_key = 0; bool fit_lists_valid = false;      // the last fit call's trees (a hash of their ids and sizes): its lists can be used again
    float last_refit_ms = 0.f;
    uint64_t n_rec_total = 0, n_tri_total = 0;
    bool meshes_dirty = false;
    float last_build_ms = 0.f;

    // mesh recording (imrcd_mesh_begin / add_primitive / end): the primitives' buffers wait in HBM until the mesh is closed
    struct RecordedPrimitive { DevBuf points, normals, indices; uint64_t n_points, n_indices, n_tri; uint32_t stride, mode; };
    std::vector<RecordedPrimitive> recording;
    bool recording_open = false;

    // frame, host side: the entry table is written straight into pinned memory (one host copy per entry) and goes to
    // HBM in chunks while the caller is still adding entries
    PinBuf p_cur, p_prev, p_mesh, p_entity, p_cb;
    uint64_t n_entries = 0;              // entries of this frame KEPT by this context (all of them unless the frame is sharded)
    uint64_t n_entries_global = 0;       // entries added this frame by the caller (every rank of a sharded frame is handed the whole list)
    uint64_t n_flagged_global = 0;       // ... of which shouldCallback
    PinBuf p_gidx; DevBuf d_gidx;        // sharded frames: caller's index of every kept entry (ascending); unsharded: unused (identity)
    std::vector<uint32_t> shard_blk_flagged;            // scratch of imrcd_frame_add_entries (sharded frames)
    uint32_t shard_rank_next = 0, shard_n_next = 1;     // imrcd_frame_set_shard takes effect at the next frame
    uint64_t n_sent = 0;                 // entries whose H2D copy has been enqueued
    bool prev_distinct = false;          // some entry of this frame carries a previous matrix different from its current one
    uint32_t shard_rank = 0, shard_n = 1;
    bool uploaded = false, ran = false, fetched = false;
    bool enqueue_only = false, async_pending = false;      // imrcd_frame_run_async .. imrcd_frame_finish
    uint64_t pending_launches = 0;
    // frame, device side
    DevBuf d_cur, d_prev, d_mesh, d_cb, d_entity, d_inv, d_ext, d_keys, d_keys2, d_idx, d_idx2, d_sorted, d_sorted_c, d_flag, d_cpos, d_wlen, d_chunks, d_chunkoff, d_cubtmp;
    DevBuf d_pairs, d_pairrec, d_pairacc, d_queue, d_combos, d_hits, d_epairs, d_ctl;
    DevBuf d_aux, d_grouped, d_lscratch, d_lpref, d_lsides, d_padded, d_padoff, d_lsmall, d_lmid, d_llarge;      // contact reduction scratch (imrcd_frame.cu)
    DevBuf d_trace;                                                                        // diagnostic timeline of k_traverse (IMRCD_TRAV_TRACE)
    DevBuf d_rays, d_resp, d_epair_pair;                                                    // response stage (imrcd_rays.cu)
    uint64_t cap_pairs = 0, cap_queue = 0, cap_combos = 0, cap_hits = 0, cap_rays = 0, cap_lscratch = 0;
    uint64_t queue_dirty = 0;            // slots whose ready flag may still be set
    // results
    PinBuf p_ctl, p_epairs, p_hits, p_pairs, p_combos;
    FrameCtl ctl_host;
    imrcd_frame_stats stats;
    bool hits_fetched = false;
    cudaEvent_t ev_merge[3] = {};       // diagnostic (IMRCD_MERGE_DEBUG): after the push / all-gather, after the compaction, after the merged records' D2H
    cudaEvent_t ev[10] = {};            // [0..6] frame stages, [6..7] build / refit, [8..9] re-pose
    cudaStream_t stream2 = nullptr, stream3 = nullptr, stream4 = nullptr;      // side streams for independent tail work of a frame
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join3 = nullptr, ev_join4 = nullptr;
    int trav_blocks = 0, trav_blocks_shard = 0, narrow_blocks = 0, trav_variant = 0, shoot_blocks = 0;
    // end-of-frame merge over NCCL (imrcd_comm.cu): one all-gather of fixed-capacity blocks on the frame's stream
    void* comm = nullptr;                // ncclComm_t
    uint32_t comm_rank = 0, comm_n = 1;
    uint64_t gcap = 0;                   // rows per rank in the gathered blocks (after the header row); the same on every rank
    DevBuf d_gather; PinBuf p_gather;    // comm_n x (gcap + 1) x 80 B
    // the merge over peer memory (imrcd_comm.cu): every rank pushes its block into every peer's buffer over NVLink, no NCCL call in a frame
    int p2p_state = 0;                   // 0 not tried yet, 1 in use, -1 not available (the NCCL all-gather is used)
    uint64_t p2p_gcap = 0;               // the capacity the buffers were laid out for
    void* p2p_buf = nullptr;             // this rank's buffer: [2 parities][comm_n blocks of (gcap + 1) rows] + [2][comm_n] arrival flags
    void* p2p_ctl = nullptr;             // device: frame sequence number, tickets, the peers' buffer addresses (P2PCtl)
    std::vector<void*> p2p_opened;       // peers' buffers opened through CUDA IPC (other processes)
    void* group = nullptr;               // imrcd_group this context belongs to (one process, several GPUs), else null
    uint64_t n_merged = 0; bool merged_valid = false;
    cudaGraphExec_t graph_exec = nullptr; uint64_t graph_key = 0, graph_seen_key = 0, graph_launches = 0, graph_spec_rows = 0; bool capturing = false; int use_graph = -1;      // the frame as a CUDA graph
    uint64_t spec_hint = 0, spec_rows_sent = 0;      // speculative D2H of the result rows (imr_frame_spec_rows)
    bool pc_attr_set = false;
    uint32_t few_flagged_max = 0xffffffffu;      // frames with at most this many flagged entries take the sort-free broad phase
    uint32_t pc_large_min = 1024;        // contact reduction: pairs with more hits go to the grid-wide passes
    const void* trav_fn = nullptr;
};

#define IMR_CUDA(ctx, call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(_e); return IMRCD_E_CUDA; } } while (0)

// ---- kernels / stages implemented in the .cu files ----------------------------------------
int imr_mesh_finalize_records(imrcd_ctx* ctx, uint32_t rec_base, uint32_t n_rec);            // surfaces
int imr_mesh_finalize_tris(imrcd_ctx* ctx, uint32_t tri_base, uint32_t n_tri);               // triangle planes (TriRec.t3)
int imr_build_mesh_device(imrcd_ctx* ctx, const float* pos, const float* nrm, const uint32_t* vid, uint64_t n_tri,
                          uint32_t mode, MeshHost* out);
int imr_mesh_assemble_device(imrcd_ctx* ctx, uint32_t build_mode, MeshHost* out);      // Triangle::CreateTriangleList on the device, then the build
int imr_frame_run_device(imrcd_ctx* ctx);
struct FrameCtl;
int imr_traverse_prepare(imrcd_ctx* ctx);                                   // mid phase (imrcd_traverse.cu)
int imr_traverse_queue_init(imrcd_ctx* ctx, FrameCtl* ctl);
int imr_traverse_launch(imrcd_ctx* ctx, FrameCtl* ctl);
int imr_frame_finish_device(imrcd_ctx* ctx);
// response stage: one thread per kept ray (Hermann passes), then one warp per colliding pair that moved (imrcd_rays.cu)
int imr_frame_shoot_device(imrcd_ctx* ctx, FrameCtl* ctl, uint64_t* launches);
int imr_test_ray_tree_device(imrcd_ctx* ctx, uint32_t mesh_id, uint64_t n, const float* mats, const float* origins, const float* dirs,
                             uint8_t* flags, float* out3, uint32_t* tri);
int imr_mesh_update_positions_device(imrcd_ctx* ctx, uint32_t mesh_id, const float* pos, const float* nrm);
int imr_meshes_refit_device(imrcd_ctx* ctx, const uint32_t* ids, uint64_t n_ids, float* ms_out);
void imr_skins_release(imrcd_ctx* ctx);
int imr_device_max_u32(imrcd_ctx* ctx, const uint32_t* d_values, uint64_t n, uint32_t* out);
