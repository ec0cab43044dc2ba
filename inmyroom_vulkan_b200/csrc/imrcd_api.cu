// imrcd_api.cu -- the extern "C" surface of libimrcd.so (declared in include/imrcd.h): context,
// mesh arena (import / export), frame entry list, upload / run / fetch, and the unit-level test hooks.
#include "imrcd_internal.cuh"
#include <algorithm>
#include <cstddef>
#include <cstring>
#include <deque>

static_assert(sizeof(imrcd_entity_pair) == 80, "imrcd_entity_pair layout");
static_assert(sizeof(imrcd_tri_hit) == 40, "imrcd_tri_hit layout");
static_assert(sizeof(TreeRec) == 64 && sizeof(TriRec) == 64 && sizeof(PairRec) == 64 && sizeof(SweepRec) == 32, "HBM layouts");

#define CHECK_CTX(ctx) do { if (!(ctx)) return IMRCD_E_ARG; } while (0)

extern "C" const char* imrcd_version(void) { return "imrcd 0.1 (sm_100a, fmad=false)"; }

// sizes and offsets of the ABI's structs as this library was compiled: a binding checks its own mirror against them (tests/test_abi.py)
extern "C" int imrcd_abi_layout(uint64_t* out, uint64_t capacity) {
    const uint64_t v[] = { sizeof(imrcd_entity_pair), sizeof(imrcd_tri_hit), sizeof(imrcd_frame_stats),
                           offsetof(imrcd_frame_stats, n_entries), offsetof(imrcd_frame_stats, n_pairs), offsetof(imrcd_frame_stats, n_sat_tests), offsetof(imrcd_frame_stats, n_combos),
                           offsetof(imrcd_frame_stats, n_tri_tests), offsetof(imrcd_frame_stats, n_hits), offsetof(imrcd_frame_stats, n_coplanar_hits), offsetof(imrcd_frame_stats, n_colliding),
                           offsetof(imrcd_frame_stats, traverse_launches), offsetof(imrcd_frame_stats, total_launches), offsetof(imrcd_frame_stats, n_queue_items),
                           offsetof(imrcd_frame_stats, n_warp_iterations), offsetof(imrcd_frame_stats, trav_busy_cycles), offsetof(imrcd_frame_stats, trav_idle_polls),
                           offsetof(imrcd_frame_stats, ms_total), offsetof(imrcd_frame_stats, ms_broad), offsetof(imrcd_frame_stats, ms_pair_setup), offsetof(imrcd_frame_stats, ms_traverse),
                           offsetof(imrcd_frame_stats, ms_narrow), offsetof(imrcd_frame_stats, ms_reduce), offsetof(imrcd_frame_stats, n_contact_pairs), offsetof(imrcd_frame_stats, n_rays),
                           offsetof(imrcd_frame_stats, n_rays_shot), offsetof(imrcd_frame_stats, n_responses), offsetof(imrcd_frame_stats, ms_response), offsetof(imrcd_frame_stats, n_merged),
                           offsetof(imrcd_frame_stats, n_entries_local) };
    const uint64_t n = sizeof(v) / sizeof(v[0]);
    if (!out || capacity < n) return (int)n;
    for (uint64_t i = 0; i < n; ++i) out[i] = v[i];
    return (int)n;
}

extern "C" const char* imrcd_last_error(const imrcd_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int imrcd_create(int device, void* cuda_stream, imrcd_ctx** out) {
    if (!out) return IMRCD_E_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return IMRCD_E_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return IMRCD_E_NODEVICE;
    if (prop.major != 10) return IMRCD_E_NODEVICE;       // the only code in this library is sm_100a SASS
    if (cudaSetDevice(device) != cudaSuccess) return IMRCD_E_CUDA;
    imrcd_ctx* ctx = new imrcd_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cuda_stream) { ctx->stream = static_cast<cudaStream_t>(cuda_stream); ctx->own_stream = false; }
    else { if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return IMRCD_E_CUDA; } ctx->own_stream = true; }
    for (auto& e : ctx->ev) if (cudaEventCreate(&e) != cudaSuccess) { delete ctx; return IMRCD_E_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream3, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join3, cudaEventDisableTiming) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream4, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join4, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return IMRCD_E_CUDA; }
    *out = ctx;
    return IMRCD_OK;
}

static void recording_clear(imrcd_ctx* ctx);
void imr_copy_pool_release(imrcd_ctx* ctx);
extern "C" void imrcd_destroy(imrcd_ctx* ctx) {
    if (!ctx) return;
    recording_clear(ctx);
    imrcd_comm_destroy(ctx);
    imr_skins_release(ctx);
    imr_copy_pool_release(ctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = { &ctx->d_recs, &ctx->d_tris, &ctx->d_tri_nrm, &ctx->d_tri_vid, &ctx->d_meshes, &ctx->d_rf_stage, &ctx->d_repose_in, &ctx->d_repose_prod, &ctx->d_repose_vtx, &ctx->d_scalar, &ctx->d_fit, &ctx->d_fit_slot, &ctx->d_fit_segs, &ctx->d_fit_scratch, &ctx->d_fit_ticket, &ctx->d_cur, &ctx->d_prev, &ctx->d_mesh,
                       &ctx->d_cb, &ctx->d_entity, &ctx->d_inv, &ctx->d_ext, &ctx->d_keys, &ctx->d_keys2, &ctx->d_idx, &ctx->d_idx2,
                       &ctx->d_sorted, &ctx->d_sorted_c, &ctx->d_flag, &ctx->d_cpos, &ctx->d_wlen, &ctx->d_chunks, &ctx->d_chunkoff, &ctx->d_cubtmp, &ctx->d_pairs, &ctx->d_pairrec, &ctx->d_pairacc, &ctx->d_queue, &ctx->d_combos,
                       &ctx->d_hits, &ctx->d_epairs, &ctx->d_ctl, &ctx->d_aux, &ctx->d_grouped, &ctx->d_lscratch, &ctx->d_lpref, &ctx->d_lsides, &ctx->d_rays, &ctx->d_resp, &ctx->d_epair_pair, &ctx->d_trace, &ctx->d_gidx, &ctx->d_gather, &ctx->d_padded, &ctx->d_padoff, &ctx->d_lsmall, &ctx->d_lmid, &ctx->d_llarge };
    for (DevBuf* b : bufs) b->release();
    PinBuf* pins[] = { &ctx->p_repose, &ctx->p_fit_segs, &ctx->p_gidx, &ctx->p_gather, &ctx->p_cur, &ctx->p_prev, &ctx->p_mesh, &ctx->p_entity, &ctx->p_cb, &ctx->p_ctl, &ctx->p_epairs, &ctx->p_hits, &ctx->p_pairs, &ctx->p_combos };
    for (PinBuf* b : pins) b->release();
    if (ctx->graph_exec) cudaGraphExecDestroy(ctx->graph_exec);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream3) cudaStreamDestroy(ctx->stream3);
    if (ctx->ev_join3) cudaEventDestroy(ctx->ev_join3);
    if (ctx->stream4) cudaStreamDestroy(ctx->stream4);
    if (ctx->ev_join4) cudaEventDestroy(ctx->ev_join4);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

// ------------------------------------------------------------------------------------------
// meshes
// ------------------------------------------------------------------------------------------
__global__ void k_rec_surface(TreeRec* recs, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    TreeRec r = recs[i];
    Box b;
    b.c = mk3(r.q0.x, r.q0.y, r.q0.z); b.u = mk3(r.q0.w, r.q1.x, r.q1.y); b.v = mk3(r.q1.z, r.q1.w, r.q2.x); b.w = mk3(r.q2.y, r.q2.z, r.q2.w);
    recs[i].q3.x = box_surface(b);
}

int imr_mesh_finalize_records(imrcd_ctx* ctx, uint32_t rec_base, uint32_t n_rec) {
    if (n_rec == 0) return IMRCD_OK;
    k_rec_surface<<<(n_rec + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_recs.as<TreeRec>() + rec_base, n_rec);
    IMR_CUDA(ctx, cudaGetLastError());
    return IMRCD_OK;
}

__global__ void k_tri_planes(TriRec* tris, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = tris[i].t0, b = tris[i].t1, c = tris[i].t2;
    V3 N; float d;
    tt_plane(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), N, d);
    tris[i].t3 = make_float4(N.x, N.y, N.z, d);
}

int imr_mesh_finalize_tris(imrcd_ctx* ctx, uint32_t tri_base, uint32_t n_tri) {
    if (n_tri == 0) return IMRCD_OK;
    k_tri_planes<<<(n_tri + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_tris.as<TriRec>() + tri_base, n_tri);
    IMR_CUDA(ctx, cudaGetLastError());
    return IMRCD_OK;
}

// reserve arena space for a mesh and register it; returns its id
static int mesh_alloc(imrcd_ctx* ctx, uint64_t n_rec, uint64_t n_tri, MeshHost* mh) {
    if (ctx->n_rec_total + n_rec >= (1ull << 32) || ctx->n_tri_total + n_tri >= (1ull << 32)) { ctx->err = "mesh arena exceeds 2^32 records"; return IMRCD_E_CAPACITY; }
    cudaStream_t s = ctx->stream;
    IMR_CUDA(ctx, ctx->d_recs.reserve(sizeof(TreeRec) * (ctx->n_rec_total + n_rec), sizeof(TreeRec) * ctx->n_rec_total, s));
    IMR_CUDA(ctx, ctx->d_tris.reserve(sizeof(TriRec) * (ctx->n_tri_total + n_tri), sizeof(TriRec) * ctx->n_tri_total, s));
    IMR_CUDA(ctx, ctx->d_tri_nrm.reserve(36ull * (ctx->n_tri_total + n_tri), 36ull * ctx->n_tri_total, s));
    IMR_CUDA(ctx, ctx->d_tri_vid.reserve(12ull * (ctx->n_tri_total + n_tri), 12ull * ctx->n_tri_total, s));
    mh->dev.rec_base = (uint32_t)ctx->n_rec_total; mh->dev.tri_base = (uint32_t)ctx->n_tri_total;
    mh->dev.n_rec = (uint32_t)n_rec; mh->dev.n_tri = (uint32_t)n_tri;
    ctx->n_rec_total += n_rec; ctx->n_tri_total += n_tri;
    return IMRCD_OK;
}

static int mesh_register(imrcd_ctx* ctx, const MeshHost& mh, uint32_t* mesh_id) {
    ctx->meshes.push_back(mh);
    ctx->meshes_dirty = true;
    *mesh_id = (uint32_t)(ctx->meshes.size() - 1);
    return IMRCD_OK;
}

static int meshes_sync(imrcd_ctx* ctx) {
    if (!ctx->meshes_dirty) return IMRCD_OK;
    std::vector<MeshDev> tab(ctx->meshes.size());
    for (size_t i = 0; i < tab.size(); ++i) tab[i] = ctx->meshes[i].dev;
    IMR_CUDA(ctx, ctx->d_meshes.reserve(sizeof(MeshDev) * tab.size(), 0, ctx->stream));
    IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_meshes.p, tab.data(), sizeof(MeshDev) * tab.size(), cudaMemcpyHostToDevice, ctx->stream));
    IMR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->meshes_dirty = false;
    return IMRCD_OK;
}

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

extern "C" int imrcd_mesh_import_tree(imrcd_ctx* ctx, uint64_t nv, const float* boxes, const int32_t* left, const int32_t* right,
                                      const uint32_t* tri_off, const uint32_t* tri_cnt, uint64_t n_tri, const float* tri_pos,
                                      const float* tri_nrm, const uint32_t* tri_vid, const uint32_t* tri_orig, uint32_t* mesh_id) {
    CHECK_CTX(ctx);
    if (!boxes || !left || !right || !tri_off || !tri_cnt || !mesh_id || nv == 0 || (n_tri && !tri_pos)) { ctx->err = "imrcd_mesh_import_tree: bad argument"; return IMRCD_E_ARG; }
    cudaSetDevice(ctx->device);
    // flat pre-order -> sibling-adjacent records: root at 0, slot 1 padding, children pairs from 2 on (BFS order)
    std::vector<uint32_t> rec_of(nv, 0xffffffffu), depth_of(nv, 0u);
    std::deque<uint32_t> bfs;
    uint64_t next = 2;
    rec_of[0] = 0; bfs.push_back(0);
    while (!bfs.empty()) {
        uint32_t v = bfs.front(); bfs.pop_front();
        if (depth_of[v] >= 120u) { ctx->err = "imrcd_mesh_import_tree: the tree is deeper than 120 levels (the ray descent of the response stage keeps one stack entry per level)"; return IMRCD_E_ARG; }
        if (left[v] >= 0) {
            if ((uint64_t)left[v] >= nv || (uint64_t)right[v] >= nv || right[v] < 0) { ctx->err = "imrcd_mesh_import_tree: child index out of range"; return IMRCD_E_ARG; }
            if (rec_of[left[v]] != 0xffffffffu || rec_of[right[v]] != 0xffffffffu || left[v] == right[v]) { ctx->err = "imrcd_mesh_import_tree: a vertex has two parents"; return IMRCD_E_ARG; }
            rec_of[left[v]] = (uint32_t)next; rec_of[right[v]] = (uint32_t)(next + 1); next += 2;
            depth_of[left[v]] = depth_of[right[v]] = depth_of[v] + 1u;
            bfs.push_back((uint32_t)left[v]); bfs.push_back((uint32_t)right[v]);
        }
    }
    const uint64_t n_rec = next;
    std::vector<TreeRec> recs(n_rec);
    memset(recs.data(), 0, sizeof(TreeRec) * n_rec);
    for (uint64_t v = 0; v < nv; ++v) {
        if (rec_of[v] == 0xffffffffu) continue;    // unreachable vertex
        const float* b = boxes + 12 * v;
        TreeRec& r = recs[rec_of[v]];
        r.q0 = make_float4(b[0], b[1], b[2], b[3]); r.q1 = make_float4(b[4], b[5], b[6], b[7]); r.q2 = make_float4(b[8], b[9], b[10], b[11]);
        if (left[v] >= 0) r.q3 = make_float4(0.f, u2f(rec_of[left[v]]), u2f(0u), u2f(0u));
        else {
            if ((uint64_t)tri_off[v] + tri_cnt[v] > n_tri || tri_cnt[v] > 4) { ctx->err = "imrcd_mesh_import_tree: bad leaf range (leaves hold <= 4 triangles, OBBtree.h:49)"; return IMRCD_E_ARG; }
            r.q3 = make_float4(0.f, u2f(tri_off[v]), u2f(tri_cnt[v]), u2f(1u));
        }
    }
    std::vector<TriRec> tris(n_tri ? n_tri : 1);
    std::vector<float> nrm(9 * (n_tri ? n_tri : 1), 0.f);
    std::vector<uint32_t> vid(3 * (n_tri ? n_tri : 1), 0u);
    for (uint64_t i = 0; i < n_tri; ++i) {
        const float* p = tri_pos + 9 * i;
        tris[i].t0 = make_float4(p[0], p[1], p[2], u2f(tri_orig ? tri_orig[i] : (uint32_t)i));
        tris[i].t1 = make_float4(p[3], p[4], p[5], 0.f);
        tris[i].t2 = make_float4(p[6], p[7], p[8], 0.f);
        tris[i].t3 = make_float4(0.f, 0.f, 0.f, 0.f);        // filled on the device (imr_mesh_finalize_tris)
        if (tri_nrm) memcpy(&nrm[9 * i], tri_nrm + 9 * i, 36);
        for (int k = 0; k < 3; ++k) vid[3 * i + k] = tri_vid ? tri_vid[3 * i + k] : (uint32_t)(3 * i + k);
    }
    MeshHost mh;
    int rc = mesh_alloc(ctx, n_rec, n_tri, &mh);
    if (rc) return rc;
    memcpy(mh.root_box, boxes, 48);
    cudaStream_t s = ctx->stream;
    IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_recs.as<TreeRec>() + mh.dev.rec_base, recs.data(), sizeof(TreeRec) * n_rec, cudaMemcpyHostToDevice, s));
    if (n_tri) {
        IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_tris.as<TriRec>() + mh.dev.tri_base, tris.data(), sizeof(TriRec) * n_tri, cudaMemcpyHostToDevice, s));
        IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_tri_nrm.as<float>() + 9ull * mh.dev.tri_base, nrm.data(), 36ull * n_tri, cudaMemcpyHostToDevice, s));
        IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_tri_vid.as<uint32_t>() + 3ull * mh.dev.tri_base, vid.data(), 12ull * n_tri, cudaMemcpyHostToDevice, s));
    }
    rc = imr_mesh_finalize_records(ctx, mh.dev.rec_base, mh.dev.n_rec);
    if (rc) return rc;
    rc = imr_mesh_finalize_tris(ctx, mh.dev.tri_base, mh.dev.n_tri);
    if (rc) return rc;
    IMR_CUDA(ctx, cudaStreamSynchronize(s));
    return mesh_register(ctx, mh, mesh_id);
}

extern "C" int imrcd_mesh_create(imrcd_ctx* ctx, const float* positions, const float* normals, const uint32_t* vertex_ids,
                                 uint64_t n_tri, uint32_t build_mode, uint32_t* mesh_id) {
    CHECK_CTX(ctx);
    if (!mesh_id || (n_tri && !positions)) { ctx->err = "imrcd_mesh_create: bad argument"; return IMRCD_E_ARG; }
    cudaSetDevice(ctx->device);
    MeshHost mh;
    int rc = imr_build_mesh_device(ctx, positions, normals, vertex_ids, n_tri, build_mode, &mh);
    if (rc) return rc;
    ctx->last_build_ms = mh.build_ms;
    return mesh_register(ctx, mh, mesh_id);
}

// ---- mesh recording: the engine's StartRecordOBBtree / AddPrimitive / GetOBBtreeAndReset (PrimitivesOfMeshes.cpp:835-863) ----
static uint64_t gltf_triangle_count(uint32_t mode, uint64_t ni) {       // CreateIndicesTriplets, Triangle.cpp:9-62
    switch (mode) {
        case 0: return ni;
        case 1: return ni / 2;
        case 3: return ni ? ni - 1 : 0;
        case 4: return ni / 3;
        case 5: case 6: return ni >= 2 ? ni - 2 : 0;
        default: return 0;                                               // line loop: not handled by the reference's switch
    }
}
static void recording_clear(imrcd_ctx* ctx) {
    for (auto& r : ctx->recording) { r.points.release(); r.normals.release(); r.indices.release(); }
    ctx->recording.clear(); ctx->recording_open = false;
}
// used by imrcd_gltf.cpp, which sees the context only through the ABI
void imr_ctx_set_error(imrcd_ctx* ctx, const char* msg) { if (ctx) ctx->err = msg ? msg : ""; }

extern "C" int imrcd_mesh_begin(imrcd_ctx* ctx) {
    CHECK_CTX(ctx);
    recording_clear(ctx);
    ctx->recording_open = true;
    return IMRCD_OK;
}
extern "C" int imrcd_mesh_add_primitive(imrcd_ctx* ctx, const float* points, uint64_t n_points, uint32_t stride_floats, const float* normals,
                                        const uint32_t* indices, uint64_t n_indices, uint32_t gltf_mode) {
    CHECK_CTX(ctx);
    if (!ctx->recording_open) { ctx->err = "imrcd_mesh_add_primitive before imrcd_mesh_begin"; return IMRCD_E_STATE; }
    if ((n_points && !points) || (stride_floats != 3 && stride_floats != 4) || gltf_mode > 6) { ctx->err = "imrcd_mesh_add_primitive: bad argument"; return IMRCD_E_ARG; }
    if (n_points == 0 && indices && n_indices) { ctx->err = "imrcd_mesh_add_primitive: indices without points"; return IMRCD_E_ARG; }
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    ctx->recording.emplace_back();
    imrcd_ctx::RecordedPrimitive& r = ctx->recording.back();
    r.n_points = n_points; r.stride = stride_floats; r.mode = gltf_mode;
    r.n_indices = indices ? n_indices : n_points;
    r.n_tri = gltf_triangle_count(gltf_mode, r.n_indices);
    const size_t pbytes = 4ull * stride_floats * n_points;
    if (pbytes) {
        IMR_CUDA(ctx, r.points.reserve(pbytes, 0, s));
        IMR_CUDA(ctx, cudaMemcpyAsync(r.points.p, points, pbytes, cudaMemcpyDefault, s));
        if (normals) { IMR_CUDA(ctx, r.normals.reserve(pbytes, 0, s)); IMR_CUDA(ctx, cudaMemcpyAsync(r.normals.p, normals, pbytes, cudaMemcpyDefault, s)); }
    }
    if (indices && n_indices) { IMR_CUDA(ctx, r.indices.reserve(4ull * n_indices, 0, s)); IMR_CUDA(ctx, cudaMemcpyAsync(r.indices.p, indices, 4ull * n_indices, cudaMemcpyDefault, s)); }
    IMR_CUDA(ctx, cudaStreamSynchronize(s));                            // the caller may reuse its buffers
    if (indices && n_indices) {       // the engine forwards raw glTF index buffers: an index past the points would read outside the primitive
        uint32_t mx = 0;
        const int rc = imr_device_max_u32(ctx, r.indices.as<uint32_t>(), n_indices, &mx);
        if (rc) { ctx->recording.back().points.release(); ctx->recording.back().normals.release(); ctx->recording.back().indices.release(); ctx->recording.pop_back(); return rc; }
        if (mx >= n_points) {
            ctx->recording.back().points.release(); ctx->recording.back().normals.release(); ctx->recording.back().indices.release(); ctx->recording.pop_back();
            ctx->err = "imrcd_mesh_add_primitive: an index exceeds the number of points"; return IMRCD_E_ARG;
        }
    }
    return IMRCD_OK;
}
extern "C" int imrcd_mesh_end(imrcd_ctx* ctx, uint32_t build_mode, uint32_t* mesh_id) {
    CHECK_CTX(ctx);
    if (!ctx->recording_open || !mesh_id) { ctx->err = "imrcd_mesh_end: no mesh is being recorded"; return IMRCD_E_STATE; }
    cudaSetDevice(ctx->device);
    MeshHost mh;
    int rc = imr_mesh_assemble_device(ctx, build_mode, &mh);
    recording_clear(ctx);
    if (rc) return rc;
    ctx->last_build_ms = mh.build_ms;
    return mesh_register(ctx, mh, mesh_id);
}

// ---- refit (BASELINE config 5) ----
extern "C" int imrcd_mesh_update_positions(imrcd_ctx* ctx, uint32_t mesh_id, const float* positions, const float* normals) {
    CHECK_CTX(ctx);
    if (mesh_id >= ctx->meshes.size() || !positions) { ctx->err = "imrcd_mesh_update_positions: bad argument"; return IMRCD_E_ARG; }
    cudaSetDevice(ctx->device);
    return imr_mesh_update_positions_device(ctx, mesh_id, positions, normals);
}
extern "C" int imrcd_mesh_refit(imrcd_ctx* ctx, const uint32_t* mesh_ids, uint64_t n) {
    CHECK_CTX(ctx);
    cudaSetDevice(ctx->device);
    std::vector<uint32_t> all;
    if (!mesh_ids) {                                    // every mesh whose positions changed since its last refit
        for (uint32_t i = 0; i < ctx->meshes.size(); ++i) if (ctx->meshes[i].needs_refit) all.push_back(i);
        mesh_ids = all.data(); n = all.size();
    }
    for (uint64_t i = 0; i < n; ++i) if (mesh_ids[i] >= ctx->meshes.size()) { ctx->err = "imrcd_mesh_refit: bad mesh id"; return IMRCD_E_ARG; }
    return imr_meshes_refit_device(ctx, mesh_ids, n, &ctx->last_refit_ms);
}
extern "C" int imrcd_mesh_last_refit_ms(imrcd_ctx* ctx, float* ms) { CHECK_CTX(ctx); if (ms) *ms = ctx->last_refit_ms; return IMRCD_OK; }

extern "C" int imrcd_mesh_last_build_ms(imrcd_ctx* ctx, float* ms) { CHECK_CTX(ctx); if (ms) *ms = ctx->last_build_ms; return IMRCD_OK; }

// used by imr_build_mesh_device
int imr_mesh_arena_alloc(imrcd_ctx* ctx, uint64_t n_rec, uint64_t n_tri, MeshHost* mh) { return mesh_alloc(ctx, n_rec, n_tri, mh); }
// grow the arena's capacity for a mesh of at most n_rec records / n_tri triangles without placing it (keeps allocation out of timed regions)
int imr_mesh_arena_reserve(imrcd_ctx* ctx, uint64_t n_rec, uint64_t n_tri) {
    cudaStream_t s = ctx->stream;
    IMR_CUDA(ctx, ctx->d_recs.reserve(sizeof(TreeRec) * (ctx->n_rec_total + n_rec), sizeof(TreeRec) * ctx->n_rec_total, s));
    IMR_CUDA(ctx, ctx->d_tris.reserve(sizeof(TriRec) * (ctx->n_tri_total + n_tri), sizeof(TriRec) * ctx->n_tri_total, s));
    IMR_CUDA(ctx, ctx->d_tri_nrm.reserve(36ull * (ctx->n_tri_total + n_tri), 36ull * ctx->n_tri_total, s));
    IMR_CUDA(ctx, ctx->d_tri_vid.reserve(12ull * (ctx->n_tri_total + n_tri), 12ull * ctx->n_tri_total, s));
    return IMRCD_OK;
}

extern "C" int imrcd_mesh_info(imrcd_ctx* ctx, uint32_t mesh_id, uint64_t* n_tri, uint64_t* n_vertices) {
    CHECK_CTX(ctx);
    if (mesh_id >= ctx->meshes.size()) { ctx->err = "bad mesh id"; return IMRCD_E_ARG; }
    const MeshDev& m = ctx->meshes[mesh_id].dev;
    if (n_tri) *n_tri = m.n_tri;
    if (n_vertices) *n_vertices = m.n_rec >= 2 ? m.n_rec - 1 : 1;   // slot 1 is padding
    return IMRCD_OK;
}

extern "C" int imrcd_mesh_export_tree(imrcd_ctx* ctx, uint32_t mesh_id, float* boxes, int32_t* left, int32_t* right,
                                      uint32_t* tri_off, uint32_t* tri_cnt, float* tri_pos, float* tri_nrm, uint32_t* tri_vid,
                                      uint32_t* tri_orig) {
    CHECK_CTX(ctx);
    if (mesh_id >= ctx->meshes.size()) { ctx->err = "bad mesh id"; return IMRCD_E_ARG; }
    cudaSetDevice(ctx->device);
    const MeshDev m = ctx->meshes[mesh_id].dev;
    std::vector<TreeRec> recs(m.n_rec);
    std::vector<TriRec> tris(m.n_tri ? m.n_tri : 1);
    cudaStream_t s = ctx->stream;
    IMR_CUDA(ctx, cudaMemcpyAsync(recs.data(), ctx->d_recs.as<TreeRec>() + m.rec_base, sizeof(TreeRec) * m.n_rec, cudaMemcpyDeviceToHost, s));
    if (m.n_tri) IMR_CUDA(ctx, cudaMemcpyAsync(tris.data(), ctx->d_tris.as<TriRec>() + m.tri_base, sizeof(TriRec) * m.n_tri, cudaMemcpyDeviceToHost, s));
    if (tri_nrm && m.n_tri) IMR_CUDA(ctx, cudaMemcpyAsync(tri_nrm, ctx->d_tri_nrm.as<float>() + 9ull * m.tri_base, 36ull * m.n_tri, cudaMemcpyDeviceToHost, s));
    if (tri_vid && m.n_tri) IMR_CUDA(ctx, cudaMemcpyAsync(tri_vid, ctx->d_tri_vid.as<uint32_t>() + 3ull * m.tri_base, 12ull * m.n_tri, cudaMemcpyDeviceToHost, s));
    IMR_CUDA(ctx, cudaStreamSynchronize(s));
    // iterative pre-order walk (left first) over the sibling-adjacent layout
    struct Fr { uint32_t rec; int32_t parent; bool is_right; };
    std::vector<Fr> stack; stack.push_back({0u, -1, false});
    uint64_t nvx = 0;
    while (!stack.empty()) {
        Fr f = stack.back(); stack.pop_back();
        const TreeRec& r = recs[f.rec];
        uint64_t me = nvx++;
        if (boxes) { float* b = boxes + 12 * me; b[0] = r.q0.x; b[1] = r.q0.y; b[2] = r.q0.z; b[3] = r.q0.w; b[4] = r.q1.x; b[5] = r.q1.y; b[6] = r.q1.z; b[7] = r.q1.w; b[8] = r.q2.x; b[9] = r.q2.y; b[10] = r.q2.z; b[11] = r.q2.w; }
        if (f.parent >= 0) { if (f.is_right) { if (right) right[f.parent] = (int32_t)me; } else { if (left) left[f.parent] = (int32_t)me; } }
        const bool leaf = f2u(r.q3.w) != 0u;
        if (leaf) {
            if (left) left[me] = -1; if (right) right[me] = -1;
            if (tri_off) tri_off[me] = f2u(r.q3.y); if (tri_cnt) tri_cnt[me] = f2u(r.q3.z);
        } else {
            if (tri_off) tri_off[me] = 0; if (tri_cnt) tri_cnt[me] = 0;
            uint32_t c = f2u(r.q3.y);
            stack.push_back({c + 1u, (int32_t)me, true});
            stack.push_back({c, (int32_t)me, false});
        }
    }
    for (uint64_t i = 0; i < m.n_tri; ++i) {
        if (tri_pos) { float* p = tri_pos + 9 * i; p[0] = tris[i].t0.x; p[1] = tris[i].t0.y; p[2] = tris[i].t0.z; p[3] = tris[i].t1.x; p[4] = tris[i].t1.y; p[5] = tris[i].t1.z; p[6] = tris[i].t2.x; p[7] = tris[i].t2.y; p[8] = tris[i].t2.z; }
        if (tri_orig) tri_orig[i] = f2u(tris[i].t0.w);
    }
    return IMRCD_OK;
}

// ------------------------------------------------------------------------------------------
// frame
// ------------------------------------------------------------------------------------------
extern "C" int imrcd_frame_reset(imrcd_ctx* ctx) {
    CHECK_CTX(ctx);
    ctx->n_entries = 0; ctx->n_sent = 0; ctx->prev_distinct = false;
    ctx->n_entries_global = 0; ctx->n_flagged_global = 0;
    ctx->shard_rank = ctx->shard_rank_next; ctx->shard_n = ctx->shard_n_next;
    ctx->uploaded = ctx->ran = ctx->fetched = false; ctx->merged_valid = false;
    return IMRCD_OK;
}

#define ENTRY_CHUNK 16384u     // entries per H2D chunk of the pipelined upload (1 MiB of matrices)
// sharded frames: entries without shouldCallback are dealt to the ranks in blocks of this many caller indices (every rank must use the same
// value; IMRCD_SHARD_BLOCK in the environment overrides it for experiments)
static const uint64_t SHARD_BLOCK = []() -> uint64_t { const char* ev = getenv("IMRCD_SHARD_BLOCK"); const long v = ev ? atol(ev) : 0; return v > 0 ? (uint64_t)v : 256ull; }();

// Sharded frames (imrcd_frame_set_shard, SURVEY 8e "sharded by entity").  A pair needs shouldCallback on one side at least
// (SweepAndPrune.cpp:60), so entries WITH the flag go to every rank and entries WITHOUT it to exactly one: a pair with an unflagged entity is
// found by the rank that owns that entity, a pair of two flagged entities by one rank picked from the pair itself (k_sweep).  Every rank is
// handed the whole entry list and keeps its share, in the caller's order, with the caller's index of every kept entry (gidx).
static inline bool shard_owns(const imrcd_ctx* ctx, uint64_t g) { return (g / SHARD_BLOCK) % ctx->shard_n == ctx->shard_rank; }

// enqueue the H2D copies of entries [n_sent, upto)
static int entries_send(imrcd_ctx* ctx, uint64_t upto) {
    const uint64_t a = ctx->n_sent;
    if (upto <= a) return IMRCD_OK;
    cudaStream_t s = ctx->stream;
    const uint64_t k = upto - a;
    IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_cur.as<char>() + 64 * a, ctx->p_cur.as<char>() + 64 * a, 64 * k, cudaMemcpyHostToDevice, s));
    IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_mesh.as<char>() + 4 * a, ctx->p_mesh.as<char>() + 4 * a, 4 * k, cudaMemcpyHostToDevice, s));
    IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_entity.as<char>() + 4 * a, ctx->p_entity.as<char>() + 4 * a, 4 * k, cudaMemcpyHostToDevice, s));
    IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_cb.as<char>() + a, ctx->p_cb.as<char>() + a, k, cudaMemcpyHostToDevice, s));
    if (ctx->shard_n > 1) IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_gidx.as<char>() + 4 * a, ctx->p_gidx.as<char>() + 4 * a, 4 * k, cudaMemcpyHostToDevice, s));
    ctx->n_sent = upto;
    return IMRCD_OK;
}

static int entries_reserve(imrcd_ctx* ctx, uint64_t total, bool device_too) {
    const uint64_t base = ctx->n_entries;
    cudaStream_t s = ctx->stream;
    IMR_CUDA(ctx, ctx->p_cur.reserve(64 * total, 64 * base, s));
    IMR_CUDA(ctx, ctx->p_prev.reserve(64 * total, 64 * base, s));
    IMR_CUDA(ctx, ctx->p_mesh.reserve(4 * total, 4 * base, s));
    IMR_CUDA(ctx, ctx->p_entity.reserve(4 * total, 4 * base, s));
    IMR_CUDA(ctx, ctx->p_cb.reserve(total, base, s));
    if (ctx->shard_n > 1) IMR_CUDA(ctx, ctx->p_gidx.reserve(4 * total, 4 * base, s));
    if (device_too) {
        IMR_CUDA(ctx, ctx->d_cur.reserve(64 * total, 64 * ctx->n_sent, s));
        IMR_CUDA(ctx, ctx->d_mesh.reserve(4 * total, 4 * ctx->n_sent, s));
        IMR_CUDA(ctx, ctx->d_entity.reserve(4 * total, 4 * ctx->n_sent, s));
        IMR_CUDA(ctx, ctx->d_cb.reserve(total, ctx->n_sent, s));
        if (ctx->shard_n > 1) IMR_CUDA(ctx, ctx->d_gidx.reserve(4 * total, 4 * ctx->n_sent, s));
    }
    return IMRCD_OK;
}

// ---- the staging copy of a large submission on a few host threads -----------------------------------------------------------------------
// Writing 100,000 entries into the pinned staging is 6.4 MB of memcpy: ~0.65 ms on one host thread, a third of the end-to-end frame.  A small
// persistent pool (created at the first large submission) copies a chunk in slices while the previous chunk's DMA runs.
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
struct CopyPool {
    std::vector<std::thread> workers;
    std::mutex m; std::condition_variable cv_go, cv_done;
    char* dst = nullptr; const char* src = nullptr; size_t bytes = 0; uint64_t epoch = 0; int pending = 0; bool quit = false;
    explicit CopyPool(int n) {
        for (int w = 0; w < n; ++w) workers.emplace_back([this, w, n]() {
            uint64_t seen = 0;
            for (;;) {
                char* d; const char* sp; size_t b;
                { std::unique_lock<std::mutex> lk(m); cv_go.wait(lk, [&] { return quit || epoch != seen; }); if (quit) return; seen = epoch; d = dst; sp = src; b = bytes; }
                const size_t per = ((b / (size_t)(n + 1)) + 63) & ~size_t(63), lo = std::min(b, per * (size_t)(w + 1)), hi = std::min(b, lo + per);
                if (hi > lo) memcpy(d + lo, sp + lo, hi - lo);
                { std::lock_guard<std::mutex> lk(m); if (--pending == 0) cv_done.notify_one(); }
            }
        });
    }
    void copy(void* d, const void* sp, size_t b) {
        const int n = (int)workers.size();
        { std::lock_guard<std::mutex> lk(m); dst = (char*)d; src = (const char*)sp; bytes = b; pending = n; ++epoch; }
        cv_go.notify_all();
        const size_t per = ((b / (size_t)(n + 1)) + 63) & ~size_t(63);
        memcpy(d, sp, std::min(b, per));                                  // the caller's own slice, then whatever the slices of the workers left over
        if (per * (size_t)(n + 1) < b) memcpy((char*)d + per * (size_t)(n + 1), (const char*)sp + per * (size_t)(n + 1), b - per * (size_t)(n + 1));
        std::unique_lock<std::mutex> lk(m); cv_done.wait(lk, [&] { return pending == 0; });
    }
    ~CopyPool() { { std::lock_guard<std::mutex> lk(m); quit = true; } cv_go.notify_all(); for (auto& t : workers) t.join(); }
};
void imr_copy_pool_release(imrcd_ctx* ctx) { delete static_cast<CopyPool*>(ctx->copy_pool); ctx->copy_pool = nullptr; }
static inline void staging_copy(imrcd_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes < (512u << 10)) { memcpy(dst, src, bytes); return; }
    if (!ctx->copy_pool) {
        const unsigned hw = std::thread::hardware_concurrency();
        const char* ev = getenv("IMRCD_COPY_THREADS");
        const int n = ev ? atoi(ev) : (hw >= 8 ? 3 : (hw >= 4 ? 1 : 0));
        ctx->copy_pool = new CopyPool(n < 0 ? 0 : n);
    }
    static_cast<CopyPool*>(ctx->copy_pool)->copy(dst, src, bytes);
}

// copy k consecutive caller entries, starting at i of the call's arrays (caller index g0 + i), behind the entries kept so far
static void entries_append_run(imrcd_ctx* ctx, uint64_t i, uint64_t k, uint64_t g0, const float* current, const float* previous,
                               const uint32_t* mesh_ids, const uint8_t* should_callback, const uint32_t* entities) {
    float* pc = ctx->p_cur.as<float>(); float* pp = ctx->p_prev.as<float>();
    uint32_t* pm = ctx->p_mesh.as<uint32_t>(); uint32_t* pe = ctx->p_entity.as<uint32_t>(); uint8_t* pb = ctx->p_cb.as<uint8_t>();
    const uint64_t at = ctx->n_entries;
    staging_copy(ctx, pc + 16 * at, current + 16 * i, 64 * k);
    if (previous && previous != current) {
        if (!ctx->prev_distinct && at) memcpy(pp, pc, 64 * at);      // earlier entries of this frame had previous == current
        staging_copy(ctx, pp + 16 * at, previous + 16 * i, 64 * k); ctx->prev_distinct = true;
    }
    else if (ctx->prev_distinct) staging_copy(ctx, pp + 16 * at, current + 16 * i, 64 * k);
    memcpy(pm + at, mesh_ids + i, 4 * k);
    if (entities) memcpy(pe + at, entities + i, 4 * k); else for (uint64_t q = 0; q < k; ++q) pe[at + q] = (uint32_t)(g0 + i + q);
    if (should_callback) for (uint64_t q = 0; q < k; ++q) pb[at + q] = should_callback[i + q] ? 1 : 0; else memset(pb + at, 1, k);
    if (ctx->shard_n > 1) { uint32_t* pg = ctx->p_gidx.as<uint32_t>(); for (uint64_t q = 0; q < k; ++q) pg[at + q] = (uint32_t)(g0 + i + q); }
    ctx->n_entries = at + k;
}

extern "C" int imrcd_frame_add_entries(imrcd_ctx* ctx, uint64_t n, const float* current, const float* previous,
                                       const uint32_t* mesh_ids, const uint8_t* should_callback, const uint32_t* entities) {
    CHECK_CTX(ctx);
    if (n && (!current || !mesh_ids)) { ctx->err = "imrcd_frame_add_entries: bad argument"; return IMRCD_E_ARG; }
    const uint32_t n_meshes = (uint32_t)ctx->meshes.size();
    const uint64_t g0 = ctx->n_entries_global;
    if (g0 + n >= (1ull << 32) - 1) { ctx->err = "too many entries"; return IMRCD_E_ARG; }
    if (n == 0) return IMRCD_OK;
    for (uint64_t i = 0; i < n; ++i) if (mesh_ids[i] >= n_meshes) { ctx->err = "imrcd_frame_add_entries: unknown mesh id"; return IMRCD_E_ARG; }
    cudaSetDevice(ctx->device);
    const bool sharded = ctx->shard_n > 1;
    const bool bulk = n >= ENTRY_CHUNK / 4;      // single entries (the reference's call pattern) are sent by imrcd_frame_upload
    uint64_t n_flagged = 0;
    if (!sharded) {
        int rc = entries_reserve(ctx, ctx->n_entries + n, bulk);
        if (rc) return rc;
        for (uint64_t c0 = 0; c0 < n; c0 += ENTRY_CHUNK) {       // chunks: the DMA of one runs while the next is being written
            const uint64_t k = std::min<uint64_t>(n, c0 + ENTRY_CHUNK) - c0;
            entries_append_run(ctx, c0, k, g0, current, previous, mesh_ids, should_callback, entities);
            if (bulk) { rc = entries_send(ctx, ctx->n_entries); if (rc) return rc; }
        }
        if (should_callback) for (uint64_t i = 0; i < n; ++i) n_flagged += should_callback[i] ? 1 : 0; else n_flagged = n;
    } else {
        // count what this rank keeps, reserve once, then copy run by run: whole blocks it owns, and the flagged entries of the other blocks
        uint64_t keep = 0;
        std::vector<uint32_t>& blk_flagged = ctx->shard_blk_flagged;      // flagged entries per block of this call: the copy pass skips foreign blocks without any
        blk_flagged.clear();
        for (uint64_t i = 0; i < n; ) {
            const uint64_t g = g0 + i, blk_end = std::min<uint64_t>(n, i + (SHARD_BLOCK - g % SHARD_BLOCK));
            uint64_t f = 0;
            if (should_callback) for (uint64_t q = i; q < blk_end; ++q) f += should_callback[q] ? 1 : 0; else f = blk_end - i;
            n_flagged += f;
            keep += shard_owns(ctx, g) ? blk_end - i : f;
            blk_flagged.push_back((uint32_t)f);
            i = blk_end;
        }
        int rc = entries_reserve(ctx, ctx->n_entries + keep, bulk);
        if (rc) return rc;
        size_t bi = 0;
        for (uint64_t i = 0; i < n; ++bi) {
            const uint64_t g = g0 + i, blk_end = std::min<uint64_t>(n, i + (SHARD_BLOCK - g % SHARD_BLOCK));
            if (shard_owns(ctx, g)) entries_append_run(ctx, i, blk_end - i, g0, current, previous, mesh_ids, should_callback, entities);
            else if (blk_flagged[bi]) for (uint64_t q = i; q < blk_end; ++q) {
                if (should_callback && !should_callback[q]) continue;
                uint64_t r = q + 1;
                while (r < blk_end && (!should_callback || should_callback[r])) ++r;
                entries_append_run(ctx, q, r - q, g0, current, previous, mesh_ids, should_callback, entities);
                q = r - 1;
            }
            i = blk_end;
            if (bulk && ctx->n_entries - ctx->n_sent >= ENTRY_CHUNK) { rc = entries_send(ctx, ctx->n_entries); if (rc) return rc; }
        }
        if (bulk) { rc = entries_send(ctx, ctx->n_entries); if (rc) return rc; }
    }
    ctx->n_entries_global = g0 + n; ctx->n_flagged_global += n_flagged;
    ctx->uploaded = ctx->ran = ctx->fetched = false;
    return IMRCD_OK;
}

extern "C" int imrcd_frame_map_entries(imrcd_ctx* ctx, uint64_t n, float** current, float** previous, uint32_t** mesh_ids,
                                       uint8_t** should_callback, uint32_t** entities) {
    CHECK_CTX(ctx);
    const uint64_t base = ctx->n_entries, total = base + n;
    if (ctx->n_entries_global + n >= (1ull << 32) - 1) { ctx->err = "too many entries"; return IMRCD_E_ARG; }
    cudaSetDevice(ctx->device);
    int rc = entries_reserve(ctx, total ? total : 1, false);
    if (rc) return rc;
    if (current) *current = ctx->p_cur.as<float>() + 16 * base;
    if (previous) *previous = ctx->p_prev.as<float>() + 16 * base;
    if (mesh_ids) *mesh_ids = ctx->p_mesh.as<uint32_t>() + base;
    if (should_callback) *should_callback = ctx->p_cb.as<uint8_t>() + base;
    if (entities) *entities = ctx->p_entity.as<uint32_t>() + base;
    return IMRCD_OK;
}

extern "C" int imrcd_frame_commit_entries(imrcd_ctx* ctx, uint64_t n, int previous_valid) {
    CHECK_CTX(ctx);
    const uint64_t base = ctx->n_entries, g0 = ctx->n_entries_global;
    if (64 * (base + n) > ctx->p_cur.cap) { ctx->err = "imrcd_frame_commit_entries: more entries than were mapped"; return IMRCD_E_STATE; }
    if (n == 0) return IMRCD_OK;
    cudaSetDevice(ctx->device);
    const uint32_t n_meshes = (uint32_t)ctx->meshes.size();
    uint32_t* pm = ctx->p_mesh.as<uint32_t>(); uint8_t* pb = ctx->p_cb.as<uint8_t>();
    for (uint64_t i = base; i < base + n; ++i) if (pm[i] >= n_meshes) { ctx->err = "imrcd_frame_commit_entries: unknown mesh id"; return IMRCD_E_ARG; }
    if (previous_valid) {
        if (!ctx->prev_distinct && base) memcpy(ctx->p_prev.p, ctx->p_cur.p, 64 * base);
        ctx->prev_distinct = true;
    } else if (ctx->prev_distinct) memcpy(ctx->p_prev.as<char>() + 64 * base, ctx->p_cur.as<char>() + 64 * base, 64 * n);
    uint64_t n_flagged = 0, kept = n;
    for (uint64_t i = base; i < base + n; ++i) n_flagged += pb[i] ? 1 : 0;
    if (ctx->shard_n > 1) {
        // the caller wrote all n entries; keep this rank's share by compacting the staging in place (destinations never pass sources)
        char* pc = ctx->p_cur.as<char>(); char* pp = ctx->p_prev.as<char>(); uint32_t* pe = ctx->p_entity.as<uint32_t>(); uint32_t* pg = ctx->p_gidx.as<uint32_t>();
        uint64_t w = base;
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t r = base + i;
            if (!(pb[r] || shard_owns(ctx, g0 + i))) continue;
            if (w != r) { memmove(pc + 64 * w, pc + 64 * r, 64); if (ctx->prev_distinct) memmove(pp + 64 * w, pp + 64 * r, 64); pm[w] = pm[r]; pe[w] = pe[r]; pb[w] = pb[r]; }
            pg[w] = (uint32_t)(g0 + i);
            ++w;
        }
        kept = w - base;
    }
    int rc = entries_reserve(ctx, base + kept, true);
    if (rc) return rc;
    ctx->n_entries = base + kept; ctx->n_entries_global = g0 + n; ctx->n_flagged_global += n_flagged;
    ctx->uploaded = ctx->ran = ctx->fetched = false;
    return entries_send(ctx, ctx->n_entries);          // the DMA starts now; imrcd_frame_upload has nothing left to copy
}

extern "C" int imrcd_frame_add_entry(imrcd_ctx* ctx, const float current[16], const float previous[16], uint32_t mesh_id,
                                     uint8_t should_callback, uint32_t entity) {
    return imrcd_frame_add_entries(ctx, 1, current, previous, &mesh_id, &should_callback, &entity);
}

extern "C" int imrcd_frame_set_shard(imrcd_ctx* ctx, uint32_t rank, uint32_t n_ranks) {
    CHECK_CTX(ctx);
    if (n_ranks == 0 || rank >= n_ranks) { ctx->err = "imrcd_frame_set_shard: bad rank"; return IMRCD_E_ARG; }
    if (ctx->comm && (n_ranks != ctx->comm_n || rank != ctx->comm_rank)) { ctx->err = "imrcd_frame_set_shard: the context's communicator fixes the shard"; return IMRCD_E_STATE; }
    ctx->shard_rank_next = rank; ctx->shard_n_next = n_ranks;
    if (ctx->n_entries_global == 0) { ctx->shard_rank = rank; ctx->shard_n = n_ranks; }      // else: from the next imrcd_frame_reset on
    return IMRCD_OK;
}

extern "C" int imrcd_frame_upload(imrcd_ctx* ctx) {
    CHECK_CTX(ctx);
    cudaSetDevice(ctx->device);
    int rc = meshes_sync(ctx);
    if (rc) return rc;
    const uint64_t n = ctx->n_entries;
    cudaStream_t s = ctx->stream;
    if (n) {
        IMR_CUDA(ctx, ctx->d_cur.reserve(64 * n, 64 * ctx->n_sent, s));
        IMR_CUDA(ctx, ctx->d_mesh.reserve(4 * n, 4 * ctx->n_sent, s));
        IMR_CUDA(ctx, ctx->d_entity.reserve(4 * n, 4 * ctx->n_sent, s));
        IMR_CUDA(ctx, ctx->d_cb.reserve(n, ctx->n_sent, s));
        if (ctx->shard_n > 1) IMR_CUDA(ctx, ctx->d_gidx.reserve(4 * n, 4 * ctx->n_sent, s));
        rc = entries_send(ctx, n);
        if (rc) return rc;
        // previous matrices feed only the response stage (CollisionDetection.cpp:80-98); an entry added with previous == NULL
        // or == current has not moved, and a frame where that holds for every entry uploads nothing for them
        if (ctx->prev_distinct) {
            IMR_CUDA(ctx, ctx->d_prev.reserve(64 * n, 0, s));
            IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_prev.p, ctx->p_prev.p, 64 * n, cudaMemcpyHostToDevice, s));
        }
    }
    ctx->uploaded = true; ctx->ran = ctx->fetched = false;
    return IMRCD_OK;
}

extern "C" int imrcd_frame_run(imrcd_ctx* ctx) {
    CHECK_CTX(ctx);
    if (!ctx->uploaded) { ctx->err = "imrcd_frame_run before imrcd_frame_upload"; return IMRCD_E_STATE; }
    cudaSetDevice(ctx->device);
    int rc = imr_frame_run_device(ctx);
    if (rc) return rc;
    ctx->ran = true; ctx->fetched = false;
    return IMRCD_OK;
}

extern "C" int imrcd_frame_run_async(imrcd_ctx* ctx) {
    CHECK_CTX(ctx);
    if (!ctx->uploaded) { ctx->err = "imrcd_frame_run_async before imrcd_frame_upload"; return IMRCD_E_STATE; }
    cudaSetDevice(ctx->device);
    ctx->enqueue_only = true;
    const int rc = imr_frame_run_device(ctx);
    ctx->enqueue_only = false;
    if (rc) return rc;
    ctx->async_pending = true; ctx->ran = false; ctx->fetched = false;
    return IMRCD_OK;
}
extern "C" int imrcd_frame_finish(imrcd_ctx* ctx) {
    CHECK_CTX(ctx);
    if (!ctx->async_pending) { ctx->err = "imrcd_frame_finish without imrcd_frame_run_async"; return IMRCD_E_STATE; }
    cudaSetDevice(ctx->device);
    ctx->async_pending = false;
    const int rc = imr_frame_finish_device(ctx);
    if (rc < 0) return rc;
    ctx->ran = true; ctx->fetched = false;
    return rc;                                            // 1: the frame was re-run with larger buffers (anything enqueued behind it saw stale results)
}

const imrcd_entity_pair* imr_comm_merged(const imrcd_ctx* ctx);
const imrcd_entity_pair* imr_comm_merged_device(const imrcd_ctx* ctx);

extern "C" int imrcd_frame_fetch(imrcd_ctx* ctx) {
    CHECK_CTX(ctx);
    if (!ctx->ran) { ctx->err = "imrcd_frame_fetch before imrcd_frame_run"; return IMRCD_E_STATE; }
    cudaSetDevice(ctx->device);
    const uint64_t nc = ctx->ctl_host.n_colliding;
    if (!ctx->comm && nc > ctx->spec_rows_sent) {        // more records than the speculative copy behind the frame brought: the rest
        const uint64_t have = ctx->spec_rows_sent;
        IMR_CUDA(ctx, ctx->p_epairs.reserve(sizeof(imrcd_entity_pair) * nc, sizeof(imrcd_entity_pair) * have, ctx->stream));
        IMR_CUDA(ctx, cudaMemcpyAsync(ctx->p_epairs.as<imrcd_entity_pair>() + have, ctx->d_epairs.as<imrcd_entity_pair>() + 1 + have, sizeof(imrcd_entity_pair) * (nc - have), cudaMemcpyDeviceToHost, ctx->stream));
        IMR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->spec_rows_sent = nc;
    }
    ctx->fetched = true;
    return IMRCD_OK;
}

extern "C" int imrcd_frame_execute(imrcd_ctx* ctx) {
    int rc = imrcd_frame_upload(ctx); if (rc) return rc;
    rc = imrcd_frame_run(ctx); if (rc) return rc;
    return imrcd_frame_fetch(ctx);
}

extern "C" int imrcd_frame_results(imrcd_ctx* ctx, const imrcd_entity_pair** pairs, uint64_t* n_pairs,
                                   const imrcd_tri_hit** hits, uint64_t* n_hits) {
    CHECK_CTX(ctx);
    if (!ctx->fetched) { ctx->err = "imrcd_frame_results before imrcd_frame_fetch/execute"; return IMRCD_E_STATE; }
    // with a communicator: the merged records of ALL ranks (in rank order); hits stay this rank's own
    if (pairs) *pairs = ctx->comm ? (ctx->n_merged ? imr_comm_merged(ctx) : nullptr) : ctx->p_epairs.as<imrcd_entity_pair>();
    if (n_pairs) *n_pairs = ctx->comm ? ctx->n_merged : ctx->ctl_host.n_colliding;
    if (hits) {   // the hit list is an intermediate of the contact reduction; copied to the host only on demand
        const uint64_t nh = ctx->ctl_host.n_hits;
        if (!ctx->hits_fetched && nh) {
            cudaSetDevice(ctx->device);
            IMR_CUDA(ctx, ctx->p_hits.reserve(sizeof(imrcd_tri_hit) * nh));
            IMR_CUDA(ctx, cudaMemcpyAsync(ctx->p_hits.p, ctx->d_hits.p, sizeof(imrcd_tri_hit) * nh, cudaMemcpyDeviceToHost, ctx->stream));
            IMR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->hits_fetched = true;
        }
        *hits = ctx->p_hits.as<imrcd_tri_hit>();
    }
    if (n_hits) *n_hits = ctx->ctl_host.n_hits;
    return IMRCD_OK;
}

extern "C" int imrcd_frame_results_local(imrcd_ctx* ctx, const imrcd_entity_pair** pairs, uint64_t* n_pairs) {
    CHECK_CTX(ctx);
    if (!ctx->ran) { ctx->err = "imrcd_frame_results_local before imrcd_frame_run"; return IMRCD_E_STATE; }
    cudaSetDevice(ctx->device);
    const uint64_t nc = ctx->ctl_host.n_colliding;
    if (ctx->comm && nc) {        // the speculative copy of a frame with a communicator carries the merged records, not the local ones
        IMR_CUDA(ctx, ctx->p_epairs.reserve(sizeof(imrcd_entity_pair) * nc));
        IMR_CUDA(ctx, cudaMemcpyAsync(ctx->p_epairs.p, ctx->d_epairs.as<imrcd_entity_pair>() + 1, sizeof(imrcd_entity_pair) * nc, cudaMemcpyDeviceToHost, ctx->stream));
        IMR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else if (!ctx->comm) { const int rc = imrcd_frame_fetch(ctx); if (rc) return rc; }
    if (pairs) *pairs = ctx->p_epairs.as<imrcd_entity_pair>();
    if (n_pairs) *n_pairs = nc;
    return IMRCD_OK;
}

extern "C" int imrcd_frame_pairs(imrcd_ctx* ctx, const uint32_t** pairs, uint64_t* n_pairs) {
    CHECK_CTX(ctx);
    if (!ctx->ran) { ctx->err = "imrcd_frame_pairs before imrcd_frame_run"; return IMRCD_E_STATE; }
    const uint64_t n = ctx->ctl_host.n_pairs;
    if (n) {
        cudaSetDevice(ctx->device);
        IMR_CUDA(ctx, ctx->p_pairs.reserve(8 * n));
        IMR_CUDA(ctx, cudaMemcpyAsync(ctx->p_pairs.p, ctx->d_pairs.p, 8 * n, cudaMemcpyDeviceToHost, ctx->stream));
        IMR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->shard_n > 1) {                            // entry slots of this shard -> the caller's entry indices
            uint32_t* q = ctx->p_pairs.as<uint32_t>(); const uint32_t* g = ctx->p_gidx.as<uint32_t>();
            for (uint64_t i = 0; i < 2 * n; ++i) q[i] = g[q[i]];
        }
    }
    if (pairs) *pairs = ctx->p_pairs.as<uint32_t>();
    if (n_pairs) *n_pairs = n;
    return IMRCD_OK;
}

extern "C" int imrcd_frame_combos(imrcd_ctx* ctx, const uint32_t** combos, uint64_t* n_combos) {
    CHECK_CTX(ctx);
    if (!ctx->ran) { ctx->err = "imrcd_frame_combos before imrcd_frame_run"; return IMRCD_E_STATE; }
    const uint64_t n = ctx->ctl_host.n_combos;
    if (n) {
        cudaSetDevice(ctx->device);
        IMR_CUDA(ctx, ctx->p_combos.reserve(16 * n));
        IMR_CUDA(ctx, cudaMemcpyAsync(ctx->p_combos.p, ctx->d_combos.p, 16 * n, cudaMemcpyDeviceToHost, ctx->stream));
        IMR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (combos) *combos = ctx->p_combos.as<uint32_t>();
    if (n_combos) *n_combos = n;
    return IMRCD_OK;
}

extern "C" int imrcd_frame_get_stats(imrcd_ctx* ctx, imrcd_frame_stats* out) {
    CHECK_CTX(ctx);
    if (!out) return IMRCD_E_ARG;
    *out = ctx->stats;
    return IMRCD_OK;
}

extern "C" int imrcd_frame_results_device(imrcd_ctx* ctx, void** d_pairs, uint64_t* n_pairs, void** d_hits, uint64_t* n_hits) {
    CHECK_CTX(ctx);
    if (!ctx->ran) { ctx->err = "imrcd_frame_results_device before imrcd_frame_run"; return IMRCD_E_STATE; }
    if (d_pairs) *d_pairs = ctx->comm ? (void*)imr_comm_merged_device(ctx) : (void*)(ctx->d_epairs.as<imrcd_entity_pair>() + 1);      // merged records of all ranks when the context has a communicator
    if (n_pairs) *n_pairs = ctx->comm ? ctx->n_merged : ctx->ctl_host.n_colliding;
    if (d_hits) *d_hits = ctx->d_hits.p;
    if (n_hits) *n_hits = ctx->ctl_host.n_hits;
    return IMRCD_OK;
}

extern "C" int imrcd_frame_results_block(imrcd_ctx* ctx, void** d_block, uint64_t* n_pairs, uint64_t* capacity) {
    CHECK_CTX(ctx);
    if (!ctx->ran && !ctx->async_pending) { ctx->err = "imrcd_frame_results_block before imrcd_frame_run"; return IMRCD_E_STATE; }
    if (d_block) *d_block = ctx->d_epairs.p;
    if (n_pairs) *n_pairs = ctx->ctl_host.n_colliding;
    if (capacity) *capacity = ctx->d_epairs.p ? ctx->d_epairs.cap / sizeof(imrcd_entity_pair) - 1 : 0;
    return IMRCD_OK;
}

// ------------------------------------------------------------------------------------------
// unit-level test hooks: the same device functions the pipeline uses, on flat arrays
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ Box load_box12(const float* f) {
    Box b; b.c = mk3(f[0], f[1], f[2]); b.u = mk3(f[3], f[4], f[5]); b.v = mk3(f[6], f[7], f[8]); b.w = mk3(f[9], f[10], f[11]); return b;
}
__device__ __forceinline__ Rel load_rel16(const float* m) {
    Rel r; r.r0 = make_float4(m[0], m[4], m[8], m[12]); r.r1 = make_float4(m[1], m[5], m[9], m[13]); r.r2 = make_float4(m[2], m[6], m[10], m[14]); return r;
}
__global__ void k_test_sat(uint64_t n, const float* a, const float* b, const float* mats, uint8_t* verdict, float* sa, float* sb) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Box A = load_box12(a + 12 * i), B = load_box12(b + 12 * i);
    if (mats) B = box_transform(load_rel16(mats + 16 * i), B);
    verdict[i] = box_sat(A, B) ? 1 : 0;
    if (sa) sa[i] = box_surface(A);
    if (sb) sb[i] = box_surface(B);
}
__global__ void k_test_tri(uint64_t n, const float* a, const float* b, const float* mat, uint8_t* flags, float* seg) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = a + 9 * i; const float* q = b + 9 * i;
    V3 V0 = mk3(p[0], p[1], p[2]), V1 = mk3(p[3], p[4], p[5]), V2 = mk3(p[6], p[7], p[8]);
    V3 U0 = mk3(q[0], q[1], q[2]), U1 = mk3(q[3], q[4], q[5]), U2 = mk3(q[6], q[7], q[8]);
    if (mat) { Rel r = load_rel16(mat); U0 = rel_mul(r, U0, 1.f); U1 = rel_mul(r, U1, 1.f); U2 = rel_mul(r, U2, 1.f); }
    V3 s = mk3(0, 0, 0), t = mk3(0, 0, 0);
    int f = tri_tri_isectline(V0, V1, V2, U0, U1, U2, s, t);
    flags[i] = (uint8_t)f;    // bit0 doIntersept, bit1 areCoplanar
    if (f == 1) { float* o = seg + 6 * i; o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = t.x; o[4] = t.y; o[5] = t.z; }
}
__global__ void k_test_pair_matrix(uint64_t n, const float* a, const float* b, float* out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float inv[16];
    mat4_inverse(a + 16 * i, inv);
    mat4_mul(inv, b + 16 * i, out + 16 * i);
}

template <class F>
static int with_device_arrays(imrcd_ctx* ctx, std::vector<std::pair<const void*, size_t>> ins, std::vector<std::pair<void*, size_t>> outs, F launch) {
    cudaSetDevice(ctx->device);
    std::vector<void*> din(ins.size(), nullptr), dout(outs.size(), nullptr);
    int rc = IMRCD_OK;
    auto cleanup = [&]() { for (void* p : din) if (p) cudaFree(p); for (void* p : dout) if (p) cudaFree(p); };
    for (size_t i = 0; i < ins.size(); ++i) {
        if (!ins[i].first || !ins[i].second) continue;
        if (cudaMalloc(&din[i], ins[i].second) != cudaSuccess || cudaMemcpyAsync(din[i], ins[i].first, ins[i].second, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { ctx->err = "test hook: alloc/copy failed"; cleanup(); return IMRCD_E_CUDA; }
    }
    for (size_t i = 0; i < outs.size(); ++i) {
        if (!outs[i].first || !outs[i].second) continue;
        if (cudaMalloc(&dout[i], outs[i].second) != cudaSuccess || cudaMemsetAsync(dout[i], 0, outs[i].second, ctx->stream) != cudaSuccess) { ctx->err = "test hook: alloc failed"; cleanup(); return IMRCD_E_CUDA; }
    }
    launch(din, dout);
    if (cudaGetLastError() != cudaSuccess) { ctx->err = "test hook: launch failed"; rc = IMRCD_E_CUDA; }
    for (size_t i = 0; i < outs.size() && rc == IMRCD_OK; ++i)
        if (dout[i] && cudaMemcpyAsync(outs[i].first, dout[i], outs[i].second, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { ctx->err = "test hook: copy back failed"; rc = IMRCD_E_CUDA; }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = std::string("test hook: ") + cudaGetErrorString(cudaGetLastError()); rc = IMRCD_E_CUDA; }
    cleanup();
    return rc;
}

extern "C" int imrcd_test_sat(imrcd_ctx* ctx, uint64_t n, const float* boxes_a, const float* boxes_b, const float* mats,
                              uint8_t* verdict, float* surface_a, float* surface_b) {
    CHECK_CTX(ctx);
    if (!n) return IMRCD_OK;
    return with_device_arrays(ctx, { {boxes_a, 48 * n}, {boxes_b, 48 * n}, {mats, mats ? 64 * n : 0} },
                              { {verdict, n}, {surface_a, surface_a ? 4 * n : 0}, {surface_b, surface_b ? 4 * n : 0} },
                              [&](std::vector<void*>& i, std::vector<void*>& o) {
                                  k_test_sat<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(n, (const float*)i[0], (const float*)i[1], (const float*)i[2],
                                                                                                   (uint8_t*)o[0], (float*)o[1], (float*)o[2]);
                              });
}
extern "C" int imrcd_test_tri_tri(imrcd_ctx* ctx, uint64_t n, const float* tris_a, const float* tris_b, const float* mat16,
                                  uint8_t* flags, float* seg) {
    CHECK_CTX(ctx);
    if (!n) return IMRCD_OK;
    return with_device_arrays(ctx, { {tris_a, 36 * n}, {tris_b, 36 * n}, {mat16, mat16 ? 64 : 0} }, { {flags, n}, {seg, 24 * n} },
                              [&](std::vector<void*>& i, std::vector<void*>& o) {
                                  k_test_tri<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(n, (const float*)i[0], (const float*)i[1], (const float*)i[2],
                                                                                                   (uint8_t*)o[0], (float*)o[1]);
                              });
}
extern "C" int imrcd_test_ray_tree(imrcd_ctx* ctx, uint32_t mesh_id, uint64_t n, const float* mats, const float* origins, const float* directions,
                                   uint8_t* flags, float* out3, uint32_t* tri) {
    CHECK_CTX(ctx);
    if (mesh_id >= ctx->meshes.size()) { ctx->err = "bad mesh id"; return IMRCD_E_ARG; }
    if (!n) return IMRCD_OK;
    int inner = IMRCD_OK;
    int rc = with_device_arrays(ctx, { {mats, 64 * n}, {origins, 12 * n}, {directions, 12 * n} }, { {flags, n}, {out3, 12 * n}, {tri, 4 * n} },
                                [&](std::vector<void*>& i, std::vector<void*>& o) {
                                    inner = imr_test_ray_tree_device(ctx, mesh_id, n, (const float*)i[0], (const float*)i[1], (const float*)i[2],
                                                                     (uint8_t*)o[0], (float*)o[1], (uint32_t*)o[2]);
                                });
    return inner != IMRCD_OK ? inner : rc;
}
extern "C" int imrcd_test_pair_matrix(imrcd_ctx* ctx, uint64_t n, const float* a, const float* b, float* out) {
    CHECK_CTX(ctx);
    if (!n) return IMRCD_OK;
    return with_device_arrays(ctx, { {a, 64 * n}, {b, 64 * n} }, { {out, 64 * n} },
                              [&](std::vector<void*>& i, std::vector<void*>& o) {
                                  k_test_pair_matrix<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(n, (const float*)i[0], (const float*)i[1], (float*)o[0]);
                              });
}
