// imrcd_comm.cu -- the end-of-frame merge of a sharded frame inside the library: ONE ncclAllGather of fixed-capacity
// blocks on the frame's own stream, right behind the frame's kernels, then a compaction of the gathered blocks and a speculative D2H of
// the merged records, so that the host waits once per frame.  The reference has no counterpart (single-threaded host loop,
// CollisionDetection.cpp:44-129); what is kept is its contract: after ExecuteCollisionDetection the caller sees the complete colliding set.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library loads and every single-GPU entry point works on a machine without NCCL,
// and inside a process that already carries an NCCL (torch's bundled one) that same copy is used.
#include "imrcd_internal.cuh"
#include <nccl.h>            // types and signatures only; no symbol of libnccl is linked
#include <dlfcn.h>
#include <algorithm>

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string err;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (const char* nm : names) { api.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (api.handle) break; }
    if (!api.handle) { api.err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : ""); return nullptr; }
#define BIND(f) api.f = reinterpret_cast<decltype(api.f)>(dlsym(api.handle, "nccl" #f)); if (!api.f) { api.err = "libnccl lacks nccl" #f; api.handle = nullptr; return nullptr; }
    BIND(GetUniqueId) BIND(CommInitRank) BIND(CommInitAll) BIND(CommDestroy) BIND(AllGather) BIND(GroupStart) BIND(GroupEnd) BIND(GetErrorString)
#undef BIND
    return &api;
}

#define IMR_NCCL(ctx, api, call) do { ncclResult_t _r = (call); if (_r != ncclSuccess) { (ctx)->err = std::string(#call) + ": " + (api)->GetErrorString(_r); return IMRCD_E_CUDA; } } while (0)

static_assert(sizeof(ncclUniqueId) == IMRCD_COMM_ID_BYTES, "IMRCD_COMM_ID_BYTES");

extern "C" int imrcd_comm_unique_id(void* id_out) {
    if (!id_out) return IMRCD_E_ARG;
    NcclApi* api = nccl_api();
    if (!api) return IMRCD_E_NODEVICE;
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return IMRCD_E_CUDA;
    memcpy(id_out, &id, sizeof(id));
    return IMRCD_OK;
}

static int comm_attach(imrcd_ctx* ctx, ncclComm_t comm, uint32_t rank, uint32_t n) {
    ctx->comm = comm; ctx->comm_rank = rank; ctx->comm_n = n;
    ctx->shard_rank_next = rank; ctx->shard_n_next = n;
    if (ctx->n_entries_global == 0) { ctx->shard_rank = rank; ctx->shard_n = n; }
    if (ctx->gcap == 0) ctx->gcap = 1024;
    return IMRCD_OK;
}

extern "C" int imrcd_comm_init(imrcd_ctx* ctx, const void* id, uint32_t rank, uint32_t n_ranks) {
    if (!ctx) return IMRCD_E_ARG;
    if (!id || n_ranks == 0 || rank >= n_ranks) { ctx->err = "imrcd_comm_init: bad argument"; return IMRCD_E_ARG; }
    if (ctx->comm) { ctx->err = "imrcd_comm_init: the context already has a communicator"; return IMRCD_E_STATE; }
    NcclApi* api = nccl_api();
    if (!api) { ctx->err = "imrcd_comm_init: NCCL is not available (libnccl.so.2 could not be loaded)"; return IMRCD_E_NODEVICE; }
    cudaSetDevice(ctx->device);
    ncclUniqueId uid; memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    IMR_NCCL(ctx, api, api->CommInitRank(&comm, (int)n_ranks, uid, (int)rank));
    return comm_attach(ctx, comm, rank, n_ranks);
}

extern "C" int imrcd_comm_destroy(imrcd_ctx* ctx) {
    if (!ctx) return IMRCD_E_ARG;
    if (!ctx->comm) return IMRCD_OK;
    NcclApi* api = nccl_api();
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    // a captured frame holds the communicator's collective as a graph node, and NCCL waits for such graphs to be gone before it lets a
    // communicator go: the graph first
    if (ctx->graph_exec) { cudaGraphExecDestroy(ctx->graph_exec); ctx->graph_exec = nullptr; ctx->graph_key = 0; ctx->graph_seen_key = 0; }
    if (api) api->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr; ctx->comm_n = 1; ctx->comm_rank = 0;
    ctx->shard_rank_next = 0; ctx->shard_n_next = 1;
    if (ctx->n_entries_global == 0) { ctx->shard_rank = 0; ctx->shard_n = 1; }
    return IMRCD_OK;
}

// ---- the gathered blocks -> one dense array -----------------------------------------------------------------------------------------
// d_gather = [ comm_n blocks of (gcap + 1) rows as they arrive | GatherHdr[comm_n], padded to whole rows | dense records ].
// Row 0 of a block: u64 record count of that rank, then its overflow bits (k_epairs_header).
struct GatherHdr { unsigned long long count; unsigned long long flags; };
static inline uint64_t hdr_rows(uint32_t n) { return (sizeof(GatherHdr) * n + sizeof(imrcd_entity_pair) - 1) / sizeof(imrcd_entity_pair); }

__global__ void __launch_bounds__(256)
k_gather_compact(const uint4* __restrict__ blocks, uint32_t n_ranks, unsigned long long gcap, GatherHdr* __restrict__ hdr, uint4* __restrict__ dense) {
    const uint32_t r = blockIdx.y;
    const unsigned long long rows_per_block = gcap + 1ull;
    unsigned long long off = 0, mine = 0, flags = 0;
    for (uint32_t q = 0; q <= r; ++q) {
        const uint4 h = blocks[q * rows_per_block * 5ull];
        const unsigned long long c = (unsigned long long)h.x | ((unsigned long long)h.y << 32);
        if (q < r) off += c < gcap ? c : gcap; else { mine = c; flags = h.z; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { hdr[r].count = mine; hdr[r].flags = flags; }
    const unsigned long long k = mine < gcap ? mine : gcap;
    const uint4* src = blocks + (r * rows_per_block + 1ull) * 5ull;
    uint4* dst = dense + off * 5ull;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < k * 5ull; i += (unsigned long long)gridDim.x * blockDim.x) dst[i] = src[i];
}

static int gather_reserve(imrcd_ctx* ctx) {
    const uint64_t rows = (uint64_t)ctx->comm_n * (ctx->gcap + 1) + hdr_rows(ctx->comm_n) + (uint64_t)ctx->comm_n * ctx->gcap;
    IMR_CUDA(ctx, ctx->d_gather.reserve(rows * sizeof(imrcd_entity_pair), 0, ctx->stream));
    return IMRCD_OK;
}

int imr_comm_reserve(imrcd_ctx* ctx) {
    int rc = gather_reserve(ctx); if (rc) return rc;
    IMR_CUDA(ctx, ctx->p_gather.reserve((hdr_rows(ctx->comm_n) + (uint64_t)ctx->comm_n * ctx->gcap) * sizeof(imrcd_entity_pair)));
    return IMRCD_OK;
}

// part 1: the collective itself (the send block is the library's own result block: header row + records, as it lies in HBM)
int imr_comm_allgather(imrcd_ctx* ctx) {
    NcclApi* api = nccl_api();
    if (!api || !ctx->comm) { ctx->err = "no communicator"; return IMRCD_E_STATE; }
    if (ctx->d_epairs.cap < sizeof(imrcd_entity_pair) * (ctx->gcap + 1)) { ctx->err = "result block smaller than the gather capacity"; return IMRCD_E_STATE; }      // imr_frame_enqueue sizes it
    IMR_NCCL(ctx, api, api->AllGather(ctx->d_epairs.p, ctx->d_gather.p, (ctx->gcap + 1) * sizeof(imrcd_entity_pair), ncclUint8, static_cast<ncclComm_t>(ctx->comm), ctx->stream));
    return IMRCD_OK;
}

// part 2: compaction + speculative D2H (headers, and as many records as the last frame had plus a margin)
int imr_comm_after_gather(imrcd_ctx* ctx, uint64_t spec_rows) {
    cudaStream_t s = ctx->stream;
    const uint64_t blocks_rows = (uint64_t)ctx->comm_n * (ctx->gcap + 1), hr = hdr_rows(ctx->comm_n);
    imrcd_entity_pair* base = ctx->d_gather.as<imrcd_entity_pair>();
    k_gather_compact<<<dim3(8, ctx->comm_n), 256, 0, s>>>(reinterpret_cast<const uint4*>(base), ctx->comm_n, ctx->gcap,
                                                          reinterpret_cast<GatherHdr*>(base + blocks_rows), reinterpret_cast<uint4*>(base + blocks_rows + hr));
    IMR_CUDA(ctx, cudaGetLastError());
    const uint64_t max_rows = (uint64_t)ctx->comm_n * ctx->gcap;
    spec_rows = std::min<uint64_t>(spec_rows, max_rows);
    IMR_CUDA(ctx, cudaMemcpyAsync(ctx->p_gather.p, base + blocks_rows, (hr + spec_rows) * sizeof(imrcd_entity_pair), cudaMemcpyDeviceToHost, s));
    return IMRCD_OK;
}

// after the host has waited: what the gathered headers say.  The decision is a function of the headers alone, so every rank takes the same one.
//   *retry    some rank's frame overflowed a buffer, or some rank has more records than the blocks hold (gcap is raised): every rank runs
//             the frame and the collective again
//   *fatal    some rank hit a limit that re-running cannot lift
int imr_comm_decide(imrcd_ctx* ctx, uint64_t spec_rows, bool* retry, bool* fatal) {
    const GatherHdr* h = ctx->p_gather.as<GatherHdr>();
    *retry = false; *fatal = false;
    unsigned long long mx = 0, total = 0;
    for (uint32_t r = 0; r < ctx->comm_n; ++r) {
        if (h[r].flags & OVF_RAYSTACK) *fatal = true;
        if (h[r].flags) *retry = true;
        mx = std::max(mx, h[r].count); total += h[r].count;
    }
    if (mx > ctx->gcap) { while (ctx->gcap < mx) ctx->gcap *= 2; *retry = true; }
    if (*retry || *fatal) return IMRCD_OK;
    const uint64_t hr = hdr_rows(ctx->comm_n);
    spec_rows = std::min<uint64_t>(spec_rows, (uint64_t)ctx->comm_n * ctx->gcap);
    if (total > spec_rows) {                 // more records than the speculative copy brought: fetch the rest
        const uint64_t blocks_rows = (uint64_t)ctx->comm_n * (ctx->gcap + 1);
        IMR_CUDA(ctx, cudaMemcpyAsync(ctx->p_gather.as<imrcd_entity_pair>() + hr + spec_rows, ctx->d_gather.as<imrcd_entity_pair>() + blocks_rows + hr + spec_rows,
                                      (total - spec_rows) * sizeof(imrcd_entity_pair), cudaMemcpyDeviceToHost, ctx->stream));
        IMR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->n_merged = total; ctx->merged_valid = true;
    return IMRCD_OK;
}
const imrcd_entity_pair* imr_comm_merged(const imrcd_ctx* ctx) { return ctx->p_gather.as<imrcd_entity_pair>() + hdr_rows(ctx->comm_n); }
const imrcd_entity_pair* imr_comm_merged_device(const imrcd_ctx* ctx) {
    return ctx->d_gather.as<imrcd_entity_pair>() + (uint64_t)ctx->comm_n * (ctx->gcap + 1) + hdr_rows(ctx->comm_n);
}

// ---- one process, several GPUs (the engine is a single process: SURVEY 8b "imrcd_create(device_ids[], n)") ---------------------------
int imr_frame_enqueue(imrcd_ctx* ctx);
int imr_frame_reserve(imrcd_ctx* ctx);
void imr_frame_begin(imrcd_ctx* ctx);
int imr_frame_complete(imrcd_ctx* ctx, bool* retry);
uint64_t imr_frame_spec_rows(const imrcd_ctx* ctx);

struct imrcd_group { std::vector<imrcd_ctx*> ctx; std::string err; };

extern "C" int imrcd_group_create(const int* device_ids, uint32_t n, imrcd_group** out) {
    if (!out || !device_ids || n == 0) return IMRCD_E_ARG;
    *out = nullptr;
    imrcd_group* g = new imrcd_group();
    for (uint32_t i = 0; i < n; ++i) {
        imrcd_ctx* c = nullptr;
        const int rc = imrcd_create(device_ids[i], nullptr, &c);
        if (rc) { for (imrcd_ctx* k : g->ctx) imrcd_destroy(k); delete g; return rc; }
        g->ctx.push_back(c);
    }
    if (n > 1) {
        NcclApi* api = nccl_api();
        std::vector<ncclComm_t> comms(n);
        if (!api || api->CommInitAll(comms.data(), (int)n, device_ids) != ncclSuccess) { for (imrcd_ctx* k : g->ctx) imrcd_destroy(k); delete g; return IMRCD_E_NODEVICE; }
        for (uint32_t i = 0; i < n; ++i) comm_attach(g->ctx[i], comms[i], i, n);
    }
    *out = g;
    return IMRCD_OK;
}
extern "C" void imrcd_group_destroy(imrcd_group* g) {
    if (!g) return;
    for (imrcd_ctx* c : g->ctx) { imrcd_comm_destroy(c); imrcd_destroy(c); }
    delete g;
}
extern "C" uint32_t imrcd_group_size(const imrcd_group* g) { return g ? (uint32_t)g->ctx.size() : 0u; }
extern "C" imrcd_ctx* imrcd_group_ctx(imrcd_group* g, uint32_t i) { return (g && i < g->ctx.size()) ? g->ctx[i] : nullptr; }
extern "C" const char* imrcd_group_last_error(const imrcd_group* g) { return g ? g->err.c_str() : "null group"; }

#define GROUP_EACH(g, call) do { for (imrcd_ctx* c : (g)->ctx) { const int _rc = (call); if (_rc) { (g)->err = imrcd_last_error(c); return _rc; } } } while (0)

extern "C" int imrcd_group_mesh_create(imrcd_group* g, const float* positions, const float* normals, const uint32_t* vertex_ids, uint64_t n_tri,
                                       uint32_t build_mode, uint32_t* mesh_id) {
    if (!g || !mesh_id) return IMRCD_E_ARG;
    uint32_t first = 0; bool have = false;
    for (imrcd_ctx* c : g->ctx) {
        uint32_t id = 0;
        const int rc = imrcd_mesh_create(c, positions, normals, vertex_ids, n_tri, build_mode, &id);
        if (rc) { g->err = imrcd_last_error(c); return rc; }
        if (!have) { first = id; have = true; } else if (id != first) { g->err = "imrcd_group_mesh_create: the contexts of the group have diverged"; return IMRCD_E_STATE; }
    }
    *mesh_id = first;
    return IMRCD_OK;
}
extern "C" int imrcd_group_gltf_load(imrcd_group* g, const char* path, uint32_t build_mode, uint32_t* mesh_ids, uint32_t capacity, uint32_t* n_meshes) {
    if (!g) return IMRCD_E_ARG;
    GROUP_EACH(g, imrcd_gltf_load(c, path, build_mode, mesh_ids, capacity, n_meshes));      // every context builds the same ids in the same order
    return IMRCD_OK;
}
extern "C" int imrcd_group_frame_reset(imrcd_group* g) { if (!g) return IMRCD_E_ARG; GROUP_EACH(g, imrcd_frame_reset(c)); return IMRCD_OK; }
extern "C" int imrcd_group_frame_add_entries(imrcd_group* g, uint64_t n, const float* current, const float* previous, const uint32_t* mesh_ids,
                                             const uint8_t* should_callback, const uint32_t* entities) {
    if (!g) return IMRCD_E_ARG;
    GROUP_EACH(g, imrcd_frame_add_entries(c, n, current, previous, mesh_ids, should_callback, entities));     // each context keeps its share
    return IMRCD_OK;
}
extern "C" int imrcd_group_frame_add_entry(imrcd_group* g, const float current[16], const float previous[16], uint32_t mesh_id, uint8_t should_callback, uint32_t entity) {
    return imrcd_group_frame_add_entries(g, 1, current, previous, &mesh_id, &should_callback, &entity);
}

// ExecuteCollisionDetection on every GPU of the group: each device's frame is enqueued, the collectives of all devices go out in one NCCL
// group call, the host waits once per device.
extern "C" int imrcd_group_frame_execute(imrcd_group* g) {
    if (!g) return IMRCD_E_ARG;
    if (g->ctx.size() == 1) { GROUP_EACH(g, imrcd_frame_execute(c)); return IMRCD_OK; }
    NcclApi* api = nccl_api();
    GROUP_EACH(g, imrcd_frame_upload(c));
    for (imrcd_ctx* c : g->ctx) imr_frame_begin(c);
    if (g->ctx[0]->n_entries_global < 2) { for (imrcd_ctx* c : g->ctx) { c->ran = true; c->fetched = true; c->merged_valid = true; c->n_merged = 0; } return IMRCD_OK; }
    for (int attempt = 0; attempt < 10; ++attempt) {
        GROUP_EACH(g, (cudaSetDevice(c->device), imr_frame_reserve(c)));
        GROUP_EACH(g, (cudaSetDevice(c->device), imr_frame_enqueue(c)));
        if (api->GroupStart() != ncclSuccess) { g->err = "ncclGroupStart"; return IMRCD_E_CUDA; }
        GROUP_EACH(g, (cudaSetDevice(c->device), imr_comm_allgather(c)));
        if (api->GroupEnd() != ncclSuccess) { g->err = "ncclGroupEnd"; return IMRCD_E_CUDA; }
        GROUP_EACH(g, (cudaSetDevice(c->device), c->spec_rows_sent = imr_frame_spec_rows(c), imr_comm_after_gather(c, c->spec_rows_sent)));
        bool any_retry = false;
        for (imrcd_ctx* c : g->ctx) {
            bool retry = false;
            cudaSetDevice(c->device);
            const int rc = imr_frame_complete(c, &retry);
            if (rc) { g->err = imrcd_last_error(c); return rc; }
            any_retry |= retry;
        }
        if (!any_retry) { for (imrcd_ctx* c : g->ctx) { c->ran = true; c->fetched = true; } return IMRCD_OK; }
    }
    g->err = "frame buffers could not be grown enough";
    return IMRCD_E_CAPACITY;
}
extern "C" int imrcd_group_frame_results(imrcd_group* g, const imrcd_entity_pair** pairs, uint64_t* n_pairs) {
    if (!g) return IMRCD_E_ARG;
    return imrcd_frame_results(g->ctx[0], pairs, n_pairs, nullptr, nullptr);
}
