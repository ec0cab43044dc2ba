// imrcd_comm.cu -- the end-of-frame merge of a sharded frame inside the library: ONE ncclAllGather of fixed-capacity
// blocks on the frame's own stream, right behind the frame's kernels, then a compaction of the gathered blocks and a speculative D2H of
// the merged records, so that the host waits once per frame.  The reference has no counterpart (single-threaded host loop,
// CollisionDetection.cpp:44-129); what is kept is its contract: after ExecuteCollisionDetection the caller sees the complete colliding set.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library loads and every single-GPU entry point works on a machine without NCCL,
// and inside a process that already carries an NCCL (torch's bundled one) that same copy is used.
//
// Where the GPUs can reach each other's memory (NVLink / NVSwitch: every B200 of a node) the per-frame exchange does not go through NCCL
// at all: the gather is part of the frame's own kernels.  Each rank owns a buffer with a slot per rank and arrival flags, mapped into
// every peer (CUDA IPC between processes, peer access inside one); k_p2p_push, right behind the frame's last kernel, stores the rank's
// block (header row + exactly the records it has, not the fixed capacity) into its slot of EVERY peer's buffer and then raises its flag
// there; k_p2p_wait_compact on each rank waits for the flags it needs and compacts the blocks in the same launch.  Two launches and one
// NVLink store latency instead of a collective's rendezvous; NCCL stays for setting the communicator up, for exchanging the buffers'
// handles, and as the path of machines without peer access (IMRCD_P2P=0 forces it).
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
void imr_comm_p2p_release(imrcd_ctx* ctx);

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
    imr_comm_p2p_release(ctx);
    if (api) api->CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr; ctx->comm_n = 1; ctx->comm_rank = 0;
    ctx->shard_rank_next = 0; ctx->shard_n_next = 1;
    if (ctx->n_entries_global == 0) { ctx->shard_rank = 0; ctx->shard_n = 1; }
    return IMRCD_OK;
}

extern "C" int imrcd_comm_transport(const imrcd_ctx* ctx) {
    if (!ctx || !ctx->comm) return 0;
    return ctx->p2p_state == 1 ? 2 : 1;
}

// ---- the gathered blocks -> one dense array -----------------------------------------------------------------------------------------
// d_gather = [ comm_n blocks of (gcap + 1) rows as they arrive | GatherHdr[comm_n], padded to whole rows | dense records ].
// Row 0 of a block: u64 record count of that rank, then its overflow bits (k_epairs_header).
struct GatherHdr { unsigned long long count; unsigned long long flags; };
#define P2P_TIMEOUT_BIT 0x80000000ull      // in GatherHdr::flags: a peer's block did not arrive within the time limit
void imr_comm_p2p_release(imrcd_ctx* ctx);
bool imr_comm_uses_p2p(const imrcd_ctx* ctx) { return ctx->comm && ctx->p2p_state == 1 && ctx->p2p_gcap == ctx->gcap; }
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

// ---- the merge over peer memory ------------------------------------------------------------------------------------------------------
#define P2P_MAX_RANKS 64u
#define P2P_FLAG_STRIDE 128u
struct P2PCtl {
    unsigned long long seq;                // frame sequence number of the communicator (the same on every rank); parity picks the buffer half
    unsigned int error, pad;
    unsigned int ticket[P2P_MAX_RANKS];    // blocks of k_p2p_push that have finished their stores to peer q
    unsigned char* peer[P2P_MAX_RANKS];    // base of rank q's buffer as THIS device addresses it
};
static inline uint64_t p2p_blocks_bytes(uint32_t n, uint64_t gcap) { return 2ull * n * (gcap + 1) * sizeof(imrcd_entity_pair); }
static inline uint64_t p2p_bytes(uint32_t n, uint64_t gcap) { return p2p_blocks_bytes(n, gcap) + 2ull * n * P2P_FLAG_STRIDE; }

__global__ void k_p2p_begin(P2PCtl* c) { c->seq += 1ull; }

// grid (x, n_ranks): the blocks of column q copy this rank's result block into rank q's buffer; the last of them to finish raises the flag
__global__ void __launch_bounds__(256)
k_p2p_push(P2PCtl* c, const uint4* __restrict__ block, uint32_t rank, uint32_t n_ranks, unsigned long long gcap) {
    const uint32_t q = blockIdx.y;
    const unsigned long long seq = c->seq, par = seq & 1ull;
    const uint4 h = block[0];
    const unsigned long long count = (unsigned long long)h.x | ((unsigned long long)h.y << 32);
    const unsigned long long n16 = ((count < gcap ? count : gcap) + 1ull) * 5ull;       // header row + records, in 16-byte words
    unsigned char* base = c->peer[q];
    uint4* dst = reinterpret_cast<uint4*>(base) + (par * n_ranks + rank) * (gcap + 1ull) * 5ull;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n16; i += (unsigned long long)gridDim.x * blockDim.x) dst[i] = block[i];
    __threadfence_system();                                   // this thread's stores are visible to the peer before the ticket is taken
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(&c->ticket[q], 1u) == gridDim.x - 1u) {
            c->ticket[q] = 0u;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long*>(base + 2ull * n_ranks * (gcap + 1ull) * sizeof(imrcd_entity_pair) + (par * n_ranks + rank) * P2P_FLAG_STRIDE) = seq;
        }
    }
}

// grid (x, n_ranks): the blocks of row r wait for the blocks of ranks 0..r (the offsets come from their headers), then copy rank r's records
// to their place in the dense array.  A rank's flag for this frame can only be followed by its flag for the frame after next in the same
// buffer half, and that one is pushed after the rank has seen THIS rank's next frame: nothing read here can be overwritten meanwhile.
__global__ void __launch_bounds__(256)
k_p2p_wait_compact(P2PCtl* c, const unsigned char* local, uint32_t n_ranks, unsigned long long gcap, GatherHdr* __restrict__ hdr, uint4* __restrict__ dense,
                   unsigned long long timeout_ns) {
    const uint32_t r = blockIdx.y;
    const unsigned long long seq = c->seq, par = seq & 1ull;
    if (threadIdx.x <= r) {
        const volatile unsigned long long* flag = reinterpret_cast<const volatile unsigned long long*>(local + 2ull * n_ranks * (gcap + 1ull) * sizeof(imrcd_entity_pair)
                                                                                                       + (par * n_ranks + threadIdx.x) * P2P_FLAG_STRIDE);
        unsigned long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (*flag < seq) {
            unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) { atomicExch(&c->error, 1u); break; }
            __nanosleep(64);
        }
    }
    __threadfence_system();
    __syncthreads();
    const uint4* blocks = reinterpret_cast<const uint4*>(local) + par * n_ranks * (gcap + 1ull) * 5ull;
    const unsigned long long rows_per_block = gcap + 1ull;
    unsigned long long off = 0, mine = 0, flags = 0;
    for (uint32_t q = 0; q <= r; ++q) {
        const uint4 h = __ldcg(blocks + q * rows_per_block * 5ull);
        const unsigned long long cq = (unsigned long long)h.x | ((unsigned long long)h.y << 32);
        if (q < r) off += cq < gcap ? cq : gcap; else { mine = cq; flags = h.z; }
    }
    if (*reinterpret_cast<volatile unsigned int*>(&c->error)) { flags |= P2P_TIMEOUT_BIT; mine = 0; }
    if (blockIdx.x == 0 && threadIdx.x == 0) { hdr[r].count = mine; hdr[r].flags = flags; }
    const unsigned long long k = mine < gcap ? mine : gcap;
    const uint4* src = blocks + (r * rows_per_block + 1ull) * 5ull;
    uint4* dst = dense + off * 5ull;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < k * 5ull; i += (unsigned long long)gridDim.x * blockDim.x) dst[i] = __ldcg(src + i);
}

static void p2p_close_peers(imrcd_ctx* ctx) {
    for (void* q : ctx->p2p_opened) cudaIpcCloseMemHandle(q);
    ctx->p2p_opened.clear();
}
static void p2p_free_own(imrcd_ctx* ctx) {
    if (ctx->p2p_buf) cudaFree(ctx->p2p_buf);
    ctx->p2p_buf = nullptr; ctx->p2p_gcap = 0;
}
void imr_comm_p2p_release(imrcd_ctx* ctx) {
    cudaSetDevice(ctx->device);
    p2p_close_peers(ctx); p2p_free_own(ctx);
    if (ctx->p2p_ctl) cudaFree(ctx->p2p_ctl);
    ctx->p2p_ctl = nullptr; ctx->p2p_state = 0;
}

// own buffer + control block for the current capacity (flags zeroed; the sequence number is kept: it counts the communicator's frames)
static int p2p_alloc_own(imrcd_ctx* ctx) {
    if (!ctx->p2p_ctl) {
        IMR_CUDA(ctx, cudaMalloc(&ctx->p2p_ctl, sizeof(P2PCtl)));
        IMR_CUDA(ctx, cudaMemset(ctx->p2p_ctl, 0, sizeof(P2PCtl)));
    }
    p2p_free_own(ctx);
    IMR_CUDA(ctx, cudaMalloc(&ctx->p2p_buf, p2p_bytes(ctx->comm_n, ctx->gcap)));
    IMR_CUDA(ctx, cudaMemset(ctx->p2p_buf, 0, p2p_bytes(ctx->comm_n, ctx->gcap)));
    ctx->p2p_gcap = ctx->gcap;
    return IMRCD_OK;
}
static int p2p_set_peers(imrcd_ctx* ctx, void* const* peer) {
    unsigned char* tab[P2P_MAX_RANKS] = {};
    for (uint32_t q = 0; q < ctx->comm_n; ++q) tab[q] = static_cast<unsigned char*>(peer[q]);
    IMR_CUDA(ctx, cudaMemcpy(static_cast<char*>(ctx->p2p_ctl) + offsetof(P2PCtl, peer), tab, sizeof(tab), cudaMemcpyHostToDevice));
    return IMRCD_OK;
}

// one all-gather of `bytes` per rank between host buffers (setup only)
static int comm_exchange(imrcd_ctx* ctx, NcclApi* api, const void* mine, void* all, size_t bytes) {
    void* d = nullptr;
    IMR_CUDA(ctx, cudaMalloc(&d, bytes * (ctx->comm_n + 1)));
    IMR_CUDA(ctx, cudaMemcpy(d, mine, bytes, cudaMemcpyHostToDevice));
    const ncclResult_t r = api->AllGather(d, static_cast<char*>(d) + bytes, bytes, ncclUint8, static_cast<ncclComm_t>(ctx->comm), ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpy(all, static_cast<char*>(d) + bytes, bytes * ctx->comm_n, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (r != ncclSuccess) { ctx->err = std::string("ncclAllGather (setup): ") + api->GetErrorString(r); return IMRCD_E_CUDA; }
    IMR_CUDA(ctx, e);
    return IMRCD_OK;
}

// Processes: (re)build the peer mapping for the current capacity.  Collective: every rank gets here in the same frame, because the capacity
// is a function of the gathered headers.  Any rank that cannot export or open a buffer makes ALL ranks fall back to NCCL for good.
static int p2p_setup_processes(imrcd_ctx* ctx) {
    NcclApi* api = nccl_api();
    if (!api) { ctx->p2p_state = -1; return IMRCD_OK; }
    const uint32_t n = ctx->comm_n;
    struct Msg { cudaIpcMemHandle_t h; int ok; int device; unsigned long long seq; };
    std::vector<Msg> all(n);
    Msg mine; memset(&mine, 0, sizeof(mine));
    // 1. nobody touches a buffer that is about to go: peers' mappings are closed first, on every rank, before any rank frees
    IMR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    p2p_close_peers(ctx);
    mine.ok = 1;
    int rc = comm_exchange(ctx, api, &mine, all.data(), sizeof(Msg)); if (rc) return rc;
    // 2. own buffer, its handle
    mine.ok = (p2p_alloc_own(ctx) == IMRCD_OK) ? 1 : 0;
    if (mine.ok && cudaIpcGetMemHandle(&mine.h, ctx->p2p_buf) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
    mine.device = ctx->device;
    if (ctx->p2p_ctl) cudaMemcpy(&mine.seq, ctx->p2p_ctl, sizeof(mine.seq), cudaMemcpyDeviceToHost);
    rc = comm_exchange(ctx, api, &mine, all.data(), sizeof(Msg)); if (rc) return rc;
    // 3. open the peers' buffers
    std::vector<void*> peer(n, nullptr);
    int ok = 1;
    for (uint32_t q = 0; q < n; ++q) {
        ok &= all[q].ok;
        if (all[q].seq != mine.seq) ok = 0;                    // the ranks' frame counters must agree (they advance together)
    }
    if (ok) for (uint32_t q = 0; q < n && ok; ++q) {
        if (q == ctx->comm_rank) { peer[q] = ctx->p2p_buf; continue; }
        void* m = nullptr;
        if (cudaIpcOpenMemHandle(&m, all[q].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
        ctx->p2p_opened.push_back(m); peer[q] = m;
    }
    // 4. everybody or nobody
    mine.ok = ok;
    rc = comm_exchange(ctx, api, &mine, all.data(), sizeof(Msg)); if (rc) return rc;
    for (uint32_t q = 0; q < n; ++q) ok &= all[q].ok;
    if (!ok) {
        p2p_close_peers(ctx);
        mine.ok = 1; rc = comm_exchange(ctx, api, &mine, all.data(), sizeof(Msg)); if (rc) return rc;      // every mapping is closed before any buffer goes
        p2p_free_own(ctx);
        ctx->p2p_state = -1;
        return IMRCD_OK;
    }
    rc = p2p_set_peers(ctx, peer.data()); if (rc) return rc;
    ctx->p2p_state = 1;
    return IMRCD_OK;
}

// called with the frame's other reservations (never inside a capture)
static int p2p_prepare(imrcd_ctx* ctx) {
    if (ctx->p2p_state < 0 || ctx->comm_n < 2 || ctx->comm_n > P2P_MAX_RANKS) { if (ctx->p2p_state == 0) ctx->p2p_state = -1; return IMRCD_OK; }
    if (ctx->p2p_state == 0) { const char* ev = getenv("IMRCD_P2P"); if (ev && atoi(ev) == 0) { ctx->p2p_state = -1; return IMRCD_OK; } }
    if (ctx->p2p_state == 1 && ctx->p2p_gcap == ctx->gcap) return IMRCD_OK;
    if (ctx->group) return IMRCD_OK;                           // one process: the group lays all its contexts out together (group_p2p_setup)
    return p2p_setup_processes(ctx);
}

// the exchange itself: two launches on the frame's stream
static int p2p_gather(imrcd_ctx* ctx) {
    cudaStream_t s = ctx->stream;
    P2PCtl* c = static_cast<P2PCtl*>(ctx->p2p_ctl);
    k_p2p_begin<<<1, 1, 0, s>>>(c);
    k_p2p_push<<<dim3(4, ctx->comm_n), 256, 0, s>>>(c, reinterpret_cast<const uint4*>(ctx->d_epairs.p), ctx->comm_rank, ctx->comm_n, ctx->gcap);
    IMR_CUDA(ctx, cudaGetLastError());
    return IMRCD_OK;
}

static int gather_reserve(imrcd_ctx* ctx) {
    const uint64_t rows = (uint64_t)ctx->comm_n * (ctx->gcap + 1) + hdr_rows(ctx->comm_n) + (uint64_t)ctx->comm_n * ctx->gcap;
    IMR_CUDA(ctx, ctx->d_gather.reserve(rows * sizeof(imrcd_entity_pair), 0, ctx->stream));
    return IMRCD_OK;
}

int imr_comm_reserve(imrcd_ctx* ctx) {
    int rc = gather_reserve(ctx); if (rc) return rc;
    rc = p2p_prepare(ctx); if (rc) return rc;
    IMR_CUDA(ctx, ctx->p_gather.reserve((hdr_rows(ctx->comm_n) + (uint64_t)ctx->comm_n * ctx->gcap) * sizeof(imrcd_entity_pair)));
    return IMRCD_OK;
}

// part 1: the collective itself (the send block is the library's own result block: header row + records, as it lies in HBM)
int imr_comm_allgather(imrcd_ctx* ctx) {
    NcclApi* api = nccl_api();
    if (!api || !ctx->comm) { ctx->err = "no communicator"; return IMRCD_E_STATE; }
    if (ctx->d_epairs.cap < sizeof(imrcd_entity_pair) * (ctx->gcap + 1)) { ctx->err = "result block smaller than the gather capacity"; return IMRCD_E_STATE; }      // imr_frame_enqueue sizes it
    if (imr_comm_uses_p2p(ctx)) return p2p_gather(ctx);
    IMR_NCCL(ctx, api, api->AllGather(ctx->d_epairs.p, ctx->d_gather.p, (ctx->gcap + 1) * sizeof(imrcd_entity_pair), ncclUint8, static_cast<ncclComm_t>(ctx->comm), ctx->stream));
    return IMRCD_OK;
}

// part 2: compaction + speculative D2H (headers, and as many records as the last frame had plus a margin)
int imr_comm_after_gather(imrcd_ctx* ctx, uint64_t spec_rows) {
    cudaStream_t s = ctx->stream;
    const uint64_t blocks_rows = (uint64_t)ctx->comm_n * (ctx->gcap + 1), hr = hdr_rows(ctx->comm_n);
    imrcd_entity_pair* base = ctx->d_gather.as<imrcd_entity_pair>();
    if (imr_comm_uses_p2p(ctx))
        k_p2p_wait_compact<<<dim3(4, ctx->comm_n), 256, 0, s>>>(static_cast<P2PCtl*>(ctx->p2p_ctl), static_cast<const unsigned char*>(ctx->p2p_buf), ctx->comm_n, ctx->gcap,
                                                                reinterpret_cast<GatherHdr*>(base + blocks_rows), reinterpret_cast<uint4*>(base + blocks_rows + hr), 5000000000ull);
    else
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
        if (h[r].flags & P2P_TIMEOUT_BIT) { *fatal = true; ctx->err = "end-of-frame merge: a peer's block did not arrive (a rank failed or left the frame)"; }
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

// One process: the contexts of a group reach each other's buffers by plain peer access.  All of them are laid out together, whenever the
// gather capacity has changed (it changes on every context at once: the decision is taken from the gathered headers).
static int group_p2p_setup(imrcd_group* g) {
    const uint32_t n = (uint32_t)g->ctx.size();
    if (n < 2 || n > P2P_MAX_RANKS || g->ctx[0]->p2p_state < 0) return IMRCD_OK;
    bool need = false;
    for (imrcd_ctx* c : g->ctx) need |= !(c->p2p_state == 1 && c->p2p_gcap == c->gcap);
    if (!need) return IMRCD_OK;
    const char* ev = getenv("IMRCD_P2P");
    bool ok = !(ev && atoi(ev) == 0);
    for (imrcd_ctx* c : g->ctx) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); }
    for (uint32_t i = 0; i < n && ok; ++i) for (uint32_t j = 0; j < n && ok; ++j) {
        if (i == j || g->ctx[i]->device == g->ctx[j]->device) continue;
        cudaSetDevice(g->ctx[i]->device);
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, g->ctx[i]->device, g->ctx[j]->device) != cudaSuccess || !can) { ok = false; break; }
        const cudaError_t e = cudaDeviceEnablePeerAccess(g->ctx[j]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
        cudaGetLastError();
    }
    if (!ok) { for (imrcd_ctx* c : g->ctx) { cudaSetDevice(c->device); imr_comm_p2p_release(c); c->p2p_state = -1; } return IMRCD_OK; }
    std::vector<void*> peer(n);
    unsigned long long seq0 = 0;
    for (uint32_t i = 0; i < n; ++i) {
        imrcd_ctx* c = g->ctx[i];
        cudaSetDevice(c->device);
        const int rc = p2p_alloc_own(c);
        if (rc) { g->err = imrcd_last_error(c); return rc; }
        peer[i] = c->p2p_buf;
        if (i == 0) cudaMemcpy(&seq0, c->p2p_ctl, sizeof(seq0), cudaMemcpyDeviceToHost);
        else cudaMemcpy(c->p2p_ctl, &seq0, sizeof(seq0), cudaMemcpyHostToDevice);      // the contexts count the group's frames together
    }
    for (imrcd_ctx* c : g->ctx) {
        cudaSetDevice(c->device);
        const int rc = p2p_set_peers(c, peer.data());
        if (rc) { g->err = imrcd_last_error(c); return rc; }
        c->p2p_state = 1;
    }
    return IMRCD_OK;
}

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
        for (uint32_t i = 0; i < n; ++i) { comm_attach(g->ctx[i], comms[i], i, n); g->ctx[i]->group = g; }
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
        { const int rc = group_p2p_setup(g); if (rc) return rc; }
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
