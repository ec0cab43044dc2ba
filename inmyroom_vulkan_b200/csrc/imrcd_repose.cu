// imrcd_repose.cu -- re-posing skinned / morphed meshes on the device (BASELINE config 5): the arithmetic of the engine's
// dynamic-mesh compute pass (IMR/shaders/dynamicMeshShader_glsl.comp:99-145, dispatched per primitive by DynamicMeshes::RecordTransformations,
// IMR/src/Graphics/DynamicMeshes.cpp:672-790) for the position stream, followed by the triangles of the collision tree taking their corners
// from the re-posed vertices.  The reference never gives such meshes a collision tree (SURVEY finding 4): this is the front half of
// "triangle recompute + OBB-tree refit then collide"; imrcd_mesh_refit is the back half.
//
//   morphed = V[x (T + 1)] + sum_i w_i * V[x (T + 1) + i + 1]                                        (:105-111, VEC = vec4)
//   result  = morphed                                  when the primitive has no joints               (:121-123)
//           = sum_groups sum_{c in xyzw} weights.c * (M[joints.c + matrixOffset + 1] * InvBind[joints.c + inverseMatricesOffset] * morphed)   (:124-132, :79-84)
// GLSL leaves the order of a matrix product's sums (and contraction) to the driver; here `M * InvBind * v` is (M * InvBind) * v in glm's
// order (type_mat4x4.inl:561-572, 630-648) without contraction, the product M * InvBind taken once per joint and frame instead of once per
// vertex and joint (the same operations on the same operands).  The checker is the plain-C restatement oracle/imr_oracle.c imro_repose.
//
// Batched: any number of meshes per call, three launches (joint products, vertices, triangles).
#include "imrcd_internal.cuh"
#include <algorithm>

struct SkinDev { const float4* verts; const ushort4* joints; const float4* weights; uint32_t n_vertices, n_targets, n_groups, pad; };

struct ReposeSeg {
    SkinDev skin;
    uint32_t vtx_prefix, joint_prefix, weight_prefix, n_joints;     // offsets into the call's concatenated vertices / joints / morph weights
    uint32_t tri_base, n_tri, tri_prefix, pad;
};

__device__ __forceinline__ uint32_t rp_locate(const ReposeSeg* __restrict__ segs, uint32_t n_seg, uint32_t g, int what) {   // last s with prefix[s] <= g
    uint32_t lo = 0, hi = n_seg;
    while (hi - lo > 1u) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint32_t p = what == 0 ? segs[mid].vtx_prefix : (what == 1 ? segs[mid].joint_prefix : segs[mid].tri_prefix);
        if (p <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// P_j = M_j * InvBind_j, glm's mat4 * mat4 (type_mat4x4.inl:630-648), one thread per joint of the call
__global__ void k_repose_joints(uint32_t total_joints, const float* __restrict__ mats, const float* __restrict__ inv_bind, float* __restrict__ prod) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= total_joints) return;
    mat4_mul(mats + 16ull * j, inv_bind + 16ull * j, prod + 16ull * j);
}

// glm's mat4 * vec4 (type_mat4x4.inl:561-572): (m0 * v0 + m1 * v1) + (m2 * v2 + m3 * v3), column-major m
__device__ __forceinline__ float4 mat4_mul_vec4(const float* __restrict__ m, float4 v) {
    float4 r;
    r.x = (m[0] * v.x + m[4] * v.y) + (m[8] * v.z + m[12] * v.w);
    r.y = (m[1] * v.x + m[5] * v.y) + (m[9] * v.z + m[13] * v.w);
    r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * v.w);
    r.w = (m[3] * v.x + m[7] * v.y) + (m[11] * v.z + m[15] * v.w);
    return r;
}

__global__ void k_repose_vertices(uint32_t total_vtx, const ReposeSeg* __restrict__ segs, uint32_t n_seg, const float* __restrict__ morph_w,
                                  const float* __restrict__ prod, float4* __restrict__ out) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_vtx) return;
    const ReposeSeg sg = segs[rp_locate(segs, n_seg, g, 0)];
    const uint32_t x = g - sg.vtx_prefix, T = sg.skin.n_targets;
    const float4* V = sg.skin.verts + (size_t)(T + 1u) * x;
    float4 m = __ldg(V);                                                       // :105
    for (uint32_t i = 0; i < T; ++i) {                                         // :106-111
        const float4 t = __ldg(V + i + 1u);
        const float w = morph_w[sg.weight_prefix + i];
        m.x += w * t.x; m.y += w * t.y; m.z += w * t.z; m.w += w * t.w;
    }
    float4 r = m;                                                              // :121-123
    if (sg.skin.n_groups != 0u) {
        r = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* P = prod + 16ull * sg.joint_prefix;
        for (uint32_t gI = 0; gI < sg.skin.n_groups; ++gI) {                   // :124-132
            const float4 w = __ldg(sg.skin.weights + (size_t)x * sg.skin.n_groups + gI);
            const ushort4 jn = sg.skin.joints[(size_t)x * sg.skin.n_groups + gI];
            float4 c;
            c = mat4_mul_vec4(P + 16u * jn.x, m); r.x += w.x * c.x; r.y += w.x * c.y; r.z += w.x * c.z; r.w += w.x * c.w;
            c = mat4_mul_vec4(P + 16u * jn.y, m); r.x += w.y * c.x; r.y += w.y * c.y; r.z += w.y * c.z; r.w += w.y * c.w;
            c = mat4_mul_vec4(P + 16u * jn.z, m); r.x += w.z * c.x; r.y += w.z * c.y; r.z += w.z * c.z; r.w += w.z * c.w;
            c = mat4_mul_vec4(P + 16u * jn.w, m); r.x += w.w * c.x; r.y += w.w * c.y; r.z += w.w * c.z; r.w += w.w * c.w;
        }
    }
    out[g] = r;                                                                // :143
}

// the tree's triangles (leaf order) take their corners from the re-posed vertices through their vertex ids (TriangleIndices,
// Triangle.cpp:242-250); the plane of the triangle (TriRec.t3) follows, the original index stays
__global__ void k_repose_triangles(uint32_t total_tri, const ReposeSeg* __restrict__ segs, uint32_t n_seg, const float4* __restrict__ posed,
                                   const uint32_t* __restrict__ tri_vid, TriRec* __restrict__ tris) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_tri) return;
    const ReposeSeg sg = segs[rp_locate(segs, n_seg, g, 2)];
    const uint32_t t = sg.tri_base + (g - sg.tri_prefix);
    const uint32_t* vid = tri_vid + 3ull * t;
    const float4 a = posed[sg.vtx_prefix + vid[0]], b = posed[sg.vtx_prefix + vid[1]], c = posed[sg.vtx_prefix + vid[2]];
    TriRec r;
    r.t0 = make_float4(a.x, a.y, a.z, tris[t].t0.w); r.t1 = make_float4(b.x, b.y, b.z, 0.f); r.t2 = make_float4(c.x, c.y, c.z, 0.f);
    V3 N; float d;
    tt_plane(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), N, d);
    r.t3 = make_float4(N.x, N.y, N.z, d);
    tris[t] = r;
}

// ---- host -----------------------------------------------------------------------------------------------------------------------------
struct SkinHost { DevBuf verts, joints, weights; uint32_t n_vertices = 0, n_targets = 0, n_groups = 0; };
static std::vector<SkinHost>& skins_of(imrcd_ctx* ctx) { return *reinterpret_cast<std::vector<SkinHost>*>(ctx->skins); }

void imr_skins_release(imrcd_ctx* ctx) {
    if (!ctx->skins) return;
    for (SkinHost& s : skins_of(ctx)) { s.verts.release(); s.joints.release(); s.weights.release(); }
    delete reinterpret_cast<std::vector<SkinHost>*>(ctx->skins);
    ctx->skins = nullptr;
}

extern "C" int imrcd_skin_create(imrcd_ctx* ctx, uint64_t n_vertices, uint32_t n_morph_targets, const float* vertices, uint32_t joints_groups,
                                 const uint16_t* joints, const float* weights, uint32_t* skin_id) {
    if (!ctx) return IMRCD_E_ARG;
    if (!skin_id || !vertices || n_vertices == 0 || n_vertices >= (1ull << 31) || (joints_groups && (!joints || !weights))) { ctx->err = "imrcd_skin_create: bad argument"; return IMRCD_E_ARG; }
    cudaSetDevice(ctx->device);
    if (!ctx->skins) ctx->skins = new std::vector<SkinHost>();
    cudaStream_t s = ctx->stream;
    SkinHost sk;
    sk.n_vertices = (uint32_t)n_vertices; sk.n_targets = n_morph_targets; sk.n_groups = joints_groups;
    const size_t vb = 16ull * n_vertices * (n_morph_targets + 1ull);
    IMR_CUDA(ctx, sk.verts.reserve(vb, 0, s));
    IMR_CUDA(ctx, cudaMemcpyAsync(sk.verts.p, vertices, vb, cudaMemcpyDefault, s));
    if (joints_groups) {
        IMR_CUDA(ctx, sk.joints.reserve(8ull * n_vertices * joints_groups, 0, s));
        IMR_CUDA(ctx, cudaMemcpyAsync(sk.joints.p, joints, 8ull * n_vertices * joints_groups, cudaMemcpyDefault, s));
        IMR_CUDA(ctx, sk.weights.reserve(16ull * n_vertices * joints_groups, 0, s));
        IMR_CUDA(ctx, cudaMemcpyAsync(sk.weights.p, weights, 16ull * n_vertices * joints_groups, cudaMemcpyDefault, s));
    }
    IMR_CUDA(ctx, cudaStreamSynchronize(s));
    skins_of(ctx).push_back(sk);
    *skin_id = (uint32_t)(skins_of(ctx).size() - 1);
    return IMRCD_OK;
}

__global__ void k_max_u32(uint64_t n, const uint32_t* __restrict__ v, uint32_t* out) {
    uint32_t m = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) m = max(m, v[i]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// largest value of a device array of n u32 (a bound check for index buffers and vertex ids that came from the caller)
int imr_device_max_u32(imrcd_ctx* ctx, const uint32_t* d_values, uint64_t n, uint32_t* out) {
    *out = 0;
    if (n == 0) return IMRCD_OK;
    IMR_CUDA(ctx, ctx->d_scalar.reserve(64, 0, ctx->stream));
    IMR_CUDA(ctx, cudaMemsetAsync(ctx->d_scalar.p, 0, 4, ctx->stream));
    k_max_u32<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 1024), 256, 0, ctx->stream>>>(n, d_values, ctx->d_scalar.as<uint32_t>());
    IMR_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_scalar.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    IMR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IMRCD_OK;
}

extern "C" int imrcd_mesh_bind_skin(imrcd_ctx* ctx, uint32_t mesh_id, uint32_t skin_id) {
    if (!ctx) return IMRCD_E_ARG;
    if (mesh_id >= ctx->meshes.size() || !ctx->skins || skin_id >= skins_of(ctx).size()) { ctx->err = "imrcd_mesh_bind_skin: bad id"; return IMRCD_E_ARG; }
    cudaSetDevice(ctx->device);
    const MeshDev md = ctx->meshes[mesh_id].dev;
    uint32_t mx = 0;                                           // every corner of the mesh must name a vertex of the skin
    const int rc = imr_device_max_u32(ctx, ctx->d_tri_vid.as<uint32_t>() + 3ull * md.tri_base, 3ull * md.n_tri, &mx);
    if (rc) return rc;
    if (md.n_tri && mx >= skins_of(ctx)[skin_id].n_vertices) { ctx->err = "imrcd_mesh_bind_skin: the mesh's vertex ids exceed the skin's vertex count"; return IMRCD_E_ARG; }
    ctx->meshes[mesh_id].skin = (int)skin_id;
    return IMRCD_OK;
}

extern "C" int imrcd_meshes_repose(imrcd_ctx* ctx, uint64_t n, const uint32_t* mesh_ids, const float* morph_weights, const float* joint_matrices,
                                   const float* inverse_bind, const uint32_t* n_joints) {
    if (!ctx) return IMRCD_E_ARG;
    if (n == 0) return IMRCD_OK;
    if (!mesh_ids) { ctx->err = "imrcd_meshes_repose: bad argument"; return IMRCD_E_ARG; }
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    std::vector<ReposeSeg> segs(n);
    uint64_t tot_v = 0, tot_j = 0, tot_w = 0, tot_t = 0;
    for (uint64_t k = 0; k < n; ++k) {
        const uint32_t id = mesh_ids[k];
        if (id >= ctx->meshes.size() || ctx->meshes[id].skin < 0) { ctx->err = "imrcd_meshes_repose: a mesh without a skin (imrcd_mesh_bind_skin)"; return IMRCD_E_ARG; }
        const SkinHost& sk = skins_of(ctx)[ctx->meshes[id].skin];
        const uint32_t nj = sk.n_groups ? (n_joints ? n_joints[k] : 0u) : 0u;
        if (sk.n_groups && (nj == 0 || !joint_matrices || !inverse_bind)) { ctx->err = "imrcd_meshes_repose: a skinned mesh needs joint matrices"; return IMRCD_E_ARG; }
        if (sk.n_targets && !morph_weights) { ctx->err = "imrcd_meshes_repose: a morphed mesh needs weights"; return IMRCD_E_ARG; }
        ReposeSeg& g = segs[k];
        g.skin.verts = sk.verts.as<float4>(); g.skin.joints = sk.joints.as<ushort4>(); g.skin.weights = sk.weights.as<float4>();
        g.skin.n_vertices = sk.n_vertices; g.skin.n_targets = sk.n_targets; g.skin.n_groups = sk.n_groups; g.skin.pad = 0;
        g.vtx_prefix = (uint32_t)tot_v; g.joint_prefix = (uint32_t)tot_j; g.weight_prefix = (uint32_t)tot_w; g.n_joints = nj;
        g.tri_base = ctx->meshes[id].dev.tri_base; g.n_tri = ctx->meshes[id].dev.n_tri; g.tri_prefix = (uint32_t)tot_t; g.pad = 0;
        tot_v += sk.n_vertices; tot_j += nj; tot_w += sk.n_targets; tot_t += g.n_tri;
        ctx->meshes[id].needs_refit = true;
    }
    if (tot_v >= (1ull << 32) || tot_t >= (1ull << 32)) { ctx->err = "imrcd_meshes_repose: too many vertices in one call"; return IMRCD_E_CAPACITY; }
    // the joint indices of a skin must stay below the joints the caller brings: checked once per (skin, count) on the device
    for (uint64_t k = 0; k < n; ++k) {
        const int sid = ctx->meshes[mesh_ids[k]].skin;
        SkinHost& sk = skins_of(ctx)[sid];
        if (!sk.n_groups) continue;
        if (ctx->skin_max_joint.size() <= (size_t)sid) ctx->skin_max_joint.resize(sid + 1, 0xffffffffu);
        if (ctx->skin_max_joint[sid] == 0xffffffffu) {
            // u16 indices, two per u32 word: the max of the words' halves
            std::vector<uint16_t> h(4ull * sk.n_vertices * sk.n_groups);
            IMR_CUDA(ctx, cudaMemcpy(h.data(), sk.joints.p, 2 * h.size(), cudaMemcpyDeviceToHost));
            ctx->skin_max_joint[sid] = *std::max_element(h.begin(), h.end());
        }
        if (ctx->skin_max_joint[sid] >= segs[k].n_joints) { ctx->err = "imrcd_meshes_repose: a joint index of the skin exceeds the joint matrices given"; return IMRCD_E_ARG; }
    }
    // inputs of the call -> HBM through pinned staging, one block: [segments | morph weights | joint matrices | inverse bind matrices]
    const size_t b_seg = sizeof(ReposeSeg) * n, b_w = 4 * tot_w, b_m = 64 * tot_j;
    const size_t o_w = (b_seg + 63) & ~size_t(63), o_m = (o_w + b_w + 63) & ~size_t(63), o_i = o_m + b_m, total = o_i + b_m;
    IMR_CUDA(ctx, ctx->p_repose.reserve(total, 0, s));
    IMR_CUDA(ctx, ctx->d_repose_in.reserve(total, 0, s));
    IMR_CUDA(ctx, ctx->d_repose_prod.reserve(std::max<size_t>(b_m, 64), 0, s));
    IMR_CUDA(ctx, ctx->d_repose_vtx.reserve(16ull * tot_v, 0, s));
    IMR_CUDA(ctx, cudaStreamSynchronize(s));                            // the staging block may still feed the previous call's copy
    char* hp = ctx->p_repose.as<char>();
    memcpy(hp, segs.data(), b_seg);
    if (b_w) memcpy(hp + o_w, morph_weights, b_w);
    if (b_m) { memcpy(hp + o_m, joint_matrices, b_m); memcpy(hp + o_i, inverse_bind, b_m); }
    IMR_CUDA(ctx, cudaMemcpyAsync(ctx->d_repose_in.p, hp, total, cudaMemcpyHostToDevice, s));
    const char* dp = ctx->d_repose_in.as<char>();
    const ReposeSeg* d_segs = reinterpret_cast<const ReposeSeg*>(dp);
    IMR_CUDA(ctx, cudaEventRecord(ctx->ev[8], s));
    if (tot_j) k_repose_joints<<<(unsigned)((tot_j + 127) / 128), 128, 0, s>>>((uint32_t)tot_j, reinterpret_cast<const float*>(dp + o_m), reinterpret_cast<const float*>(dp + o_i), ctx->d_repose_prod.as<float>());
    k_repose_vertices<<<(unsigned)((tot_v + 255) / 256), 256, 0, s>>>((uint32_t)tot_v, d_segs, (uint32_t)n, reinterpret_cast<const float*>(dp + o_w), ctx->d_repose_prod.as<float>(), ctx->d_repose_vtx.as<float4>());
    if (tot_t) k_repose_triangles<<<(unsigned)((tot_t + 255) / 256), 256, 0, s>>>((uint32_t)tot_t, d_segs, (uint32_t)n, ctx->d_repose_vtx.as<float4>(), ctx->d_tri_vid.as<uint32_t>(), ctx->d_tris.as<TriRec>());
    IMR_CUDA(ctx, cudaEventRecord(ctx->ev[9], s));
    IMR_CUDA(ctx, cudaGetLastError());
    ctx->repose_pending = true;
    return IMRCD_OK;
}

extern "C" int imrcd_mesh_last_repose_ms(imrcd_ctx* ctx, float* ms) {
    if (!ctx || !ms) return IMRCD_E_ARG;
    cudaSetDevice(ctx->device);
    *ms = 0.f;
    if (ctx->repose_pending) { IMR_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); cudaEventElapsedTime(ms, ctx->ev[8], ctx->ev[9]); ctx->last_repose_ms = *ms; ctx->repose_pending = false; }
    else *ms = ctx->last_repose_ms;
    return IMRCD_OK;
}

// test hook: the re-posed vertices (vec4) of the last imrcd_meshes_repose call, in the call's order
extern "C" int imrcd_test_reposed_vertices(imrcd_ctx* ctx, float* out, uint64_t n_vertices) {
    if (!ctx || !out) return IMRCD_E_ARG;
    cudaSetDevice(ctx->device);
    if (16 * n_vertices > ctx->d_repose_vtx.cap) { ctx->err = "imrcd_test_reposed_vertices: more vertices than the last call re-posed"; return IMRCD_E_ARG; }
    IMR_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_repose_vtx.p, 16 * n_vertices, cudaMemcpyDeviceToHost, ctx->stream));
    IMR_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return IMRCD_OK;
}
