/* imrcd.h -- C ABI of libimrcd.so: B200-native (sm_100a CUDA) replacement for the CPU
 * collision-detection path of thesmallcreeper/inMyRoom_vulkan.
 *
 * The boundary it replaces ("IMR/" = inMyRoom_vulkan/ in the reference checkout):
 *   class CollisionDetection            IMR/include/CollisionDetection/CollisionDetection.h:11-35
 *     Reset()                           IMR/src/CollisionDetection/CollisionDetection.cpp:28
 *     AddCollisionDetectionEntry(e)     IMR/src/CollisionDetection/CollisionDetection.cpp:33
 *     ExecuteCollisionDetection()       IMR/src/CollisionDetection/CollisionDetection.cpp:38-129
 *   OBBtree::OBBtree(vector<Triangle>&&) IMR/include/Geometry/OBBtree.h:105, src/Geometry/OBBtree.cpp:321
 *   struct CollisionDetectionEntry      IMR/include/ECS/ECStypes.h:149-156
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success or a negative
 * IMRCD_E_* code, with text in imrcd_last_error(); no exceptions cross the ABI; a context is
 * single-threaded like the reference (it runs under ECSwrapper::controlMutex, ECSwrapper.cpp:316).
 * Matrices are glm::mat4 layout: 16 contiguous floats, column-major.  There is no CPU fallback:
 * imrcd_create fails when no CUDA device of compute capability 10.x is usable.
 */
#ifndef IMRCD_H
#define IMRCD_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define IMRCD_OK            0
#define IMRCD_E_CUDA       -1   /* CUDA runtime / launch failure */
#define IMRCD_E_ARG        -2   /* bad argument */
#define IMRCD_E_NODEVICE   -3   /* no usable sm_100 device: the library has no CPU path */
#define IMRCD_E_CAPACITY   -4   /* an internal queue could not be grown */
#define IMRCD_E_STATE      -5   /* call sequence error */

typedef struct imrcd_ctx imrcd_ctx;

/* build modes for imrcd_mesh_create */
#define IMRCD_BUILD_MORTON     0u  /* bottom-up Morton-ordered GPU build (default, fast) */
#define IMRCD_BUILD_REFERENCE  1u  /* top-down build with the reference's split rule and summation order:
                                      bit-identical tree to OBBtree.cpp:321 (parity mode, slower) */

/* One colliding entity pair (CollisionDetection.cpp:60-78).  `first` is the entity earlier on the
 * sweep's U axis (SweepAndPrune.cpp:63); contact data live in first's model space. */
typedef struct {
    uint32_t entry_first, entry_second;    /* indices into this frame's entry list */
    uint32_t entity_first, entity_second;  /* ids given to imrcd_frame_add_entry */
    uint32_t n_hits;                       /* non-coplanar intersecting triangle pairs */
    uint32_t n_rays_first, n_rays_second;  /* "uncollide" rays per side (CreateUncollideRays.cpp:131-178) */
    uint32_t flags;                        /* bit0: colliding (>= 1 ray, CollisionDetection.cpp:63) */
    float    avg_first[3];                 /* average_point_first_modelspace  (CreateUncollideRays.cpp:185-189) */
    float    avg_second[3];                /* average_point_second_modelspace (CreateUncollideRays.cpp:191-198) */
    float    delta_first[3];               /* CollisionCallbackData.deltaVector for first (CollisionDetection.cpp:80-103): the response of
                                              ShootUncollideRays.cpp:14-93 split by movement; zero when neither entity moved (:99-103) */
    float    delta_second[3];              /* ... for second */
} imrcd_entity_pair;                       /* 80 bytes */

/* One intersecting, non-coplanar triangle pair (CreateUncollideRays.cpp:86-100). */
typedef struct {
    uint32_t pair;          /* index into the broad-phase pair list of this frame (imrcd_frame_pairs) */
    uint32_t tri_first;     /* ORIGINAL input triangle index in first's mesh (not leaf order) */
    uint32_t tri_second;    /* ORIGINAL input triangle index in second's mesh */
    float    source[3];     /* TrianglesIntersectionInfo.source (Triangle.h:13-19), first's model space */
    float    target[3];     /* TrianglesIntersectionInfo.target */
    float    weight;        /* |source - target| (CreateUncollideRays.cpp:93) */
} imrcd_tri_hit;            /* 40 bytes */

typedef struct {
    uint64_t n_entries;
    uint64_t n_pairs;           /* broad-phase pairs owned by this shard */
    uint64_t n_sat_tests;       /* OBB-OBB SAT evaluations (IntersectOBBtreesRecursive visits) */
    uint64_t n_combos;          /* leaf x leaf candidate range combinations */
    uint64_t n_tri_tests;       /* sum over combos of cntA*cntB == triangle-pair tests */
    uint64_t n_hits;            /* non-coplanar intersecting triangle pairs */
    uint64_t n_coplanar_hits;   /* intersecting but coplanar (dropped, CreateUncollideRays.cpp:88) */
    uint64_t n_colliding;       /* colliding entity pairs */
    uint64_t traverse_launches; /* kernel launches of the traversal stage */
    uint64_t total_launches;    /* all kernel launches of the frame */
    uint64_t n_queue_items;     /* work items that went through the global queue after the roots (load balancing) */
    uint64_t n_warp_iterations; /* traversal warp-iterations (<= 32 SAT visits each) */
    uint64_t trav_busy_cycles;  /* sum over traversal warps of SM cycles spent inside iterations (diagnostic) */
    uint64_t trav_idle_polls;   /* failed refill attempts of starving traversal warps (diagnostic) */
    float    ms_total;          /* device time of imrcd_frame_run (CUDA events) */
    float    ms_broad, ms_pair_setup, ms_traverse, ms_narrow, ms_reduce;
    uint64_t n_contact_pairs;   /* pairs with at least one hit that went through the contact reduction (CreateUncollideRays.cpp:117-198) */
    uint64_t n_rays;            /* "uncollide" rays of the colliding pairs, both sides (CreateUncollideRays.cpp:131-178) */
    uint64_t n_rays_shot;       /* rays of the pairs whose entities moved: each goes through the Hermann passes of ShootUncollideRays.cpp:73-89 */
    uint64_t n_responses;       /* successful Hermann passes (entries of ray_responses, ShootUncollideRays.cpp:43,61) */
    float    ms_response;       /* device time of the response stage (part of ms_reduce) */
    uint64_t n_merged;          /* colliding entity pairs of ALL ranks after the end-of-frame merge (== n_colliding without a communicator) */
    uint64_t n_entries_local;   /* entries of the frame this context kept (== n_entries unless the frame is sharded) */
} imrcd_frame_stats;

/* ---- context ---------------------------------------------------------------------------- */
/* `cuda_stream` is a cudaStream_t to launch on (e.g. torch.cuda.current_stream().cuda_stream) or NULL
 * for a stream owned by the context. */
int  imrcd_create(int device, void* cuda_stream, imrcd_ctx** out);
void imrcd_destroy(imrcd_ctx* ctx);
const char* imrcd_last_error(const imrcd_ctx* ctx);
const char* imrcd_version(void);
/* sizeof(imrcd_entity_pair), sizeof(imrcd_tri_hit), sizeof(imrcd_frame_stats), then offsetof every imrcd_frame_stats field in declaration
 * order, as the library was compiled; returns the number of values (out == NULL or capacity too small: only the count) */
int imrcd_abi_layout(uint64_t* out, uint64_t capacity);

/* ---- meshes: replaces OBBtree::OBBtree(std::vector<Triangle>&&) (OBBtree.cpp:321) --------- */
/* positions/normals: n_tri*9 floats (p0,p1,p2 per triangle; normals may be NULL -> face normals,
 * Triangle.cpp:214-234); vertex_ids: n_tri*3 u32 (TriangleIndices, Triangle.cpp:242-250) or NULL. */
int imrcd_mesh_create(imrcd_ctx* ctx, const float* positions, const float* normals, const uint32_t* vertex_ids,
                      uint64_t n_tri, uint32_t build_mode, uint32_t* mesh_id);
/* The engine's own way in: PrimitivesOfMeshes::StartRecordOBBtree / one PrimitiveOBBtreeData per glTF primitive /
 * GetOBBtreeAndReset (IMR/src/Graphics/Meshes/PrimitivesOfMeshes.cpp:637-671,835-863).  A primitive is handed over as it lies in the glTF
 * buffers -- points (stride 3 or 4 floats: the engine keeps vec4), optional normals (same stride), optional u32 indices, the glTF draw mode
 * (IMR/include/glTFenum.h:27-36: 0 points, 1 lines, 3 line strip, 4 triangles, 5 triangle strip, 6 triangle fan; 2 = line loop yields no
 * triangles, as in the reference) -- and Triangle::CreateTriangleList (IMR/src/Geometry/Triangle.cpp:9-62,214-280) runs on the device:
 * positions, normals (the face normal when the primitive has none, Triangle.cpp:141-147,223-232) and vertex ids of every triangle, the
 * primitives of a mesh concatenated in the order given.  indices == NULL means 0 .. n_points-1 (the reference sizes that iota by the FLOAT
 * count, PrimitivesOfMeshes.cpp:662, and reads out of bounds; no shipped asset has a non-indexed primitive). */
int imrcd_mesh_begin(imrcd_ctx* ctx);
int imrcd_mesh_add_primitive(imrcd_ctx* ctx, const float* points, uint64_t n_points, uint32_t stride_floats, const float* normals,
                             const uint32_t* indices, uint64_t n_indices, uint32_t gltf_mode);
int imrcd_mesh_end(imrcd_ctx* ctx, uint32_t build_mode, uint32_t* mesh_id);

/* ---- glTF files: .gltf (+ .bin, base64 data URIs) and .glb read on the host, triangles and trees made on the device ----
 * Replaces the engine's load path for collision geometry: tinygltf + PrimitiveInitializationData (IMR/src/Graphics/Meshes/
 * PrimitivesOfMeshes.cpp:44-175: float POSITION / NORMAL with y and z negated :83-87,:134-139; u16 / u32 indices :58-70 -- u8 is also taken
 * here; line loop drawn as line strip :49-55; skinned or morphed primitives give no collision triangles :641,:669) and the per-mesh loop of
 * MeshesOfNodes.cpp:34-53 (StartRecordOBBtree, the primitives "triangles first", GetOBBtreeAndReset): ONE tree per glTF mesh, mesh_ids[i]
 * belongs to meshes[i] of the file.  imrcd_gltf_open / _primitive are host-only (no device needed): they expose each mesh's primitives exactly
 * as they are handed to imrcd_mesh_add_primitive, in the order the reference would record them. */
typedef struct imrcd_gltf imrcd_gltf;
typedef struct imrcd_gltf_primitive_view {
    const float*    points;        /* n_points * 3, glTF -> engine axes (y and z negated) */
    const float*    normals;       /* n_points * 3 or NULL */
    const uint32_t* indices;       /* n_indices or NULL */
    uint64_t        n_points, n_indices;
    uint32_t        mode;          /* glTF draw mode as recorded (2 has become 3) */
    uint32_t        skipped;       /* 1 = skinned or morphed: not part of the tree, no arrays */
    uint32_t        source_index;  /* position in the file's "primitives" array */
    uint32_t        has_indices;   /* 0 = non-indexed (draws 0 .. n_points-1) */
} imrcd_gltf_primitive_view;
int  imrcd_gltf_open(const char* path, imrcd_gltf** out, char* err, uint64_t err_cap);
void imrcd_gltf_close(imrcd_gltf* g);
int  imrcd_gltf_mesh_count(const imrcd_gltf* g, uint32_t* n);
int  imrcd_gltf_primitive_count(const imrcd_gltf* g, uint32_t mesh, uint32_t* n);
int  imrcd_gltf_primitive(const imrcd_gltf* g, uint32_t mesh, uint32_t k, imrcd_gltf_primitive_view* out);
int  imrcd_gltf_build_mesh(imrcd_ctx* ctx, const imrcd_gltf* g, uint32_t mesh, uint32_t build_mode, uint32_t* mesh_id);
/* open + one tree per mesh + close.  mesh_ids == NULL: only count (*n_meshes). */
int  imrcd_gltf_load(imrcd_ctx* ctx, const char* path, uint32_t build_mode, uint32_t* mesh_ids, uint32_t capacity, uint32_t* n_meshes);

/* Test-only: upload a tree built elsewhere (flat pre-order form, see oracle/imr_oracle.h). */
int imrcd_mesh_import_tree(imrcd_ctx* ctx, uint64_t n_vertices, const float* boxes, const int32_t* left, const int32_t* right,
                           const uint32_t* tri_off, const uint32_t* tri_cnt, uint64_t n_tri, const float* tri_pos,
                           const float* tri_nrm, const uint32_t* tri_vid, const uint32_t* tri_orig, uint32_t* mesh_id);
int imrcd_mesh_info(imrcd_ctx* ctx, uint32_t mesh_id, uint64_t* n_tri, uint64_t* n_vertices);
/* Read a tree back in the flat pre-order form (any pointer may be NULL). */
int imrcd_mesh_export_tree(imrcd_ctx* ctx, uint32_t mesh_id, float* boxes, int32_t* left, int32_t* right,
                           uint32_t* tri_off, uint32_t* tri_cnt, float* tri_pos, float* tri_nrm, uint32_t* tri_vid,
                           uint32_t* tri_orig);
/* Re-posed meshes (BASELINE config 5; the engine re-poses skinned / morphed meshes on the GPU, IMR/shaders/dynamicMeshShader_glsl.comp:99-145,
 * but never gives them collision trees: SURVEY finding 4).  imrcd_mesh_update_positions replaces the triangle positions (and normals
 * when non-NULL) of a mesh, given in the ORIGINAL input order of imrcd_mesh_create / tri_orig of an imported tree; the topology of the
 * tree is kept.  imrcd_mesh_refit recomputes every box of the listed meshes (NULL = all meshes updated since their last refit) in one
 * batched pass: fresh PCA axes per node, boxes rounded outward. */
int imrcd_mesh_update_positions(imrcd_ctx* ctx, uint32_t mesh_id, const float* positions, const float* normals);
int imrcd_mesh_refit(imrcd_ctx* ctx, const uint32_t* mesh_ids, uint64_t n);
int imrcd_mesh_last_refit_ms(imrcd_ctx* ctx, float* ms);
/* Re-posing on the device (BASELINE config 5): the arithmetic of the engine's dynamic-mesh compute pass for the position stream
 * (IMR/shaders/dynamicMeshShader_glsl.comp:99-145, DynamicMeshes::RecordTransformations, IMR/src/Graphics/DynamicMeshes.cpp:672-790):
 *     morphed = V[x (T + 1)] + sum_i w_i V[x (T + 1) + i + 1]
 *     result  = morphed (no joints)  or  sum_groups sum_c weights.c * (M[joints.c] * InvBind[joints.c] * morphed)
 * A skin is what that pass reads per vertex: `vertices` = n_vertices * (n_morph_targets + 1) vec4 (the base vertex, then its morph
 * targets, as the engine lays them out), and per vertex `joints_groups` groups of 4 u16 joint indices and 4 float weights (0 groups: a
 * morph-only mesh).  imrcd_mesh_bind_skin ties a mesh to a skin: the corners of its triangles are the skin's vertices named by the
 * mesh's vertex ids (TriangleIndices, Triangle.cpp:242-250).  imrcd_meshes_repose re-poses any number of bound meshes in one batched
 * pass - morph_weights: the skins' n_morph_targets floats per mesh, concatenated; joint_matrices / inverse_bind: n_joints[k] mat4 per
 * mesh, concatenated (modelMatrices[matrixOffset + 1 + j].positionMatrix and inverseModelMatrices[inverseMatricesOffset + j] of the
 * shader) - and rewrites the triangles of their trees; imrcd_mesh_refit (NULL) then refits exactly those meshes.  Normals are not
 * re-posed (imrcd_mesh_update_positions takes new ones). */
int imrcd_skin_create(imrcd_ctx* ctx, uint64_t n_vertices, uint32_t n_morph_targets, const float* vertices, uint32_t joints_groups,
                      const uint16_t* joints, const float* weights, uint32_t* skin_id);
int imrcd_mesh_bind_skin(imrcd_ctx* ctx, uint32_t mesh_id, uint32_t skin_id);
int imrcd_meshes_repose(imrcd_ctx* ctx, uint64_t n, const uint32_t* mesh_ids, const float* morph_weights, const float* joint_matrices,
                        const float* inverse_bind, const uint32_t* n_joints);
int imrcd_mesh_last_repose_ms(imrcd_ctx* ctx, float* ms);
/* device time of the last imrcd_mesh_create in ms */
int imrcd_mesh_last_build_ms(imrcd_ctx* ctx, float* ms);

/* ---- frame: Reset / AddCollisionDetectionEntry / ExecuteCollisionDetection -------------- */
int imrcd_frame_reset(imrcd_ctx* ctx);
int imrcd_frame_add_entry(imrcd_ctx* ctx, const float current[16], const float previous[16], uint32_t mesh_id,
                          uint8_t should_callback, uint32_t entity);
int imrcd_frame_add_entries(imrcd_ctx* ctx, uint64_t n, const float* current, const float* previous,
                            const uint32_t* mesh_ids, const uint8_t* should_callback, const uint32_t* entities);
/* Zero-copy submission: the caller writes the entries of this frame straight into the context's PINNED staging buffers
 * (what AddCollisionDetectionEntry's push_back is to the reference, CollisionDetection.cpp:33-36) and then commits them.
 * imrcd_frame_map_entries returns pointers to room for `n` more entries after the ones already added (any out pointer
 * may be NULL); the memory keeps its previous contents, stays valid until the next map call that has to grow it, and
 * must be completely filled before imrcd_frame_commit_entries(n).  `previous` may be left untouched when
 * previous_valid == 0 (previous == current for every committed entry). */
int imrcd_frame_map_entries(imrcd_ctx* ctx, uint64_t n, float** current, float** previous, uint32_t** mesh_ids,
                            uint8_t** should_callback, uint32_t** entities);
int imrcd_frame_commit_entries(imrcd_ctx* ctx, uint64_t n, int previous_valid);
/* Multi-GPU ("the broad-phase pair list sharded by entity"): this context is rank `rank` of `n_ranks` working on the same frame.
 * Every rank is handed the WHOLE entry list (same calls, same order) and keeps its share: every entry with shouldCallback, and the
 * entries without it in blocks of 256 caller indices dealt round-robin.  A pair needs the flag on one side (SweepAndPrune.cpp:60), so a
 * pair with an unflagged entity lives on exactly one rank; a pair of two flagged entities is kept by one rank chosen from the pair.  The
 * pair lists of ranks 0..n_ranks-1 are disjoint and add up to the unsharded list; entry indices in every result are the caller's.
 * Only the kept entries are copied, uploaded, sorted and swept.  Takes effect for the next frame (at once when no entry has been added
 * yet).  Default (0,1) = everything.  imrcd_comm_init sets it. */
int imrcd_frame_set_shard(imrcd_ctx* ctx, uint32_t rank, uint32_t n_ranks);
/* ExecuteCollisionDetection() = upload + run + fetch.  The three steps are exposed so that a caller
 * can time the device part with inputs already resident. */
int imrcd_frame_execute(imrcd_ctx* ctx);
int imrcd_frame_upload(imrcd_ctx* ctx);   /* host entries -> HBM (async on the stream) */
int imrcd_frame_run(imrcd_ctx* ctx);      /* all kernels; returns after the stream has drained */
int imrcd_frame_fetch(imrcd_ctx* ctx);    /* results -> pinned host memory */
/* imrcd_frame_run in two halves: run_async enqueues every kernel of the frame and returns at once, so that the caller can put more work on
 * the stream behind it (the end-of-frame collective on imrcd_frame_results_block) before the host waits; finish waits, reads the frame's
 * counters and returns 0, or 1 when a buffer had overflowed and the frame was run again (work enqueued in between saw stale results). */
int imrcd_frame_run_async(imrcd_ctx* ctx);
int imrcd_frame_finish(imrcd_ctx* ctx);
/* Results stay valid until the next imrcd_frame_reset / imrcd_frame_run.  With a communicator (imrcd_comm_init, imrcd_group_create)
 * `pairs` are the merged records of ALL ranks, in rank order; hits, imrcd_frame_pairs / _combos and the counters of the statistics stay
 * this rank's own. */
int imrcd_frame_results(imrcd_ctx* ctx, const imrcd_entity_pair** pairs, uint64_t* n_pairs,
                        const imrcd_tri_hit** hits, uint64_t* n_hits);
/* this rank's own colliding pairs, whatever the communicator merged */
int imrcd_frame_results_local(imrcd_ctx* ctx, const imrcd_entity_pair** pairs, uint64_t* n_pairs);
/* Broad-phase pair list of this shard as (first,second) entry indices; host copy made on demand. */
int imrcd_frame_pairs(imrcd_ctx* ctx, const uint32_t** pairs, uint64_t* n_pairs);
/* Leaf combos (pair, offA, offB, cntA | cntB<<16) in leaf order; test/diagnostic, host copy on demand. */
int imrcd_frame_combos(imrcd_ctx* ctx, const uint32_t** combos, uint64_t* n_combos);
int imrcd_frame_get_stats(imrcd_ctx* ctx, imrcd_frame_stats* out);
/* Device pointers of the result arrays (for an NCCL gather by the caller): hits, entity pairs. */
int imrcd_frame_results_device(imrcd_ctx* ctx, void** d_pairs, uint64_t* n_pairs, void** d_hits, uint64_t* n_hits);

/* The same records as one contiguous device block for a fixed-capacity collective: row 0 (80 B) starts with the u64 record count, the
 * records follow from row 1; `capacity` = rows allocated after the header (rows past the count hold stale data). */
int imrcd_frame_results_block(imrcd_ctx* ctx, void** d_block, uint64_t* n_pairs, uint64_t* capacity);

/* ---- multi-GPU: the end-of-frame merge inside the library ----------------------------------------------------------------
 * The reference is one host thread (CollisionDetection.cpp:44-129); its contract is that after ExecuteCollisionDetection the caller sees
 * the whole colliding set.  Here N contexts (one per GPU) each work on their shard of the frame and ONE ncclAllGather of fixed-capacity
 * blocks (header row = record count + overflow bits, then the 80-B records, as they lie in HBM) runs on the frame's own stream right
 * behind its kernels, so imrcd_frame_run / _run_async + _finish / _execute return the merged set with a single host wait.  Whether a
 * frame has to be run again (a buffer overflowed somewhere, or the blocks were too small) is decided from the gathered headers, i.e.
 * identically on every rank.  NCCL is bound at run time (dlopen libnccl.so.2): without it these calls return IMRCD_E_NODEVICE and
 * everything else works.
 *   one process per GPU : rank 0 calls imrcd_comm_unique_id and hands the 128 bytes to the others (any transport), every rank calls
 *                         imrcd_comm_init (collective);
 *   one process, N GPUs : imrcd_group_create (what SURVEY 8b sketches as imrcd_create(device_ids[], n)): the engine is a single process. */
#define IMRCD_COMM_ID_BYTES 128
int imrcd_comm_unique_id(void* id_out /* IMRCD_COMM_ID_BYTES */);
int imrcd_comm_init(imrcd_ctx* ctx, const void* id, uint32_t rank, uint32_t n_ranks);
int imrcd_comm_destroy(imrcd_ctx* ctx);
/* How the end-of-frame merge of this context travels: 0 = no communicator, 1 = ncclAllGather on the frame's stream, 2 = peer memory
 * (each rank's last kernel stores its block into every peer's buffer over NVLink and raises a flag; no NCCL call inside a frame).
 * Decided at the first frame after imrcd_comm_init / imrcd_group_create, collectively; IMRCD_P2P=0 in the environment forces 1. */
int imrcd_comm_transport(const imrcd_ctx* ctx);

typedef struct imrcd_group imrcd_group;
int         imrcd_group_create(const int* device_ids, uint32_t n, imrcd_group** out);
void        imrcd_group_destroy(imrcd_group* g);
uint32_t    imrcd_group_size(const imrcd_group* g);
imrcd_ctx*  imrcd_group_ctx(imrcd_group* g, uint32_t i);           /* for per-context calls (statistics, tree export, ...) */
const char* imrcd_group_last_error(const imrcd_group* g);
/* meshes are replicated: the same id on every context of the group */
int imrcd_group_mesh_create(imrcd_group* g, const float* positions, const float* normals, const uint32_t* vertex_ids, uint64_t n_tri,
                            uint32_t build_mode, uint32_t* mesh_id);
int imrcd_group_gltf_load(imrcd_group* g, const char* path, uint32_t build_mode, uint32_t* mesh_ids, uint32_t capacity, uint32_t* n_meshes);
/* Reset / AddCollisionDetectionEntry / ExecuteCollisionDetection over all GPUs of the group; results = the merged colliding pairs */
int imrcd_group_frame_reset(imrcd_group* g);
int imrcd_group_frame_add_entry(imrcd_group* g, const float current[16], const float previous[16], uint32_t mesh_id, uint8_t should_callback, uint32_t entity);
int imrcd_group_frame_add_entries(imrcd_group* g, uint64_t n, const float* current, const float* previous, const uint32_t* mesh_ids,
                                  const uint8_t* should_callback, const uint32_t* entities);
int imrcd_group_frame_execute(imrcd_group* g);
int imrcd_group_frame_results(imrcd_group* g, const imrcd_entity_pair** pairs, uint64_t* n_pairs);

/* ---- unit-level entry points used by the parity tests (device kernels on flat arrays) ---- */
int imrcd_test_sat(imrcd_ctx* ctx, uint64_t n, const float* boxes_a, const float* boxes_b, const float* mats /* n*16 or NULL */,
                   uint8_t* verdict, float* surface_a, float* surface_b);
int imrcd_test_tri_tri(imrcd_ctx* ctx, uint64_t n, const float* tris_a, const float* tris_b, const float* mat16 /* or NULL */,
                       uint8_t* flags, float* seg /* n*6 */);
int imrcd_test_pair_matrix(imrcd_ctx* ctx, uint64_t n, const float* a, const float* b, float* out /* n*16 */);
int imrcd_test_obb_fit(imrcd_ctx* ctx, uint64_t n_points, const float* points, float* out12);
/* the re-posed vertices (vec4 each) of the last imrcd_meshes_repose call, in the call's order */
int imrcd_test_reposed_vertices(imrcd_ctx* ctx, float* out, uint64_t n_vertices);
/* Ray::IntersectOBBtree (Ray.cpp:136-236) on n rays against one mesh: mats n*16 (tree -> ray space), origins / directions n*3;
 * flags bit0 doIntersect, bit1 itBackfaces; out3 = distanceFromOrigin, baryPosition.x, .y; tri = LEAF-ORDER triangle index (0xffffffff: none) */
int imrcd_test_ray_tree(imrcd_ctx* ctx, uint32_t mesh_id, uint64_t n, const float* mats, const float* origins, const float* directions,
                        uint8_t* flags, float* out3, uint32_t* tri);

#ifdef __cplusplus
}
#endif
#endif /* IMRCD_H */
