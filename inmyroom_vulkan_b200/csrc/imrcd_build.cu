// imrcd_build.cu -- GPU OBB-tree construction (placeholder until the Morton build lands).
#include "imrcd_internal.cuh"

int imr_build_mesh_device(imrcd_ctx* ctx, const float*, const float*, const uint32_t*, uint64_t, uint32_t, MeshHost*) {
    ctx->err = "imrcd_mesh_create: GPU build not implemented yet";
    return IMRCD_E_STATE;
}
extern "C" int imrcd_test_obb_fit(imrcd_ctx* ctx, uint64_t, const float*, float*) {
    if (!ctx) return IMRCD_E_ARG;
    ctx->err = "imrcd_test_obb_fit: not implemented yet";
    return IMRCD_E_STATE;
}
