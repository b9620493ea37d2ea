// abi.cu -- error strings, ABI version and workspace sizing of liby3d_b200.
#include "y3d_common.cuh"

namespace y3d {
size_t topk_workspace_bytes(int B, int A, int nc, int D);
size_t assign_workspace_bytes(int B, int A, int M);
size_t loss_workspace_bytes(int B, int A, int M, int k);
size_t dd_loss_workspace_bytes(int B, int A, int M);
}  // namespace y3d

extern "C" const char *y3d_strerror(int rc) {
    switch (rc) {
        case Y3D_OK: return "ok";
        case Y3D_EINVAL: return "y3d: invalid argument (shape, null pointer or flag)";
        case Y3D_EUNSUPPORTED: return "y3d: configuration outside the compiled limits";
        case Y3D_EALIGN: return "y3d: pointer or stride alignment requirement not met";
        case Y3D_EWORKSPACE: return "y3d: workspace missing or too small (see y3d_workspace_bytes)";
        default: break;
    }
    if (rc > 0) return cudaGetErrorString((cudaError_t)rc);
    return "y3d: unknown error";
}

extern "C" int y3d_abi_version(void) { return 2; }

extern "C" size_t y3d_workspace_bytes(int stage, int B, int A, int nc, int M, int k, int D) {
    switch (stage) {
        case Y3D_STAGE_POSTPROCESS:
        case Y3D_STAGE_DECODE_TOPK: return y3d::topk_workspace_bytes(B, A, nc, D) + 256;
        case Y3D_STAGE_TAL_ASSIGN:
        case Y3D_STAGE_TAL_ASSIGN3D: return y3d::assign_workspace_bytes(B, A, M) + 256;
        case Y3D_STAGE_V8_LOSS: return y3d::loss_workspace_bytes(B, A, M, k) + 256;
        case Y3D_STAGE_DD_LOSS: return y3d::dd_loss_workspace_bytes(B, A, M) + 256;
        default: return 0;
    }
}
