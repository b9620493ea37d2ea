// pack.cu -- ground-truth packing for the loss entry points (sm_100a).
//   y3d_pack_targets   v8DetectionLoss.preprocess   reference ultralytics/utils/loss.py:180-195 (+ xywh2xyxy ops.py:403-422)
//                      DDDetectionLoss.preprocess   reference ultralytics/utils/loss.py:795-810
// The reference loops over the images in Python (one boolean mask, one masked copy and one host sync per image);
// here one warp walks the ragged [N, ...] rows in order, 32 at a time, and scatters them into the padded
// [B, M, 5 + E] tensor: row order inside an image is the order of appearance, exactly as `targets[matches]` gives.
#include "y3d_common.cuh"

namespace y3d {

// single warp: per-image running counters live in `counts` (zero-initialised by the caller's memset)
__global__ void __launch_bounds__(32) pack_targets_kernel(const float *__restrict__ batch_idx,
                                                          const float *__restrict__ cls,
                                                          const float *__restrict__ bboxes,
                                                          const float *__restrict__ extra, int n_extra, int N, int B,
                                                          int M, float img_w, float img_h, float *__restrict__ out,
                                                          int *__restrict__ counts) {
    const int lane = threadIdx.x;
    const unsigned lt = (1u << lane) - 1u;
    const int W = 5 + n_extra;
    for (int i0 = 0; i0 < N; i0 += 32) {
        const int i = i0 + lane;
        const bool on = i < N;
        const int b = on ? (int)batch_idx[i] : -1;
        const bool ok = on && b >= 0 && b < B;
        const unsigned act = __ballot_sync(0xffffffffu, ok);
        int rank = -1;
        if (ok) {
            const unsigned peers = __match_any_sync(act, b);  // rows of the same image in this group of 32
            const int before = counts[b];
            rank = before + __popc(peers & lt);
            __syncwarp(act);
            if (lane == __ffs(peers) - 1) counts[b] = before + __popc(peers);
        }
        __syncwarp();
        if (ok && rank < M) {
            float *o = out + ((long long)b * M + rank) * W;
            // out[..., 1:5].mul_(scale) then xywh2xyxy (loss.py:194 / ops.py:418-421)
            const float x = __fmul_rn(bboxes[4 * i], img_w), y = __fmul_rn(bboxes[4 * i + 1], img_h);
            const float dw = __fmul_rn(bboxes[4 * i + 2], img_w) / 2.0f, dh = __fmul_rn(bboxes[4 * i + 3], img_h) / 2.0f;
            o[0] = cls[i];
            o[1] = __fsub_rn(x, dw);
            o[2] = __fsub_rn(y, dh);
            o[3] = __fadd_rn(x, dw);
            o[4] = __fadd_rn(y, dh);
            for (int e = 0; e < n_extra; ++e) o[5 + e] = extra[(long long)i * n_extra + e];
        }
    }
}

}  // namespace y3d

using namespace y3d;

extern "C" int y3d_pack_targets(const float *batch_idx, const float *cls, const float *bboxes, const float *extra,
                                int n_extra, int N, int B, int M, float img_w, float img_h, float *out, int *counts,
                                void *stream) {
    if (N < 0 || B < 1 || M < 0 || n_extra < 0 || !counts) return Y3D_EINVAL;
    if (N > 0 && (!batch_idx || !cls || !bboxes || (n_extra > 0 && !extra))) return Y3D_EINVAL;
    if (M > 0 && !out) return Y3D_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)B, s);
    if (e != cudaSuccess) return (int)e;
    if (M > 0) {
        e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * M * (5 + n_extra), s);  // zero rows = padding
        if (e != cudaSuccess) return (int)e;
    }
    if (N == 0) return Y3D_OK;
    pack_targets_kernel<<<1, 32, 0, s>>>(batch_idx, cls, bboxes, extra, n_extra, N, B, M, img_w, img_h, out, counts);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}
