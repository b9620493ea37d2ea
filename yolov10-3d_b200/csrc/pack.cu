// pack.cu -- ground-truth packing for the loss entry points (sm_100a).
//   y3d_pack_targets   v8DetectionLoss.preprocess   reference ultralytics/utils/loss.py:180-195 (+ xywh2xyxy ops.py:403-422)
//                      DDDetectionLoss.preprocess   reference ultralytics/utils/loss.py:795-810
// The reference loops over the images in Python (one boolean mask, one masked copy and one host sync per image);
// here one CTA per image scans the ragged [N, ...] rows 256 at a time, keeps the rows of its image in their order of
// appearance (ballot + prefix over the warps: exactly what `targets[matches]` gives) and writes them, the zero
// padding and the image's row count -- no memset, no atomics.
#include "y3d_common.cuh"

namespace y3d {

constexpr int kPackThreads = 256;

// grid B, block 256
__global__ void __launch_bounds__(kPackThreads) pack_targets_kernel(const float *__restrict__ batch_idx,
                                                                    const float *__restrict__ cls,
                                                                    const float *__restrict__ bboxes,
                                                                    const float *__restrict__ extra, int n_extra, int N,
                                                                    int M, float img_w, float img_h,
                                                                    float *__restrict__ out, int *__restrict__ counts) {
    __shared__ int s_w[kPackThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int W = 5 + n_extra;
    float *ob = out + (long long)b * M * W;
    for (int j = tid; j < M * W; j += kPackThreads) ob[j] = 0.0f;  // zero rows = padding
    __syncthreads();
    int base = 0;  // rows of this image seen so far (uniform over the CTA)
    for (int i0 = 0; i0 < N; i0 += kPackThreads) {
        const int i = i0 + tid;
        const bool ok = i < N && (int)batch_idx[i] == b;
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_w[wid] = __popc(bal);
        __syncthreads();
        int before = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kPackThreads / 32; ++w) {
            const int v = s_w[w];
            before += w < wid ? v : 0;
            tot += v;
        }
        const int rank = base + before + __popc(bal & lt);
        if (ok && rank < M) {
            float *o = ob + (long long)rank * W;
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
        base += tot;
        __syncthreads();
    }
    if (tid == 0) counts[b] = base;
}

}  // namespace y3d

using namespace y3d;

extern "C" int y3d_pack_targets(const float *batch_idx, const float *cls, const float *bboxes, const float *extra,
                                int n_extra, int N, int B, int M, float img_w, float img_h, float *out, int *counts,
                                void *stream) {
    if (N < 0 || B < 1 || M < 0 || n_extra < 0 || !counts) return Y3D_EINVAL;
    if (N > 0 && (!batch_idx || !cls || !bboxes || (n_extra > 0 && !extra))) return Y3D_EINVAL;
    if (M > 0 && !out) return Y3D_EINVAL;
    pack_targets_kernel<<<B, kPackThreads, 0, (cudaStream_t)stream>>>(batch_idx, cls, bboxes, extra, n_extra, N, M, img_w,
                                                                      img_h, out, counts);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}
