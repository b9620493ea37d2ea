// xrank.cu -- the one cross-rank step of the image-sharded loss, as a single kernel over NVLink peer memory (sm_100a).
//   y3d_loss_allreduce_finalize : sum of the per-rank un-normalised loss partials (8 doubles for v10DetectLoss) over
//       all ranks + the target_scores_sum normalisation of reference ultralytics/utils/loss.py:240-256.
// The exchange protocol (flag-in-data words over peer memory, no fence, no separate flag) is in xrank.cuh; the fused
// loss runs the same exchange inside its finishing kernel (y3d_v10_loss_fwd_sharded, loss.cu).
#include "xrank.cuh"

namespace y3d {

struct XPeers {
    XSlot *buf[kXMaxWorld];  // buf[r] = rank r's exchange buffer: XSlot[kXRing][world]
};

// one CTA of 64 threads; thread r talks to rank r
__global__ void __launch_bounds__(64) loss_allreduce_finalize_kernel(XPeers peers, const double *__restrict__ partials,
                                                                     int n_vals, int n_branch, int rank, int world,
                                                                     unsigned long long seq, float gain_box,
                                                                     float gain_cls, float gain_dfl,
                                                                     float *__restrict__ loss_items,
                                                                     double *__restrict__ global_partials,
                                                                     int *__restrict__ status, long long timeout_cycles) {
    __shared__ double sum[kXMaxVals];
    __shared__ int failed;
    const int tid = threadIdx.x;
    xrank_allreduce(peers.buf, rank, world, seq, partials, n_vals, sum, &failed, timeout_cycles);
    if (tid < n_vals && global_partials) global_partials[tid] = sum[tid];
    if (tid < n_branch) {
        const double tss = sum[4 * tid + 3] > 1.0 ? sum[4 * tid + 3] : 1.0;  // max(target_scores.sum(), 1) loss.py:240
        loss_items[4 * tid + 0] = (float)(sum[4 * tid + 0] / tss * gain_box);
        loss_items[4 * tid + 1] = (float)(sum[4 * tid + 1] / tss * gain_cls);
        loss_items[4 * tid + 2] = (float)(sum[4 * tid + 2] / tss * gain_dfl);
        loss_items[4 * tid + 3] = (float)tss;
    }
    if (tid == 0 && status) *status = failed;
}

// the collecting half alone: the partials were posted by the loss' finishing kernel (y3d_v10_loss_fwd_sharded, defer = 1)
__global__ void __launch_bounds__(64) loss_exchange_resolve_kernel(XPeers peers, int n_vals, int n_branch, int rank, int world,
                                                                   unsigned long long seq, float gain_box, float gain_cls,
                                                                   float gain_dfl, float total_scale,
                                                                   float *__restrict__ loss_items,
                                                                   float *__restrict__ loss_total,
                                                                   double *__restrict__ global_partials,
                                                                   int *__restrict__ status, long long timeout_cycles) {
    __shared__ double sum[kXMaxVals];
    __shared__ int failed;
    const int tid = threadIdx.x;
    xrank_collect(peers.buf, rank, world, seq, n_vals, sum, &failed, timeout_cycles);
    if (tid < n_vals && global_partials) global_partials[tid] = sum[tid];
    if (tid < n_branch) {
        const double tss = sum[4 * tid + 3] > 1.0 ? sum[4 * tid + 3] : 1.0;  // max(target_scores.sum(), 1) loss.py:240
        loss_items[4 * tid + 0] = (float)(sum[4 * tid + 0] / tss * gain_box);
        loss_items[4 * tid + 1] = (float)(sum[4 * tid + 1] / tss * gain_cls);
        loss_items[4 * tid + 2] = (float)(sum[4 * tid + 2] / tss * gain_dfl);
        loss_items[4 * tid + 3] = (float)tss;
    }
    if (tid == 0 && loss_total) {  // loss.sum() * batch_size per branch, added up (loss.py:257, 736)
        float total = 0.f;
        for (int z = 0; z < n_branch; ++z) {
            const double tss = sum[4 * z + 3] > 1.0 ? sum[4 * z + 3] : 1.0;
            const float i0 = (float)(sum[4 * z + 0] / tss * gain_box), i1 = (float)(sum[4 * z + 1] / tss * gain_cls),
                        i2 = (float)(sum[4 * z + 2] / tss * gain_dfl);
            total += ((i0 + i1) + i2) * total_scale;
        }
        *loss_total = total;
    }
    if (tid == 0 && status && failed) *status = 1;
}

}  // namespace y3d

using namespace y3d;

extern "C" size_t y3d_xrank_buffer_bytes(int world) { return sizeof(XSlot) * kXRing * (size_t)(world > 0 ? world : 1); }

extern "C" int y3d_loss_allreduce_finalize(const double *partials, int n_branch, int rank, int world,
                                           void *const *peer_bufs, unsigned long long seq, float gain_box,
                                           float gain_cls, float gain_dfl, float *loss_items, double *global_partials,
                                           int *status, void *stream) {
    if (!partials || !peer_bufs || !loss_items || n_branch < 1 || 4 * n_branch > kXMaxVals) return Y3D_EINVAL;
    if (world < 1 || world > kXMaxWorld || rank < 0 || rank >= world || seq == 0) return Y3D_EINVAL;
    XPeers P{};
    for (int r = 0; r < world; ++r) {
        if (!peer_bufs[r] || ((uintptr_t)peer_bufs[r]) % 16) return Y3D_EALIGN;
        P.buf[r] = (XSlot *)peer_bufs[r];
    }
    loss_allreduce_finalize_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(P, partials, 4 * n_branch, n_branch, rank, world, seq,
                                                                      gain_box, gain_cls, gain_dfl, loss_items,
                                                                      global_partials, status, xrank_timeout_cycles());
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

extern "C" int y3d_loss_exchange_resolve(int n_branch, int rank, int world, void *const *peer_bufs, unsigned long long seq,
                                         float gain_box, float gain_cls, float gain_dfl, float total_scale,
                                         float *loss_items, float *loss_total, double *global_partials, int *status,
                                         void *stream) {
    if (!peer_bufs || !loss_items || n_branch < 1 || 4 * n_branch > kXMaxVals) return Y3D_EINVAL;
    if (world < 1 || world > kXMaxWorld || rank < 0 || rank >= world || seq == 0) return Y3D_EINVAL;
    XPeers P{};
    for (int r = 0; r < world; ++r) {
        if (!peer_bufs[r] || ((uintptr_t)peer_bufs[r]) % 16) return Y3D_EALIGN;
        P.buf[r] = (XSlot *)peer_bufs[r];
    }
    loss_exchange_resolve_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(P, 4 * n_branch, n_branch, rank, world, seq, gain_box,
                                                                    gain_cls, gain_dfl, total_scale, loss_items, loss_total,
                                                                    global_partials, status, xrank_timeout_cycles());
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}
