// xrank.cuh -- exchange of a few doubles per rank over NVLink peer memory, shared by the stand-alone reduction kernel
// (xrank.cu) and the fused loss' finishing kernel (loss.cu).  See xrank.cu for the protocol.
#pragma once
#include "y3d_common.cuh"

namespace y3d {

constexpr int kXMaxWorld = 64;
constexpr int kXMaxVals = 16;  // doubles per rank and call
struct XSlot {
    double v[kXMaxVals];
    unsigned long long seq;
    unsigned long long pad;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Called by ALL threads of one CTA (at least max(world, n_vals) of them; contains barriers).  bufs[r] = rank r's
// exchange buffer XSlot[2][world].  vals: this rank's n_vals doubles (shared or global memory, written before the call
// and made visible by a barrier); sum: shared double[n_vals] receiving the rank-ordered sums (NaN when a peer never
// arrived); *failed: shared int.
__device__ __forceinline__ void xrank_allreduce(XSlot *const *bufs, int rank, int world, unsigned long long seq,
                                                const double *vals, int n_vals, double *sum, int *failed) {
    const int tid = threadIdx.x;
    const int par = (int)(seq & 1ull);
    if (tid == 0) *failed = 0;
    __syncthreads();
    if (tid < world) {  // thread r talks to rank r
        XSlot *dst = bufs[tid] + (size_t)par * world + rank;  // my slot in rank `tid`'s buffer
        for (int j = 0; j < n_vals; ++j) dst->v[j] = vals[j];
        __threadfence_system();
        st_release_sys(&dst->seq, seq);
        const XSlot *src = bufs[rank] + (size_t)par * world + tid;  // rank `tid`'s slot in my buffer
        const long long t0 = clock64();
        while (ld_acquire_sys(&src->seq) != seq) {
            if (clock64() - t0 > (1ll << 34)) {  // ~8 s: a peer never arrived; fail loudly instead of hanging the GPU
                *failed = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (tid < n_vals) {
        double s = 0.0;
        const XSlot *mine = bufs[rank] + (size_t)par * world;
        for (int r = 0; r < world; ++r) s += mine[r].v[tid];  // rank order: the same sum on every rank
        sum[tid] = *failed ? __longlong_as_double(0x7ff8000000000000ll) : s;
    }
    __syncthreads();
}

}  // namespace y3d
