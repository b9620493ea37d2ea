// xrank.cuh -- exchange of a few doubles per rank over NVLink peer memory, shared by the stand-alone reduction kernel
// (xrank.cu) and the fused loss' finishing kernel (loss.cu).
//
// Every rank owns one exchange buffer that all peers can address (symmetric memory), XSlot[2][world].  A call pushes
// this rank's values straight into its slot of every peer's buffer (P2P stores through NVSwitch) and waits until the
// slots of all peers have arrived in its own buffer.  Flag-in-data protocol: every 64-bit word carries 32 bits of
// payload and the low 32 bits of the call's sequence number, and a naturally aligned 64-bit store is single-copy
// atomic -- so a word is either old or complete, no word depends on another, and neither a fence nor a separate flag
// (one more NVLink round trip) is needed.  A double travels as two words.  Slots are double-buffered by the parity of
// the sequence number: a peer can be at most one call ahead, because it needs this rank's next words to finish that
// call.
#pragma once
#include <cstdlib>

#include "y3d_common.cuh"

namespace y3d {

constexpr int kXMaxWorld = 64;
constexpr int kXMaxVals = 16;  // doubles per rank and call
struct XSlot {
    unsigned long long w[2 * kXMaxVals];  // (payload << 32) | (uint32) seq
};

__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// How long a rank waits for its peers before it gives up (sums = NaN, *failed = 1).  Like an NCCL collective the call
// is meant to wait: ranks legitimately arrive far apart (rank 0 validates and checkpoints while the others are already
// in the next step's loss), so the default is ten minutes; Y3D_XRANK_TIMEOUT_S overrides it (seconds, host side).
inline long long xrank_timeout_cycles() {
    static long long cycles = 0;
    if (cycles == 0) {
        double sec = 600.0;
        if (const char *e = getenv("Y3D_XRANK_TIMEOUT_S")) {
            const double v = atof(e);
            if (v > 0.0) sec = v;
        }
        cycles = (long long)(sec * 2.0e9);  // clock64 ticks at the SM clock (< 2 GHz)
    }
    return cycles;
}

// Called by ALL threads of one CTA (at least max(world, n_vals) of them; contains barriers).  bufs[r] = rank r's
// exchange buffer.  vals: this rank's n_vals doubles (shared or global memory, written before the call and made
// visible by a barrier); sum: shared double[n_vals] receiving the rank-ordered sums (NaN when a peer never arrived);
// *failed: shared int.
__device__ __forceinline__ void xrank_allreduce(XSlot *const *bufs, int rank, int world, unsigned long long seq,
                                                const double *vals, int n_vals, double *sum, int *failed,
                                                long long timeout_cycles) {
    __shared__ double recv[kXMaxWorld][kXMaxVals];
    const int tid = threadIdx.x;
    const int par = (int)(seq & 1ull);
    const unsigned long long flag = seq & 0xffffffffull;
    if (tid == 0) *failed = 0;
    __syncthreads();
    if (tid < world) {  // thread r talks to rank r
        XSlot *dst = bufs[tid] + (size_t)par * world + rank;  // my slot in rank `tid`'s buffer
        for (int j = 0; j < n_vals; ++j) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[j]);
            st_relaxed_sys(&dst->w[2 * j], (bits & 0xffffffff00000000ull) | flag);
            st_relaxed_sys(&dst->w[2 * j + 1], (bits << 32) | flag);
        }
        const XSlot *src = bufs[rank] + (size_t)par * world + tid;  // rank `tid`'s slot in my buffer
        const long long t0 = clock64();
        bool dead = false;
        for (int j = 0; j < n_vals && !dead; ++j) {
            unsigned long long hi, lo;
            while (((hi = ld_relaxed_sys(&src->w[2 * j])) & 0xffffffffull) != flag ||
                   ((lo = ld_relaxed_sys(&src->w[2 * j + 1])) & 0xffffffffull) != flag) {
                if (clock64() - t0 > timeout_cycles) {  // a peer never arrived (default: ten minutes, xrank_timeout_cycles)
                    dead = true;
                    break;
                }
            }
            if (!dead) recv[tid][j] = __longlong_as_double((long long)((hi & 0xffffffff00000000ull) | (lo >> 32)));
        }
        if (dead) *failed = 1;
    }
    __syncthreads();
    if (tid < n_vals) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += recv[r][tid];  // rank order: the same sum on every rank
        sum[tid] = *failed ? __longlong_as_double(0x7ff8000000000000ll) : s;
    }
    __syncthreads();
}

}  // namespace y3d
