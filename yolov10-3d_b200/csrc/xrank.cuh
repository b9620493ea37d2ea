// xrank.cuh -- exchange of a few doubles per rank over NVLink peer memory, shared by the stand-alone reduction kernel
// (xrank.cu) and the fused loss' finishing kernel (loss.cu).
//
// Every rank owns one exchange buffer that all peers can address (symmetric memory), XSlot[kXRing][world].  A call pushes
// this rank's values straight into its slot of every peer's buffer (P2P stores through NVSwitch) and waits until the
// slots of all peers have arrived in its own buffer.  Flag-in-data protocol: every 64-bit word carries 32 bits of
// payload and the low 32 bits of the call's sequence number, and a naturally aligned 64-bit store is single-copy
// atomic -- so a word is either old or complete, no word depends on another, and neither a fence nor a separate flag
// (one more NVLink round trip) is needed.  A double travels as two words.  The buffer holds the slots of kXRing = 4
// consecutive calls (slot = seq % 4).  Undeferred calls (post + collect in one kernel) keep the ranks within one call of
// each other, because a rank needs its peers' words of call j to finish call j.  Deferred calls (post in the loss' last
// kernel, collect later on a side stream) run under the rule "a rank posts call j only after its own collect of call
// j - 2 has completed" (dist.PeerLossReducer.pre_post): the collect of j - 2 needs every peer's post of j - 2, which that
// peer issued after ITS collect of j - 4 -- so when call j's words overwrite those of call j - 4, every peer has read
// them.  Four slots are what this one call of slack costs.
#pragma once
#include <cstdlib>

#include "y3d_common.cuh"

namespace y3d {

constexpr int kXMaxWorld = 64;
constexpr int kXMaxVals = 16;  // doubles per rank and call
constexpr int kXRing = 4;      // calls whose slots coexist in an exchange buffer (see below)
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

// The exchange in two halves, so that a caller can post early and collect late (the round trip and the wait for the slowest
// rank then overlap whatever runs in between -- the next step's streaming pass, or the host's way to the backward call).
//
// xrank_post: threads tid < world; thread r stores this rank's n_vals doubles into its slot of rank r's buffer.
__device__ __forceinline__ void xrank_post(XSlot *const *bufs, int rank, int world, unsigned long long seq, const double *vals,
                                           int n_vals) {
    const int tid = threadIdx.x;
    if (tid >= world) return;
    const int par = (int)(seq % kXRing);
    const unsigned long long flag = seq & 0xffffffffull;
    XSlot *dst = bufs[tid] + (size_t)par * world + rank;  // my slot in rank `tid`'s buffer
    for (int j = 0; j < n_vals; ++j) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[j]);
        st_relaxed_sys(&dst->w[2 * j], (bits & 0xffffffff00000000ull) | flag);
        st_relaxed_sys(&dst->w[2 * j + 1], (bits << 32) | flag);
    }
}

// xrank_collect: called by ALL threads of one CTA (at least max(world, n_vals) of them; contains barriers).  Waits until
// the slots of all ranks have arrived in this rank's buffer and sums them in rank order (the same sum on every rank; NaN
// and *failed = 1 when a peer never arrived).  sum: shared double[n_vals]; failed: shared int.
__device__ __forceinline__ void xrank_collect(XSlot *const *bufs, int rank, int world, unsigned long long seq, int n_vals,
                                              double *sum, int *failed, long long timeout_cycles) {
    __shared__ double recv[kXMaxWorld][kXMaxVals];
    const int tid = threadIdx.x;
    const int par = (int)(seq % kXRing);
    const unsigned long long flag = seq & 0xffffffffull;
    if (tid == 0) *failed = 0;
    __syncthreads();
    if (tid < world) {
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

// post + collect in one go.  vals: this rank's n_vals doubles (shared or global memory, written before the call and made
// visible by a barrier).
__device__ __forceinline__ void xrank_allreduce(XSlot *const *bufs, int rank, int world, unsigned long long seq,
                                                const double *vals, int n_vals, double *sum, int *failed,
                                                long long timeout_cycles) {
    xrank_post(bufs, rank, world, seq, vals, n_vals);
    xrank_collect(bufs, rank, world, seq, n_vals, sum, failed, timeout_cycles);
}

}  // namespace y3d
