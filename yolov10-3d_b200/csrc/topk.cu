// topk.cu -- NMS-free two-stage top-k (sm_100a).
//   y3d_postprocess    ops.v10postprocess / ops.v10_3Dpostprocess   reference ultralytics/utils/ops.py:852-880
//   y3d_decode_topk2d  v10Detect.forward export branch               reference ultralytics/nn/modules/head.py:526-531
//
// Stage 0 (streaming, many CTAs): per-anchor class max  -> keys [B,A]           (scores.amax(-1), ops.py:855)
// Stage 1 (one CTA per image)   : exact top-D of the A keys by 4-pass MSB radix select on order-preserving
//                                 uint32 keys, ties resolved lowest-index-first, then a bitonic sort of the D
//                                 winners on (key desc, index asc)                 (torch.topk, ops.py:856)
// Stage 2 (same CTA)            : the D x nc scores of the winners are gathered into shared memory, the same
//                                 select + sort runs over the flattened index i*nc + c (ops.py:861), then
//                                 labels = idx % nc, anchor = winners[idx // nc], gather of the regression channels.
// Fused export branch (y3d_decode_topk2d): stage 0 reads the class logits of the head levels (cls_max_kernel), stages 1-2
// rank by the exactly specified sigmoid, and the boxes of the D winners are decoded by box_decode_kernel, a machine-wide
// grid behind the selection kernel -- which also carries the peer-memory epilogue of the image-sharded path.
// Results are identical to a stable descending sort (lowest index wins ties) -- the order BASELINE.json mandates.
#include <cstdlib>

#include "y3d_common.cuh"

namespace y3d {

constexpr int kTopkThreads = 1024;
constexpr int kKeys2SmemCap = 40960;  // uint32 keys kept in shared memory (160 KB; + 64 KB of lists / histograms at D = 1024)

__device__ __forceinline__ uint32_t float_key(float x) {
    if (x != x) return 0xFFFFFFFFu;  // NaN sorts first, like torch.topk
    x = x + 0.0f;                    // -0 -> +0 (equal values must tie)
    uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(u);
}

constexpr int kGatherMaxWorld = 16;
// Image-sharded detection path: every rank keeps the detections of ALL ranks in a buffer its peers
// can address (symmetric memory): [2 parities][world * n_local][D][6] floats (the result), then a staging area
// [2][world * n_local][6 D] 64-bit words the peers write into.  Flag-in-data, as in xrank.cuh: a finished row travels as
// six aligned 64-bit stores (float bits << 32 | sequence number of the call) -- single-copy atomic, so a word is either
// old or complete and neither a system-scope fence nor a flag store follows the data (both were measured: the fences
// cost 15 us per step, the release stores of the flags 11 us).  gather_collect_kernel, one CTA per remote image, polls
// the staged words of the call and unpacks them into the result.  Parity = seq & 1: a peer can be at most one call ahead
// of a rank that is still reading.
struct GatherOut {
    float *data[kGatherMaxWorld];                // per rank: its buffer's result rows of this call's parity
    unsigned long long *stage[kGatherMaxWorld];  // per rank: its buffer's staging words of this call's parity
    int rank, world, n_local;
    unsigned seq;
};

// grid ((world - 1) * n_local), block 256: CTA = one image of one peer.  Launched as a programmatic dependent of
// box_decode_kernel (it reads nothing of this rank's own kernels: it can poll while they still run); it waits for that
// grid at its very end, so that "this kernel has completed" keeps meaning "all rows, own and remote, are in place".
__global__ void __launch_bounds__(256) gather_collect_kernel(const __grid_constant__ GatherOut G, int D,
                                                             long long timeout_cycles, int *status) {
    const int pi = blockIdx.x / G.n_local, b = blockIdx.x - pi * G.n_local;
    const int pr = pi + (pi >= G.rank ? 1 : 0);  // the peers in rank order, this rank left out
    const long long slot = (long long)pr * G.n_local + b;
    const unsigned long long *src = G.stage[G.rank] + slot * 6 * D;
    float *dst = G.data[G.rank] + slot * 6 * D;
    const long long t0 = clock64();
    bool dead = false;
    for (int i = threadIdx.x; i < 6 * D && !dead; i += blockDim.x) {
        unsigned long long w;
        for (;;) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(src + i) : "memory");
            if ((unsigned)(w & 0xffffffffull) == G.seq) break;
            if (clock64() - t0 > timeout_cycles) { dead = true; break; }
        }
        if (!dead) dst[i] = __uint_as_float((unsigned)(w >> 32));
    }
    if (dead && status) *status = 1;
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

struct TopkSrc {
    int mode;  // 0: strided preds tensor; 1: head levels (sigmoid + DFL decode on the fly)
    const float *preds;
    long long sB, sA, sC;
    int soff, roff;  // first score / first regression channel
    LevelTable t;
    int xywh;
};

__device__ __forceinline__ float src_score(const TopkSrc &s, int b, int a, int c) {
    if (s.mode == 0) return s.preds[b * s.sB + a * s.sA + (long long)(s.soff + c) * s.sC];
    int l = level_of(s.t, a);
    const float *p = s.t.ptr[l] + (long long)b * s.t.sB[l] + (a - s.t.start[l]);
    return dm::sigmoid_(p[(long long)(64 + c) * s.t.sC[l]]);  // the exactly specified sigmoid: scores decide indices here
}

// ------------------------------------------------------------------------------------------ stage 0 kernels
// Per anchor: the largest class key (amax of ops.py:855), its class (first maximum) and the second largest key.
// With the second key the selection kernel knows, without touching the scores again, whether an anchor can contribute
// more than its best class to the final top-D: only then are its nc scores gathered.
struct Top2 {
    uint32_t k1 = 0, k2 = 0;  // keys are order-preserving uint32 images of the floats (0 is below every real key)
    int arg = 0;
};
__device__ __forceinline__ void top2_push(Top2 &t, uint32_t k, int c) {
    if (k > t.k1) { t.k2 = t.k1; t.k1 = k; t.arg = c; }
    else if (k > t.k2) t.k2 = k;
}
__device__ __forceinline__ void top2_store(const Top2 &t, uint32_t *keys, int2 *aux, long long o) {
    keys[o] = t.k1;
    aux[o] = make_int2((int)t.k2, t.arg);
}

// generic strided layout: one warp per anchor, lanes over classes (coalesced when sC == 1)
__global__ void amax_warp_kernel(const float *__restrict__ preds, long long sB, long long sA, long long sC, int soff,
                                 int B, int A, int nc, uint32_t *__restrict__ keys, int2 *__restrict__ aux) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B * A) return;
    int b = warp / A, a = warp % A;
    const float *p = preds + b * sB + a * sA + (long long)soff * sC;
    Top2 t;
    for (int c = lane; c < nc; c += 32) top2_push(t, float_key(p[c * sC]), c);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {  // merge: larger k1 wins, lower class on ties; k2 = best of the rest
        const uint32_t ok1 = __shfl_xor_sync(0xffffffffu, t.k1, o), ok2 = __shfl_xor_sync(0xffffffffu, t.k2, o);
        const int oarg = __shfl_xor_sync(0xffffffffu, t.arg, o);
        const bool mine = t.k1 > ok1 || (t.k1 == ok1 && t.arg < oarg);
        const uint32_t loser = mine ? ok1 : t.k1;
        const uint32_t second = max(max(t.k2, ok2), loser);
        if (!mine) { t.k1 = ok1; t.arg = oarg; }
        t.k2 = second;
    }
    if (lane == 0) top2_store(t, keys, aux, warp);
}
// anchor-contiguous layout (sA == 1, the reference's permuted view of [B, C, A]): one thread per anchor
__global__ void amax_anchor_kernel(const float *__restrict__ preds, long long sB, long long sC, int soff, int B, int A,
                                   int nc, uint32_t *__restrict__ keys, int2 *__restrict__ aux) {
    int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (a >= A) return;
    const float *p = preds + b * sB + a + (long long)soff * sC;
    Top2 t;
    int c = 0;
    for (; c + 4 <= nc; c += 4) {
        float v0 = ldg_stream1(p + (c + 0) * sC), v1 = ldg_stream1(p + (c + 1) * sC);
        float v2 = ldg_stream1(p + (c + 2) * sC), v3 = ldg_stream1(p + (c + 3) * sC);
        top2_push(t, float_key(v0), c); top2_push(t, float_key(v1), c + 1);
        top2_push(t, float_key(v2), c + 2); top2_push(t, float_key(v3), c + 3);
    }
    for (; c < nc; ++c) top2_push(t, float_key(ldg_stream1(p + c * sC)), c);
    top2_store(t, keys, aux, (long long)b * A + a);
}
// head levels -> sigmoid(logit) keys: 4 anchors per thread, 128-bit loads; class channels only are read.
// The two largest LOGITS of an anchor are tracked and only those go through the (exactly specified, weakly monotone)
// sigmoid: largest score = sigmoid(largest logit), second largest score = sigmoid(second largest logit).  The class of
// the largest logit is the first class of the largest score unless another class ties with it in score -- and then the
// second key equals the first, the anchor counts as one that can contribute twice, and the selection kernel ranks its
// classes from their scores without looking at this argmax.
// The two largest LOGITS of an anchor in the float domain (first maximum, like the keys: -0 == +0; NaN never wins here --
// the caller notes it and falls back to the key path, where NaN sorts first like torch.topk).
struct Top2f {
    float k1 = -INFINITY, k2 = -INFINITY;
    int arg = 0;
};
__device__ __forceinline__ void top2f_push(Top2f &t, float v, int c) {
    const bool first = v > t.k1;
    const float lo = first ? t.k1 : v;  // the smaller of (old maximum, v) when v takes over, v itself otherwise
    t.k2 = fmaxf(t.k2, lo);
    t.k1 = first ? v : t.k1;
    t.arg = first ? c : t.arg;
}

// grid (ceil(quads / 32), B), block 128 = 32 units of VEC anchors x 4 class parts (warp = part: classes [p*nc/4, ...)),
// up to ten channel rows in flight per thread; the parts' top-2 meet in shared memory, part 0 merges them in class order
// (strict comparisons keep the first maximum) and writes the keys.
template <int VEC>
__global__ void __launch_bounds__(128) cls_max_kernel(const __grid_constant__ LevelTable t, int nq_total, int nc, int A,
                                                      uint32_t *__restrict__ keys, int2 *__restrict__ aux) {
    __shared__ float s_k1[3][32 * VEC], s_k2[3][32 * VEC];
    __shared__ int s_arg[3][32 * VEC];
    __shared__ int s_nan;
    if (threadIdx.x == 0) s_nan = 0;
    const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
    const int q = blockIdx.x * 32 + lane, b = blockIdx.y;
    const bool on = q < nq_total;
    int l = 0, qs = 0;
    for (int i = 0; i + 1 < t.nl; ++i) {  // quads are enumerated level by level; start[] is in anchors = VEC * quads
        if (q >= t.start[i + 1] / VEC) { l = i + 1; qs = t.start[i + 1] / VEC; }
    }
    const int cell = (q - qs) * VEC;
    const long long cs = t.sC[l];
    const int cpp = (nc + 3) >> 2;
    const int c_lo = part * cpp, c_hi = min(nc, c_lo + cpp);
    Top2f tt[VEC];
    bool nan_seen = false;
    __syncthreads();
    if (on) {
        const float *p = t.ptr[l] + (long long)b * t.sB[l] + cell + (64LL + c_lo) * cs;
        constexpr int CB = 10;
        int c = c_lo;
        for (; c + CB <= c_hi; c += CB, p += CB * cs) {
            float v[CB][VEC];
#pragma unroll
            for (int jj = 0; jj < CB; ++jj) {
                if constexpr (VEC == 4) {
                    const float4 r = ldg_stream4(p + jj * cs);
                    v[jj][0] = r.x; v[jj][1] = r.y; v[jj][2] = r.z; v[jj][3] = r.w;
                } else {
                    v[jj][0] = ldg_stream1(p + jj * cs);
                }
            }
#pragma unroll
            for (int jj = 0; jj < CB; ++jj)
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    top2f_push(tt[e], v[jj][e], c + jj);
                    nan_seen |= v[jj][e] != v[jj][e];
                }
        }
        for (; c < c_hi; ++c, p += cs) {
            float v[VEC];
            if constexpr (VEC == 4) {
                const float4 r = ldg_stream4(p);
                v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
            } else {
                v[0] = ldg_stream1(p);
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                top2f_push(tt[e], v[e], c);
                nan_seen |= v[e] != v[e];
            }
        }
    }
    if (nan_seen) s_nan = 1;
    if (part > 0) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            s_k1[part - 1][lane * VEC + e] = tt[e].k1;
            s_k2[part - 1][lane * VEC + e] = tt[e].k2;
            s_arg[part - 1][lane * VEC + e] = tt[e].arg;
        }
    }
    __syncthreads();
    if (part != 0 || !on) return;
    const long long o = (long long)b * A + t.start[l] + cell;
    if (s_nan) {  // some logit of this CTA's anchors is NaN: the order-preserving keys decide (NaN first, like torch.topk)
        const float *p0 = t.ptr[l] + (long long)b * t.sB[l] + cell + 64LL * cs;
#pragma unroll 1
        for (int e = 0; e < VEC; ++e) {
            Top2 tk;
            for (int c = 0; c < nc; ++c) top2_push(tk, float_key(p0[(long long)c * cs + e]), c);
            tk.k1 = float_key(dm::sigmoid_(key_float(tk.k1)));
            if (tk.k2) tk.k2 = float_key(dm::sigmoid_(key_float(tk.k2)));
            top2_store(tk, keys, aux, o + e);
        }
        return;
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
#pragma unroll
        for (int pp = 0; pp < 3; ++pp) {  // later parts hold higher classes: they win only when strictly larger
            const float ok1 = s_k1[pp][lane * VEC + e], ok2 = s_k2[pp][lane * VEC + e];
            const bool theirs = ok1 > tt[e].k1;
            const float loser = theirs ? tt[e].k1 : ok1;
            tt[e].k2 = fmaxf(fmaxf(tt[e].k2, ok2), loser);
            if (theirs) { tt[e].k1 = ok1; tt[e].arg = s_arg[pp][lane * VEC + e]; }
        }
        Top2 tk;  // only the two survivors go through the (exactly specified, weakly monotone) sigmoid
        tk.k1 = float_key(dm::sigmoid_(tt[e].k1));
        tk.k2 = nc > 1 ? float_key(dm::sigmoid_(tt[e].k2)) : 0u;
        tk.arg = tt[e].arg;
        top2_store(tk, keys, aux, o + e);
    }
}

// ------------------------------------------------------------------------------------------ block select
struct SelShared {
    unsigned hist[256];
    unsigned warp_tot[32];
    unsigned prefix, need, n_gt, n_eq, eq_base, cnt;
};

#ifdef Y3D_TIMING
__device__ long long g_sel_stamps[16];
#define SEL_STAMP(i)                                                    \
    do {                                                                \
        __syncthreads();                                                \
        if (threadIdx.x == 0 && blockIdx.x == 0) g_sel_stamps[i] = clock64(); \
    } while (0)
extern "C" int y3d_debug_read_sel_stamps(long long *host) {
    return (int)cudaMemcpyFromSymbol(host, g_sel_stamps, sizeof(long long) * 16);
}
#else
#define SEL_STAMP(i)
#endif

// Bitonic sort of buf[0..L) (L a power of two), descending.  One element per thread while L <= blockDim.x: partners less
// than a warp apart are exchanged with shuffles (no barrier), only the strides >= 32 go through shared memory.
__device__ void block_sort_desc(unsigned long long *buf, int L) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (L <= nt) {
        // only the L participating threads (whole warps) synchronise, on a named barrier: a full-block barrier per
        // step costs several hundred cycles with 32 warps
        __syncthreads();
        const int nbar = (L + 31) & ~31;
        if (tid < nbar) {
            unsigned long long v = tid < L ? buf[tid] : 0ull;
            for (int k = 2; k <= L; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    unsigned long long o;
                    if (j >= 32) {
                        asm volatile("bar.sync 1, %0;" ::"r"(nbar) : "memory");
                        buf[tid] = v;
                        asm volatile("bar.sync 1, %0;" ::"r"(nbar) : "memory");
                        o = buf[tid ^ j];
                    } else {
                        o = __shfl_xor_sync(0xffffffffu, v, j);
                    }
                    const bool desc = (tid & k) == 0, lower = (tid & j) == 0;
                    const bool take_max = desc == lower;
                    v = take_max ? (v > o ? v : o) : (v < o ? v : o);
                }
            if (L >= 64) asm volatile("bar.sync 1, %0;" ::"r"(nbar) : "memory");
            if (tid < L) buf[tid] = v;
        }
        __syncthreads();
        return;
    }
    for (int k = 2; k <= L; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < L; i += nt) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long x = buf[i], y = buf[p];
                    const bool desc = (i & k) == 0;
                    if (desc ? (x < y) : (x > y)) { buf[i] = y; buf[p] = x; }
                }
            }
            __syncthreads();
        }
}

// Exact top-D of n keys by (key desc, index asc).  key_at(i) must be pure.  `out` has room for 2 * Dpad composites
// (key << 32 | ~index); on return out[0..D) holds the winners sorted descending.
// MSD radix select, 8 bits per pass, that stops as soon as the boundary bucket is small: everything above the bucket
// plus the whole bucket (at most 2 * Dpad composites) is sorted directly, which also resolves ties in index order.
// Only when even the full 32-bit key leaves too many equal keys does the ordered tie pass run.
// whist (optional): 32 x 256 counters, one private histogram per warp -- plain shared-memory atomics without the
// cross-warp contention on a few hot digits (score keys cluster in a handful of exponents) and without match_any.
template <class KeyAt>
__device__ void block_topk(KeyAt key_at, int n, int D, int Dpad, unsigned long long *out, SelShared &sh,
                           unsigned *whist = nullptr) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int cap = 2 * Dpad;
    if (tid == 0) { sh.prefix = 0; sh.need = D; sh.n_eq = (unsigned)n; }
    unsigned mask = 0;
    __syncthreads();
    SEL_STAMP(8);
    for (int pass = 3; pass >= 0; --pass) {
        if ((D - (int)sh.need) + (int)sh.n_eq <= cap) break;  // block-uniform: read after a barrier
        const int shift = pass * 8;
        for (int i = tid; i < 256; i += nt) sh.hist[i] = 0;
        __syncthreads();
        const unsigned prefix = sh.prefix;
        if (whist) {
            for (int i = tid; i < 32 * 256; i += nt) whist[i] = 0;
            __syncthreads();
            unsigned *mine = whist + wid * 256;
            for (int i = tid; i < n; i += nt) {
                const unsigned k = key_at(i);
                if ((k & mask) == prefix) atomicAdd(&mine[(k >> shift) & 255u], 1u);
            }
            __syncthreads();
            for (int d = tid; d < 256; d += nt) {
                unsigned sum = 0;
                for (int w = 0; w < (nt >> 5); ++w) sum += whist[w * 256 + d];
                sh.hist[d] = sum;
            }
        } else {
            for (int i0 = 0; i0 < n; i0 += nt) {
                int i = i0 + tid;
                bool on = i < n;
                unsigned k = on ? key_at(i) : 0u;
                on = on && ((k & mask) == prefix);
                unsigned digit = (k >> shift) & 255u;
                // warp-aggregated shared atomics: one add per distinct digit per warp
                unsigned act = __ballot_sync(0xffffffffu, on);
                if (on) {
                    unsigned peers = __match_any_sync(act, digit);
                    if (lane == __ffs(peers) - 1) atomicAdd(&sh.hist[digit], (unsigned)__popc(peers));
                }
            }
        }
        __syncthreads();
        if (wid == 0) {  // lane l owns digits 255-8l .. 248-8l (descending)
            unsigned loc[8], s = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { loc[j] = sh.hist[255 - 8 * lane - j]; s += loc[j]; }
            unsigned inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            unsigned exc = inc - s, need = sh.need;
            if (exc < need && need <= inc) {
                unsigned cum = exc;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (cum < need && need <= cum + loc[j]) {
                        sh.prefix = prefix | ((unsigned)(255 - 8 * lane - j) << shift);
                        sh.need = need - cum;
                        sh.n_eq = loc[j];
                    }
                    cum += loc[j];
                }
            }
        }
        mask |= 255u << shift;
        __syncthreads();
    }
    const unsigned T = sh.prefix, need_eq = sh.need, n_eq = sh.n_eq, n_gt = D - need_eq;
    SEL_STAMP(9);
    if (tid == 0) { sh.cnt = 0; sh.eq_base = 0; }
    __syncthreads();
    const bool fits = (int)(n_gt + n_eq) <= cap;
    // keys above the boundary bucket, plus the whole bucket when it fits: unordered append, warp-aggregated
    for (int i0 = 0; i0 < n; i0 += nt) {
        int i = i0 + tid;
        unsigned k = i < n ? key_at(i) : 0u;
        const unsigned km = k & mask;
        bool sel = i < n && (km > T || (fits && km == T));
        unsigned bal = __ballot_sync(0xffffffffu, sel);
        if (bal) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&sh.cnt, (unsigned)__popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (sel) out[base + __popc(bal & ((1u << lane) - 1u))] = ((unsigned long long)k << 32) | (0xFFFFFFFFu - (unsigned)i);
        }
    }
    if (!fits) {  // (only reached with mask == ~0: more than Dpad exact ties) the first need_eq keys == T in index order
        for (int i0 = 0; i0 < n; i0 += nt) {
            __syncthreads();
            if (sh.eq_base >= need_eq) break;
            int i = i0 + tid;
            bool eq = i < n && key_at(i) == T;
            unsigned bal = __ballot_sync(0xffffffffu, eq);
            if (lane == 0) sh.warp_tot[wid] = __popc(bal);
            __syncthreads();
            unsigned before = 0, total = 0;
            for (int w = 0; w < (nt >> 5); ++w) {
                unsigned v = sh.warp_tot[w];
                before += (w < wid) ? v : 0u;
                total += v;
            }
            unsigned rank = sh.eq_base + before + __popc(bal & ((1u << lane) - 1u));
            if (eq && rank < need_eq) out[n_gt + rank] = ((unsigned long long)T << 32) | (0xFFFFFFFFu - (unsigned)i);
            __syncthreads();
            if (tid == 0) sh.eq_base += total;
        }
    }
    SEL_STAMP(10);
    const int filled = fits ? (int)(n_gt + n_eq) : D;
    int L = Dpad;
    while (L < filled) L <<= 1;  // <= 2 * Dpad
    for (int i = filled + tid; i < L; i += nt) out[i] = 0ull;
    __syncthreads();
    block_sort_desc(out, L);
    SEL_STAMP(11);
#ifdef Y3D_TIMING
    if (threadIdx.x == 0 && blockIdx.x == 0) { g_sel_stamps[12] = L; g_sel_stamps[13] = (long long)mask; }
#endif
}

// one CTA per image
__global__ void __launch_bounds__(kTopkThreads) topk_select_kernel(TopkSrc src, const uint32_t *__restrict__ keys,
                                                                    const int2 *__restrict__ aux, int A, int keys1_smem,
                                                                    int nc, int nreg, int D, int Dpad, int keys2_smem,
                                                                    uint32_t *__restrict__ keys2_ws, float *reg,
                                                                    float *scores, int64_t *labels,
                                                                    int32_t *anchor_idx, int out_mode, int32_t *win_anchor) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *win1 = (unsigned long long *)smem_raw;
    unsigned long long *win2 = win1 + 2 * Dpad;
    unsigned *whist = (unsigned *)(win2 + 2 * Dpad);         // 32 per-warp histograms of the radix passes
    uint32_t *k2s = (uint32_t *)(whist + 32 * 256);
    __shared__ SelShared sh;
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const uint32_t *kb = keys + (long long)b * A;
    asm volatile("griddepcontrol.launch_dependents;");  // (fused mode: box_decode_kernel is scheduled as this grid drains)
    SEL_STAMP(0);
    if (keys1_smem) {  // the radix passes re-read the keys: stage them in shared memory once (aliases the stage-2 buffer)
        for (int i = tid; i < A; i += nt) k2s[i] = kb[i];
        __syncthreads();
        block_topk([&](int i) { return k2s[i]; }, A, D, Dpad, win1, sh, whist);
    } else {
        block_topk([&](int i) { return kb[i]; }, A, D, Dpad, win1, sh, whist);
    }
    SEL_STAMP(1);
    // stage 2: D x nc candidate scores, flattened index j = i*nc + c   (ops.py:858-861).
    // Every selected anchor contributes its own maximum, so at least D candidates are >= tau, the smallest selected
    // maximum: the final top-D lives among the candidates with key >= tau.  Those are compacted (typically little more
    // than D of the D*nc) and sorted directly; the full radix select over D*nc keys is only the overflow fallback.
    const int n2 = D * nc;
    uint32_t *k2 = keys2_smem ? k2s : keys2_ws + (long long)b * n2;
    unsigned long long *cand = reinterpret_cast<unsigned long long *>(k2);
    const int cap = (((uintptr_t)cand & 7) == 0) ? min(n2 / 2, 8192) : 0;  // u64 slots available in the key buffer
    const uint32_t tau = (uint32_t)(win1[D - 1] >> 32);
    __shared__ unsigned n_cand;
    if (tid == 0) n_cand = 0;
    __syncthreads();
    // every selected anchor contributes its best class; only anchors whose SECOND best key also reaches tau can
    // contribute more, and only their nc scores are gathered
    __shared__ unsigned n_rich;
    int *rich = reinterpret_cast<int *>(win2);  // ranks of those anchors (win2 is free until the final copy)
    if (tid == 0) n_rich = 0;
    __syncthreads();
    const int2 *ab = aux + (long long)b * A;
    for (int i = tid; i < D; i += nt) {
        const uint32_t k1 = (uint32_t)(win1[i] >> 32);
        const int a = (int)(0xFFFFFFFFu - (unsigned)(win1[i] & 0xFFFFFFFFull));
        const int2 ax = ab[a];
        if ((uint32_t)ax.x >= tau) {
            rich[atomicAdd(&n_rich, 1u)] = i;
        } else {
            const unsigned pos = atomicAdd(&n_cand, 1u);
            if (pos < (unsigned)cap) cand[pos] = ((unsigned long long)k1 << 32) | (0xFFFFFFFFu - (unsigned)(i * nc + ax.y));
        }
    }
    __syncthreads();
    const int n2r = (int)n_rich * nc;
    constexpr int GU = 8;  // gathers in flight per thread: latency-bound sector gathers
    for (int j0 = 0; j0 < n2r; j0 += nt * GU) {
        uint32_t key[GU];
        int jj[GU];
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            const int t = j0 + u * nt + tid;
            key[u] = 0;
            jj[u] = -1;
            if (t < n2r) {
                const int ri = t / nc, c = t - ri * nc;
                const int i = rich[ri];
                const int a = (int)(0xFFFFFFFFu - (unsigned)(win1[i] & 0xFFFFFFFFull));
                key[u] = float_key(src_score(src, b, a, c));
                jj[u] = i * nc + c;
            }
        }
#pragma unroll
        for (int u = 0; u < GU; ++u) {
            const bool keep = jj[u] >= 0 && key[u] >= tau;
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (bal) {
                unsigned base = 0;
                if ((tid & 31) == 0) base = atomicAdd(&n_cand, (unsigned)__popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                const unsigned pos = base + __popc(bal & ((1u << (tid & 31)) - 1u));
                if (keep && pos < (unsigned)cap) cand[pos] = ((unsigned long long)key[u] << 32) | (0xFFFFFFFFu - (unsigned)jj[u]);
            }
        }
    }
    __syncthreads();
    SEL_STAMP(2);
    const int nc2 = (int)n_cand;
    int L = Dpad;
    while (L < nc2 && L < (1 << 20)) L <<= 1;
    if (nc2 <= cap && L <= cap) {
        for (int i = nc2 + tid; i < L; i += nt) cand[i] = 0ull;
        __syncthreads();
        block_sort_desc(cand, L);
        for (int r = tid; r < D; r += nt) win2[r] = cand[r];
        __syncthreads();
    } else {  // overflow (massive ties at tau): exact radix select over all D*nc keys
        __syncthreads();
        for (int j = tid; j < n2; j += nt) {
            int i = j / nc, c = j - i * nc;
            int a = (int)(0xFFFFFFFFu - (unsigned)(win1[i] & 0xFFFFFFFFull));
            k2[j] = float_key(src_score(src, b, a, c));
        }
        __syncthreads();
        block_topk([&](int j) { return k2[j]; }, n2, D, Dpad, win2, sh, whist);
    }
    SEL_STAMP(3);
    for (int r = tid; r < D; r += nt) {
        unsigned long long w = win2[r];
        int j = (int)(0xFFFFFFFFu - (unsigned)(w & 0xFFFFFFFFull));
        int i = j / nc, c = j - i * nc;
        int a = (int)(0xFFFFFFFFu - (unsigned)(win1[i] & 0xFFFFFFFFull));
        float sc = key_float((uint32_t)(w >> 32));
        long long o = (long long)b * D + r;
        if (anchor_idx) anchor_idx[o] = a;
        if (win_anchor) win_anchor[o] = a;
        if (out_mode == 0) {
            scores[o] = sc;
            labels[o] = c;
        } else {  // fused export layout [B,D,6] = box, score, label (head.py:531)
            float *q = reg + o * 6;
            q[4] = sc; q[5] = (float)c;
        }
    }
    if (out_mode == 0) {
        for (int e = tid; e < D * nreg; e += nt) {
            int r = e / nreg, rr = e - r * nreg;
            unsigned long long w = win2[r];
            int j = (int)(0xFFFFFFFFu - (unsigned)(w & 0xFFFFFFFFull));
            int a = (int)(0xFFFFFFFFu - (unsigned)(win1[j / nc] & 0xFFFFFFFFull));
            reg[((long long)b * D + r) * nreg + rr] =
                src.preds[b * src.sB + a * src.sA + (long long)(src.roff + rr) * src.sC];
        }
    }
    SEL_STAMP(4);
}

// Boxes of the winners of the fused decode + top-k, spread over the whole machine (the selection kernel is one CTA per
// image: the 64 bin gathers per detection -- random 32-byte sectors of the head tensor -- would all queue behind that one
// SM's load pipe; 25 us at 32 images x 1280^2, against ~5 us here).  grid (ceil(4 D / 256), B), thread = (detection,
// side): 16 gathers in flight, softmax expectation, the four lanes of a detection meet through shuffles; arithmetic of
// decode2d_kernel.  Launched programmatically dependent on the selection kernel, which left (score, label) in the
// row and the anchor in win_anchor.  Image-sharded detection path (G.world > 1): the CTA's 64 finished rows also go straight
// into every peer's staging area (coalesced NVLink stores, sequence number inside every word: see GatherOut).
__global__ void __launch_bounds__(256) box_decode_kernel(const __grid_constant__ TopkSrc src, const int32_t *__restrict__ win_anchor,
                                                         int D, float *reg, const __grid_constant__ GatherOut G) {
    __shared__ float2 s_rows[3 * 64];
    const int b = blockIdx.y, e = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = e >> 2, side = e & 3;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    float d = 0.f;
    int a = 0, l = 0;
    if (r < D) {
        a = __ldcg(win_anchor + (long long)b * D + r);
        l = level_of(src.t, a);
        const float *p = src.t.ptr[l] + (long long)b * src.t.sB[l] + (a - src.t.start[l]) + (long long)(side * 16) * src.t.sC[l];
        const long long cs = src.t.sC[l];
        float x[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) x[jj] = p[jj * cs];
        d = im::dfl16(x);
    }
    const int base = (threadIdx.x & 31) & ~3;
    const float d0 = __shfl_sync(0xffffffffu, d, base), d1 = __shfl_sync(0xffffffffu, d, base + 1),
                d2 = __shfl_sync(0xffffffffu, d, base + 2), d3 = __shfl_sync(0xffffffffu, d, base + 3);
    if (r < D && side == 0) {
        const int cell = a - src.t.start[l], wd = src.t.w[l];
        float *q = reg + ((long long)b * D + r) * 6;
        float bx[4];
        im::box_axis((float)(cell % wd) + 0.5f, d0, d2, src.t.stride[l], src.xywh, bx[0], bx[2]);
        im::box_axis((float)(cell / wd) + 0.5f, d1, d3, src.t.stride[l], src.xywh, bx[1], bx[3]);
        const float2 r0 = make_float2(bx[0], bx[1]), r1 = make_float2(bx[2], bx[3]);
        reinterpret_cast<float2 *>(q)[0] = r0;
        reinterpret_cast<float2 *>(q)[1] = r1;
        if (G.world > 1) {  // the CTA's 64 rows are contiguous in the output: staged, then stored to the peers coalesced
            const int lr = threadIdx.x >> 2;
            s_rows[3 * lr] = r0;
            s_rows[3 * lr + 1] = r1;
            s_rows[3 * lr + 2] = __ldcg(reinterpret_cast<const float2 *>(q) + 2);  // score, label: the selection kernel's
        }
    }
    if (G.world > 1) {  // the staged rows -> every peer's staging area, value and sequence number in one 64-bit word
        __syncthreads();
        const int r_first = (blockIdx.x * blockDim.x) >> 2;
        const int nw = 6 * max(0, min(D - r_first, (int)blockDim.x >> 2));  // floats of this CTA's rows
        const long long w0 = ((long long)(G.rank * G.n_local + b) * D + r_first) * 6;
        const float *sf = reinterpret_cast<const float *>(s_rows);
        for (int pr = 0; pr < G.world; ++pr) {
            if (pr == G.rank) continue;
            unsigned long long *dst = G.stage[pr] + w0;
            for (int i = threadIdx.x; i < nw; i += blockDim.x)
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + i),
                             "l"(((unsigned long long)__float_as_uint(sf[i]) << 32) | (unsigned long long)G.seq) : "memory");
        }
    }
}

// ------------------------------------------------------------------------------------------ sparse 3D head glue
// v10Detect3d.select_candidates (head.py:681-687): per image, top-K cells of one level by max_c(raw class logit), as
// (row, col) pairs in descending score order (lowest flat index first among equals).  One CTA per image; the per-cell
// maxima are staged in shared memory (dynamic: H*W uint32 keys + 2 * 2 * Kpad composites).
__global__ void __launch_bounds__(kTopkThreads) select_candidates_kernel(const float *__restrict__ cls, long long sB,
                                                                          long long sC, int nc, int HW, int Wd, int K,
                                                                          int Kpad, int64_t *__restrict__ idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *win = (unsigned long long *)smem_raw;
    uint32_t *keys = (uint32_t *)(win + 2 * Kpad);
    __shared__ SelShared sh;
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const float *p = cls + (long long)b * sB;
    for (int i = tid; i < HW; i += nt) {
        uint32_t k = 0;
        for (int c = 0; c < nc; ++c) k = max(k, float_key(p[(long long)c * sC + i]));  // torch.max(scores, dim=1)[0]
        keys[i] = k;
    }
    __syncthreads();
    block_topk([&](int i) { return keys[i]; }, HW, K, Kpad, win, sh);
    for (int r = tid; r < K; r += nt) {
        const int i = (int)(0xFFFFFFFFu - (unsigned)(win[r] & 0xFFFFFFFFull));
        idx[((long long)b * K + r) * 2 + 0] = i / Wd;  // unravel_index (head.py:652-657): (row, col)
        idx[((long long)b * K + r) * 2 + 1] = i % Wd;
    }
}

// v10Detect3d.extract_patches (head.py:659-679): zero-padded P x P patch of every channel around each candidate.
// out [B*K, C, P, P]; one thread per output element
__global__ void __launch_bounds__(256) extract_patches_kernel(const float *__restrict__ x, long long sB, long long sC,
                                                              const int64_t *__restrict__ idx, int C, int H, int Wd,
                                                              int K, int P, long long total, float *__restrict__ out) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= total) return;
    const int dx = (int)(o % P);
    long long r = o / P;
    const int dy = (int)(r % P);
    r /= P;
    const int c = (int)(r % C);
    const long long bk = r / C;
    const int b = (int)(bk / K);
    const int pad = P / 2;
    const int row = (int)idx[bk * 2] + dy - pad, col = (int)idx[bk * 2 + 1] + dx - pad;
    float v = 0.f;
    if (row >= 0 && row < H && col >= 0 && col < Wd) v = x[b * sB + c * sC + (long long)row * Wd + col];
    out[o] = v;
}

// head.py:709-714: head_output[b, :, row_k, col_k] = conv_out[b*K + k, :] into a zero-filled [B, Cout, H, W]
__global__ void __launch_bounds__(256) scatter_candidates_kernel(const float *__restrict__ vals,
                                                                 const int64_t *__restrict__ idx, int Cout, int H,
                                                                 int Wd, int K, long long total, float *__restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int c = (int)(t % Cout);
    const long long bk = t / Cout;
    const int b = (int)(bk / K);
    const long long row = idx[bk * 2], col = idx[bk * 2 + 1];
    if (row < 0 || row >= H || col < 0 || col >= Wd) return;
    out[(((long long)b * Cout + c) * H + row) * Wd + col] = vals[t];
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

size_t topk_workspace_bytes(int B, int A, int nc, int D) {
    size_t keys = align256(sizeof(uint32_t) * (size_t)B * A) + align256(sizeof(int2) * (size_t)B * A);
    size_t k2 = (D * nc > kKeys2SmemCap) ? align256(sizeof(uint32_t) * (size_t)B * D * nc) : 0;
    // fused decode + top-k: anchors of the winners [B, D] (read by box_decode_kernel)
    return keys + k2 + align256(sizeof(int32_t) * (size_t)B * D);
}

static int launch_select(const TopkSrc &src, const uint32_t *keys, const int2 *aux, int B, int A, int nc, int nreg, int D, float *reg,
                         float *scores, int64_t *labels, int32_t *anchor_idx, int out_mode, uint32_t *keys2_ws,
                         cudaStream_t s, const GatherOut *gather = nullptr, int32_t *win_anchor = nullptr) {
    int Dpad = next_pow2(D);
    int n2 = D * nc;
    int k2smem = n2 <= kKeys2SmemCap;
    int k1smem = A <= kKeys2SmemCap;
    size_t nkeys = 0;
    if (k2smem) nkeys = (size_t)n2;
    if (k1smem && (size_t)A > nkeys) nkeys = (size_t)A;
    size_t smem = sizeof(unsigned long long) * 4 * (size_t)Dpad + sizeof(unsigned) * 32 * 256 + sizeof(uint32_t) * nkeys;
    static size_t smem_limit = 48 * 1024;  // raised once per process to the largest size asked for
    if (smem > smem_limit) {
        cudaError_t e = cudaFuncSetAttribute(topk_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        smem_limit = smem;
    }
    topk_select_kernel<<<B, kTopkThreads, smem, s>>>(src, keys, aux, A, k1smem, nc, nreg, D, Dpad, k2smem, keys2_ws, reg, scores,
                                                    labels, anchor_idx, out_mode, win_anchor);
    Y3D_CHECK_LAUNCH();
    if (out_mode != 0) {  // boxes of the winners (+ the peer copies of the image-sharded path), over all SMs
        GatherOut G{};
        if (gather) G = *gather;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((4 * D + 255) / 256), (unsigned)B);
        cfg.blockDim = dim3(256);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t le = cudaLaunchKernelEx(&cfg, box_decode_kernel, src, (const int32_t *)win_anchor, D, reg, G);
        if (le != cudaSuccess) return (int)le;
        Y3D_CHECK_LAUNCH();
    }
    return Y3D_OK;
}

}  // namespace y3d

using namespace y3d;

extern "C" int y3d_postprocess(const float *preds, int64_t sB, int64_t sA, int64_t sC, int B, int A, int nc, int nreg,
                               int scores_first, int D, float *reg, float *scores, int64_t *labels,
                               int32_t *anchor_idx, void *ws, size_t ws_bytes, void *stream) {
    if (!preds || !reg || !scores || !labels || B < 0 || A < 1 || nc < 1 || nreg < 0) return Y3D_EINVAL;
    if (D < 1 || D > A) return Y3D_EINVAL;  // torch.topk raises when k > A (ops.py:856)
    if (D > Y3D_MAX_DET) return Y3D_EUNSUPPORTED;
    size_t need = topk_workspace_bytes(B, A, nc, D);
    if (!ws || ws_bytes < need) return Y3D_EWORKSPACE;
    if (((uintptr_t)ws) % 256) return Y3D_EALIGN;
    if (B == 0) return Y3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t *keys = (uint32_t *)ws;
    int2 *aux = (int2 *)((char *)ws + align256(sizeof(uint32_t) * (size_t)B * A));
    uint32_t *k2 = (uint32_t *)((char *)aux + align256(sizeof(int2) * (size_t)B * A));
    const int soff = scores_first ? 0 : nreg, roff = scores_first ? nc : 0;
    if (sA == 1) {
        dim3 grid((A + 255) / 256, B);
        amax_anchor_kernel<<<grid, 256, 0, s>>>(preds, sB, sC, soff, B, A, nc, keys, aux);
    } else {
        long long threads = (long long)B * A * 32;
        amax_warp_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(preds, sB, sA, sC, soff, B, A, nc, keys, aux);
    }
    Y3D_CHECK_LAUNCH();
    TopkSrc src{};
    src.mode = 0;
    src.preds = preds;
    src.sB = sB; src.sA = sA; src.sC = sC;
    src.soff = soff; src.roff = roff;
    return launch_select(src, keys, aux, B, A, nc, nreg, D, reg, scores, labels, anchor_idx, 0, k2, s);
}

static int decode_topk2d_run(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC, const int *lvl_hw,
                            const float *lvl_stride, int nl, int B, int nc, int reg_max, int xywh, int D, float *out,
                            int32_t *anchor_idx, void *ws, size_t ws_bytes, cudaStream_t s, const GatherOut *gather) {
    if (!lvl_ptr || !lvl_sB || !lvl_sC || !out || B < 0 || nc < 1) return Y3D_EINVAL;
    if (reg_max != 16) return Y3D_EUNSUPPORTED;
    TopkSrc src{};
    int A = make_level_table(src.t, lvl_ptr, lvl_sB, lvl_sC, lvl_hw, lvl_stride, nl);
    if (A < 0) return A;
    for (int l = 0; l < nl; ++l)
        if (!lvl_ptr[l]) return Y3D_EINVAL;
    if (D < 1 || D > A) return Y3D_EINVAL;
    if (D > Y3D_MAX_DET) return Y3D_EUNSUPPORTED;
    size_t need = topk_workspace_bytes(B, A, nc, D);
    if (!ws || ws_bytes < need) return Y3D_EWORKSPACE;
    if (((uintptr_t)ws) % 256) return Y3D_EALIGN;
    if (B == 0) return Y3D_OK;
    uint32_t *keys = (uint32_t *)ws;
    int2 *aux = (int2 *)((char *)ws + align256(sizeof(uint32_t) * (size_t)B * A));
    uint32_t *k2 = (uint32_t *)((char *)aux + align256(sizeof(int2) * (size_t)B * A));
    int32_t *win_anchor = (int32_t *)((char *)k2 + ((D * nc > kKeys2SmemCap) ? align256(sizeof(uint32_t) * (size_t)B * D * nc) : 0));
    bool v4 = true;
    for (int l = 0; l < nl; ++l)
        v4 = v4 && (src.t.h[l] * src.t.w[l]) % 4 == 0 && ((uintptr_t)lvl_ptr[l]) % 16 == 0 && lvl_sB[l] % 4 == 0 &&
             lvl_sC[l] % 4 == 0;
    if (v4) {
        int nq = A / 4;
        dim3 grid((nq + 31) / 32, B);
        cls_max_kernel<4><<<grid, 128, 0, s>>>(src.t, nq, nc, A, keys, aux);
    } else {
        dim3 grid((A + 31) / 32, B);
        cls_max_kernel<1><<<grid, 128, 0, s>>>(src.t, A, nc, A, keys, aux);
    }
    Y3D_CHECK_LAUNCH();
    src.mode = 1;
    src.xywh = xywh;
    return launch_select(src, keys, aux, B, A, nc, 4, D, out, nullptr, nullptr, anchor_idx, 1, k2, s, gather, win_anchor);
}

extern "C" int y3d_decode_topk2d(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC,
                                 const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                                 int xywh, int D, float *out, int32_t *anchor_idx, void *ws, size_t ws_bytes,
                                 void *stream) {
    return decode_topk2d_run(lvl_ptr, lvl_sB, lvl_sC, lvl_hw, lvl_stride, nl, B, nc, reg_max, xywh, D, out, anchor_idx, ws,
                             ws_bytes, (cudaStream_t)stream, nullptr);
}

static size_t gather_data_floats(int world, int n_local, int D) { return (size_t)world * n_local * D * 6; }
static size_t gather_stage_off(int world, int n_local, int D) { return align256(sizeof(float) * 2 * gather_data_floats(world, n_local, D)); }

extern "C" size_t y3d_gather_buffer_bytes(int world, int n_local, int D) {
    if (world < 1 || n_local < 1 || D < 1) return 0;
    return gather_stage_off(world, n_local, D) + align256(sizeof(unsigned long long) * 2 * gather_data_floats(world, n_local, D));
}

extern "C" int y3d_decode_topk2d_sharded(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC,
                                         const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                                         int xywh, int D, int rank, int world, void *const *peer_bufs,
                                         unsigned long long seq, int *status, void *ws, size_t ws_bytes, void *stream) {
    if (!peer_bufs || world < 1 || world > kGatherMaxWorld || rank < 0 || rank >= world || seq == 0 || B < 1) return Y3D_EINVAL;
    GatherOut G{};
    const int par = (int)(seq & 1ull);
    const size_t nfl = gather_data_floats(world, B, D), soff = gather_stage_off(world, B, D);
    for (int r = 0; r < world; ++r) {
        if (!peer_bufs[r] || ((uintptr_t)peer_bufs[r]) % 256) return Y3D_EALIGN;
        G.data[r] = (float *)peer_bufs[r] + (size_t)par * nfl;
        G.stage[r] = (unsigned long long *)((char *)peer_bufs[r] + soff) + (size_t)par * nfl;
    }
    G.rank = rank; G.world = world; G.n_local = B;
    G.seq = (unsigned)(seq & 0xffffffffull);
    if (G.seq == 0) return Y3D_EINVAL;  // (the zero-filled staging area reads as sequence number 0)
    cudaStream_t s = (cudaStream_t)stream;
    float *own = G.data[rank] + (size_t)rank * B * D * 6;  // this rank's images inside its own buffer
    int rc = decode_topk2d_run(lvl_ptr, lvl_sB, lvl_sC, lvl_hw, lvl_stride, nl, B, nc, reg_max, xywh, D, own, nullptr, ws,
                               ws_bytes, s, &G);
    if (rc) return rc;
    if (world > 1) {
        static long long timeout = 0;
        if (timeout == 0) {
            double sec = 600.0;
            if (const char *e = getenv("Y3D_XRANK_TIMEOUT_S")) { const double v = atof(e); if (v > 0.0) sec = v; }
            timeout = (long long)(sec * 2.0e9);
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((world - 1) * B));
        cfg.blockDim = dim3(256);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t le = cudaLaunchKernelEx(&cfg, gather_collect_kernel, G, D, timeout, status);
        if (le != cudaSuccess) return (int)le;
        Y3D_CHECK_LAUNCH();
    }
    return Y3D_OK;
}

extern "C" int y3d_select_candidates(const float *cls, int64_t sB, int64_t sC, int B, int nc, int H, int W, int K,
                                     int64_t *idx, void *stream) {
    if (!cls || !idx || B < 0 || nc < 1 || H < 1 || W < 1) return Y3D_EINVAL;
    const long long HW = (long long)H * W;
    if (K < 1 || K > HW) return Y3D_EINVAL;  // torch.topk raises when k exceeds the number of cells
    if (K > Y3D_MAX_DET) return Y3D_EUNSUPPORTED;
    if (B == 0) return Y3D_OK;
    const int Kpad = next_pow2(K);
    const size_t smem = sizeof(unsigned long long) * 2 * (size_t)Kpad + sizeof(uint32_t) * (size_t)HW;
    if (smem > 200 * 1024) return Y3D_EUNSUPPORTED;  // level larger than ~48 k cells
    static size_t smem_limit = 48 * 1024;  // raised once per process to the largest size asked for
    if (smem > smem_limit) {
        cudaError_t e = cudaFuncSetAttribute(select_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        smem_limit = smem;
    }
    select_candidates_kernel<<<B, kTopkThreads, smem, (cudaStream_t)stream>>>(cls, sB, sC, nc, (int)HW, W, K, Kpad, idx);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

extern "C" int y3d_extract_patches(const float *x, int64_t sB, int64_t sC, const int64_t *idx, int B, int C, int H, int W,
                                   int K, int P, float *out, void *stream) {
    if (!x || !idx || !out || B < 0 || C < 1 || H < 1 || W < 1 || K < 0 || P < 1 || !(P & 1)) return Y3D_EINVAL;
    const long long total = (long long)B * K * C * P * P;
    if (total == 0) return Y3D_OK;
    extract_patches_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, sB, sC, idx, C, H, W, K, P,
                                                                                             total, out);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

extern "C" int y3d_scatter_candidates(const float *vals, const int64_t *idx, int B, int Cout, int H, int W, int K,
                                      float *out, void *stream) {
    if (!out || B < 0 || Cout < 1 || H < 1 || W < 1 || K < 0) return Y3D_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t bytes = sizeof(float) * (size_t)B * Cout * H * W;
    if (bytes == 0) return Y3D_OK;
    cudaError_t e = cudaMemsetAsync(out, 0, bytes, s);  // torch.zeros(output_shape) head.py:709
    if (e != cudaSuccess) return (int)e;
    const long long total = (long long)B * K * Cout;
    if (total == 0) return Y3D_OK;
    if (!vals || !idx) return Y3D_EINVAL;
    scatter_candidates_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(vals, idx, Cout, H, W, K, total, out);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}
