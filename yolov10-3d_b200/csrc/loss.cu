// loss.cu -- fused v8DetectionLoss / v10DetectLoss forward (sm_100a).
//   y3d_v10_loss_fwd     v10DetectLoss.__call__     reference ultralytics/utils/loss.py:727-737
//   y3d_v8_loss_fwd      v8DetectionLoss.__call__   reference ultralytics/utils/loss.py:206-257
//                        bbox_decode                reference ultralytics/utils/loss.py:197-204
//                        BboxLoss.forward/_df_loss  reference ultralytics/utils/loss.py:82-113
//                        bbox2dist                  reference ultralytics/utils/tal.py:328-331
//   y3d_train_decode     bbox_decode + permute/sigmoid (loss.py:214-215,232)
//   y3d_v8_loss_finalize normalisation by target_scores_sum (loss.py:240-256)
//
// Both branches of the dual assignment (one2many top-k 10, one2one top-k 1) run in the SAME launches
// (blockIdx.z / tile index selects the branch); a step is one memset + four kernels:
//   1. loss_stream_tma_kernel : the ONE pass over the two head tensors (4*(4R+nc)*A bytes per image and branch).
//        Persistent CTAs (one per SM); a producer warp streams [C x 128-anchor] tiles into a 3-stage shared-memory
//        ring with TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) and eight consumer warps turn each tile
//        into DFL softmax-integral boxes (xyxy, grid units, 16 B/anchor: the only dense write) and
//        sum softplus(logit) = sum BCE(logit, 0).  pd_scores and the dense target_scores of the reference are never
//        materialised: sum BCE(x,t) = sum BCE(x,0) - sum_fg x[label]*t.
//   2. tal_topk_kernel / 3. tal_resolve_kernel : assign.cuh; scores are read as logits straight from the head.
//   4. loss_fg_kernel : warp-cooperative CIoU / DFL / BCE-correction terms of the foreground anchors; the last CTA
//        to finish reduces all per-CTA partials in a fixed order (deterministic) and writes the loss items.
#include "assign.cuh"

namespace y3d {

constexpr int kR = 16;
constexpr int kTileA = 128;          // anchors per tile
constexpr int kStages = 3;
constexpr int kConsumerWarps = 8;
constexpr int kStreamThreads = (kConsumerWarps + 1) * 32;
constexpr float kLog2e = 1.4426950408889634f;

// ---------------------------------------------------------------------------------------------- shared arithmetic
// Explicitly rounded (no contraction freedom) so that every kernel using these produces identical bits.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// softmax over the 16 DFL bins followed by the expectation sum_j j*p_j (bbox_decode, loss.py:199-201)
__device__ __forceinline__ float dfl_expect16(const float (&x)[kR]) {
    float m = x[0];
#pragma unroll
    for (int j = 1; j < kR; ++j) m = fmaxf(m, x[j]);
    float s = 0.f, acc = 0.f;
#pragma unroll
    for (int j = 0; j < kR; ++j) {
        float e = ex2_approx(__fmul_rn(__fsub_rn(x[j], m), kLog2e));
        s = __fadd_rn(s, e);
        acc = __fmaf_rn((float)j, e, acc);
    }
    return __fdiv_rn(acc, s);
}
// BCEWithLogits(x, 0) = max(x,0) + log1p(exp(-|x|)); log1p(t) = 2 atanh(t/(2+t)) as an odd series in s = t/(2+t)
// (s <= 1/3): relative error ~1e-6 over the whole range, no cancellation for very negative logits.
__device__ __forceinline__ float softplus_fast(float x) {
    float t = ex2_approx(__fmul_rn(-fabsf(x), kLog2e));
    float s = __fdividef(t, __fadd_rn(2.0f, t));
    float s2 = __fmul_rn(s, s);
    float p = __fmaf_rn(s2, 0.1111111111f, 0.1428571429f);
    p = __fmaf_rn(p, s2, 0.2f);
    p = __fmaf_rn(p, s2, 0.3333333333f);
    p = __fmaf_rn(p, s2, 1.0f);
    return __fmaf_rn(__fadd_rn(s, s), p, fmaxf(x, 0.f));
}
__device__ __forceinline__ float sigmoid_acc(float v) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v))); }

// ---------------------------------------------------------------------------------------------- mbarrier / TMA PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------- stream kernel
struct StreamParams {
    LevelTable t[2];
    float *pd_bboxes[2];   // [B,A,4] xyxy grid units
    float *pd_scores[2];   // optional [B,A,nc] sigmoid (y3d_train_decode only)
    double *part_bce;      // [n_branch][gridDim.x] or nullptr
    int tile0[Y3D_MAX_LEVELS + 1];  // first tile index of each level inside one image
    int n_branch, B, nc, A;
};

// persistent: CTA x handles tiles x, x + gridDim.x, ...; tile id = ((branch * B + b) * tiles_per_image + ti)
__global__ void __launch_bounds__(kStreamThreads, 1) loss_stream_tma_kernel(StreamParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ __align__(8) uint64_t empty_bar[kStages];
    __shared__ double red[2][kConsumerWarps];
    const int C = 4 * kR + P.nc;
    const int stage_floats = C * kTileA;
    float *ring = reinterpret_cast<float *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tpi = P.tile0[P.t[0].nl];
    const long long n_tiles = (long long)P.n_branch * P.B * tpi;

    auto locate = [&](long long tile, int &br, int &b, int &l, int &cell0, int &nv) {
        int ti = (int)(tile % tpi);
        long long r = tile / tpi;
        b = (int)(r % P.B);
        br = (int)(r / P.B);
        l = 0;
#pragma unroll
        for (int i = 1; i < Y3D_MAX_LEVELS; ++i) l += (i < P.t[0].nl && ti >= P.tile0[i]) ? 1 : 0;
        cell0 = (ti - P.tile0[l]) * kTileA;
        nv = min(kTileA, P.t[0].h[l] * P.t[0].w[l] - cell0);
    };

    if (wid == kConsumerWarps) {
        // ---------------- producer warp: TMA bulk copies, one 512-byte row segment per (channel, tile)
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int s = it % kStages, round = it / kStages;
            int br, b, l, cell0, nv;
            locate(tile, br, b, l, cell0, nv);
            mbar_wait(&empty_bar[s], (round & 1) ^ 1);
            if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], (unsigned)(C * nv * 4));
            __syncwarp();
            const LevelTable &t = P.t[br];
            const float *src = t.ptr[l] + (long long)b * t.sB[l] + cell0;
            float *dst = ring + (long long)s * stage_floats;
            for (int r = lane; r < C; r += 32)
                tma_bulk_g2s(dst + r * kTileA, src + (long long)r * t.sC[l], (unsigned)(nv * 4), &full_bar[s]);
        }
    } else {
        // ---------------- consumer warps
        constexpr int G = kConsumerWarps * 32 / kTileA;  // threads per anchor (2)
        const int a = tid % kTileA, h = tid / kTileA;
        const int cls_chunk = (P.nc + G - 1) / G;
        double dacc[2] = {0.0, 0.0};
        int it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int s = it % kStages, round = it / kStages;
            int br, b, l, cell0, nv;
            locate(tile, br, b, l, cell0, nv);
            mbar_wait(&full_bar[s], round & 1);
            const float *sm = ring + (long long)s * stage_floats + a;
            if (a < nv) {
                const LevelTable &t = P.t[br];
                const int cell = cell0 + a;
                const long long ga = (long long)b * P.A + t.start[l] + cell;
                const int w = t.w[l];
                float *ob = P.pd_bboxes[br] + ga * 4;
#pragma unroll
                for (int side = h; side < 4; side += G) {
                    float x[kR];
#pragma unroll
                    for (int j = 0; j < kR; ++j) x[j] = sm[(side * kR + j) * kTileA];
                    const float d = dfl_expect16(x);
                    const float anc = ((side & 1) ? (float)(cell / w) : (float)(cell % w)) + 0.5f;
                    ob[side] = side < 2 ? __fsub_rn(anc, d) : __fadd_rn(anc, d);  // dist2bbox xyxy tal.py:319-325
                }
                const int c0 = h * cls_chunk, c1 = min(P.nc, c0 + cls_chunk);
                float acc = 0.f;
                float *ps = P.pd_scores[br];
                for (int c = c0; c < c1; ++c) {
                    const float v = sm[(4 * kR + c) * kTileA];
                    acc += softplus_fast(v);
                    if (ps) ps[ga * P.nc + c] = sigmoid_acc(v);
                }
                dacc[br] += (double)acc;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
        dacc[0] = warp_sum(dacc[0]);
        dacc[1] = warp_sum(dacc[1]);
        if (lane == 0) { red[0][wid] = dacc[0]; red[1][wid] = dacc[1]; }
    }
    __syncthreads();
    if (tid < P.n_branch && P.part_bce) {  // fixed-order block reduction -> one partial per CTA and branch
        double s = 0.0;
        for (int i = 0; i < kConsumerWarps; ++i) s += red[tid][i];
        P.part_bce[(long long)tid * gridDim.x + blockIdx.x] = s;
    }
}

// fallback for shapes TMA cannot take (level sizes not a multiple of 4, unaligned pointers): one thread per anchor
__global__ void __launch_bounds__(128) loss_stream_simple_kernel(StreamParams P) {
    __shared__ double red[4];
    const int br = blockIdx.z, b = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    double local = 0.0;
    if (a < P.A) {
        const LevelTable &t = P.t[br];
        const int l = level_of(t, a);
        const int cell = a - t.start[l];
        const float *base = t.ptr[l] + (long long)b * t.sB[l] + cell;
        const long long cs = t.sC[l];
        const long long ga = (long long)b * P.A + a;
        const int w = t.w[l];
        float *ob = P.pd_bboxes[br] + ga * 4;
        for (int side = 0; side < 4; ++side) {
            float x[kR];
#pragma unroll
            for (int j = 0; j < kR; ++j) x[j] = base[(long long)(side * kR + j) * cs];
            const float d = dfl_expect16(x);
            const float anc = ((side & 1) ? (float)(cell / w) : (float)(cell % w)) + 0.5f;
            ob[side] = side < 2 ? __fsub_rn(anc, d) : __fadd_rn(anc, d);
        }
        float acc = 0.f;
        float *ps = P.pd_scores[br];
        for (int c = 0; c < P.nc; ++c) {
            const float v = base[(long long)(4 * kR + c) * cs];
            acc += softplus_fast(v);
            if (ps) ps[ga * P.nc + c] = sigmoid_acc(v);
        }
        local = (double)acc;
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0 && P.part_bce) {
        const long long nb = (long long)gridDim.x * gridDim.y;
        P.part_bce[br * nb + (long long)b * gridDim.x + blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
    }
}

// ---------------------------------------------------------------------------------------------- foreground terms
struct FgParams {
    const float *gt5;          // [B,M,5]
    const double *part_bce;    // [n_branch][n_bce]
    double *part_fg;           // [n_branch][4][n_fg]
    unsigned *counter;         // zero-initialised ticket
    double *partials;          // optional out [n_branch][4]
    float *loss_items;         // optional out [n_branch][4]
    uint8_t *dbg_fg[2];
    int32_t *dbg_gi[2];
    int n_bce, n_fg, n_branch, normalise;
    float gain_box, gain_cls, gain_dfl;
};

// grid (ceil(A/256), B, n_branch).  Each warp owns 32 anchors and walks its foreground ones cooperatively.
__global__ void __launch_bounds__(256) loss_fg_kernel(AssignCtx2 cc, FgParams F) {
    __shared__ double red[5][256];
    __shared__ unsigned s_ticket;
    const AssignCtx &c = cc.c[blockIdx.z];
    const int b = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const long long o = (long long)b * c.A + a;
    const int gi = (a < c.A && c.M > 0) ? c.tgi[o] : -1;
    const float alv = gi >= 0 ? c.alignv[o] : 0.f;
    if (a < c.A) {
        if (F.dbg_fg[blockIdx.z]) F.dbg_fg[blockIdx.z][o] = (uint8_t)(gi >= 0);
        if (F.dbg_gi[blockIdx.z]) F.dbg_gi[blockIdx.z][o] = gi >= 0 ? gi : 0;
    }
    double s_iou = 0.0, s_dfl = 0.0, s_ts = 0.0, s_xt = 0.0;
    unsigned bal = __ballot_sync(0xffffffffu, gi >= 0);
    while (bal) {
        const int j = __ffs(bal) - 1;
        bal &= bal - 1;
        const int aj = a - lane + j;
        const int gj = __shfl_sync(0xffffffffu, gi, j);
        const float wgt = assigned_norm(c, b, gj, __shfl_sync(0xffffffffu, alv, j));  // = target_scores.sum(-1)
        const int l = level_of(c.t, aj);
        const int cell = aj - c.t.start[l];
        const float st = c.t.stride[l];
        const float ax = (float)(cell % c.t.w[l]) + 0.5f, ay = (float)(cell / c.t.w[l]) + 0.5f;
        const float *g = F.gt5 + ((long long)b * c.M + gj) * 5;
        int lab = (int)g[0];
        lab = lab < 0 ? 0 : lab;
        // target_bboxes /= stride_tensor (loss.py:248)
        const float4 tb = make_float4(dm::div(g[1], st), dm::div(g[2], st), dm::div(g[3], st), dm::div(g[4], st));
        const float4 pb = *reinterpret_cast<const float4 *>(c.pd_bboxes + ((long long)b * c.A + aj) * 4);
        const float iou = dm::ciou(pb, tb, dm::box1_atan(pb));  // BboxLoss.forward loss.py:85 (box1 = pred)
        // DFL (loss.py:90-113): lane holds bins of sides (lane/16) and 2 + (lane/16)
        const float *hp = c.t.ptr[l] + (long long)b * c.t.sB[l] + cell;
        const long long cs = c.t.sC[l];
        const float x0 = hp[(long long)lane * cs], x1 = hp[(long long)(32 + lane) * cs];
        float m0 = x0, m1 = x1;
#pragma unroll
        for (int of = 8; of > 0; of >>= 1) {
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, of));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, of));
        }
        float e0 = expf(x0 - m0), e1 = expf(x1 - m1);
#pragma unroll
        for (int of = 8; of > 0; of >>= 1) {
            e0 += __shfl_xor_sync(0xffffffffu, e0, of);
            e1 += __shfl_xor_sync(0xffffffffu, e1, of);
        }
        const float lse0 = m0 + logf(e0), lse1 = m1 + logf(e1);
        const float ltrb[4] = {ax - tb.x, ay - tb.y, tb.z - ax, tb.w - ay};  // bbox2dist tal.py:328-331
        float dfl = 0.f;
#pragma unroll
        for (int side = 0; side < 4; ++side) {
            const float tt = fminf(fmaxf(ltrb[side], 0.0f), (float)(kR - 1) - 0.01f);
            const int tl = (int)tt;
            const float wl = (float)(tl + 1) - tt, wr = 1.0f - wl;
            const int src = (side & 1) * 16 + tl;
            const float xs = side < 2 ? x0 : x1, ls = side < 2 ? lse0 : lse1;
            const float xl = __shfl_sync(0xffffffffu, xs, src), xr = __shfl_sync(0xffffffffu, xs, src + 1);
            const float lse = __shfl_sync(0xffffffffu, ls, (side & 1) * 16);
            dfl += (lse - xl) * wl + (lse - xr) * wr;
        }
        const float xlab = hp[(long long)(4 * kR + lab) * cs];
        s_iou += (double)(1.0f - iou) * (double)wgt;
        s_dfl += (double)(dfl * 0.25f) * (double)wgt;  // .mean(-1) over the 4 sides
        s_ts += (double)wgt;
        s_xt += (double)xlab * (double)wgt;  // BCE(x,t) - BCE(x,0) = -x*t
    }
    // all lanes hold identical sums; lane 0 of each warp publishes them
    double(*r4)[256] = red;
    if (lane == 0) { r4[0][wid] = s_iou; r4[1][wid] = s_dfl; r4[2][wid] = s_ts; r4[3][wid] = s_xt; }
    __syncthreads();
    const long long n_fg = F.n_fg;
    const long long blk = (long long)blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += r4[threadIdx.x][i];
        F.part_fg[((long long)blockIdx.z * 4 + threadIdx.x) * n_fg + blk] = s;
    }
    // last CTA done: fixed-order reduction of every partial (deterministic whichever CTA it is)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(F.counter, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x * gridDim.y * gridDim.z - 1) return;
    __threadfence();
    for (int z = 0; z < F.n_branch; ++z) {
        double acc[5] = {0, 0, 0, 0, 0};
        for (int i = threadIdx.x; i < F.n_bce; i += 256) acc[0] += __ldcg(F.part_bce + (long long)z * F.n_bce + i);
        for (int k = 0; k < 4; ++k)
            for (int i = threadIdx.x; i < F.n_fg; i += 256) acc[1 + k] += __ldcg(F.part_fg + ((long long)z * 4 + k) * n_fg + i);
        __syncthreads();
        for (int k = 0; k < 5; ++k) red[k][threadIdx.x] = acc[k];
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s)
                for (int k = 0; k < 5; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const double bce = red[0][0] - red[4][0], s_i = red[1][0], s_d = red[2][0], s_t = red[3][0];
            if (F.partials) {
                F.partials[4 * z + 0] = s_i; F.partials[4 * z + 1] = bce;
                F.partials[4 * z + 2] = s_d; F.partials[4 * z + 3] = s_t;
            }
            if (F.normalise && F.loss_items) {
                const double tss = s_t > 1.0 ? s_t : 1.0;  // max(target_scores.sum(), 1) loss.py:240
                F.loss_items[4 * z + 0] = (float)(s_i / tss * F.gain_box);
                F.loss_items[4 * z + 1] = (float)(bce / tss * F.gain_cls);
                F.loss_items[4 * z + 2] = (float)(s_d / tss * F.gain_dfl);
                F.loss_items[4 * z + 3] = (float)tss;
            }
        }
    }
}

__global__ void loss_finalize_partials_kernel(const double *__restrict__ partials, int n_branch, float gain_box,
                                              float gain_cls, float gain_dfl, float *__restrict__ loss_items) {
    const int z = threadIdx.x;
    if (z >= n_branch) return;
    const double tss = partials[4 * z + 3] > 1.0 ? partials[4 * z + 3] : 1.0;
    loss_items[4 * z + 0] = (float)(partials[4 * z + 0] / tss * gain_box);
    loss_items[4 * z + 1] = (float)(partials[4 * z + 1] / tss * gain_cls);
    loss_items[4 * z + 2] = (float)(partials[4 * z + 2] / tss * gain_dfl);
    loss_items[4 * z + 3] = (float)tss;
}

// ---------------------------------------------------------------------------------------------- host side
static bool tma_ok(const LevelTable &t) {
    for (int l = 0; l < t.nl; ++l) {
        if ((t.h[l] * t.w[l]) % 4) return false;
        if (((uintptr_t)t.ptr[l]) % 16) return false;
        if (t.sB[l] % 4 || t.sC[l] % 4) return false;
    }
    return true;
}

struct LossWs {
    size_t zero_per_branch, off_counter, zero_total;
    size_t off_scratch, scratch_per_branch;  // tgi | alignv | boxes
    size_t off_pbce, off_pfg, total;
    size_t ba4, bm;
    int n_bce, n_fg;
};
static LossWs loss_ws_layout(int nb, int B, int A, int M, int n_bce) {
    LossWs w;
    w.ba4 = a256(sizeof(int) * (size_t)B * A);
    w.bm = a256(sizeof(int) * (size_t)B * (M > 0 ? M : 1));
    w.zero_per_branch = 2 * w.ba4 + 2 * w.bm;  // claim (8 B / anchor) | pos_align | pos_ov
    w.off_counter = nb * w.zero_per_branch;
    w.zero_total = w.off_counter + 256;
    w.off_scratch = w.zero_total;
    w.scratch_per_branch = 2 * w.ba4 + 4 * w.ba4;
    w.n_bce = n_bce;
    w.n_fg = ((A + 255) / 256) * B;
    w.off_pbce = w.off_scratch + nb * w.scratch_per_branch;
    w.off_pfg = w.off_pbce + a256(sizeof(double) * (size_t)nb * n_bce);
    w.total = w.off_pfg + a256(sizeof(double) * (size_t)nb * 4 * w.n_fg);
    return w;
}
static int simple_blocks(int A, int B) { return ((A + 127) / 128) * B; }
size_t loss_workspace_bytes(int B, int A, int M) {
    int n_bce = simple_blocks(A, B);
    if (n_bce < 4 * kNumSMs) n_bce = 4 * kNumSMs;
    return loss_ws_layout(2, B, A, M, n_bce).total;
}

static int launch_stream(StreamParams &P, int *n_bce, cudaStream_t s) {
    const int C = 4 * kR + P.nc;
    bool tma = tma_ok(P.t[0]) && (P.n_branch < 2 || tma_ok(P.t[1]));
    size_t smem = sizeof(float) * (size_t)kStages * C * kTileA;
    if (smem > 220 * 1024) tma = false;
    if (tma) {
        int tiles = 0;
        for (int l = 0; l <= Y3D_MAX_LEVELS; ++l) {
            P.tile0[l] = tiles;
            if (l < P.t[0].nl) tiles += (P.t[0].h[l] * P.t[0].w[l] + kTileA - 1) / kTileA;
        }
        long long n_tiles = (long long)P.n_branch * P.B * tiles;
        int dev = 0, sms = kNumSMs;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int grid = (int)(n_tiles < sms ? n_tiles : sms);
        cudaError_t e = cudaFuncSetAttribute(loss_stream_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        *n_bce = grid;
        loss_stream_tma_kernel<<<grid, kStreamThreads, smem, s>>>(P);
    } else {
        dim3 grid((P.A + 127) / 128, P.B, P.n_branch);
        *n_bce = grid.x * grid.y;
        loss_stream_simple_kernel<<<grid, 128, 0, s>>>(P);
    }
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

struct BranchIn {
    const float *const *lvl_ptr;
    const int64_t *sB, *sC;
    int topk;
};

static int loss_run(int nb, const BranchIn *br, const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc,
                    int reg_max, const float *gt, int M, float gain_box, float gain_cls, float gain_dfl, int normalise,
                    float *loss_items, double *partials, uint8_t *dbg_fg_mask, int32_t *dbg_target_gt_idx,
                    void *const *prof_events, void *ws, size_t ws_bytes, void *stream) {
    if (!lvl_hw || !lvl_stride || B < 1 || nc < 1 || M < 0 || (M > 0 && !gt)) return Y3D_EINVAL;
    if (!loss_items && !partials) return Y3D_EINVAL;
    if (reg_max != kR) return Y3D_EUNSUPPORTED;
    AssignCtx2 cc{};
    StreamParams P{};
    int A = 0;
    for (int z = 0; z < nb; ++z) {
        if (!br[z].lvl_ptr || !br[z].sB || !br[z].sC) return Y3D_EINVAL;
        A = make_level_table(cc.c[z].t, br[z].lvl_ptr, br[z].sB, br[z].sC, lvl_hw, lvl_stride, nl);
        if (A < 0) return A;
        for (int l = 0; l < nl; ++l)
            if (!br[z].lvl_ptr[l]) return Y3D_EINVAL;
        if (br[z].topk < 1 || br[z].topk > A) return Y3D_EINVAL;
        if (br[z].topk > Y3D_MAX_TOPK) return Y3D_EUNSUPPORTED;
        P.t[z] = cc.c[z].t;
    }
    int n_bce_bound = simple_blocks(A, B);
    if (n_bce_bound < 4 * kNumSMs) n_bce_bound = 4 * kNumSMs;
    LossWs w = loss_ws_layout(nb, B, A, M, n_bce_bound);
    if (!ws || ws_bytes < w.total) return Y3D_EWORKSPACE;
    if (((uintptr_t)ws) % 256) return Y3D_EALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    char *p = (char *)ws;
    auto mark = [&](int i) {
        if (prof_events && prof_events[i]) cudaEventRecord((cudaEvent_t)prof_events[i], s);
    };
    mark(0);
    cudaError_t e = cudaMemsetAsync(p, 0, w.zero_total, s);
    if (e != cudaSuccess) return (int)e;
    P.n_branch = nb; P.B = B; P.nc = nc; P.A = A;
    P.part_bce = (double *)(p + w.off_pbce);
    for (int z = 0; z < nb; ++z) {
        char *sc = p + w.off_scratch + z * w.scratch_per_branch;
        AssignCtx &c = cc.c[z];
        c.claim = (unsigned long long *)(p + z * w.zero_per_branch);
        c.pos_align = (int *)(p + z * w.zero_per_branch + 2 * w.ba4);
        c.pos_ov = (int *)(p + z * w.zero_per_branch + 2 * w.ba4 + w.bm);
        c.tgi = (int *)sc;
        c.alignv = (float *)(sc + w.ba4);
        float *boxes = (float *)(sc + 2 * w.ba4);
        P.pd_bboxes[z] = boxes;
        P.pd_scores[z] = nullptr;
        c.score_mode = 1;
        c.cls_ch0 = 4 * kR;
        c.pd_bboxes = boxes; c.box_grid_units = 1;
        c.use_grid = 1;
        c.gt_labels = gt; c.gl_stride = 5;
        c.gt_bboxes = gt ? gt + 1 : nullptr; c.gb_stride = 5;
        c.mask_gt = nullptr;
        c.B = B; c.A = A; c.nc = nc; c.M = M; c.k = br[z].topk;
        c.alpha = 0.5f; c.beta = 6.0f; c.gamma = 1.0f; c.eps = 1e-9f;  // loss.py:176
        c.use_2d = 1; c.use_3d = 0; c.kps_l2 = 0; c.constrain = 1;
    }
    int n_bce = 0;
    int rc = launch_stream(P, &n_bce, s);
    if (rc) return rc;
    mark(1);
    if (M > 0) {
        rc = assign_run_core(cc, nb, s, prof_events ? (cudaEvent_t)prof_events[2] : nullptr);
        if (rc) return rc;
    } else {
        mark(2);
    }
    mark(3);
    FgParams F{};
    F.gt5 = gt;
    F.part_bce = P.part_bce;
    F.part_fg = (double *)(p + w.off_pfg);
    F.counter = (unsigned *)(p + w.off_counter);
    F.partials = partials;
    F.loss_items = loss_items;
    for (int z = 0; z < nb; ++z) {
        F.dbg_fg[z] = dbg_fg_mask ? dbg_fg_mask + (size_t)z * B * A : nullptr;
        F.dbg_gi[z] = dbg_target_gt_idx ? dbg_target_gt_idx + (size_t)z * B * A : nullptr;
    }
    F.n_bce = n_bce; F.n_fg = w.n_fg; F.n_branch = nb; F.normalise = normalise;
    F.gain_box = gain_box; F.gain_cls = gain_cls; F.gain_dfl = gain_dfl;
    dim3 grid((A + 255) / 256, B, nb);
    loss_fg_kernel<<<grid, 256, 0, s>>>(cc, F);
    Y3D_CHECK_LAUNCH();
    mark(4);
    return Y3D_OK;
}

}  // namespace y3d

using namespace y3d;

extern "C" int y3d_train_decode(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC,
                                const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                                float *pd_bboxes, float *pd_scores, void *stream) {
    if (!lvl_ptr || !lvl_sB || !lvl_sC || !pd_bboxes || B < 0 || nc < 1) return Y3D_EINVAL;
    if (reg_max != kR) return Y3D_EUNSUPPORTED;
    StreamParams P{};
    int A = make_level_table(P.t[0], lvl_ptr, lvl_sB, lvl_sC, lvl_hw, lvl_stride, nl);
    if (A < 0) return A;
    for (int l = 0; l < nl; ++l)
        if (!lvl_ptr[l]) return Y3D_EINVAL;
    if (B == 0) return Y3D_OK;
    P.t[1] = P.t[0];
    P.n_branch = 1; P.B = B; P.nc = nc; P.A = A;
    P.pd_bboxes[0] = pd_bboxes;
    P.pd_scores[0] = pd_scores;
    P.part_bce = nullptr;
    int n_bce;
    return launch_stream(P, &n_bce, (cudaStream_t)stream);
}

extern "C" int y3d_v8_loss_fwd(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC,
                               const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                               const float *gt, int M, int topk, float gain_box, float gain_cls, float gain_dfl,
                               int normalise, float *loss_items, double *partials, uint8_t *dbg_fg_mask,
                               int32_t *dbg_target_gt_idx, void *const *prof_events, void *ws, size_t ws_bytes,
                               void *stream) {
    BranchIn br[1] = {{lvl_ptr, lvl_sB, lvl_sC, topk}};
    return loss_run(1, br, lvl_hw, lvl_stride, nl, B, nc, reg_max, gt, M, gain_box, gain_cls, gain_dfl, normalise,
                    loss_items, partials, dbg_fg_mask, dbg_target_gt_idx, prof_events, ws, ws_bytes, stream);
}

extern "C" int y3d_v10_loss_fwd(const float *const *o2m_ptr, const int64_t *o2m_sB, const int64_t *o2m_sC,
                                const float *const *o2o_ptr, const int64_t *o2o_sB, const int64_t *o2o_sC,
                                const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                                const float *gt, int M, int topk_o2m, int topk_o2o, float gain_box, float gain_cls,
                                float gain_dfl, int normalise, float *loss_items, double *partials,
                                uint8_t *dbg_fg_mask, int32_t *dbg_target_gt_idx, void *const *prof_events, void *ws,
                                size_t ws_bytes, void *stream) {
    BranchIn br[2] = {{o2m_ptr, o2m_sB, o2m_sC, topk_o2m}, {o2o_ptr, o2o_sB, o2o_sC, topk_o2o}};
    return loss_run(2, br, lvl_hw, lvl_stride, nl, B, nc, reg_max, gt, M, gain_box, gain_cls, gain_dfl, normalise,
                    loss_items, partials, dbg_fg_mask, dbg_target_gt_idx, prof_events, ws, ws_bytes, stream);
}

extern "C" int y3d_v8_loss_finalize(const double *partials, int n_branch, float gain_box, float gain_cls,
                                    float gain_dfl, float *loss_items, void *stream) {
    if (!partials || !loss_items || n_branch < 1 || n_branch > 2) return Y3D_EINVAL;
    loss_finalize_partials_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partials, n_branch, gain_box, gain_cls, gain_dfl,
                                                                     loss_items);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}
