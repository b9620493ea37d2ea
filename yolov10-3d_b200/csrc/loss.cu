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
// (blockIdx.z / .y selects the branch); a step is three kernels and no memset:
//   1. head_stream_kernel : the ONE pass over the two head tensors (4*(4R+nc)*A bytes per image and branch).
//        A thread owns 4 consecutive anchors (128-bit loads along the anchor axis, 512 contiguous bytes per warp and
//        channel row) and one quarter of the channels: DFL side `part` (16 bins) and nc/4 classes.  Out: boxes
//        (xyxy, grid units) and the per-side log-sum-exp as [B,A,4] records (coalesced 128-bit stores; 32 B/anchor,
//        the only dense writes), zeroed claim words, and sum softplus(logit) = sum BCE(logit, 0) per CTA.
//        pd_scores and the dense target_scores of the reference are never materialised:
//        sum BCE(x,t) = sum BCE(x,0) - sum_fg x[label]*t.  One warp per image also sorts the image's valid GTs into
//        size classes: the work order of the next kernel.
//   2. tal_topk_fused_kernel (topk_fused.cu; the generic tal_topk_kernel of assign.cu when there are few GTs): per-GT
//        candidate walk + top-k + claims, biggest GTs first.  Scores are read as logits straight from the head.  With up
//        to kFinishApMaxM GTs per image the claiming lane also evaluates the pair's loss terms and appends a claim
//        record to the image's list; a per-image counter tells kernel 3 when an image is complete.  Launched
//        programmatically dependent on 1, as 3 is on 2.
//   3. loss_finish_ap_kernel (M <= kFinishApMaxM): the claim records of an image spread over the machine, 128 per CTA:
//        conflict resolution (select_highest_overlaps), per-GT maxima, weights, exact (fixed-point, order-independent)
//        sums; the CTA that finishes last for the batch normalises and writes the loss items -- after summing them over
//        the ranks through NVLink peer memory when the batch is sharded over several GPUs (y3d_v10_loss_fwd_sharded).
//      loss_finish_kernel (more GTs per image): one CTA per (image, branch) over the image's list of claimed anchors.
//      Either way the claim word of a foreground anchor ends up as (GT index, alignment weight): what the backward
//      pass reads.
#include "loss.cuh"
#include "xrank.cuh"

namespace y3d {

constexpr int kStreamThreads = 128;  // 32 anchor-quads x 4 channel parts
constexpr int kFinishThreads = 1024;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ---------------------------------------------------------------------------------------------- shared arithmetic
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_acc(float v) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v))); }

template <int V>
__device__ __forceinline__ void ldv(const float *p, float (&o)[V], unsigned long long pol) {
    if constexpr (V == 4) {
        const float4 r = ldg_stream4(p, pol);
        o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
    } else {
        o[0] = ldg_stream1(p, pol);
    }
}
template <int V>
__device__ __forceinline__ void stv(float *p, const float (&o)[V]) {
    if constexpr (V == 4) {
        *reinterpret_cast<float4 *>(p) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
        p[0] = o[0];
    }
}

#ifdef Y3D_TIMING
// step timeline (developer builds): [0] ~first stream CTA start [1] last stream CTA end [2] ~first finish CTA start
// [3] ~first finish CTA past its image wait [4] last phase-R end [5] final reduction end   (~t under atomicMax = minimum)
// one row of 8 per step, 16 steps kept (g_y3d_step counts the steps: bumped by the finishing kernel's very last CTA)
__device__ unsigned long long g_y3d_tl[16 * 8];
__device__ unsigned g_y3d_step;
__device__ __forceinline__ unsigned long long gtimer0() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define Y3D_TL_MIN(i) do { if (threadIdx.x == 0) atomicMax(&g_y3d_tl[(*(volatile unsigned *)&g_y3d_step & 15) * 8 + (i)], ~gtimer0()); } while (0)
#define Y3D_TL_MAX(i) do { if (threadIdx.x == 0) atomicMax(&g_y3d_tl[(*(volatile unsigned *)&g_y3d_step & 15) * 8 + (i)], gtimer0()); } while (0)
#define Y3D_TL_STEP() do { if (threadIdx.x == 0) { __threadfence(); atomicAdd(&g_y3d_step, 1u); } } while (0)
extern "C" int y3d_debug_read_timeline(unsigned long long *host, int reset) {
    int rc = (int)cudaMemcpyFromSymbol(host, g_y3d_tl, sizeof(unsigned long long) * 16 * 8);
    if (reset) {
        unsigned long long z[16 * 8] = {};
        cudaMemcpyToSymbol(g_y3d_tl, z, sizeof(z));
        unsigned zero = 0;
        cudaMemcpyToSymbol(g_y3d_step, &zero, sizeof(zero));
    }
    return rc;
}
__device__ unsigned long long g_y3d_stamps[2048 * 8];
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define Y3D_STAMP(i)                                                                                  \
    do {                                                                                              \
        __syncthreads();                                                                              \
        if (threadIdx.x == 0) g_y3d_stamps[((blockIdx.y * gridDim.x + blockIdx.x) & 2047) * 8 + (i)] = gtimer(); \
    } while (0)
#else
#define Y3D_STAMP(i)
#define Y3D_TL_MIN(i)
#define Y3D_TL_MAX(i)
#define Y3D_TL_STEP()
#endif

// ---------------------------------------------------------------------------------------------- stream kernel
struct StreamParams {
    LevelTable t[2];
    float *boxes[2];       // [B,A,4] xyxy grid units (the layout of the reference's pred_bboxes, loss.py:233)
    float *lse[2];         // optional [B,A,4]: log-sum-exp of the 16 DFL bins of every side
    float *pd_scores[2];   // optional [B,A,nc] sigmoid (y3d_train_decode only)
    unsigned long long *claim[2];  // optional [B,A], zeroed here
    int *list_count[2];    // optional [B], zeroed here
    int *img_cnt[2];       // optional [B], zeroed here: finished chunks of the image (anchor-parallel finishing kernel)
    unsigned *topk_done[2];  // optional [B], zeroed here: GTs of the image the top-k kernel has finished
    int *pos[2];           // optional [B,M,2], zeroed here: per-GT maxima of alignment metric and overlap
    unsigned *counter;     // optional ticket of the finishing kernel (+1: work counter of the top-k kernel), zeroed here
    double *part_bce;      // [n_branch][gridDim.x * B] or nullptr
    const float *gt5;      // optional [B,M,5] (with ord_cnt / ord_list)
    int *ord_cnt, *ord_list;  // optional [B][4], [B][M] x 32-byte records: see AssignCtx2
    float ord_cells_per_px2;  // sum over the levels of 1 / stride^2: candidate cells of a GT per px^2 of its area
    int M;
    int n_branch, B, nc, A;
};

// One warp sorts the valid GTs of image b into kOrdClasses classes by their number of candidate cells (stable within a
// class, so the result does not depend on timing).  Out of line: the streaming kernel's register budget is not touched.
static __device__ __noinline__ void gt_order_image(const float *gt5, int M, int b, float cells_per_px2, int *ord_cnt,
                                                   int *ord_list, int lane) {
    const float *g = gt5 + (long long)b * M * 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    int base = 0;
    for (int c = 0; c < kOrdClasses; ++c) {
        int cnt = 0;
        for (int m0 = 0; m0 < M; m0 += 32) {
            const int m = m0 + lane;
            int cls = -1;
            if (m < M) {
                const float x1 = g[m * 5 + 1], y1 = g[m * 5 + 2], x2 = g[m * 5 + 3], y2 = g[m * 5 + 4];
                if (dm::add(dm::add(dm::add(x1, y1), x2), y2) > 0.0f) {  // gt_valid (assign.cuh)
                    const float cells = (x2 - x1) * (y2 - y1) * cells_per_px2;
                    cls = cells >= 650.0f ? 0 : cells >= 450.0f ? 1 : cells >= 180.0f ? 2 : 3;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, cls == c);
            if (cls == c) {  // the record the top-k kernel loads: GT index, label, box (32 bytes)
                int4 *rec = reinterpret_cast<int4 *>(ord_list) + ((long long)b * M + base + cnt + __popc(bal & lt_mask)) * 2;
                rec[0] = make_int4(m, (int)g[m * 5], __float_as_int(g[m * 5 + 1]), __float_as_int(g[m * 5 + 2]));
                rec[1] = make_int4(__float_as_int(g[m * 5 + 3]), __float_as_int(g[m * 5 + 4]), 0, 0);
            }
            cnt += __popc(bal);
        }
        if (lane == 0) ord_cnt[b * 4 + c] = cnt;
        base += cnt;
    }
}

static __device__ __noinline__ void zero_ints(int *p, int n, int lane) {
    for (int i = lane; i < n; i += 32) p[i] = 0;
}

// grid (ceil(A/V/32), B, n_branch), block 128 = 32 units of V anchors x 4 channel parts (warp = part)
template <int V, bool PS>
__global__ void __launch_bounds__(kStreamThreads, 4) head_stream_kernel(const __grid_constant__ StreamParams P) {
    __shared__ double red[4];
    // the four warps of a CTA each decode one side of the same 32 * V anchors: the sides meet here so that boxes and
    // log-sum-exps leave as one 16-byte record per anchor (what every later gather wants), in coalesced stores
    __shared__ __align__(16) float s_box[4][32 * V];
    __shared__ __align__(16) float s_lse[4][32 * V];
    const int z = blockIdx.z, b = blockIdx.y;
    const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
    const int q = blockIdx.x * 32 + lane;
    const LevelTable &t = P.t[z];
    const int A = P.A;
    const unsigned long long pol = l2_evict_first_policy();
    // lets a programmatically dependent launch (the fused loss' top-k kernel) be scheduled as this grid drains.  (This
    // grid itself is launched the ordinary way: parked behind the previous step's last kernel, its first wave starts in
    // lockstep and the whole pass ran 20 % slower -- measured.)
    asm volatile("griddepcontrol.launch_dependents;");
    Y3D_TL_MIN(0);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (P.list_count[z]) P.list_count[z][b] = 0;
        if (P.img_cnt[z]) P.img_cnt[z][b] = 0;
        if (P.topk_done[z]) P.topk_done[z][b] = 0u;
        if (b == 0 && z == 0 && P.counter) {  // tickets, and the fixed-point accumulators of the finishing kernel
            P.counter[0] = 0u; P.counter[1] = 0u;
            for (int i = 16; i < 16 + 2 * 2 * 5; ++i) P.counter[i] = 0u;
        }
    }
    float bce = 0.f;
    if (q * V < A) {
        const int a0 = q * V;
        const int l = level_of(t, a0);
        const int cell = a0 - t.start[l];
        const float *base = t.ptr[l] + (long long)b * t.sB[l] + cell;
        const long long cs = t.sC[l];
        // ---- DFL side `part`: softmax over 16 bins, expectation (bbox_decode loss.py:199-201) and log-sum-exp
        {
            float x[kR][V];
            const float *pb = base + (long long)(part * kR) * cs;
#pragma unroll
            for (int j = 0; j < kR; ++j) ldv<V>(pb + (long long)j * cs, x[j], pol);
            const int w = t.w[l];
            int cx = cell % w, cy = cell / w;
            float ob[V], ol[V];
#pragma unroll
            for (int i = 0; i < V; ++i) {
                float m = x[0][i];
#pragma unroll
                for (int j = 1; j < kR; ++j) m = fmaxf(m, x[j][i]);
                const float mo = -m * kLog2e;
                float s = 0.f, acc = 0.f;
#pragma unroll
                for (int j = 0; j < kR; ++j) {
                    const float e = ex2_approx(__fmaf_rn(x[j][i], kLog2e, mo));
                    s = __fadd_rn(s, e);
                    acc = __fmaf_rn((float)j, e, acc);
                }
                const float d = __fdiv_rn(acc, s);
                ol[i] = __fmaf_rn(lg2_approx(s), kLn2, m);
                const float anc = ((part & 1) ? (float)cy : (float)cx) + 0.5f;
                ob[i] = part < 2 ? __fsub_rn(anc, d) : __fadd_rn(anc, d);  // dist2bbox xyxy tal.py:319-325
                if (++cx >= w) { cx = 0; ++cy; }
            }
            stv<V>(&s_box[part][lane * V], ob);
            stv<V>(&s_lse[part][lane * V], ol);
        }
        if (part == 0 && P.claim[z]) {
            unsigned long long *cl = P.claim[z] + (long long)b * A + a0;
            if constexpr (V == 4) {
                reinterpret_cast<ulonglong2 *>(cl)[0] = make_ulonglong2(0ull, 0ull);
                reinterpret_cast<ulonglong2 *>(cl)[1] = make_ulonglong2(0ull, 0ull);
            } else {
                cl[0] = 0ull;
            }
        }
        // ---- classes of this part: sum_c softplus(x_c) = sum_c max(x_c,0) + log prod_c (1 + exp(-|x_c|)).
        // The product is carried as u = prod - 1 (u' = u + t + u*t), so tiny terms keep full relative precision and a
        // single log1p per anchor and segment replaces one log per element.  Segments of 64 keep u < 2^64.
        const int cpp = (P.nc + 3) >> 2;
        const int c_lo = part * cpp, c_hi = min(P.nc, c_lo + cpp);
        float tot[V];
#pragma unroll
        for (int i = 0; i < V; ++i) tot[i] = 0.f;
        for (int s0 = c_lo; s0 < c_hi; s0 += 64) {
            const int s1 = min(c_hi, s0 + 64);
            float u[V], pos[V];
#pragma unroll
            for (int i = 0; i < V; ++i) u[i] = pos[i] = 0.f;
            const float *pc = base + (long long)(4 * kR + s0) * cs;
            auto fold = [&](const float (&v)[V], int c) {
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float tt = ex2_approx(-fabsf(v[i]) * kLog2e);
                    u[i] = __fmaf_rn(u[i], tt, u[i] + tt);
                    pos[i] += fmaxf(v[i], 0.f);
                    if constexpr (PS) P.pd_scores[z][((long long)b * A + a0 + i) * P.nc + c] = sigmoid_acc(v[i]);
                }
            };
            int c = s0;
            constexpr int CB = 10;  // channel rows in flight per thread (nc = 80: two batches)
            for (; c + CB <= s1; c += CB, pc += CB * cs) {
                float v[CB][V];
#pragma unroll
                for (int jj = 0; jj < CB; ++jj) ldv<V>(pc + jj * cs, v[jj], pol);
#pragma unroll
                for (int jj = 0; jj < CB; ++jj) fold(v[jj], c + jj);
            }
            for (; c < s1; ++c, pc += cs) {
                float v[V];
                ldv<V>(pc, v, pol);
                fold(v, c);
            }
#pragma unroll
            for (int i = 0; i < V; ++i) tot[i] += pos[i] + log1pf(u[i]);
        }
#pragma unroll
        for (int i = 0; i < V; ++i) bce += tot[i];
    }
    {  // fixed-order block reduction -> one partial per CTA
        const double w = warp_sum((double)bce);
        if (lane == 0) red[part] = w;
        __syncthreads();
        if (P.part_bce && threadIdx.x == 0)
            P.part_bce[((long long)z * gridDim.y + b) * gridDim.x + blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
    }
    {  // anchor t of this CTA's 32 * V: its four sides as one float4
        const int t = threadIdx.x;
        const int a = blockIdx.x * 32 * V + t;
        if (t < 32 * V && a < A) {
            reinterpret_cast<float4 *>(P.boxes[z])[(long long)b * A + a] =
                make_float4(s_box[0][t], s_box[1][t], s_box[2][t], s_box[3][t]);
            if (P.lse[z])
                reinterpret_cast<float4 *>(P.lse[z])[(long long)b * A + a] =
                    make_float4(s_lse[0][t], s_lse[1][t], s_lse[2][t], s_lse[3][t]);
        }
    }
    // one warp per image: the processing order of the image's GTs for the top-k kernel
    if (P.ord_cnt && blockIdx.x == 0 && z == 0 && part == 0)
        gt_order_image(P.gt5, P.M, b, P.ord_cells_per_px2, P.ord_cnt, P.ord_list, lane);
    // one warp per image and branch: zero the per-GT maxima the finishing kernel folds its atomicMax into
    if (P.pos[z] && blockIdx.x == 0 && part == 1) zero_ints(P.pos[z] + (long long)b * P.M * 2, 2 * P.M, lane);
    Y3D_TL_MAX(1);
}

// ---------------------------------------------------------------------------------------------- finishing kernel
struct FinishParams {
    const float *gt5;          // [B,M,5]
    const float *lse[2];       // [B,4,A]
    int *list_gi[2];           // [B,cap] scratch (per-image kernel)
    float *list_al[2];         // [B,cap] scratch (per-image kernel)
    int *img_cnt[2];           // [B] zeroed by the stream kernel: chunks of the image that have finished
    int *pos[2];               // [B,M,2] zeroed by the stream kernel: per-GT max alignment metric / max overlap (float bits)
    long long *acc;            // [n_branch][5] zeroed by the stream kernel: fixed-point sums over the images (anchor-parallel
                               // kernel): iou, dfl, target_scores, x*t, softplus
    double *img_bce;           // [n_branch][B] scratch: softplus partials of an image, summed ahead of time
    const double *part_bce;    // [n_branch][n_bce]
    double *part_fg;           // [n_branch][B][5]: iou, dfl, target_scores, x*t, softplus
    unsigned *counter;         // zeroed by the stream kernel
    double *partials;          // optional out [n_branch][4]
    float *loss_items;         // optional out [n_branch][4]
    float *loss_total;         // optional out [1]: total_scale * sum over the branches of (box + cls + dfl)
    float total_scale;
    uint8_t *dbg_fg[2];
    int32_t *dbg_gi[2];
    int n_bce_x, n_branch, normalise;  // n_bce_x: BCE partials per image (= gridDim.x of the stream kernel)
    float gain_box, gain_cls, gain_dfl;
    // image-sharded loss over several GPUs (optional, x_world > 1): the last CTA exchanges the partial sums with the
    // peer ranks over NVLink peer memory (xrank.cuh) before it normalises -- compute and collective in one kernel
    XSlot *const *x_bufs;      // DEVICE table of x_world exchange buffers
    int x_rank, x_world;
    unsigned long long x_seq;
    long long x_timeout;       // clock cycles (xrank_timeout_cycles)
    int x_defer;               // post this rank's partial sums only; y3d_loss_exchange_resolve collects and normalises
    int *x_status;             // optional: 1 when a peer never arrived
};

constexpr double kFix = 4294967296.0;  // 2^32: loss terms are summed as 64-bit fixed point (exact, order-independent)
__device__ __forceinline__ long long to_fix(float v) { return __double2ll_rn((double)v * kFix); }
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}



constexpr int kConfMax = 1024;  // multiply-claimed anchors resolved cooperatively per image (more: serial fallback)
constexpr int kFinishWarps = kFinishThreads / 32;
constexpr int kBucketPx = 32;       // coarse cell of the per-image GT index (pixels)
constexpr int kBucketCells = 2048;  // at most this many coarse cells (e.g. 1280 x 1280 -> 1600)
constexpr int kBucketCap = 12288;   // (GT, coarse cell) incidences kept; larger sets fall back to the all-GT scan

struct FinishSmem {  // dynamic shared memory of loss_finish_kernel, followed by GtRec[M], int pos_a[M], int pos_o[M]
    float4 conf_box[kConfMax];            // predicted box (px) of a conflicted anchor
    unsigned long long conf_best[kConfMax];  // (overlap bits << 32) | (0xffffffff - m): atomicMax = first maximum
    float2 conf_xy[kConfMax];             // its anchor point (px)
    int2 queue[kFinishWarps][64];         // per-warp compaction queue of in-GT (conflict, GT) pairs
    int bucket_start[kBucketCells + 1];   // coarse spatial index of the image's GT boxes: CSR over 32 px cells
    int bucket_fill[kBucketCells];
    unsigned short bucket_items[kBucketCap];
    int bucket_total, bucket_ok;
    long long redl[4][kFinishWarps];
    double redd[kFinishWarps];
    double fin[2][5];
    double xv[kXMaxVals], xs[kXMaxVals];  // this rank's partial sums, the sums over the ranks
    int xfail;
    int n_conf;
    unsigned ticket;
};
size_t finish_smem_bytes(int M) { return sizeof(FinishSmem) + (sizeof(GtRec) + 2 * sizeof(int)) * (size_t)M; }

// grid (B, n_branch), block kFinishThreads; one CTA per (image, branch) over the image's claimed anchors only
__global__ void __launch_bounds__(kFinishThreads) loss_finish_kernel(AssignCtx2 cc, FinishParams F) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FinishSmem &S = *reinterpret_cast<FinishSmem *>(smem_raw);
    const int z = blockIdx.y, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const AssignCtx &c = cc.c[z];
    const int M = c.M, A = c.A;
    Y3D_STAMP(0);
    GtRec *gts = reinterpret_cast<GtRec *>(smem_raw + sizeof(FinishSmem));
    int *pos_a = reinterpret_cast<int *>(gts + M);
    int *pos_o = pos_a + M;
    for (int m = tid; m < M; m += kFinishThreads) {
        gts[m] = load_gt(c, b, m);
        pos_a[m] = 0;
        pos_o[m] = 0;
    }
    if (tid == 0) S.n_conf = 0;
    if (F.dbg_fg[z])
        for (int a = tid; a < A; a += kFinishThreads) {
            F.dbg_fg[z][(long long)b * A + a] = 0;
            F.dbg_gi[z][(long long)b * A + a] = 0;
        }
    // Everything above reads only the caller's inputs; what follows reads what the top-k kernel wrote.  With
    // programmatic dependent launch this CTA may have started while that kernel was still draining: wait here.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int n = M > 0 ? min(__ldcg(c.list_count + b), c.list_cap) : 0;
    const int *la = c.list_a + (long long)b * c.list_cap;
    int *lgi = F.list_gi[z] + (long long)b * c.list_cap;
    float *lal = F.list_al[z] + (long long)b * c.list_cap;
    __syncthreads();
    // ---- coarse spatial index of the GT boxes (32 px cells): a contested anchor then tests the few GTs registered in
    //      its cell instead of all M (dense crowds: 500 GTs, thousands of contested anchors per image)
    const int gw = (int)ceilf(c.t.w[0] * c.t.stride[0] / kBucketPx), gh = (int)ceilf(c.t.h[0] * c.t.stride[0] / kBucketPx);
    const int ncell = gw * gh;
    auto cell_range = [&](const float4 &bx, int &x0, int &x1, int &y0, int &y1) {
        x0 = min(max((int)floorf(bx.x / kBucketPx), 0), gw - 1); x1 = min(max((int)floorf(bx.z / kBucketPx), 0), gw - 1);
        y0 = min(max((int)floorf(bx.y / kBucketPx), 0), gh - 1); y1 = min(max((int)floorf(bx.w / kBucketPx), 0), gh - 1);
    };
    // (worth its construction only when the all-GT scan is long: with M <= 128 the cooperative scan below is faster)
    if (tid == 0) { S.bucket_total = 0; S.bucket_ok = (ncell <= kBucketCells && M > 128 && M <= 65535 && n > 0) ? 1 : 0; }
    __syncthreads();
    if (S.bucket_ok) {
        for (int i = tid; i < ncell; i += kFinishThreads) { S.bucket_start[i] = 0; S.bucket_fill[i] = 0; }
        __syncthreads();
        for (int m = tid; m < M; m += kFinishThreads)
            if (gts[m].valid) {
                int x0, x1, y0, y1;
                cell_range(gts[m].box, x0, x1, y0, y1);
                atomicAdd(&S.bucket_total, (x1 - x0 + 1) * (y1 - y0 + 1));
                for (int y = y0; y <= y1; ++y)
                    for (int x = x0; x <= x1; ++x) atomicAdd(&S.bucket_start[y * gw + x], 1);
            }
        __syncthreads();
        if (S.bucket_total > kBucketCap) {
            if (tid == 0) S.bucket_ok = 0;
        } else if (wid == 0) {  // exclusive prefix over the cells (one warp, lane-strided chunks)
            const int per = (ncell + 31) / 32;
            const int lo = min(lane * per, ncell), hi = min(lo + per, ncell);
            int s_ = 0;
            for (int i = lo; i < hi; ++i) s_ += S.bucket_start[i];
            int inc = s_;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            int run = inc - s_;
            for (int i = lo; i < hi; ++i) {
                const int cnt_i = S.bucket_start[i];
                S.bucket_start[i] = run;
                run += cnt_i;
            }
            if (lane == 31) S.bucket_start[ncell] = run;
        }
        __syncthreads();
        if (S.bucket_ok)
            for (int m = tid; m < M; m += kFinishThreads)
                if (gts[m].valid) {
                    int x0, x1, y0, y1;
                    cell_range(gts[m].box, x0, x1, y0, y1);
                    for (int y = y0; y <= y1; ++y)
                        for (int x = x0; x <= x1; ++x) {
                            const int cell = y * gw + x;
                            S.bucket_items[S.bucket_start[cell] + atomicAdd(&S.bucket_fill[cell], 1)] = (unsigned short)m;
                        }
                }
        __syncthreads();
    }
    const bool use_buckets = S.bucket_ok != 0;
    Y3D_STAMP(1);
    // ---- resolve, step 1: singly-claimed anchors are final; multiply-claimed ones (select_highest_overlaps
    //      tal.py:237-264: argmax over ALL GTs of the overlap, first maximum) are parked for the cooperative step
    auto settle = [&](int e, int a, int gi, float ax, float ay, const PairRaw &raw, float x) {
        const GtRec g = gts[gi];
        float metric = 0.0f, ovl = 0.0f;
        if (g.valid && dm::in_gt(ax, ay, g.box))
            metric = pair_metric(c, b, gi, g, a, raw, dm::pow_(pair_score(c, x), c.alpha), ovl);
        lgi[e] = gi;
        lal[e] = metric;
        atomicMax(pos_a + gi, __float_as_int(metric));  // values >= 0: int order == float order
        atomicMax(pos_o + gi, __float_as_int(ovl));
    };
    for (int e = tid; e < n; e += kFinishThreads) {
        const int a = __ldcg(la + e);
        const unsigned long long cl = __ldcg(c.claim + (long long)b * A + a);
        const int cnt = (int)(cl >> 32);
        float ax, ay, st;
        anchor_px(c, a, ax, ay, st);
        int gi = (int)(cl & 0xffffffffull);
        const PairRaw raw = pair_load_box(c, b, a);
        if (cnt > 1 && use_buckets) {  // only the GTs registered in the anchor's coarse cell can contain it
            const float4 pbox = pair_box(c, raw, a);
            const int cell = min(max((int)floorf(ay / kBucketPx), 0), gh - 1) * gw + min(max((int)floorf(ax / kBucketPx), 0), gw - 1);
            unsigned long long best = 0xffffffffull;  // overlap 0 at GT 0
            for (int i = S.bucket_start[cell]; i < S.bucket_start[cell + 1]; ++i) {
                const int m = S.bucket_items[i];
                const GtRec g = gts[m];
                if (dm::in_gt(ax, ay, g.box)) {
                    const float ovl = dm::ciou(g.box, pbox, g.at1);
                    if (ovl > 0.0f) {
                        const unsigned long long key =
                            ((unsigned long long)__float_as_uint(ovl) << 32) | (unsigned long long)(0xffffffffu - (unsigned)m);
                        best = key > best ? key : best;
                    }
                }
            }
            gi = (int)(0xffffffffu - (unsigned)(best & 0xffffffffull));
        } else if (cnt > 1) {
            const float4 pbox = pair_box(c, raw, a);
            const int ci = atomicAdd(&S.n_conf, 1);
            if (ci < kConfMax) {
                S.conf_box[ci] = pbox;
                S.conf_xy[ci] = make_float2(ax, ay);
                S.conf_best[ci] = 0xffffffffull;  // overlap 0 at GT 0: what the scan starts from
                lgi[e] = -1 - ci;
                continue;
            }
            float bv = -1.0f;  // overflow of the cooperative buffers: serial scan
            gi = 0;
            for (int m = 0; m < M; ++m) {
                const GtRec g = gts[m];
                float ovl = 0.0f;
                if (g.valid && dm::in_gt(ax, ay, g.box)) {
                    ovl = dm::ciou(g.box, pbox, g.at1);
                    ovl = ovl < 0.0f ? 0.0f : ovl;
                }
                if (ovl > bv) { bv = ovl; gi = m; }
            }
        }
        settle(e, a, gi, ax, ay, raw, pair_load_score(c, b, a, gts[gi].label));
    }
    __syncthreads();
    // ---- resolve, step 2: warp w takes conflicts w, w+32, ...; lanes test the GTs, in-GT pairs are compacted and
    //      their CIoU evaluated on full lanes
    const int n_conf = min(S.n_conf, kConfMax);
    if (n_conf > 0) {
        int qn = 0;
        const unsigned lt_mask = (1u << lane) - 1u;
        auto drain = [&](int take) {
            qn -= take;
            if (lane < take) {
                const int2 pr = S.queue[wid][qn + lane];
                const GtRec g = gts[pr.y];
                float ovl = dm::ciou(g.box, S.conf_box[pr.x], g.at1);
                if (ovl > 0.0f)
                    atomicMax(&S.conf_best[pr.x],
                              ((unsigned long long)__float_as_uint(ovl) << 32) | (unsigned long long)(0xffffffffu - (unsigned)pr.y));
            }
            __syncwarp();
        };
        for (int ci = wid; ci < n_conf; ci += kFinishWarps) {
            const float2 xy = S.conf_xy[ci];
            for (int m0 = 0; m0 < M; m0 += 32) {
                const int m = m0 + lane;
                bool in = false;
                if (m < M) {
                    const GtRec g = gts[m];
                    in = g.valid && dm::in_gt(xy.x, xy.y, g.box);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                if (in) S.queue[wid][qn + __popc(bal & lt_mask)] = make_int2(ci, m);
                qn += __popc(bal);
                __syncwarp();
                if (qn >= 32) drain(32);
            }
        }
        if (qn > 0) drain(qn);
    }
    __syncthreads();
    // ---- resolve, step 3: settle the conflicted anchors on their winning GT
    if (n_conf > 0)
        for (int e = tid; e < n; e += kFinishThreads) {
            const int code = lgi[e];
            if (code >= 0) continue;
            const int ci = -1 - code;
            const int a = __ldcg(la + e);
            const int gi = (int)(0xffffffffu - (unsigned)(S.conf_best[ci] & 0xffffffffull));
            const float2 xy = S.conf_xy[ci];
            settle(e, a, gi, xy.x, xy.y, pair_load_box(c, b, a), pair_load_score(c, b, a, gts[gi].label));
        }
    __syncthreads();
    Y3D_STAMP(2);
    // ---- loss terms of the foreground anchors (BboxLoss.forward loss.py:82-113; BCE correction -x[label]*t)
    long long s_iou = 0, s_dfl = 0, s_ts = 0, s_xt = 0;
    for (int e = tid; e < n; e += kFinishThreads) {
        const int a = __ldcg(la + e);
        const int gi = lgi[e];
        const float alv = lal[e];
        const float pa = __int_as_float(pos_a[gi]), po = __int_as_float(pos_o[gi]);
        const float wgt = dm::div(dm::mul(alv, po), dm::add(pa, c.eps));  // = target_scores.sum(-1), tal.py:89-92
        // kept for the backward pass: the claim word of a foreground anchor becomes (1 << 63 | GT index << 32 | weight)
        c.claim[(long long)b * A + a] =
            0x8000000000000000ull | ((unsigned long long)(unsigned)gi << 32) | (unsigned long long)__float_as_uint(wgt);
        const int l = level_of(c.t, a);
        const int cell = a - c.t.start[l];
        const float st = c.t.stride[l];
        const float ax = (float)(cell % c.t.w[l]) + 0.5f, ay = (float)(cell / c.t.w[l]) + 0.5f;
        const GtRec g = gts[gi];
        const int lab = g.label < 0 ? 0 : g.label;
        const float *hp = c.t.ptr[l] + (long long)b * c.t.sB[l] + cell;
        const long long cs = c.t.sC[l];
        // target_bboxes /= stride_tensor (loss.py:248)
        const float4 tb = make_float4(dm::div(g.box.x, st), dm::div(g.box.y, st), dm::div(g.box.z, st), dm::div(g.box.w, st));
        const float ltrb[4] = {ax - tb.x, ay - tb.y, tb.z - ax, tb.w - ay};  // bbox2dist tal.py:328-331
        const float4 pb4 = __ldcg(reinterpret_cast<const float4 *>(c.pd_bboxes) + (long long)b * A + a);
        const float4 ls4 = __ldcg(reinterpret_cast<const float4 *>(F.lse[z]) + (long long)b * A + a);
        const float ls[4] = {ls4.x, ls4.y, ls4.z, ls4.w};
        float xl[4], xr[4], wl[4];
#pragma unroll
        for (int side = 0; side < 4; ++side) {  // all gathers first
            const float tt = fminf(fmaxf(ltrb[side], 0.0f), (float)(kR - 1) - 0.01f);
            const int tl = (int)tt;
            wl[side] = (float)(tl + 1) - tt;
            xl[side] = hp[(long long)(side * kR + tl) * cs];
            xr[side] = hp[(long long)(side * kR + tl + 1) * cs];
        }
        const float xlab = hp[(long long)(4 * kR + lab) * cs];
        const float4 pb = pb4;
        const float iou = ciou_fast(pb, tb);  // BboxLoss.forward loss.py:85 (box1 = pred)
        float dfl = 0.f;
#pragma unroll
        for (int side = 0; side < 4; ++side)  // _df_loss loss.py:99-113
            dfl += (ls[side] - xl[side]) * wl[side] + (ls[side] - xr[side]) * (1.0f - wl[side]);
        s_iou += to_fix((1.0f - iou) * wgt);
        s_dfl += to_fix(dfl * 0.25f * wgt);  // .mean(-1) over the 4 sides
        s_ts += to_fix(wgt);
        s_xt += to_fix(xlab * wgt);  // BCE(x,t) - BCE(x,0) = -x*t
        if (F.dbg_fg[z]) {
            F.dbg_fg[z][(long long)b * A + a] = 1;
            F.dbg_gi[z][(long long)b * A + a] = gi;
        }
    }
    Y3D_STAMP(3);
    // ---- per-image partials: exact integer sums of the foreground terms + this image's BCE partials (fixed order)
    s_iou = warp_sum_ll(s_iou); s_dfl = warp_sum_ll(s_dfl); s_ts = warp_sum_ll(s_ts); s_xt = warp_sum_ll(s_xt);
    double bce = 0.0;
    for (int i = tid; i < F.n_bce_x; i += kFinishThreads)
        bce += __ldcg(F.part_bce + ((long long)z * gridDim.x + b) * F.n_bce_x + i);
    bce = warp_sum(bce);
    if (lane == 0) {
        S.redl[0][wid] = s_iou; S.redl[1][wid] = s_dfl; S.redl[2][wid] = s_ts; S.redl[3][wid] = s_xt;
        S.redd[wid] = bce;
    }
    __syncthreads();
    double *pimg = F.part_fg + ((long long)z * gridDim.x + b) * 5;
    if (tid < 4) {
        long long s = 0;
        for (int i = 0; i < kFinishWarps; ++i) s += S.redl[tid][i];
        pimg[tid] = (double)s / kFix;
    } else if (tid == 4) {
        double s = 0.0;
        for (int i = 0; i < kFinishWarps; ++i) s += S.redd[i];
        pimg[4] = s;
    }
    // last CTA done: fixed-order reduction of the per-image partials (deterministic whichever CTA it is)
    __threadfence();
    __syncthreads();
    if (tid == 0) S.ticket = atomicAdd(F.counter, 1u);
    __syncthreads();
    Y3D_STAMP(4);
    if (S.ticket != gridDim.x * gridDim.y - 1) return;
    __threadfence();
    const int B = gridDim.x;
    if (wid < 5 * F.n_branch) {  // warp (zz, k): sum over the images, lane-strided then a fixed shuffle tree
        const int zz = wid / 5, k = wid % 5;
        double acc = 0.0;
        for (int i = lane; i < B; i += 32) acc += __ldcg(F.part_fg + ((long long)zz * B + i) * 5 + k);
        acc = warp_sum(acc);
        if (lane == 0) S.fin[zz][k] = acc;
    }
    __syncthreads();
    if (tid < 4 * F.n_branch) {  // partials of a branch: iou, bce, dfl, target_scores_sum
        const int zz = tid >> 2, j = tid & 3;
        S.xv[tid] = j == 0 ? S.fin[zz][0] : j == 1 ? S.fin[zz][4] - S.fin[zz][3] : j == 2 ? S.fin[zz][1] : S.fin[zz][2];
    }
    __syncthreads();
    const double *tot = S.xv;
    if (F.x_world > 1 && F.x_defer) {  // post only: the collecting kernel runs later, on the caller's side stream
        xrank_post(F.x_bufs, F.x_rank, F.x_world, F.x_seq, S.xv, 4 * F.n_branch);
        return;
    }
    if (F.x_world > 1) {  // sum over the ranks: stores into every peer's buffer, flag, wait, rank-ordered sum
        xrank_allreduce(F.x_bufs, F.x_rank, F.x_world, F.x_seq, S.xv, 4 * F.n_branch, S.xs, &S.xfail, F.x_timeout);
        tot = S.xs;
        if (tid == 0 && F.x_status) *F.x_status = S.xfail;
    }
    if (tid < 4 * F.n_branch && F.partials) F.partials[tid] = tot[tid];
    if (tid < F.n_branch && F.normalise && F.loss_items) {
        const int zz = tid;
        const double tss = tot[4 * zz + 3] > 1.0 ? tot[4 * zz + 3] : 1.0;  // max(target_scores.sum(), 1) loss.py:240
        F.loss_items[4 * zz + 0] = (float)(tot[4 * zz + 0] / tss * F.gain_box);
        F.loss_items[4 * zz + 1] = (float)(tot[4 * zz + 1] / tss * F.gain_cls);
        F.loss_items[4 * zz + 2] = (float)(tot[4 * zz + 2] / tss * F.gain_dfl);
        F.loss_items[4 * zz + 3] = (float)tss;
    }
    if (tid == 0 && F.normalise && F.loss_total) {  // loss.sum() * batch_size per branch, added up (loss.py:257, 736)
        float total = 0.f;
        for (int zz = 0; zz < F.n_branch; ++zz) {
            const double tss = tot[4 * zz + 3] > 1.0 ? tot[4 * zz + 3] : 1.0;
            const float i0 = (float)(tot[4 * zz + 0] / tss * F.gain_box), i1 = (float)(tot[4 * zz + 1] / tss * F.gain_cls),
                        i2 = (float)(tot[4 * zz + 2] / tss * F.gain_dfl);
            total += ((i0 + i1) + i2) * F.total_scale;
        }
        *F.loss_total = total;
    }
    Y3D_STAMP(5);
}
#ifdef Y3D_TIMING
extern "C" int y3d_debug_read_stamps(unsigned long long *host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_y3d_stamps, sizeof(unsigned long long) * n);
}
#endif

// ---------------------------------------------------------------------------------------------- anchor-parallel finish
// The per-image kernel above keeps a whole image on one SM: its ~10 scattered gathers per claimed anchor (two DFL bins
// per side, box, log-sum-exp, label logit) queue up behind one SM's load pipe (16 us for an image with 1000 claimed
// anchors, the critical path of the step).  For M <= kFinishApMaxM the claimed anchors of an image are therefore
// spread over the machine: CTA (chunk, image, branch) takes kFinApThreads consecutive list entries, one per thread.
//   phase R (every chunk): conflict resolution (select_highest_overlaps), the pair's exact alignment metric / overlap
//       folded into the image's per-GT maxima (global atomicMax), and everything of the loss terms that does not need
//       those maxima: 1 - CIoU, the DFL cross-entropy and the label logit -> one 20-byte record per anchor.
//   phase S (the chunk that finishes last for its image, found with a counter -- nobody ever waits): weight =
//       align * max overlap / (max align + eps) per record (tal.py:89-92), exact fixed-point sums, the claim word the
//       backward pass reads, the image's BCE partials; the globally last image then reduces all images in a fixed order,
//       exchanges with the peer ranks (sharded batch) and normalises.
constexpr int kFinApThreads = 128;
constexpr int kFinApWarps = kFinApThreads / 32;
constexpr int kFinApPairs = 1024;
#ifdef Y3D_TIMING
#define Y3D_APSTAMP(i)                                                                                              \
    do {                                                                                                            \
        if (threadIdx.x == 0)                                                                                       \
            g_y3d_stamps[(((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) & 2047) * 8 + (i)] = gtimer(); \
    } while (0)
#else
#define Y3D_APSTAMP(i)
#endif

__global__ void __launch_bounds__(kFinApThreads) loss_finish_ap_kernel(AssignCtx2 cc, FinishParams F) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GtRec *gts = reinterpret_cast<GtRec *>(smem_raw);
    __shared__ float2 s_cxy[kFinApThreads];                // contested anchors of the chunk: anchor point (px)
    __shared__ float4 s_cbox[kFinApThreads];               // predicted box (px)
    __shared__ unsigned long long s_cbest[kFinApThreads];  // (overlap bits << 32) | (0xffffffff - m): atomicMax = first maximum
    __shared__ int2 s_pairs[kFinApPairs];                  // (contested anchor, GT) with the anchor inside the GT
    __shared__ int s_ncf, s_np;
    __shared__ int s_pos[2 * kFinishApMaxM];               // single-chunk images: per-GT max alignment / max overlap (float bits)
    __shared__ long long s_redl[4][kFinApWarps];
    __shared__ double s_fin[2][5];
    __shared__ double s_xv[kXMaxVals], s_xs[kXMaxVals];
    __shared__ int s_xfail, s_flag;
    const int z = blockIdx.z, b = blockIdx.y, chunk = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const AssignCtx &c = cc.c[z];
    const int M = c.M, A = c.A;
    Y3D_APSTAMP(0);
    Y3D_TL_MIN(2);
    if (tid == 0) { s_ncf = 0; s_np = 0; }
    s_pos[tid] = 0;
    s_pos[tid + kFinApThreads] = 0;
    // Without GTs there is no top-k grid in between (whose own wait orders this grid behind the streaming kernel): the
    // streaming grid itself is then the programmatic primary, and its softplus partials and zeroed counters are read below.
    if (M == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
    bool my_valid = false;
    for (int m = tid; m < M; m += kFinApThreads) {
        gts[m] = load_gt(c, b, m);
        my_valid = gts[m].valid;  // M <= kFinishApMaxM == blockDim.x: one GT per thread
    }
    const int n_valid = __syncthreads_count(my_valid);
    if (chunk > 0 && chunk * kFinApThreads >= n_valid * c.k) return;  // more records than this image can have
    if (chunk == 0 && wid == 0) {  // off the critical path: this image's softplus partials (written by the stream kernel)
        double bsum = 0.0;
        for (int i = lane; i < F.n_bce_x; i += 32)
            bsum += __ldcg(F.part_bce + ((long long)z * gridDim.y + b) * F.n_bce_x + i);
        bsum = warp_sum(bsum);
        if (lane == 0) __stcg(F.img_bce + (long long)z * gridDim.y + b, bsum);
    }
    // Everything above reads only the caller's inputs.  This grid is launched programmatically dependent on the top-k
    // kernel and becomes resident as that kernel's CTAs retire; it does not wait for the whole grid but for its own
    // image: the top-k kernel counts the GTs it has finished per image (release), all of them = the records are there.
    if (tid == 0 && n_valid > 0) {
        const unsigned *done = c.topk_done + b;
        while (ld_acquire_u32(done) < (unsigned)n_valid) __nanosleep(200);
    }
    __syncthreads();
    const int n = M > 0 ? min(__ldcg(c.list_count + b), c.rec_cap) : 0;
    const int n_chunks = max(1, (n + kFinApThreads - 1) / kFinApThreads);
    Y3D_APSTAMP(1);
    Y3D_TL_MIN(3);
    if (chunk >= n_chunks) return;
    __syncthreads();
    float4 *rec = c.rec + (long long)b * c.rec_cap * kRecF4;
    int *pos = F.pos[z] + (long long)b * M * 2;
    // An image whose records fit one chunk (every top-k 1 image, the small top-k 10 ones) never leaves this CTA: its per-GT
    // maxima live in shared memory and phase S runs from registers -- no record rewrite, no fence, no counter.
    const bool single = n_chunks == 1;
    bool my_active = false;
    int my_a = 0, my_gi = 0;
    float my_alv = 0.f, my_xlab = 0.f, my_tiou = 0.f, my_tdfl = 0.f;
    // ---- phase R: one claim record per thread; only the first claimer of an anchor carries it on
    {
        const int e = chunk * kFinApThreads + tid;
        bool active = e < n;
        int a = 0, cnt = 0, gi = 0, m0 = 0;
        float ax = 0.f, ay = 0.f, st = 1.f, metric = 0.f, xlab = 0.f;
        PairRaw raw;
        raw.box = make_float4(0.f, 0.f, 0.f, 0.f);
        raw.s = 0.f;
        float4 terms = raw.box;
        if (active) {
            const float4 r0 = __ldcg(rec + (long long)e * kRecF4);
            terms = __ldcg(rec + (long long)e * kRecF4 + 1);
            raw.box = __ldcg(rec + (long long)e * kRecF4 + 2);
            const int aw = __float_as_int(r0.x);
            active = aw < 0;  // first-claimer bit
            a = aw & 0x7fffffff;
            m0 = __float_as_int(r0.y);
            metric = r0.z;
            xlab = r0.w;
        }
        if (active) {
            const unsigned long long cl = __ldcg(c.claim + (long long)b * A + a);
            cnt = (int)(cl >> 32);
            gi = m0;
            anchor_px(c, a, ax, ay, st);
        }
        const float4 pbox = pair_box(c, raw, a);
        // multiply-claimed anchors (select_highest_overlaps tal.py:237-264: argmax over ALL GTs of the overlap, first
        // maximum), resolved by the whole CTA: thread m tests GT m against every contested anchor of the chunk, the
        // (anchor, GT) pairs with the anchor inside the GT are collected, and their CIoU is evaluated one pair per thread
        const bool contested = active && cnt > 1;
        int my_ci = -1;
        if (contested) {
            my_ci = atomicAdd(&s_ncf, 1);
            s_cxy[my_ci] = make_float2(ax, ay);
            s_cbox[my_ci] = pbox;
            s_cbest[my_ci] = 0xffffffffull;  // overlap 0 at GT 0: what the reference's argmax of zeros gives
        }
        __syncthreads();
        const int ncf = s_ncf;
        if (ncf > 0) {
            auto eval_pair = [&](int ci, int m) {
                const GtRec g = gts[m];
                const float ovl = dm::ciou(g.box, s_cbox[ci], g.at1);
                if (ovl > 0.0f)
                    atomicMax(&s_cbest[ci],
                              ((unsigned long long)__float_as_uint(ovl) << 32) | (unsigned long long)(0xffffffffu - (unsigned)m));
            };
            for (int m = tid; m < M; m += kFinApThreads) {
                const GtRec g = gts[m];
                if (!g.valid) continue;
                for (int ci = 0; ci < ncf; ++ci) {
                    const float2 xy = s_cxy[ci];
                    if (dm::in_gt(xy.x, xy.y, g.box)) {
                        const int q = atomicAdd(&s_np, 1);
                        if (q < kFinApPairs) s_pairs[q] = make_int2(ci, m);
                        else eval_pair(ci, m);  // pair buffer full (every GT contains every contested anchor): in place
                    }
                }
            }
            __syncthreads();
            const int np = min(s_np, kFinApPairs);
            for (int q = tid; q < np; q += kFinApThreads) eval_pair(s_pairs[q].x, s_pairs[q].y);
            __syncthreads();
            if (contested) gi = (int)(0xffffffffu - (unsigned)(s_cbest[my_ci] & 0xffffffffull));
        }
        if (active) {
            float ovl = terms.x;
            if (gi != m0) {
                // the anchor went to a GT other than its first claimer (possibly one that never selected it): the terms of
                // this pair were not evaluated by the top-k kernel
                const GtRec g = gts[gi];
                const ClaimTerms ct = claim_terms(c, b, a, g);
                terms = ct.terms;
                xlab = ct.xlab;
                const float xs = g.label < 0 ? pair_load_score(c, b, a, g.label) : xlab;  // the score the assigner saw
                metric = 0.0f;
                ovl = 0.0f;
                if (g.valid && dm::in_gt(ax, ay, g.box))
                    metric = pair_metric(c, b, gi, g, a, raw, dm::pow_(pair_score(c, xs), c.alpha), ovl);
            }
            int *pm = single ? s_pos : pos;  // values >= 0: int order == float order
            atomicMax(pm + 2 * gi, __float_as_int(metric));
            atomicMax(pm + 2 * gi + 1, __float_as_int(ovl));
            my_a = a; my_gi = gi; my_alv = metric; my_xlab = xlab; my_tiou = terms.y; my_tdfl = terms.z;
            if (!single) {  // phase S (whichever chunk runs it) reads words 0 and 1 of the record
                __stcg(rec + (long long)e * kRecF4, make_float4(__int_as_float(a | (int)0x80000000), __int_as_float(gi), metric, xlab));
                __stcg(rec + (long long)e * kRecF4 + 1, make_float4(my_tiou, my_tdfl, 0.f, 0.f));
            }
        }
        my_active = active;
    }
    long long s_iou = 0, s_dfl = 0, s_ts = 0, s_xt = 0;
    auto fold = [&](int a, int gi, float alv, float xlab, float tiou, float tdfl, float pa, float po) {
        const float wgt = dm::div(dm::mul(alv, po), dm::add(pa, c.eps));  // = target_scores.sum(-1), tal.py:89-92
        // kept for the backward pass: the claim word of a foreground anchor becomes (1 << 63 | GT index << 32 | weight)
        c.claim[(long long)b * A + a] =
            0x8000000000000000ull | ((unsigned long long)(unsigned)gi << 32) | (unsigned long long)__float_as_uint(wgt);
        s_iou += to_fix(tiou * wgt);
        s_dfl += to_fix(tdfl * wgt);  // .mean(-1) over the 4 sides is in the term
        s_ts += to_fix(wgt);
        s_xt += to_fix(xlab * wgt);  // BCE(x,t) - BCE(x,0) = -x*t
        if (F.dbg_fg[z]) {
            F.dbg_fg[z][(long long)b * A + a] = 1;
            F.dbg_gi[z][(long long)b * A + a] = gi;
        }
    };
    __syncthreads();
    Y3D_APSTAMP(2);
    Y3D_TL_MAX(4);
    if (single) {
        // ---- phase S from registers
        if (my_active) fold(my_a, my_gi, my_alv, my_xlab, my_tiou, my_tdfl, __int_as_float(s_pos[2 * my_gi]), __int_as_float(s_pos[2 * my_gi + 1]));
        Y3D_APSTAMP(3);
    } else {
        // ---- which chunk of the image finishes last?
        __threadfence();
        __syncthreads();
        if (tid == 0) s_flag = atomicAdd(F.img_cnt[z] + b, 1) == n_chunks - 1;
        __syncthreads();
        Y3D_APSTAMP(3);
        if (!s_flag) return;
        __threadfence();
        // ---- phase S: weights and sums of the whole image
        constexpr int SU = 8;  // records in flight per thread: the loop is two dependent round trips per batch
        for (int e0 = tid; e0 < n; e0 += kFinApThreads * SU) {
            float4 r0[SU], r1[SU];
            int2 pp[SU];
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                const int e = e0 + u * kFinApThreads;
                r0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                r1[u] = r0[u];
                if (e < n) {
                    r0[u] = __ldcg(rec + (long long)e * kRecF4);
                    r1[u] = __ldcg(rec + (long long)e * kRecF4 + 1);
                }
            }
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                pp[u] = make_int2(0, 0);
                if (__float_as_int(r0[u].x) < 0) pp[u] = __ldcg(reinterpret_cast<const int2 *>(pos) + __float_as_int(r0[u].y));
            }
#pragma unroll
            for (int u = 0; u < SU; ++u) {
                const int aw = __float_as_int(r0[u].x);
                if (aw >= 0) continue;  // past the end, or not the anchor's first claimer
                fold(aw & 0x7fffffff, __float_as_int(r0[u].y), r0[u].z, r0[u].w, r1[u].x, r1[u].y, __int_as_float(pp[u].x),
                     __int_as_float(pp[u].y));
            }
        }
    }
    // ---- the image's sums join the batch sums: 64-bit fixed point, so the order of arrival does not matter
    s_iou = warp_sum_ll(s_iou); s_dfl = warp_sum_ll(s_dfl); s_ts = warp_sum_ll(s_ts); s_xt = warp_sum_ll(s_xt);
    if (lane == 0) { s_redl[0][wid] = s_iou; s_redl[1][wid] = s_dfl; s_redl[2][wid] = s_ts; s_redl[3][wid] = s_xt; }
    __syncthreads();
    const int B = gridDim.y;
    if (tid < 4) {
        long long sum = 0;
        for (int i = 0; i < kFinApWarps; ++i) sum += s_redl[tid][i];
        atomicAdd(reinterpret_cast<unsigned long long *>(F.acc + 5 * z + tid), (unsigned long long)sum);
    } else if (tid == 4) {
        const double bce = __ldcg(F.img_bce + (long long)z * B + b);  // summed by the image's first chunk before it waited
        atomicAdd(reinterpret_cast<unsigned long long *>(F.acc + 5 * z + 4), (unsigned long long)__double2ll_rn(bce * kFix));
    }
    // last image done?
    Y3D_APSTAMP(4);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_flag = atomicAdd(F.counter, 1u) == (unsigned)(B * gridDim.z) - 1u;
    __syncthreads();
    Y3D_APSTAMP(5);
    if (!s_flag) return;
    __threadfence();
    if (tid < 5 * F.n_branch) s_fin[tid / 5][tid % 5] = (double)__ldcg(F.acc + tid) / kFix;
    __syncthreads();
    if (tid < 4 * F.n_branch) {  // partials of a branch: iou, bce, dfl, target_scores_sum
        const int zz = tid >> 2, j = tid & 3;
        s_xv[tid] = j == 0 ? s_fin[zz][0] : j == 1 ? s_fin[zz][4] - s_fin[zz][3] : j == 2 ? s_fin[zz][1] : s_fin[zz][2];
    }
    __syncthreads();
    const double *tot = s_xv;
    if (F.x_world > 1 && F.x_defer) {  // post only: the collecting kernel runs later, on the caller's side stream
        xrank_post(F.x_bufs, F.x_rank, F.x_world, F.x_seq, s_xv, 4 * F.n_branch);
        return;
    }
    if (F.x_world > 1) {  // sum over the ranks: stores into every peer's buffer, wait, rank-ordered sum
        xrank_allreduce(F.x_bufs, F.x_rank, F.x_world, F.x_seq, s_xv, 4 * F.n_branch, s_xs, &s_xfail, F.x_timeout);
        tot = s_xs;
        if (tid == 0 && F.x_status) *F.x_status = s_xfail;
    }
    if (tid < 4 * F.n_branch && F.partials) F.partials[tid] = tot[tid];
    if (tid < F.n_branch && F.normalise && F.loss_items) {
        const int zz = tid;
        const double tss = tot[4 * zz + 3] > 1.0 ? tot[4 * zz + 3] : 1.0;  // max(target_scores.sum(), 1) loss.py:240
        F.loss_items[4 * zz + 0] = (float)(tot[4 * zz + 0] / tss * F.gain_box);
        F.loss_items[4 * zz + 1] = (float)(tot[4 * zz + 1] / tss * F.gain_cls);
        F.loss_items[4 * zz + 2] = (float)(tot[4 * zz + 2] / tss * F.gain_dfl);
        F.loss_items[4 * zz + 3] = (float)tss;
    }
    if (tid == 0 && F.normalise && F.loss_total) {  // loss.sum() * batch_size per branch, added up (loss.py:257, 736)
        float total = 0.f;
        for (int zz = 0; zz < F.n_branch; ++zz) {
            const double tss = tot[4 * zz + 3] > 1.0 ? tot[4 * zz + 3] : 1.0;
            const float i0 = (float)(tot[4 * zz + 0] / tss * F.gain_box), i1 = (float)(tot[4 * zz + 1] / tss * F.gain_cls),
                        i2 = (float)(tot[4 * zz + 2] / tss * F.gain_dfl);
            total += ((i0 + i1) + i2) * F.total_scale;
        }
        *F.loss_total = total;
    }
    Y3D_APSTAMP(6);
    Y3D_TL_MAX(5);
    Y3D_TL_STEP();
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
// 128-bit path: every level holds a multiple of 4 cells and all row starts are 16-byte aligned
static bool vec4_ok(const LevelTable &t) {
    for (int l = 0; l < t.nl; ++l) {
        if ((t.h[l] * t.w[l]) % 4) return false;
        if (((uintptr_t)t.ptr[l]) % 16) return false;
        if (t.sB[l] % 4 || t.sC[l] % 4) return false;
    }
    return true;
}

size_t loss_workspace_bytes(int B, int A, int M, int k) { return loss_ws_layout(2, B, A, M, k > 0 ? k : Y3D_MAX_TOPK).total; }

// returns the number of BCE partials per branch through *n_bce
static int launch_stream(const StreamParams &P, int *n_bce, cudaStream_t s) {
    const bool v4 = vec4_ok(P.t[0]) && (P.n_branch < 2 || vec4_ok(P.t[1]));
    const int units = v4 ? P.A / 4 : P.A;
    dim3 grid((units + 31) / 32, P.B, P.n_branch);
    *n_bce = (int)(grid.x * grid.y);
    const bool ps = P.pd_scores[0] != nullptr;
    if (v4 && !ps)
        head_stream_kernel<4, false><<<grid, kStreamThreads, 0, s>>>(P);
    else if (v4)
        head_stream_kernel<4, true><<<grid, kStreamThreads, 0, s>>>(P);
    else if (!ps)
        head_stream_kernel<1, false><<<grid, kStreamThreads, 0, s>>>(P);
    else
        head_stream_kernel<1, true><<<grid, kStreamThreads, 0, s>>>(P);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

struct BranchIn {
    const float *const *lvl_ptr;
    const int64_t *sB, *sC;
    int topk;
};
struct XRankIn {  // cross-rank exchange of the fused loss (world <= 1: none)
    void *const *bufs_dev;
    int rank, world;
    unsigned long long seq;
    int *status;
    int defer;
};

static int loss_run(int nb, const BranchIn *br, const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc,
                    int reg_max, const float *gt, int M, float gain_box, float gain_cls, float gain_dfl, int normalise,
                    float *loss_items, double *partials, uint8_t *dbg_fg_mask, int32_t *dbg_target_gt_idx,
                    void *const *prof_events, void *ws, size_t ws_bytes, void *stream, const XRankIn *xr = nullptr,
                    float total_scale = 0.f, float *loss_total = nullptr) {
    if (!lvl_hw || !lvl_stride || B < 1 || nc < 1 || M < 0 || (M > 0 && !gt)) return Y3D_EINVAL;
    if (xr && xr->world > 1 &&
        (!xr->bufs_dev || xr->world > kXMaxWorld || xr->rank < 0 || xr->rank >= xr->world || xr->seq == 0))
        return Y3D_EINVAL;
    if (!loss_items && !partials && !(xr && xr->world > 1 && xr->defer)) return Y3D_EINVAL;
    if ((dbg_fg_mask == nullptr) != (dbg_target_gt_idx == nullptr)) return Y3D_EINVAL;
    if (reg_max != kR) return Y3D_EUNSUPPORTED;
    AssignCtx2 cc{};
    StreamParams P{};
    int A = 0, kmax = 1;
    for (int z = 0; z < nb; ++z) {
        if (!br[z].lvl_ptr || !br[z].sB || !br[z].sC) return Y3D_EINVAL;
        A = make_level_table(cc.c[z].t, br[z].lvl_ptr, br[z].sB, br[z].sC, lvl_hw, lvl_stride, nl);
        if (A < 0) return A;
        for (int l = 0; l < nl; ++l)
            if (!br[z].lvl_ptr[l]) return Y3D_EINVAL;
        if (br[z].topk < 1 || br[z].topk > A) return Y3D_EINVAL;
        if (br[z].topk > Y3D_MAX_TOPK) return Y3D_EUNSUPPORTED;
        if (br[z].topk > kmax) kmax = br[z].topk;
        P.t[z] = cc.c[z].t;
    }
    const LossWs w = loss_ws_layout(nb, B, A, M, kmax);
    if (!ws || ws_bytes < w.total) return Y3D_EWORKSPACE;
    if (((uintptr_t)ws) % 256) return Y3D_EALIGN;
    const size_t A0 = (size_t)A;
    const bool use_ap = M <= kFinishApMaxM;
    const size_t fin_smem = finish_smem_bytes(M);
    if (!use_ap && fin_smem > 212 * 1024) return Y3D_EUNSUPPORTED;  // + 8 KB static (exchange scratch) <= 227 KB
    cudaStream_t s = (cudaStream_t)stream;
    if (use_ap && dbg_fg_mask) {  // the anchor-parallel kernel writes the foreground entries only
        cudaError_t e1 = cudaMemsetAsync(dbg_fg_mask, 0, (size_t)nb * B * A0, s);
        cudaError_t e2 = cudaMemsetAsync(dbg_target_gt_idx, 0, sizeof(int32_t) * (size_t)nb * B * A0, s);
        if (e1 != cudaSuccess || e2 != cudaSuccess) return (int)(e1 != cudaSuccess ? e1 : e2);
    }
    char *p = (char *)ws;
    auto mark = [&](int i) {
        if (prof_events && prof_events[i]) cudaEventRecord((cudaEvent_t)prof_events[i], s);
    };
    mark(0);
    P.n_branch = nb; P.B = B; P.nc = nc; P.A = A;
    P.part_bce = (double *)(p + w.off_pbce);
    P.counter = (unsigned *)(p + w.off_counter);
    FinishParams F{};
    for (int z = 0; z < nb; ++z) {
        char *q = p + z * w.per_branch;
        AssignCtx &c = cc.c[z];
        c.claim = (unsigned long long *)(q + w.claim);
        c.list_count = (int *)(q + w.list_count);
        c.list_a = (int *)(q + w.list_a);
        c.list_cap = w.cap;
        if (M <= kFinishApMaxM) {  // anchor-parallel finish: claim records instead of the list
            c.list_a = nullptr;
            c.rec = (float4 *)(q + w.rec);
            c.rec_cap = w.rcap;
            c.lse = (const float *)(q + w.lse);
            c.topk_done = (unsigned *)(q + w.topk_done);
            P.topk_done[z] = c.topk_done;
        }
        float *boxes = (float *)(q + w.boxes);
        P.boxes[z] = boxes;
        P.lse[z] = (float *)(q + w.lse);
        P.pd_scores[z] = nullptr;
        P.claim[z] = M > 0 ? c.claim : nullptr;
        P.list_count[z] = c.list_count;
        P.img_cnt[z] = (int *)(q + w.img_cnt);
        P.pos[z] = M > 0 ? (int *)(q + w.pos) : nullptr;
        c.score_mode = 1;
        c.cls_ch0 = 4 * kR;
        c.pd_bboxes = boxes; c.box_grid_units = 1; c.box_soa = 0;
        c.use_grid = 1;
        c.gt_labels = gt; c.gl_stride = 5;
        c.gt_bboxes = gt ? gt + 1 : nullptr; c.gb_stride = 5;
        c.mask_gt = nullptr;
        c.B = B; c.A = A; c.nc = nc; c.M = M; c.k = br[z].topk;
        c.alpha = 0.5f; c.beta = 6.0f; c.gamma = 1.0f; c.eps = 1e-9f;  // loss.py:176
        c.use_2d = 1; c.use_3d = 0; c.kps_l2 = 0; c.constrain = 1;
        F.lse[z] = P.lse[z];
        F.list_gi[z] = (int *)(q + w.list_gi);
        F.list_al[z] = (float *)(q + w.list_al);
        F.img_cnt[z] = P.img_cnt[z];
        F.pos[z] = (int *)(q + w.pos);
        F.dbg_fg[z] = dbg_fg_mask ? dbg_fg_mask + (size_t)z * B * A : nullptr;
        F.dbg_gi[z] = dbg_target_gt_idx ? dbg_target_gt_idx + (size_t)z * B * A : nullptr;
    }
    // Longest-first work order for the top-k kernel: pays when a persistent warp gets only a few GTs (cfg2: 1.7 valid
    // GTs per warp -- the kernel's tail was one or two big GTs picked up late); with dozens of GTs per warp (crowds) the
    // dynamic distribution balances by itself and the per-item table lookup would only cost (measured: +6 % at cfg5).
    if (M > 0 && kOrdClasses * nb * B <= kOrdMaxSeg && (long long)nb * B * M <= 8LL * 6 * kTopkWarps * device_sm_count()) {
        P.gt5 = gt; P.M = M;
        P.ord_cnt = (int *)(p + w.off_ord_cnt);
        P.ord_list = (int *)(p + w.off_ord_list);
        P.ord_cells_per_px2 = 0.0f;
        for (int l = 0; l < nl; ++l) P.ord_cells_per_px2 += 1.0f / (lvl_stride[l] * lvl_stride[l]);
        cc.ord_cnt = P.ord_cnt;
        cc.ord_list = P.ord_list;
    }
    int n_bce = 0;
    int rc = launch_stream(P, &n_bce, s);
    if (rc) return rc;
    mark(1);
    cc.work_counter = (int *)(P.counter + 1);
    if (M > 0) {
        rc = assign_run_topk_fused(cc, nb, s, /*pdl=*/true);  // scheduled while the streaming kernel drains
        if (rc == Y3D_EUNSUPPORTED) rc = assign_run_topk(cc, nb, s, /*pdl=*/true);  // few GTs: several warps per GT
        if (rc) return rc;
    }
    mark(2);
    F.gt5 = gt;
    F.part_bce = P.part_bce;
    F.part_fg = (double *)(p + w.off_pfg);
    F.acc = (long long *)(P.counter + 16);
    F.img_bce = F.part_fg;  // the anchor-parallel kernel does not use the per-image partials: [n_branch][B] doubles fit
    F.counter = P.counter;
    F.partials = partials;
    F.loss_items = loss_items;
    F.loss_total = loss_total;
    F.total_scale = total_scale;
    F.n_bce_x = n_bce / B; F.n_branch = nb; F.normalise = normalise;
    F.gain_box = gain_box; F.gain_cls = gain_cls; F.gain_dfl = gain_dfl;
    if (xr && xr->world > 1) {
        F.x_bufs = (XSlot *const *)xr->bufs_dev;
        F.x_rank = xr->rank; F.x_world = xr->world; F.x_seq = xr->seq; F.x_status = xr->status;
        F.x_timeout = xrank_timeout_cycles();
        F.x_defer = xr->defer;
    }
    // programmatic dependent launch: the prologue (GT records) overlaps the top-k kernel's tail
    cudaLaunchConfig_t cfg = {};
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (use_ap) {  // claimed anchors spread over the machine, kFinApThreads per CTA
        cfg.gridDim = dim3((unsigned)((w.rcap + kFinApThreads - 1) / kFinApThreads), B, nb);
        cfg.blockDim = dim3(kFinApThreads);
        cfg.dynamicSmemBytes = sizeof(GtRec) * (size_t)M;
        cudaError_t le = cudaLaunchKernelEx(&cfg, loss_finish_ap_kernel, cc, F);
        if (le != cudaSuccess) return (int)le;
    } else {  // dense crowds: one CTA per (image, branch) with a spatial index of the GT boxes
        static bool attr_set = false;  // once per process: the largest size the kernel is ever launched with
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(loss_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
            if (e != cudaSuccess) return (int)e;
            attr_set = true;
        }
        cfg.gridDim = dim3(B, nb);
        cfg.blockDim = dim3(kFinishThreads);
        cfg.dynamicSmemBytes = fin_smem;
        cudaError_t le = cudaLaunchKernelEx(&cfg, loss_finish_kernel, cc, F);
        if (le != cudaSuccess) return (int)le;
    }
    Y3D_CHECK_LAUNCH();
    mark(3);
    return Y3D_OK;
}

}  // namespace y3d

using namespace y3d;

extern "C" int y3d_train_decode(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC,
                                const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                                float *pd_bboxes, float *pd_scores, void *stream) {
    if (!lvl_ptr || !lvl_sB || !lvl_sC || !pd_bboxes || B < 0 || nc < 1) return Y3D_EINVAL;
    if (reg_max != kR) return Y3D_EUNSUPPORTED;
    if (((uintptr_t)pd_bboxes) % 16) return Y3D_EALIGN;  // one 16-byte store per anchor
    StreamParams P{};
    int A = make_level_table(P.t[0], lvl_ptr, lvl_sB, lvl_sC, lvl_hw, lvl_stride, nl);
    if (A < 0) return A;
    for (int l = 0; l < nl; ++l)
        if (!lvl_ptr[l]) return Y3D_EINVAL;
    if (B == 0) return Y3D_OK;
    P.t[1] = P.t[0];
    P.n_branch = 1; P.B = B; P.nc = nc; P.A = A;
    P.boxes[0] = pd_bboxes;  // the reference layout [B,A,4]
    P.pd_scores[0] = pd_scores;
    int n_bce;
    return launch_stream(P, &n_bce, (cudaStream_t)stream);
}

extern "C" int y3d_v8_loss_fwd(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC,
                               const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                               const float *gt, int M, int topk, float gain_box, float gain_cls, float gain_dfl,
                               int normalise, float *loss_items, double *partials, float total_scale,
                               float *loss_total, uint8_t *dbg_fg_mask, int32_t *dbg_target_gt_idx,
                               void *const *prof_events, void *ws, size_t ws_bytes, void *stream) {
    BranchIn br[1] = {{lvl_ptr, lvl_sB, lvl_sC, topk}};
    return loss_run(1, br, lvl_hw, lvl_stride, nl, B, nc, reg_max, gt, M, gain_box, gain_cls, gain_dfl, normalise,
                    loss_items, partials, dbg_fg_mask, dbg_target_gt_idx, prof_events, ws, ws_bytes, stream, nullptr,
                    total_scale, loss_total);
}

extern "C" int y3d_v10_loss_fwd(const float *const *o2m_ptr, const int64_t *o2m_sB, const int64_t *o2m_sC,
                                const float *const *o2o_ptr, const int64_t *o2o_sB, const int64_t *o2o_sC,
                                const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                                const float *gt, int M, int topk_o2m, int topk_o2o, float gain_box, float gain_cls,
                                float gain_dfl, int normalise, float *loss_items, double *partials,
                                float total_scale, float *loss_total, uint8_t *dbg_fg_mask,
                                int32_t *dbg_target_gt_idx, void *const *prof_events, void *ws, size_t ws_bytes,
                                void *stream) {
    BranchIn br[2] = {{o2m_ptr, o2m_sB, o2m_sC, topk_o2m}, {o2o_ptr, o2o_sB, o2o_sC, topk_o2o}};
    return loss_run(2, br, lvl_hw, lvl_stride, nl, B, nc, reg_max, gt, M, gain_box, gain_cls, gain_dfl, normalise,
                    loss_items, partials, dbg_fg_mask, dbg_target_gt_idx, prof_events, ws, ws_bytes, stream, nullptr,
                    total_scale, loss_total);
}

extern "C" int y3d_v10_loss_fwd_sharded(const float *const *o2m_ptr, const int64_t *o2m_sB, const int64_t *o2m_sC,
                                        const float *const *o2o_ptr, const int64_t *o2o_sB, const int64_t *o2o_sC,
                                        const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                                        const float *gt, int M, int topk_o2m, int topk_o2o, float gain_box,
                                        float gain_cls, float gain_dfl, float *loss_items, double *partials,
                                        float total_scale, float *loss_total, int rank, int world,
                                        void *const *peer_bufs_dev, unsigned long long seq, int defer,
                                        int *status, void *const *prof_events, void *ws, size_t ws_bytes,
                                        void *stream) {
    if (!loss_items && !defer) return Y3D_EINVAL;
    BranchIn br[2] = {{o2m_ptr, o2m_sB, o2m_sC, topk_o2m}, {o2o_ptr, o2o_sB, o2o_sC, topk_o2o}};
    XRankIn xr = {peer_bufs_dev, rank, world, seq, status, defer};
    return loss_run(2, br, lvl_hw, lvl_stride, nl, B, nc, reg_max, gt, M, gain_box, gain_cls, gain_dfl, 1, loss_items,
                    partials, nullptr, nullptr, prof_events, ws, ws_bytes, stream, &xr, total_scale, loss_total);
}

extern "C" int y3d_v8_loss_finalize(const double *partials, int n_branch, float gain_box, float gain_cls,
                                    float gain_dfl, float *loss_items, void *stream) {
    if (!partials || !loss_items || n_branch < 1 || n_branch > 2) return Y3D_EINVAL;
    loss_finalize_partials_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partials, n_branch, gain_box, gain_cls, gain_dfl,
                                                                     loss_items);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}
