// loss3d.cu -- fused forward of the fork's 3D detection loss (sm_100a).
//   y3d_dd_loss_fwd   DDDetectionLoss.__call__   reference ultralytics/utils/loss.py:821-900
//                     bbox_decode                reference ultralytics/utils/loss.py:812-819
//                     compute_box2d_loss         reference ultralytics/utils/loss.py:913-926
//                     compute_box3d_loss         reference ultralytics/utils/loss.py:928-963
//                     laplacian_aleatoric_uncertainty_loss_new / compute_heading_loss   loss.py:1112-1136
//   (one branch; DetectLoss3d, loss.py:741-771, calls it for one2one and, in training, one2many)
//
// Head layout [B, nc+35, h, w]: cls(nc) | o2d(2) s2d(2) | o3d(2) s3d(3) hd(24) dep(1) dep_un(1)  (loss.py:825-829).
// Kernels of a call:
//   1. dd_stream_kernel : one pass over the head (4*(nc+35)*A bytes per image).  Per anchor: sigmoid scores -> argmax
//      class -> mean size; 2D box (px); 3D centre / size / heading / depth -> the 8 box corners in camera
//      coordinates (TaskAlignedAssigner3d.forward tal.py:425-447); sum softplus(logit).  Out: boxes [B,A,4] px and
//      keypoints [B,A,24] (the assigner's inputs), zeroed claim words.  pd_scores / pd_3d are never materialised.
//   2. GT keypoints, per-GT top-k, conflict resolution: the assigner core of assign.cu (scores read as logits from the
//      head).
//   3. dd_fg_kernel : the foreground sums (L1 terms, Laplacian depth, heading CE + L1, BCE correction, sum of target
//      scores, foreground count) per CTA; the last CTA of an image reduces the image's rows, the last image reduces
//      the images and writes the partials and the loss items (fixed orders: deterministic whichever CTA is last).
#include "assign.cuh"

namespace y3d {

constexpr int kNSum = 10;  // soft+, off2d, size2d, depth, off3d, size3d, hd_ce, hd_l1, x*t, sum t  (+ n_fg kept apart)

struct DDParams {
    LevelTable t;
    const float *gts;         // [B,M,17] label bbox4(px) c2d2 s2d2 c3d2 s3d3 depth hbin hres
    const float *calibs;      // [B,6]
    const float *mean_sizes;  // [nc,3]
    float *boxes;             // [B,A,4] px
    float *pd_kps;            // [B,A,24]
    unsigned long long *claim;
    double *part;             // [n_blocks][kNSum + 1]
    double *part_img;         // [B][kNSum + 1]
    unsigned *tickets;        // [B + 1], zeroed by the streaming kernel
    double *partials;         // out: float64[kNSum + 1]
    float *items;             // out (optional): float[8]
    float gain[6];
    int B, nc, A, M;
};

struct DDParams2 {
    DDParams p[2];  // one2many / one2one of DetectLoss3d in the same launches (blockIdx.z)
};

// GT keypoints of image b (add_cls_mean_size tal.py:605-609 + get_3d_keypoints): out of line, run by one CTA per image
static __device__ __noinline__ void dd_gt_keypoints(const float *gts, const float *calibs, const float *mean_sizes, int b,
                                                    int M, int nc, float *gt_kps) {
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const float *g = gts + ((long long)b * M + m) * 17;
        int lab = (int)g[0];
        lab = lab < 0 ? 0 : (lab >= nc ? nc - 1 : lab);
        const float sh_ = dm::add(mean_sizes[3 * lab], g[11]), sw_ = dm::add(mean_sizes[3 * lab + 1], g[12]),
                    sl_ = dm::add(mean_sizes[3 * lab + 2], g[13]);
        float out[24];
        keypoints24(g[9], g[10], g[14], sh_, sw_, sl_, (int)g[15], g[16], calibs + 6 * b, out);
        for (int j = 0; j < 24; ++j) gt_kps[((long long)b * M + m) * 24 + j] = out[j];
    }
}

// grid (ceil(A/128), B, n_branch), one thread per anchor.  Besides the pass over the head: zeroes what the assigner core
// expects zeroed (claim words, per-GT maxima, work counter) and, in branch 0, computes the GT keypoints -- no memset and
// no separate small kernel in front of the top-k kernel.
__global__ void __launch_bounds__(128) dd_stream_kernel(const __grid_constant__ DDParams2 PP, int *work_counter,
                                                        int *pos_align0, int *pos_ov0, int *pos_align1, int *pos_ov1,
                                                        float *gt_kps) {
    __shared__ double red[4];
    const DDParams &P = PP.p[blockIdx.z];
    const int b = blockIdx.y, a = blockIdx.x * blockDim.x + threadIdx.x;
    // the next kernel of the call is launched programmatically dependent on this one: scheduled as this grid drains
    asm volatile("griddepcontrol.launch_dependents;");
    if (blockIdx.x == 0 && b == 0 && blockIdx.z == 0 && threadIdx.x == 0 && work_counter) *work_counter = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.tickets[b] = 0u;
        if (b == 0) P.tickets[P.B] = 0u;
    }
    if (blockIdx.x == 0) {
        int *pa = blockIdx.z ? pos_align1 : pos_align0, *po = blockIdx.z ? pos_ov1 : pos_ov0;
        if (pa)
            for (int m = threadIdx.x; m < P.M; m += blockDim.x) {
                pa[(long long)b * P.M + m] = 0;
                po[(long long)b * P.M + m] = 0;
            }
        if (gt_kps && blockIdx.z == 0) dd_gt_keypoints(P.gts, P.calibs, P.mean_sizes, b, P.M, P.nc, gt_kps);
    }
    double soft = 0.0;
    if (a < P.A) {
        const LevelTable &t = P.t;
        const int l = level_of(t, a);
        const int cell = a - t.start[l];
        const float st = t.stride[l];
        const float *x = t.ptr[l] + (long long)b * t.sB[l] + cell;
        const long long cs = t.sC[l];
        const int w = t.w[l];
        const float gx = (float)(cell % w) + 0.5f, gy = (float)(cell / w) + 0.5f;  // make_anchors tal.py:300-312
        const int nc = P.nc;
        // classes: first maximum of the sigmoid scores (decode_3d_size tal.py:458-462), sum softplus
        int cls = 0;
        float best = -1.0f, acc = 0.f;
        for (int c = 0; c < nc; ++c) {
            const float v = x[(long long)c * cs];
            const float s = 1.0f / (1.0f + expf(-v));
            if (s > best) { best = s; cls = c; }
            acc += fmaxf(v, 0.f) + log1pf(expf(-fabsf(v)));
        }
        soft = (double)acc;
        const float *r = x + (long long)nc * cs;  // 35 regression channels
        // 2D box: centres = anchor + offset; xy1 = centres - size/2; xy2 = centres + size/2; * stride (loss.py:812-819)
        const float cx = dm::add(gx, r[0]), cy = dm::add(gy, r[cs]);
        const float hw_ = dm::mul(r[2 * cs], 0.5f), hh_ = dm::mul(r[3 * cs], 0.5f);
        const long long o = (long long)b * P.A + a;
        *reinterpret_cast<float4 *>(P.boxes + o * 4) =
            make_float4(dm::mul(dm::sub(cx, hw_), st), dm::mul(dm::sub(cy, hh_), st), dm::mul(dm::add(cx, hw_), st),
                        dm::mul(dm::add(cy, hh_), st));
        // 3D: decode_3d_center tal.py:454-456 (anchor px + offset * stride), size = mean + residual, heading argmax
        const float ax = dm::mul(gx, st), ay = dm::mul(gy, st);
        const float c3x = dm::add(ax, dm::mul(r[4 * cs], st)), c3y = dm::add(ay, dm::mul(r[5 * cs], st));
        const float sh_ = dm::add(P.mean_sizes[3 * cls], r[6 * cs]), sw_ = dm::add(P.mean_sizes[3 * cls + 1], r[7 * cs]),
                    sl_ = dm::add(P.mean_sizes[3 * cls + 2], r[8 * cs]);
        int hb = 0;
        float hbv = r[9 * cs];
        for (int j = 1; j < 12; ++j) {
            const float v = r[(9 + j) * cs];
            if (v > hbv) { hbv = v; hb = j; }
        }
        float out[24];
        keypoints24(c3x, c3y, r[33 * cs], sh_, sw_, sl_, hb, r[(9 + 12 + hb) * cs], P.calibs + 6 * b, out);
        float4 *ok = reinterpret_cast<float4 *>(P.pd_kps + o * 24);
#pragma unroll
        for (int j = 0; j < 6; ++j) ok[j] = make_float4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
        if (P.claim) P.claim[o] = 0ull;
    }
    soft = warp_sum(soft);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = soft;
    __syncthreads();
    if (threadIdx.x == 0) {
        double *p = P.part + ((long long)b * gridDim.x + blockIdx.x) * (kNSum + 1);
        p[0] = (red[0] + red[1]) + (red[2] + red[3]);
    }
}

// loss.py:879-888: items = loss2d, cls, depth, offset3d, size3d, heading (+ target_scores_sum, n_fg) from the partials
__device__ __forceinline__ void dd_items(const double *p, int M, const float *g, float *items) {
    if (M == 0) {  // no targets: the reference returns its zero-initialised loss (loss.py:873-876)
        for (int i = 0; i < 8; ++i) items[i] = 0.f;
        items[6] = 1.f;
        return;
    }
    const double tss = p[9] > 1.0 ? p[9] : 1.0, nfg = p[10];
    // F.l1_loss(..., reduction="mean") over [n_fg, 2] elements (NaN when nothing is assigned, like the reference)
    items[0] = (float)((p[2] / (2.0 * nfg) + p[1] / (2.0 * nfg)) / tss * g[0]);
    items[1] = (float)((p[0] - p[8]) / tss * g[1]);
    items[2] = (float)(p[3] / tss * g[2]);
    items[3] = (float)(p[4] / (2.0 * nfg) / tss * g[3]);
    items[4] = (float)(p[5] / tss * g[4]);
    items[5] = (float)((p[6] + p[7]) / tss * g[5]);
    items[6] = (float)tss;
    items[7] = (float)nfg;
}

// grid (ceil(ceil(A/128) / kFgRows), B, n_branch): a CTA covers the anchors of kFgRows CTAs of the streaming kernel and puts
// their foreground sums into slots 1..kNSum of the first of those partial rows (n_rows = the streaming kernel's grid.x).
// Only a few per cent of the anchors are foreground: the CTA first compacts its foreground anchors (in anchor order, so
// the sums do not depend on timing) and then gives every one of them its own thread -- a quarter of the CTAs, fences and
// tickets of a one-anchor-per-thread grid, and no thread runs two latency chains back to back.
constexpr int kFgRows = 4;
__global__ void __launch_bounds__(128) dd_fg_kernel(const __grid_constant__ AssignCtx2 cc, const __grid_constant__ DDParams2 PP,
                                                    int n_rows) {
    __shared__ double red[kNSum][4];
    __shared__ int s_cnt[kFgRows][4];
    __shared__ int s_fa[kFgRows * 128], s_fgi[kFgRows * 128];
    const AssignCtx &c = cc.c[blockIdx.z];
    const DDParams &P = PP.p[blockIdx.z];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    asm volatile("griddepcontrol.wait;" ::: "memory");  // programmatic dependent of the resolve (or the streaming) kernel
    double s[kNSum];
#pragma unroll
    for (int i = 0; i < kNSum; ++i) s[i] = 0.0;
    int gis[kFgRows];
    unsigned bal[kFgRows];
#pragma unroll
    for (int i = 0; i < kFgRows; ++i) {
        const int a = (blockIdx.x * kFgRows + i) * 128 + threadIdx.x;
        gis[i] = (a < P.A && P.M > 0) ? c.tgi[(long long)b * P.A + a] : -1;
    }
#pragma unroll
    for (int i = 0; i < kFgRows; ++i) {
        bal[i] = __ballot_sync(0xffffffffu, gis[i] >= 0);
        if (lane == 0) s_cnt[i][wid] = __popc(bal[i]);
    }
    __syncthreads();
    int n_fg = 0;
    {   // position of (row i, warp w) in anchor order = everything before it
        int before[kFgRows];
        int run = 0;
#pragma unroll
        for (int i = 0; i < kFgRows; ++i) {
            before[i] = 0;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                if (w == wid) before[i] = run;
                run += s_cnt[i][w];
            }
        }
        n_fg = run;
#pragma unroll
        for (int i = 0; i < kFgRows; ++i)
            if (gis[i] >= 0) {
                const int pos = before[i] + __popc(bal[i] & ((1u << lane) - 1u));
                s_fa[pos] = (blockIdx.x * kFgRows + i) * 128 + threadIdx.x;
                s_fgi[pos] = gis[i];
            }
    }
    __syncthreads();
#pragma unroll 1
    for (int e = threadIdx.x; e < n_fg; e += 128) {
        const int a = s_fa[e];
        const long long o = (long long)b * P.A + a;
        const int gi = s_fgi[e];
        {
            const LevelTable &t = P.t;
            const int l = level_of(t, a);
            const int cell = a - t.start[l];
            const float st = t.stride[l];
            const float *x = t.ptr[l] + (long long)b * t.sB[l] + cell;
            const long long cs = t.sC[l];
            const int w = t.w[l];
            const float ax = ((float)(cell % w) + 0.5f) * st, ay = ((float)(cell / w) + 0.5f) * st;
            const float *g = P.gts + ((long long)b * P.M + gi) * 17;
            const float *r = x + (long long)P.nc * cs;
            const float norm = assigned_norm(c, b, gi, c.alignv[o]);  // target score at the assigned label
            int lab = (int)g[0];
            lab = lab < 0 ? 0 : lab;
            // compute_box2d_loss loss.py:913-926 (px): offset vs (center_2d - anchor), size vs size_2d
            s[1] += (double)fabsf(r[0] * st - (g[5] - ax)) + (double)fabsf(r[cs] * st - (g[6] - ay));
            s[2] += (double)fabsf(r[2 * cs] * st - g[7]) + (double)fabsf(r[3 * cs] * st - g[8]);
            // compute_box3d_loss loss.py:928-963
            const float dep = r[33 * cs], un = r[34 * cs];
            s[3] += (double)(1.4142f * expf(-0.5f * un) * fabsf(dep - g[14]) + 0.5f * un);  // loss.py:1118
            s[4] += (double)fabsf(r[4 * cs] * st - (g[9] - ax)) + (double)fabsf(r[5 * cs] * st - (g[10] - ay));
            s[5] += (double)fabsf(r[6 * cs] - g[11]) + (double)fabsf(r[7 * cs] - g[12]) + (double)fabsf(r[8 * cs] - g[13]);
            // compute_heading_loss loss.py:1122-1136: CE over the 12 bins + L1 of the residual of the target bin
            const int tb = (int)g[15];
            float m = -3.4e38f, hv[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                hv[j] = r[(9 + j) * cs];
                m = fmaxf(m, hv[j]);
            }
            float se = 0.f;
#pragma unroll
            for (int j = 0; j < 12; ++j) se += expf(hv[j] - m);
            s[6] += (double)(m + logf(se) - r[(9 + tb) * cs]);
            s[7] += (double)fabsf(r[(9 + 12 + tb) * cs] - g[16]);
            s[8] += (double)x[(long long)lab * cs] * (double)norm;  // BCE(x,t) - BCE(x,0) = -x*t
            s[9] += (double)norm;
            s[0] += 1.0;  // foreground count
        }
    }
    if (__syncthreads_or(s[0] != 0.0)) {
#pragma unroll
        for (int i = 0; i < kNSum; ++i) {
            const double v = warp_sum(s[i]);
            if ((threadIdx.x & 31) == 0) red[i][threadIdx.x >> 5] = v;
        }
        __syncthreads();
    } else if (threadIdx.x < 4 * kNSum) {  // most CTAs hold no foreground anchor: their row is zero
        red[threadIdx.x >> 2][threadIdx.x & 3] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x < kNSum) {
        double *p = P.part + ((long long)b * n_rows + blockIdx.x * kFgRows) * (kNSum + 1);
        const int i = threadIdx.x;
        const double v = (red[i][0] + red[i][1]) + (red[i][2] + red[i][3]);
        if (i == 0) p[kNSum] = v;  // n_fg
        else p[i] = v;
    }
    // ---- last CTA of the image: the image's rows -> part_img[b]; last image: part_img -> partials (+ items)
    __shared__ unsigned s_ticket;
    __shared__ double s_fin[kNSum + 1];
    constexpr int W = kNSum + 1;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(P.tickets + b, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    {  // warp w takes columns w, w + 4, w + 8; lanes stride the rows; fixed shuffle tree
        double acc[3] = {0.0, 0.0, 0.0};
        for (int r = lane; r < n_rows; r += 32) {  // column 0 (soft+) lives in every row, the others in every kFgRows-th
            const double *row = P.part + ((long long)b * n_rows + r) * W;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int col = wid + 4 * j;
                if (col < W && (col == 0 || r % kFgRows == 0)) acc[j] += __ldcg(row + col);
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double v = warp_sum(acc[j]);
            if (lane == 0 && wid + 4 * j < W) P.part_img[(long long)b * W + wid + 4 * j] = v;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(P.tickets + P.B, 1u);
    __syncthreads();
    if (s_ticket != gridDim.y - 1) return;
    __threadfence();
    {
        double acc[3] = {0.0, 0.0, 0.0};
        for (int r = lane; r < P.B; r += 32) {
            const double *row = P.part_img + (long long)r * W;
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (wid + 4 * j < W) acc[j] += __ldcg(row + wid + 4 * j);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double v = warp_sum(acc[j]);
            if (lane == 0 && wid + 4 * j < W) { s_fin[wid + 4 * j] = v; P.partials[wid + 4 * j] = v; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && P.items) dd_items(s_fin, P.M, P.gain, P.items);
}

// partials float64[kNSum + 1] = soft+, off2d, size2d, depth, off3d, size3d, hd_ce, hd_l1, x*t, sum t, n_fg
// loss.py:879-888: items = loss2d, cls, depth, offset3d, size3d, heading (+ target_scores_sum, n_fg)
__global__ void dd_finalize_kernel(const double *p, int M, float g2d, float gcls, float gdep, float go3d, float gs3d,
                                   float ghd, float *items) {
    if (threadIdx.x != 0) return;
    const float g[6] = {g2d, gcls, gdep, go3d, gs3d, ghd};
    dd_items(p, M, g, items);
}

// ------------------------------------------------------------------------------------------------ backward
struct DDBwdParams {
    LevelTable t;                    // head tensors (forward inputs)
    float *g_ptr[Y3D_MAX_LEVELS];    // gradient tensors, same indexing
    long long g_sB[Y3D_MAX_LEVELS], g_sC[Y3D_MAX_LEVELS];
    const float *gts;
    const float *items;              // DEVICE float[8] of the forward pass: [6] = target_scores_sum, [7] = n_fg
    const float *gitems;             // DEVICE float[6]: d total / d item
    float gain[6];
    int B, nc, A, M;
};

// d (loss items) / d head, reference autograd of loss.py:879-888 with compute_box2d_loss :913-926, compute_box3d_loss
// :928-963, laplacian_aleatoric_uncertainty_loss_new :1118 and compute_heading_loss :1122-1136.  One thread per
// anchor writes all nc+35 gradient rows of that anchor (coalesced along the anchor axis): the class rows are dense
// (sigmoid(x) - t), the regression rows are zero except at foreground anchors.
__global__ void __launch_bounds__(128) dd_bwd_kernel(AssignCtx c, DDBwdParams P) {
    const int b = blockIdx.y, a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= P.A) return;
    const LevelTable &t = P.t;
    const int l = level_of(t, a);
    const int cell = a - t.start[l];
    const float st = t.stride[l];
    const float *x = t.ptr[l] + (long long)b * t.sB[l] + cell;
    float *g = P.g_ptr[l] + (long long)b * P.g_sB[l] + cell;
    const long long cs = t.sC[l], gs = P.g_sC[l];
    const int nc = P.nc;
    const float tss = P.items[6], nfg = P.items[7];
    float k[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) k[i] = P.M > 0 ? P.gitems[i] * P.gain[i] / tss : 0.0f;  // no targets: constant zero loss
    const long long o = (long long)b * P.A + a;
    const int gi = P.M > 0 ? c.tgi[o] : -1;
    int lab = -1;
    float norm = 0.f;
    const float *gt = nullptr;
    if (gi >= 0) {
        gt = P.gts + ((long long)b * P.M + gi) * 17;
        lab = (int)gt[0];
        lab = lab < 0 ? 0 : lab;
        norm = assigned_norm(c, b, gi, c.alignv[o]);
    }
    for (int cc = 0; cc < nc; ++cc) {  // BCEWithLogits: sigmoid(x) - t
        const float s = 1.0f / (1.0f + __expf(-x[(long long)cc * cs]));
        g[(long long)cc * gs] = k[1] * (s - (cc == lab ? norm : 0.f));
    }
    const float *r = x + (long long)nc * cs;
    float *gr = g + (long long)nc * gs;
    if (gi < 0) {
#pragma unroll 5
        for (int j = 0; j < 35; ++j) gr[(long long)j * gs] = 0.f;
        return;
    }
    auto sgn = [](float v) { return v > 0.f ? 1.0f : (v < 0.f ? -1.0f : 0.0f); };
    const int w = t.w[l];
    const float ax = ((float)(cell % w) + 0.5f) * st, ay = ((float)(cell / w) + 0.5f) * st;
    const float m2 = 1.0f / (2.0f * nfg);  // F.l1_loss(reduction="mean") over [n_fg, 2]
    gr[0] = k[0] * m2 * st * sgn(r[0] * st - (gt[5] - ax));
    gr[gs] = k[0] * m2 * st * sgn(r[cs] * st - (gt[6] - ay));
    gr[2 * gs] = k[0] * m2 * st * sgn(r[2 * cs] * st - gt[7]);
    gr[3 * gs] = k[0] * m2 * st * sgn(r[3 * cs] * st - gt[8]);
    gr[4 * gs] = k[3] * m2 * st * sgn(r[4 * cs] * st - (gt[9] - ax));
    gr[5 * gs] = k[3] * m2 * st * sgn(r[5 * cs] * st - (gt[10] - ay));
    gr[6 * gs] = k[4] * sgn(r[6 * cs] - gt[11]);
    gr[7 * gs] = k[4] * sgn(r[7 * cs] - gt[12]);
    gr[8 * gs] = k[4] * sgn(r[8 * cs] - gt[13]);
    const int tb = (int)gt[15];
    float hv[12], m = -3.4e38f;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        hv[j] = r[(9 + j) * cs];
        m = fmaxf(m, hv[j]);
    }
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        hv[j] = __expf(hv[j] - m);
        se += hv[j];
    }
    const float inv = 1.0f / se;
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        gr[(9 + j) * gs] = k[5] * (hv[j] * inv - (j == tb ? 1.0f : 0.0f));
        gr[(21 + j) * gs] = j == tb ? k[5] * sgn(r[(21 + j) * cs] - gt[16]) : 0.0f;
    }
    const float dep = r[33 * cs], un = r[34 * cs];
    const float e = 1.4142f * __expf(-0.5f * un), d = dep - gt[14];
    gr[33 * gs] = k[2] * e * sgn(d);
    gr[34 * gs] = k[2] * (-0.5f * e * fabsf(d) + 0.5f);
}

struct DDWs {
    size_t assign, boxes, pd_kps, gt_kps, part, part_img, tickets, total;
    int n_rows;
};
static DDWs dd_ws_layout(int B, int A, int M) {
    DDWs w;
    size_t o = 0;
    w.assign = o; o += a256(assign_ws_layout(B, A, M).total);
    w.boxes = o;  o += a256(sizeof(float) * 4 * (size_t)B * A);
    w.pd_kps = o; o += a256(sizeof(float) * 24 * (size_t)B * A);
    w.gt_kps = o; o += a256(sizeof(float) * 24 * (size_t)B * (M > 0 ? M : 1));
    w.n_rows = ((A + 127) / 128) * B;
    w.part = o;   o += a256(sizeof(double) * (kNSum + 1) * (size_t)w.n_rows);
    w.part_img = o; o += a256(sizeof(double) * (kNSum + 1) * (size_t)B);
    w.tickets = o;  o += a256(sizeof(unsigned) * ((size_t)B + 1));
    w.total = o;
    return w;
}
size_t dd_loss_workspace_bytes(int B, int A, int M) { return dd_ws_layout(B, A, M).total; }

}  // namespace y3d

using namespace y3d;

struct DDBranchIn {
    const float *const *lvl_ptr;
    const int64_t *sB, *sC;
    int topk;
};

// nb = 1: one branch (DDDetectionLoss); nb = 2: both branches of DetectLoss3d in the same four launches.  Branch z uses
// the workspace block [z * dd_ws_layout().total, ...); loss_items [nb][8], partials [nb][kNSum + 1], dbg [nb][B][A].
static int dd_loss_run(int nb, const DDBranchIn *br, const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc,
                       const float *gts, int M, const float *calibs, const float *mean_sizes, float alpha, float beta,
                       float gamma, int flags, const float *gains, int normalise, float *loss_items, double *partials,
                       int32_t *dbg_target_gt_idx, void *ws, size_t ws_bytes, void *stream) {
    if (!lvl_hw || !lvl_stride || !calibs || !mean_sizes || !gains) return Y3D_EINVAL;
    if (B < 1 || nc < 1 || M < 0 || (M > 0 && !gts) || !partials || (normalise && !loss_items)) return Y3D_EINVAL;
    const int use_2d = flags & 1, use_3d = (flags >> 1) & 1, kps_l2 = (flags >> 2) & 1, constrain = (flags >> 3) & 1;
    if (!use_2d && !use_3d) return Y3D_EINVAL;  // tal.py:486
    DDParams2 PP{};
    AssignCtx2 cc{};
    int A = 0;
    for (int z = 0; z < nb; ++z) {
        if (!br[z].lvl_ptr || !br[z].sB || !br[z].sC) return Y3D_EINVAL;
        A = make_level_table(PP.p[z].t, br[z].lvl_ptr, br[z].sB, br[z].sC, lvl_hw, lvl_stride, nl);
        if (A < 0) return A;
        for (int l = 0; l < nl; ++l)
            if (!br[z].lvl_ptr[l]) return Y3D_EINVAL;
        if (br[z].topk < 1 || br[z].topk > A) return Y3D_EINVAL;
        if (br[z].topk > Y3D_MAX_TOPK) return Y3D_EUNSUPPORTED;
    }
    const DDWs w = dd_ws_layout(B, A, M);
    if (!ws || ws_bytes < (size_t)nb * w.total) return Y3D_EWORKSPACE;
    if (((uintptr_t)ws) % 256) return Y3D_EALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    const AssignWs aw = assign_ws_layout(B, A, M);
    float *gt_kps = (float *)((char *)ws + w.gt_kps);  // shared by the branches (same GT)
    for (int z = 0; z < nb; ++z) {
        char *p = (char *)ws + (size_t)z * w.total;
        DDParams &P = PP.p[z];
        AssignCtx &c = cc.c[z];
        c.t = P.t;
        assign_bind_ws(c, p + w.assign, aw);
        P.gts = gts; P.calibs = calibs; P.mean_sizes = mean_sizes;
        P.boxes = (float *)(p + w.boxes);
        P.pd_kps = (float *)(p + w.pd_kps);
        P.claim = M > 0 ? c.claim : nullptr;  // zeroed by the streaming kernel, like the per-GT maxima and the work counter
        P.part = (double *)(p + w.part);
        P.part_img = (double *)(p + w.part_img);
        P.tickets = (unsigned *)(p + w.tickets);
        P.partials = partials + (size_t)z * (kNSum + 1);
        P.items = normalise ? loss_items + 8 * z : nullptr;
        for (int i = 0; i < 6; ++i) P.gain[i] = gains[i];
        P.B = B; P.nc = nc; P.A = A; P.M = M;
        c.score_mode = 1; c.cls_ch0 = 0;  // class logits are the first nc channels of the head
        c.pd_bboxes = P.boxes; c.box_grid_units = 0; c.box_soa = 0;
        c.use_grid = 1;
        c.gt_labels = gts; c.gl_stride = 17;
        c.gt_bboxes = gts ? gts + 1 : nullptr; c.gb_stride = 17;
        c.mask_gt = nullptr;  // valid iff the box sums to > 0 (loss.py:857)
        c.B = B; c.A = A; c.nc = nc; c.M = M; c.k = br[z].topk;
        c.alpha = alpha; c.beta = beta; c.gamma = gamma; c.eps = 1e-9f;
        c.use_2d = use_2d; c.use_3d = use_3d; c.kps_l2 = kps_l2; c.constrain = constrain;
        c.pd_kps = P.pd_kps; c.gt_kps = gt_kps;
    }
    cc.work_counter = (int *)((char *)ws + w.assign + aw.off_work);
    dim3 grid((A + 127) / 128, B, nb);
    dd_stream_kernel<<<grid, 128, 0, s>>>(PP, M > 0 ? cc.work_counter : nullptr, M > 0 ? cc.c[0].pos_align : nullptr,
                                          cc.c[0].pos_ov, nb > 1 ? cc.c[1].pos_align : nullptr, cc.c[1].pos_ov,
                                          M > 0 ? gt_kps : nullptr);
    Y3D_CHECK_LAUNCH();
    if (M > 0) {
        const int rc = assign_run_core(cc, nb, s, nullptr, /*pdl=*/true);
        if (rc) return rc;
    }
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((grid.x + kFgRows - 1) / kFgRows, B, nb);
        cfg.blockDim = dim3(128);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t le = cudaLaunchKernelEx(&cfg, dd_fg_kernel, cc, PP, (int)grid.x);
        if (le != cudaSuccess) return (int)le;
    }
    Y3D_CHECK_LAUNCH();
    if (dbg_target_gt_idx) {  // assigned GT per anchor, -1 = background
        for (int z = 0; z < nb; ++z) {
            cudaError_t e;
            int32_t *dst = dbg_target_gt_idx + (size_t)z * B * A;
            if (M == 0) e = cudaMemsetAsync(dst, 0xff, sizeof(int32_t) * (size_t)B * A, s);
            else e = cudaMemcpyAsync(dst, cc.c[z].tgi, sizeof(int32_t) * (size_t)B * A, cudaMemcpyDeviceToDevice, s);
            if (e != cudaSuccess) return (int)e;
        }
    }
    return Y3D_OK;
}

extern "C" int y3d_dd_loss_fwd(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC,
                               const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, const float *gts, int M,
                               const float *calibs, const float *mean_sizes, int topk, float alpha, float beta,
                               float gamma, int flags, const float *gains, int normalise, float *loss_items,
                               double *partials, int32_t *dbg_target_gt_idx, void *ws, size_t ws_bytes,
                               void *stream) {
    DDBranchIn br[1] = {{lvl_ptr, lvl_sB, lvl_sC, topk}};
    return dd_loss_run(1, br, lvl_hw, lvl_stride, nl, B, nc, gts, M, calibs, mean_sizes, alpha, beta, gamma, flags, gains,
                       normalise, loss_items, partials, dbg_target_gt_idx, ws, ws_bytes, stream);
}

extern "C" int y3d_dd_loss_dual_fwd(const float *const *o2m_ptr, const int64_t *o2m_sB, const int64_t *o2m_sC,
                                    const float *const *o2o_ptr, const int64_t *o2o_sB, const int64_t *o2o_sC,
                                    const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, const float *gts,
                                    int M, const float *calibs, const float *mean_sizes, int topk_o2m, int topk_o2o,
                                    float alpha, float beta, float gamma, int flags, const float *gains, int normalise,
                                    float *loss_items, double *partials, int32_t *dbg_target_gt_idx, void *ws,
                                    size_t ws_bytes, void *stream) {
    DDBranchIn br[2] = {{o2m_ptr, o2m_sB, o2m_sC, topk_o2m}, {o2o_ptr, o2o_sB, o2o_sC, topk_o2o}};
    return dd_loss_run(2, br, lvl_hw, lvl_stride, nl, B, nc, gts, M, calibs, mean_sizes, alpha, beta, gamma, flags, gains,
                       normalise, loss_items, partials, dbg_target_gt_idx, ws, ws_bytes, stream);
}

extern "C" int y3d_dd_loss_bwd(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC,
                               float *const *grad_ptr, const int64_t *grad_sB, const int64_t *grad_sC,
                               const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, const float *gts, int M,
                               const float *gains, const float *loss_items, const float *grad_items, const void *ws,
                               size_t ws_bytes, void *stream) {
    if (!lvl_ptr || !lvl_sB || !lvl_sC || !grad_ptr || !grad_sB || !grad_sC || !gains || !loss_items || !grad_items)
        return Y3D_EINVAL;
    if (B < 1 || nc < 1 || M < 0 || (M > 0 && !gts)) return Y3D_EINVAL;
    DDBwdParams P{};
    const int A = make_level_table(P.t, lvl_ptr, lvl_sB, lvl_sC, lvl_hw, lvl_stride, nl);
    if (A < 0) return A;
    for (int l = 0; l < nl; ++l) {
        if (!lvl_ptr[l] || !grad_ptr[l]) return Y3D_EINVAL;
        P.g_ptr[l] = grad_ptr[l];
        P.g_sB[l] = grad_sB[l];
        P.g_sC[l] = grad_sC[l];
    }
    const DDWs w = dd_ws_layout(B, A, M);  // must be the untouched workspace of the forward call
    if (!ws || ws_bytes < w.total) return Y3D_EWORKSPACE;
    if (((uintptr_t)ws) % 256) return Y3D_EALIGN;
    AssignCtx c{};
    assign_bind_ws(c, (char *)ws + w.assign, assign_ws_layout(B, A, M));
    c.B = B; c.A = A; c.nc = nc; c.M = M; c.eps = 1e-9f;
    P.gts = gts; P.items = loss_items; P.gitems = grad_items;
    for (int i = 0; i < 6; ++i) P.gain[i] = gains[i];
    P.B = B; P.nc = nc; P.A = A; P.M = M;
    dd_bwd_kernel<<<dim3((A + 127) / 128, B), 128, 0, (cudaStream_t)stream>>>(c, P);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

extern "C" int y3d_dd_loss_finalize(const double *partials, int M, const float *gains, float *loss_items, void *stream) {
    if (!partials || !gains || !loss_items) return Y3D_EINVAL;
    dd_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partials, M, gains[0], gains[1], gains[2], gains[3], gains[4],
                                                          gains[5], loss_items);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}
