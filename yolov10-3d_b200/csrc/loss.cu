// loss.cu -- fused v8DetectionLoss forward for one branch (sm_100a).
//   y3d_v8_loss_fwd      v8DetectionLoss.__call__   reference ultralytics/utils/loss.py:206-257
//                        bbox_decode                reference ultralytics/utils/loss.py:197-204
//                        BboxLoss.forward/_df_loss  reference ultralytics/utils/loss.py:82-113
//                        bbox2dist                  reference ultralytics/utils/tal.py:328-331
//   y3d_train_decode     bbox_decode + permute/sigmoid (loss.py:214-215,232)
//   y3d_v8_loss_finalize normalisation by target_scores_sum (loss.py:240-256)
//
// Pipeline per branch (v10DetectLoss runs it twice: top-k 10 on one2many, top-k 1 on one2one, loss.py:727-737):
//   1. loss_stream_kernel : the ONE pass over the head tensor (4*(4R+nc)*A bytes per image): DFL softmax-integral
//                           -> xyxy boxes in grid units [B,A,4] (16 B/anchor, the only dense write), and
//                           sum softplus(logit) = sum BCE(logit, 0) over all class logits.  pd_scores and the dense
//                           target_scores of the reference are never materialised:
//                           sum BCE(x,t) = sum BCE(x,0) - sum_fg x[label]*t.
//   2. assignment core    : assign.cuh, scores read as logits straight from the head (sigmoid on the fly).
//   3. loss_fg_kernel     : per foreground anchor CIoU / DFL / BCE-correction terms.
//   4. loss_finalize_kernel: fixed-order (deterministic) reduction of the per-block partials in float64.
#include "assign.cuh"

namespace y3d {

constexpr int kR = 16;

struct Quads {
    int qstart[Y3D_MAX_LEVELS + 1];
};

template <int VEC>
__device__ __forceinline__ void ld(const float *p, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        float4 t = ldg_stream4(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = ldg_stream1(p);
    }
}

template <int VEC>
__device__ __forceinline__ void dfl_side(const float *p, long long cs, float (&out)[VEC]) {
    float x[kR][VEC];
#pragma unroll
    for (int j = 0; j < kR; ++j) ld<VEC>(p + j * cs, x[j]);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
        float m = x[0][e];
#pragma unroll
        for (int j = 1; j < kR; ++j) m = fmaxf(m, x[j][e]);
        float s = 0.f, acc = 0.f;
#pragma unroll
        for (int j = 0; j < kR; ++j) {
            float ex = expf(x[j][e] - m);
            s += ex;
            acc += (float)j * ex;
        }
        out[e] = acc / s;
    }
}

__device__ __forceinline__ float softplusf_(float x) {  // BCEWithLogits(x, 0) = max(x,0) + log1p(exp(-|x|))
    return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}

// grid (ceil(Q/32), B), block (32, 2 + n_cls_roles); same warp-role layout as decode2d_kernel.
template <int VEC>
__global__ void __launch_bounds__(256) loss_stream_kernel(LevelTable t, Quads qm, int nc, int cls_chunk, int A,
                                                          float *__restrict__ pd_bboxes, float *__restrict__ pd_scores,
                                                          double *__restrict__ part_bce) {
    __shared__ double red[8];
    const int b = blockIdx.y;
    const int q = blockIdx.x * 32 + threadIdx.x;
    const int role = threadIdx.y;
    double local = 0.0;
    if (q < qm.qstart[t.nl]) {
        int l = 0;
#pragma unroll
        for (int i = 1; i < Y3D_MAX_LEVELS; ++i) l += (i < t.nl && q >= qm.qstart[i]) ? 1 : 0;
        const int cell = (q - qm.qstart[l]) * VEC;
        const float *base = t.ptr[l] + (long long)b * t.sB[l] + cell;
        const long long cs = t.sC[l];
        const long long a0 = (long long)b * A + t.start[l] + cell;
        if (role < 2) {
            float d_lo[VEC], d_hi[VEC];
            dfl_side<VEC>(base + (long long)(role * kR) * cs, cs, d_lo);
            dfl_side<VEC>(base + (long long)((role + 2) * kR) * cs, cs, d_hi);
            const int w = t.w[l];
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                int c = cell + e;
                float anc = (role == 0 ? (float)(c % w) : (float)(c / w)) + 0.5f;
                float *o = pd_bboxes + (a0 + e) * 4 + role;
                o[0] = anc - d_lo[e];  // x1 / y1   (dist2bbox xyxy, tal.py:319-325)
                o[2] = anc + d_hi[e];  // x2 / y2
            }
        } else {
            const int c0 = (role - 2) * cls_chunk;
            const int c1 = min(nc, c0 + cls_chunk);
            const float *p = base + (long long)(4 * kR + c0) * cs;
            float acc = 0.f;
            int c = c0;
            for (; c + 4 <= c1; c += 4) {
                float v[4][VEC];
#pragma unroll
                for (int u = 0; u < 4; ++u) ld<VEC>(p + (long long)u * cs, v[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        acc += softplusf_(v[u][e]);
                        if (pd_scores) pd_scores[(a0 + e) * nc + c + u] = 1.0f / (1.0f + expf(-v[u][e]));
                    }
                p += 4 * cs;
            }
            for (; c < c1; ++c) {
                float v[VEC];
                ld<VEC>(p, v);
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    acc += softplusf_(v[e]);
                    if (pd_scores) pd_scores[(a0 + e) * nc + c] = 1.0f / (1.0f + expf(-v[e]));
                }
                p += cs;
            }
            local = (double)acc;
        }
    }
    // deterministic block reduction -> one partial per block
    local = warp_sum(local);
    if (threadIdx.x == 0) red[threadIdx.y] = local;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        double s = 0.0;
        for (int i = 2; i < (int)blockDim.y; ++i) s += red[i];
        if (part_bce) part_bce[(long long)blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}

// per foreground anchor loss terms; grid (ceil(A/256), B); partials [4][nblocks]
__global__ void __launch_bounds__(256) loss_fg_kernel(AssignCtx c, const float *__restrict__ gt5,
                                                      double *__restrict__ part, uint8_t *__restrict__ dbg_fg,
                                                      int32_t *__restrict__ dbg_gi) {
    __shared__ double red[4][8];
    const int b = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    double s_iou = 0.0, s_dfl = 0.0, s_ts = 0.0, s_xt = 0.0;
    if (a < c.A) {
        const long long o = (long long)b * c.A + a;
        const int gi = c.tgi[o];
        if (dbg_fg) dbg_fg[o] = (uint8_t)(gi >= 0);
        if (dbg_gi) dbg_gi[o] = gi >= 0 ? gi : 0;
        if (gi >= 0) {
            const float wgt = assigned_norm(c, b, gi, c.alignv[o]);  // = target_scores.sum(-1) (one-hot * norm)
            const int l = level_of(c.t, a);
            const int cell = a - c.t.start[l];
            const float st = c.t.stride[l];
            const float ax = (float)(cell % c.t.w[l]) + 0.5f, ay = (float)(cell / c.t.w[l]) + 0.5f;
            const float *g = gt5 + ((long long)b * c.M + gi) * 5;
            int lab = (int)g[0];
            lab = lab < 0 ? 0 : lab;
            // target_bboxes /= stride_tensor (loss.py:248)
            const float4 tb = make_float4(dm::div(g[1], st), dm::div(g[2], st), dm::div(g[3], st), dm::div(g[4], st));
            const float4 pb = *reinterpret_cast<const float4 *>(c.pd_bboxes + o * 4);
            const float iou = dm::ciou(pb, tb, dm::box1_atan(pb));  // BboxLoss.forward loss.py:85 (box1 = pred)
            s_iou = (double)(1.0f - iou) * (double)wgt;
            // DFL (loss.py:90-113): targets bbox2dist(...).clamp(0, reg_max-1-0.01)
            const float *hp = c.t.ptr[l] + (long long)b * c.t.sB[l] + cell;
            const long long cs = c.t.sC[l];
            const float ltrb[4] = {ax - tb.x, ay - tb.y, tb.z - ax, tb.w - ay};
            float dfl = 0.f;
#pragma unroll
            for (int side = 0; side < 4; ++side) {
                float tt = fminf(fmaxf(ltrb[side], 0.0f), (float)(kR - 1) - 0.01f);
                int tl = (int)tt;
                float wl = (float)(tl + 1) - tt, wr = 1.0f - wl;
                float x[kR];
                float m = -INFINITY;
#pragma unroll
                for (int j = 0; j < kR; ++j) {
                    x[j] = hp[(long long)(side * kR + j) * cs];
                    m = fmaxf(m, x[j]);
                }
                float se = 0.f, xl = 0.f, xr = 0.f;
#pragma unroll
                for (int j = 0; j < kR; ++j) {
                    se += expf(x[j] - m);
                    xl = (j == tl) ? x[j] : xl;
                    xr = (j == tl + 1) ? x[j] : xr;
                }
                float lse = m + logf(se);
                dfl += (lse - xl) * wl + (lse - xr) * wr;
            }
            s_dfl = (double)(dfl * 0.25f) * (double)wgt;  // .mean(-1) over the 4 sides
            s_ts = (double)wgt;
            const float xlab = hp[(long long)(4 * kR + lab) * cs];
            s_xt = (double)xlab * (double)wgt;  // BCE(x,t) - BCE(x,0) = -x*t
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    s_iou = warp_sum(s_iou); s_dfl = warp_sum(s_dfl); s_ts = warp_sum(s_ts); s_xt = warp_sum(s_xt);
    if (lane == 0) { red[0][wid] = s_iou; red[1][wid] = s_dfl; red[2][wid] = s_ts; red[3][wid] = s_xt; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
        const long long nb = (long long)gridDim.x * gridDim.y;
        part[threadIdx.x * nb + (long long)blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}

// one block: fixed-order tree over the per-block partials; out: partials[4] = sum (1-ciou)w, sum bce, sum dfl w, sum ts
__global__ void __launch_bounds__(256) loss_finalize_kernel(const double *__restrict__ part_bce, int n_bce,
                                                            const double *__restrict__ part_fg, int n_fg,
                                                            float gain_box, float gain_cls, float gain_dfl,
                                                            int normalise, double *__restrict__ partials,
                                                            float *__restrict__ loss_items) {
    __shared__ double red[5][256];
    double acc[5] = {0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < n_bce; i += 256) acc[0] += part_bce[i];
    if (part_fg)
        for (int k = 0; k < 4; ++k)
            for (int i = threadIdx.x; i < n_fg; i += 256) acc[1 + k] += part_fg[(long long)k * n_fg + i];
    for (int k = 0; k < 5; ++k) red[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int k = 0; k < 5; ++k) red[k][threadIdx.x] += red[k][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double bce0 = red[0][0], s_iou = red[1][0], s_dfl = red[2][0], s_ts = red[3][0], s_xt = red[4][0];
        const double bce = bce0 - s_xt;
        if (partials) { partials[0] = s_iou; partials[1] = bce; partials[2] = s_dfl; partials[3] = s_ts; }
        if (normalise && loss_items) {
            const double tss = s_ts > 1.0 ? s_ts : 1.0;  // max(target_scores.sum(), 1) loss.py:240
            loss_items[0] = (float)(s_iou / tss * gain_box);
            loss_items[1] = (float)(bce / tss * gain_cls);
            loss_items[2] = (float)(s_dfl / tss * gain_dfl);
            loss_items[3] = (float)tss;
        }
    }
}

__global__ void loss_finalize_partials_kernel(const double *__restrict__ partials, float gain_box, float gain_cls,
                                              float gain_dfl, float *__restrict__ loss_items) {
    const double tss = partials[3] > 1.0 ? partials[3] : 1.0;
    loss_items[0] = (float)(partials[0] / tss * gain_box);
    loss_items[1] = (float)(partials[1] / tss * gain_cls);
    loss_items[2] = (float)(partials[2] / tss * gain_dfl);
    loss_items[3] = (float)tss;
}

struct LossWs {
    AssignWs aw;
    size_t off_boxes, off_pbce, off_pfg, total;
    int nb_stream, nb_fg;
};
static LossWs loss_ws_layout(int B, int A, int M, int nq_blocks) {
    LossWs w;
    w.aw = assign_ws_layout(B, A, M);
    w.nb_stream = nq_blocks * B;
    w.nb_fg = ((A + 255) / 256) * B;
    w.off_boxes = w.aw.total;
    w.off_pbce = w.off_boxes + a256(sizeof(float) * 4 * (size_t)B * A);
    w.off_pfg = w.off_pbce + a256(sizeof(double) * (size_t)w.nb_stream);
    w.total = w.off_pfg + a256(sizeof(double) * 4 * (size_t)w.nb_fg);
    return w;
}
size_t loss_workspace_bytes(int B, int A, int M) {
    // upper bound on stream blocks: scalar path, one quad per anchor, + one partial block per level
    return loss_ws_layout(B, A, M, (A + 31) / 32 + Y3D_MAX_LEVELS).total;
}

static bool vec4_ok(const LevelTable &t) {
    for (int l = 0; l < t.nl; ++l) {
        if ((t.h[l] * t.w[l]) % 4) return false;
        if (((uintptr_t)t.ptr[l]) % 16) return false;
        if (t.sB[l] % 4 || t.sC[l] % 4) return false;
    }
    return true;
}
static Quads make_quads(const LevelTable &t, int vec) {
    Quads qm;
    int q = 0;
    for (int l = 0; l <= Y3D_MAX_LEVELS; ++l) {
        qm.qstart[l] = q;
        if (l < t.nl) q += t.h[l] * t.w[l] / vec;
    }
    return qm;
}

static int launch_stream(const LevelTable &t, int B, int nc, int A, float *pd_bboxes, float *pd_scores,
                         double *part_bce, int *nblocks_x, cudaStream_t s) {
    int n_cls_roles = (nc + 15) / 16;
    if (n_cls_roles > 6) n_cls_roles = 6;
    int chunk = (nc + n_cls_roles - 1) / n_cls_roles;
    dim3 block(32, 2 + n_cls_roles);
    if (vec4_ok(t)) {
        Quads qm = make_quads(t, 4);
        dim3 grid((qm.qstart[t.nl] + 31) / 32, B);
        *nblocks_x = grid.x;
        loss_stream_kernel<4><<<grid, block, 0, s>>>(t, qm, nc, chunk, A, pd_bboxes, pd_scores, part_bce);
    } else {
        Quads qm = make_quads(t, 1);
        dim3 grid((qm.qstart[t.nl] + 31) / 32, B);
        *nblocks_x = grid.x;
        loss_stream_kernel<1><<<grid, block, 0, s>>>(t, qm, nc, chunk, A, pd_bboxes, pd_scores, part_bce);
    }
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

}  // namespace y3d

using namespace y3d;

extern "C" int y3d_train_decode(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC,
                                const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                                float *pd_bboxes, float *pd_scores, void *stream) {
    if (!lvl_ptr || !lvl_sB || !lvl_sC || !pd_bboxes || B < 0 || nc < 1) return Y3D_EINVAL;
    if (reg_max != kR) return Y3D_EUNSUPPORTED;
    LevelTable t;
    int A = make_level_table(t, lvl_ptr, lvl_sB, lvl_sC, lvl_hw, lvl_stride, nl);
    if (A < 0) return A;
    for (int l = 0; l < nl; ++l)
        if (!lvl_ptr[l]) return Y3D_EINVAL;
    if (B == 0) return Y3D_OK;
    int nbx;
    return launch_stream(t, B, nc, A, pd_bboxes, pd_scores, nullptr, &nbx, (cudaStream_t)stream);
}

extern "C" int y3d_v8_loss_fwd(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC,
                               const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                               const float *gt, int M, int topk, float gain_box, float gain_cls, float gain_dfl,
                               int normalise, float *loss_items, double *partials, uint8_t *dbg_fg_mask,
                               int32_t *dbg_target_gt_idx, void *const *prof_events, void *ws, size_t ws_bytes,
                               void *stream) {
    if (!lvl_ptr || !lvl_sB || !lvl_sC || B < 1 || nc < 1 || M < 0 || (M > 0 && !gt)) return Y3D_EINVAL;
    if (!loss_items && !partials) return Y3D_EINVAL;
    if (reg_max != kR) return Y3D_EUNSUPPORTED;
    AssignCtx c{};
    int A = make_level_table(c.t, lvl_ptr, lvl_sB, lvl_sC, lvl_hw, lvl_stride, nl);
    if (A < 0) return A;
    for (int l = 0; l < nl; ++l)
        if (!lvl_ptr[l]) return Y3D_EINVAL;
    if (topk < 1 || topk > A) return Y3D_EINVAL;
    if (topk > Y3D_MAX_TOPK) return Y3D_EUNSUPPORTED;
    const bool v4 = vec4_ok(c.t);
    const int nq = v4 ? A / 4 : A;
    LossWs w = loss_ws_layout(B, A, M, (nq + 31) / 32);
    if (!ws || ws_bytes < w.total) return Y3D_EWORKSPACE;
    if (((uintptr_t)ws) % 256) return Y3D_EALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    char *p = (char *)ws;
    float *pd_bboxes = (float *)(p + w.off_boxes);
    double *part_bce = (double *)(p + w.off_pbce);
    double *part_fg = (double *)(p + w.off_pfg);
    int nbx = 0;
    auto mark = [&](int i) {
        if (prof_events && prof_events[i]) cudaEventRecord((cudaEvent_t)prof_events[i], s);
    };
    mark(0);
    int rc = launch_stream(c.t, B, nc, A, pd_bboxes, nullptr, part_bce, &nbx, s);
    if (rc) return rc;
    mark(1);
    const int n_bce = nbx * B;
    if (M > 0) {
        c.score_mode = 1;
        c.cls_ch0 = 4 * kR;
        c.pd_bboxes = pd_bboxes; c.box_grid_units = 1;
        c.use_grid = 1;
        c.gt_labels = gt; c.gl_stride = 5;
        c.gt_bboxes = gt + 1; c.gb_stride = 5;
        c.mask_gt = nullptr;
        c.B = B; c.A = A; c.nc = nc; c.M = M; c.k = topk;
        c.alpha = 0.5f; c.beta = 6.0f; c.gamma = 1.0f; c.eps = 1e-9f;  // loss.py:176
        c.use_2d = 1; c.use_3d = 0; c.kps_l2 = 0; c.constrain = 1;
        assign_bind_ws(c, ws, w.aw);
        rc = assign_run_core(c, ws, w.aw, s, prof_events ? (cudaEvent_t)prof_events[2] : nullptr);
        if (rc) return rc;
        mark(3);
        dim3 grid((A + 255) / 256, B);
        loss_fg_kernel<<<grid, 256, 0, s>>>(c, gt, part_fg, dbg_fg_mask, dbg_target_gt_idx);
        Y3D_CHECK_LAUNCH();
        mark(4);
    } else {
        mark(2); mark(3); mark(4);
        if (dbg_fg_mask) cudaMemsetAsync(dbg_fg_mask, 0, (size_t)B * A, s);
        if (dbg_target_gt_idx) cudaMemsetAsync(dbg_target_gt_idx, 0, sizeof(int32_t) * (size_t)B * A, s);
    }
    loss_finalize_kernel<<<1, 256, 0, s>>>(part_bce, n_bce, M > 0 ? part_fg : nullptr, w.nb_fg, gain_box, gain_cls,
                                           gain_dfl, normalise, partials, loss_items);
    Y3D_CHECK_LAUNCH();
    mark(5);
    return Y3D_OK;
}

extern "C" int y3d_v8_loss_finalize(const double *partials, float gain_box, float gain_cls, float gain_dfl,
                                    float *loss_items, void *stream) {
    if (!partials || !loss_items) return Y3D_EINVAL;
    loss_finalize_partials_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(partials, gain_box, gain_cls, gain_dfl,
                                                                    loss_items);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}
