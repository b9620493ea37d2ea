// loss_bwd.cu -- backward of the fused v8DetectionLoss / v10DetectLoss (sm_100a).
//   y3d_v10_loss_bwd / y3d_v8_loss_bwd : d(loss items) / d(head tensors), i.e. what autograd produces for
//       reference ultralytics/utils/loss.py:206-257 (BCEWithLogits :240, BboxLoss.forward :82-96, _df_loss :99-113,
//       bbox_decode :197-204) with bbox_iou(..., CIoU=True) of ultralytics/utils/metrics.py:78-134 (alpha under
//       no_grad :128-129).  The assignment (tal.py:44, @torch.no_grad) is a constant of the backward pass.
//
// The gradient has the shape of the head tensors, [B, 4R+nc, h_l, w_l] per level and branch:
//   * class rows (dense):  g_cls * gain_cls / tss * (sigmoid(x) - t),  t = alignment weight at the assigned label
//   * box rows: zero except at foreground anchors, where the CIoU term flows through dist2bbox and the softmax
//     expectation of the 16 DFL bins, plus the DFL cross-entropy term.
// One streaming kernel (reads the class logits, writes every row: 4*nc*A read + 4*(4R+nc)*A written per image and
// branch).  The forward pass left, in the claim word of every foreground anchor, its GT index and alignment weight
// (loss.cu, loss_finish_kernel): the thread that owns an anchor's rows looks the word up and patches the few
// foreground values itself, right after its own full-width stores, so that no second pass has to read-modify-write
// 4-byte pieces of sectors that have already left the L2.
#include "loss.cuh"

namespace y3d {

constexpr int kBwdThreads = 128;
constexpr int kBwdRows = 10;     // class rows in flight per thread (as in the forward streaming kernel)

struct BwdParams {
    LevelTable t[2];                       // head tensors (inputs of the forward pass)
    float *g_ptr[2][Y3D_MAX_LEVELS];       // gradient tensors, one per level and branch
    long long g_sB[2][Y3D_MAX_LEVELS], g_sC[2][Y3D_MAX_LEVELS];
    const float *items;                    // DEVICE float[4 * n_branch] of the forward pass: tss at [4z + 3]
    const float *gitems;                   // DEVICE float[3 * n_branch]: d total / d (box, cls, dfl) item
    float gain_box, gain_cls, gain_dfl;
    int n_branch, B, nc, A, M, cap;
    const float *gt5;                      // [B,M,5]
    const float *boxes[2];                 // forward workspace: [B,A,4] xyxy grid units
    const unsigned long long *claim[2];    // forward workspace: [B,A]; bit 63 = foreground, GT index << 32 | weight bits
};

template <int V>
__device__ __forceinline__ void st_zero(float *p) {
    if constexpr (V == 4) *reinterpret_cast<float4 *>(p) = make_float4(0.f, 0.f, 0.f, 0.f);
    else p[0] = 0.f;
}

// d CIoU(box1 = pred, box2 = target) / d pred, alpha constant (metrics.py:96-131).  min/max ties split the gradient
// like torch.minimum / torch.maximum; clamp_(0) passes the gradient where its input is >= 0.
__device__ __forceinline__ void ciou_grad(float4 p, float4 t, float (&d)[4]) {
    const float eps = 1e-7f;
    const float w1 = p.z - p.x, h1 = p.w - p.y + eps;
    const float w2 = t.z - t.x, h2 = t.w - t.y + eps;
    const float iwr = fminf(p.z, t.z) - fmaxf(p.x, t.x), ihr = fminf(p.w, t.w) - fmaxf(p.y, t.y);
    const float iw = fmaxf(iwr, 0.f), ih = fmaxf(ihr, 0.f);
    const float inter = iw * ih;
    const float uni = w1 * h1 + w2 * h2 - inter + eps;
    const float iou = inter / uni;
    const float cw = fmaxf(p.z, t.z) - fminf(p.x, t.x), ch = fmaxf(p.w, t.w) - fminf(p.y, t.y);
    const float c2 = cw * cw + ch * ch + eps;
    const float sx = t.x + t.z - p.x - p.z, sy = t.y + t.w - p.y - p.w;
    const float rho2 = (sx * sx + sy * sy) * 0.25f;
    const float D = atanf(w2 / h2) - atanf(w1 / h1);
    const float kv = 0.4052847345693511f;  // 4 / pi^2
    const float v = kv * D * D;
    const float alpha = v / (v - iou + (1.0f + eps));
    auto gt_w = [](float a, float b) { return a > b ? 1.0f : (a == b ? 0.5f : 0.0f); };  // d max(a,b)/da, d min(b,a)... see use
    const float miw = iwr >= 0.f ? 1.f : 0.f, mih = ihr >= 0.f ? 1.f : 0.f;
    // d iw / d (x1, x2), d ih / d (y1, y2)
    const float diw_x1 = -gt_w(p.x, t.x) * miw, diw_x2 = gt_w(t.z, p.z) * miw;
    const float dih_y1 = -gt_w(p.y, t.y) * mih, dih_y2 = gt_w(t.w, p.w) * mih;
    const float dint[4] = {ih * diw_x1, iw * dih_y1, ih * diw_x2, iw * dih_y2};
    const float dwh[4] = {-h1, -w1, h1, w1};  // d (w1*h1)
    // d cw / d (x1, x2), d ch / d (y1, y2): cw = max(x2,tx2) - min(x1,tx1)
    const float dcw_x1 = -gt_w(t.x, p.x), dcw_x2 = gt_w(p.z, t.z);
    const float dch_y1 = -gt_w(t.y, p.y), dch_y2 = gt_w(p.w, t.w);
    const float dc2[4] = {2.f * cw * dcw_x1, 2.f * ch * dch_y1, 2.f * cw * dcw_x2, 2.f * ch * dch_y2};
    const float drho[4] = {-0.5f * sx, -0.5f * sy, -0.5f * sx, -0.5f * sy};
    const float wh2 = w1 * w1 + h1 * h1;
    const float dvc = 2.f * kv * D / wh2;
    const float dv[4] = {dvc * h1, -dvc * w1, -dvc * h1, dvc * w1};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float duni = dwh[i] - dint[i];
        const float diou = (dint[i] * uni - inter * duni) / (uni * uni);
        const float dpen = (drho[i] * c2 - rho2 * dc2[i]) / (c2 * c2);
        d[i] = diou - dpen - alpha * dv[i];
    }
}

// Gradient of the rows of foreground anchor `a` that belong to `side`: the 16 DFL bins of that side and, for side 0,
// the class row of the assigned label.
__device__ __noinline__ void bwd_fg_patch(const BwdParams &P, int z, int b, int a, int side, unsigned long long cl) {
    const LevelTable &t = P.t[z];
    const int l = level_of(t, a);
    const int cell = a - t.start[l];
    const long long cs = t.sC[l], gs = P.g_sC[z][l];
    const float *x = t.ptr[l] + (long long)b * t.sB[l] + cell;
    float *g = P.g_ptr[z][l] + (long long)b * P.g_sB[z][l] + cell;
    const int A = P.A;
    const int gi = (int)((cl >> 32) & 0x7fffffffull);
    const float wgt = __uint_as_float((unsigned)(cl & 0xffffffffull));
    const float st = t.stride[l];
    const float ax = (float)(cell % t.w[l]) + 0.5f, ay = (float)(cell / t.w[l]) + 0.5f;
    const float *g5 = P.gt5 + ((long long)b * P.M + gi) * 5;
    int lab = (int)g5[0];
    lab = lab < 0 ? 0 : lab;
    const float4 tb = make_float4(g5[1] / st, g5[2] / st, g5[3] / st, g5[4] / st);
    const float4 pb = reinterpret_cast<const float4 *>(P.boxes[z])[(long long)b * A + a];
    const float tss = P.items[4 * z + 3];
    const float c_box = P.gitems[3 * z + 0] * P.gain_box / tss;
    const float c_cls = P.gitems[3 * z + 1] * P.gain_cls / tss;
    const float c_dfl = P.gitems[3 * z + 2] * P.gain_dfl / tss;
    // this side's 16 bins: softmax, expectation
    const float *xs = x + (long long)(side * kR) * cs;
    float v[kR];
    float m = -3.4e38f;
#pragma unroll
    for (int j = 0; j < kR; ++j) {
        v[j] = xs[(long long)j * cs];
        m = fmaxf(m, v[j]);
    }
    float s = 0.f, ex = 0.f;
#pragma unroll
    for (int j = 0; j < kR; ++j) {
        v[j] = __expf(v[j] - m);
        s += v[j];
        ex += (float)j * v[j];
    }
    const float inv = 1.0f / s;
    ex *= inv;
    // d loss / d dist_side through (1 - CIoU) * w: x1 = ax - d0, y1 = ay - d1, x2 = ax + d2, y2 = ay + d3
    float dc[4];
    ciou_grad(pb, tb, dc);
    const float dcs = side == 0 ? dc[0] : side == 1 ? dc[1] : side == 2 ? dc[2] : dc[3];
    const float gd = -wgt * c_box * dcs * (side < 2 ? -1.0f : 1.0f);
    // DFL target of this side (bbox2dist tal.py:328-331, _df_loss loss.py:99-113)
    const float ltrb = side == 0 ? ax - tb.x : side == 1 ? ay - tb.y : side == 2 ? tb.z - ax : tb.w - ay;
    const float tt = fminf(fmaxf(ltrb, 0.0f), (float)(kR - 1) - 0.01f);
    const int tl = (int)tt;
    const float wl = (float)(tl + 1) - tt, wr = 1.0f - wl;
    const float cd = c_dfl * wgt * 0.25f;
    float *gp = g + (long long)(side * kR) * gs;
#pragma unroll
    for (int j = 0; j < kR; ++j) {
        const float p = v[j] * inv;
        float gx = gd * p * ((float)j - ex) + cd * p;
        if (j == tl) gx -= cd * wl;
        if (j == tl + 1) gx -= cd * wr;
        gp[(long long)j * gs] = gx;
    }
    if (side == 0) {  // BCE target: t = w at the assigned label
        const float xv = x[(long long)(4 * kR + lab) * cs];
        g[(long long)(4 * kR + lab) * gs] = c_cls * __fdividef(1.0f, 1.0f + __expf(-xv)) - c_cls * wgt;
    }
}

// grid (ceil(A/V/32), B, n_branch), block 128 = 32 units of V anchors x 4 channel parts (same decomposition as the
// forward streaming kernel): part p writes the 16 DFL rows of side p (zeros) and a quarter of the class rows.  The CTA's
// foreground anchors are collected in shared memory and patched after a barrier, one (anchor, side) pair per thread, so
// that the rare path runs on full warps.  (Several tiles per CTA, or persistent CTAs striding over the tiles, measured
// 10-30 % slower: consecutive CTAs on consecutive 512-byte pieces of each row is what keeps the DRAM pages open.)
template <int V>
__global__ void __launch_bounds__(kBwdThreads) loss_bwd_kernel(const __grid_constant__ BwdParams P) {
    __shared__ int s_n;
    __shared__ int s_a[32 * V];
    __shared__ unsigned long long s_cl[32 * V];
    const int z = blockIdx.z, b = blockIdx.y;
    const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int q = blockIdx.x * 32 + lane;
    if (q * V < P.A) {
        const LevelTable &t = P.t[z];
        const int a0 = q * V;
        const int l = level_of(t, a0);
        const int cell = a0 - t.start[l];
        const float *x = t.ptr[l] + (long long)b * t.sB[l] + cell;
        float *g = P.g_ptr[z][l] + (long long)b * P.g_sB[z][l] + cell;
        const long long cs = t.sC[l], gs = P.g_sC[z][l];
        const float coef = P.gitems[3 * z + 1] * P.gain_cls / P.items[4 * z + 3];
        if (part == 0 && P.claim[z]) {
            const unsigned long long *cp = P.claim[z] + (long long)b * P.A + a0;
            unsigned long long cl[V];
            if constexpr (V == 4) {
                const ulonglong2 c01 = __ldg(reinterpret_cast<const ulonglong2 *>(cp));
                const ulonglong2 c23 = __ldg(reinterpret_cast<const ulonglong2 *>(cp) + 1);
                cl[0] = c01.x; cl[1] = c01.y; cl[2] = c23.x; cl[3] = c23.y;
            } else {
                cl[0] = __ldg(cp);
            }
#pragma unroll
            for (int i = 0; i < V; ++i)
                if (cl[i] >> 63) {
                    const int k = atomicAdd(&s_n, 1);
                    s_a[k] = a0 + i;
                    s_cl[k] = cl[i];
                }
        }
#pragma unroll
        for (int j = 0; j < kR; ++j) st_zero<V>(g + (long long)(part * kR + j) * gs);
        const int cpp = (P.nc + 3) >> 2;
        const int c_lo = part * cpp, c_hi = min(P.nc, c_lo + cpp);
        constexpr int CB = kBwdRows;  // class rows in flight per thread
        for (int c0 = c_lo; c0 < c_hi; c0 += CB) {
            const float *px = x + (long long)(4 * kR + c0) * cs;
            float *pg = g + (long long)(4 * kR + c0) * gs;
            const int n = min(CB, c_hi - c0);
            if constexpr (V == 4) {
                float4 v[CB];
                if (n == CB) {
#pragma unroll
                    for (int j = 0; j < CB; ++j) v[j] = ldg_stream4(px + (long long)j * cs);
                } else {
#pragma unroll
                    for (int j = 0; j < CB; ++j)
                        if (j < n) v[j] = ldg_stream4(px + (long long)j * cs);
                }
#pragma unroll
                for (int j = 0; j < CB; ++j) {
                    if (j < n) {
                        float4 o;
                        o.x = coef * __fdividef(1.0f, 1.0f + __expf(-v[j].x));
                        o.y = coef * __fdividef(1.0f, 1.0f + __expf(-v[j].y));
                        o.z = coef * __fdividef(1.0f, 1.0f + __expf(-v[j].z));
                        o.w = coef * __fdividef(1.0f, 1.0f + __expf(-v[j].w));
                        *reinterpret_cast<float4 *>(pg + (long long)j * gs) = o;
                    }
                }
            } else {
                for (int j = 0; j < n; ++j)
                    pg[(long long)j * gs] = coef * __fdividef(1.0f, 1.0f + __expf(-ldg_stream1(px + (long long)j * cs)));
            }
        }
    }
    __syncthreads();  // orders the patches below after every thread's full-width stores above
    const int n_items = 4 * s_n;
    for (int it = threadIdx.x; it < n_items; it += kBwdThreads) bwd_fg_patch(P, z, b, s_a[it >> 2], it & 3, s_cl[it >> 2]);
}

static bool vec4_ok_bwd(const BwdParams &P, int z) {
    const LevelTable &t = P.t[z];
    for (int l = 0; l < t.nl; ++l) {
        if ((t.h[l] * t.w[l]) % 4) return false;
        if (((uintptr_t)t.ptr[l]) % 16 || ((uintptr_t)P.g_ptr[z][l]) % 16) return false;
        if (t.sB[l] % 4 || t.sC[l] % 4 || P.g_sB[z][l] % 4 || P.g_sC[z][l] % 4) return false;
    }
    return true;
}

struct BranchBwd {
    const float *const *lvl_ptr;
    const int64_t *sB, *sC;
    float *const *g_ptr;
    const int64_t *g_sB, *g_sC;
    int topk;
};

static int loss_bwd_run(int nb, const BranchBwd *br, const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc,
                        int reg_max, const float *gt, int M, float gain_box, float gain_cls, float gain_dfl,
                        const float *loss_items, const float *grad_items, const void *ws, size_t ws_bytes,
                        void *stream) {
    if (!lvl_hw || !lvl_stride || B < 1 || nc < 1 || M < 0 || (M > 0 && !gt)) return Y3D_EINVAL;
    if (!loss_items || !grad_items) return Y3D_EINVAL;
    if (reg_max != kR) return Y3D_EUNSUPPORTED;
    BwdParams P{};
    int A = 0, kmax = 1;
    for (int z = 0; z < nb; ++z) {
        if (!br[z].lvl_ptr || !br[z].sB || !br[z].sC || !br[z].g_ptr || !br[z].g_sB || !br[z].g_sC) return Y3D_EINVAL;
        A = make_level_table(P.t[z], br[z].lvl_ptr, br[z].sB, br[z].sC, lvl_hw, lvl_stride, nl);
        if (A < 0) return A;
        for (int l = 0; l < nl; ++l) {
            if (!br[z].lvl_ptr[l] || !br[z].g_ptr[l]) return Y3D_EINVAL;
            P.g_ptr[z][l] = br[z].g_ptr[l];
            P.g_sB[z][l] = br[z].g_sB[l];
            P.g_sC[z][l] = br[z].g_sC[l];
        }
        if (br[z].topk > kmax) kmax = br[z].topk;
    }
    const LossWs w = loss_ws_layout(nb, B, A, M, kmax);  // must be the layout of the forward call
    if (!ws || ws_bytes < w.total) return Y3D_EWORKSPACE;
    if (((uintptr_t)ws) % 256) return Y3D_EALIGN;
    const char *p = (const char *)ws;
    for (int z = 0; z < nb; ++z) {
        const char *q = p + z * w.per_branch;
        P.boxes[z] = (const float *)(q + w.boxes);
        P.claim[z] = M > 0 ? (const unsigned long long *)(q + w.claim) : nullptr;
    }
    P.items = loss_items; P.gitems = grad_items;
    P.gain_box = gain_box; P.gain_cls = gain_cls; P.gain_dfl = gain_dfl;
    P.n_branch = nb; P.B = B; P.nc = nc; P.A = A; P.M = M; P.cap = w.cap; P.gt5 = gt;
    cudaStream_t s = (cudaStream_t)stream;
    const bool v4 = vec4_ok_bwd(P, 0) && (nb < 2 || vec4_ok_bwd(P, 1));
    const int units = v4 ? A / 4 : A;
    dim3 grid((units + 31) / 32, B, nb);
    if (v4) loss_bwd_kernel<4><<<grid, kBwdThreads, 0, s>>>(P);
    else loss_bwd_kernel<1><<<grid, kBwdThreads, 0, s>>>(P);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

}  // namespace y3d

using namespace y3d;

extern "C" int y3d_v8_loss_bwd(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC,
                               float *const *grad_ptr, const int64_t *grad_sB, const int64_t *grad_sC,
                               const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                               const float *gt, int M, int topk, float gain_box, float gain_cls, float gain_dfl,
                               const float *loss_items, const float *grad_items, const void *ws, size_t ws_bytes,
                               void *stream) {
    BranchBwd br[1] = {{lvl_ptr, lvl_sB, lvl_sC, grad_ptr, grad_sB, grad_sC, topk}};
    return loss_bwd_run(1, br, lvl_hw, lvl_stride, nl, B, nc, reg_max, gt, M, gain_box, gain_cls, gain_dfl, loss_items,
                        grad_items, ws, ws_bytes, stream);
}

extern "C" int y3d_v10_loss_bwd(const float *const *o2m_ptr, const int64_t *o2m_sB, const int64_t *o2m_sC,
                                float *const *o2m_grad, const int64_t *o2m_gsB, const int64_t *o2m_gsC,
                                const float *const *o2o_ptr, const int64_t *o2o_sB, const int64_t *o2o_sC,
                                float *const *o2o_grad, const int64_t *o2o_gsB, const int64_t *o2o_gsC,
                                const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max,
                                const float *gt, int M, int topk_o2m, int topk_o2o, float gain_box, float gain_cls,
                                float gain_dfl, const float *loss_items, const float *grad_items, const void *ws,
                                size_t ws_bytes, void *stream) {
    BranchBwd br[2] = {{o2m_ptr, o2m_sB, o2m_sC, o2m_grad, o2m_gsB, o2m_gsC, topk_o2m},
                       {o2o_ptr, o2o_sB, o2o_sC, o2o_grad, o2o_gsB, o2o_gsC, topk_o2o}};
    return loss_bwd_run(2, br, lvl_hw, lvl_stride, nl, B, nc, reg_max, gt, M, gain_box, gain_cls, gain_dfl, loss_items,
                        grad_items, ws, ws_bytes, stream);
}
