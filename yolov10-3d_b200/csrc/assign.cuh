// assign.cuh -- task-aligned assignment core shared by assign.cu (API-faithful assigners) and loss.cu (fused loss).
//
// Restates TaskAlignedAssigner / TaskAlignedAssigner3d (reference ultralytics/utils/tal.py:44-264, 391-700) without
// ever materialising the reference's dense [B, M, A] tensors:
//
//   tal_topk_kernel    one warp per (image, GT) -- or kTopkWarps warps per GT when every anchor is a candidate
//                      (no grid promised / constrain_anchors off).  Exact in-GT rectangle per level, flat centre-out
//                      walk, score / fast-IoU upper bounds before the exactly rounded CIoU, survivors compacted and
//                      evaluated on full lanes; the per-GT top-k is a lane-distributed sorted list of 64-bit keys
//                      (value desc, index asc); winners are claimed with one 64-bit atomic per (GT, anchor).  Details
//                      in assign.cu.
//                      Zero-metric ties are exact: anchors 0..k-1 always enter the list (they are what a dense
//                      stable top-k would pick among zeros), zero-metric anchors >= k can never be selected.
//   tal_resolve_kernel one thread per (image, anchor): 0 claims -> background; 1 claim -> that GT; >1 claims ->
//                      argmax over ALL M GTs of the overlap (select_highest_overlaps, tal.py:252-263, including its
//                      quirk that a GT which never selected the anchor can win); then the pair's (align, overlap)
//                      is recomputed and folded into per-GT maxima with integer atomicMax (values are >= 0).
//   emit kernels       norm = align * pos_overlap / (pos_align + eps) (tal.py:88-92) and the reference-format outputs.
#pragma once
#include "y3d_common.cuh"

namespace y3d {

constexpr int kTopkWarps = 4;  // warps (= GTs) per CTA of tal_topk_kernel
constexpr int kRecF4 = 3;     // float4 words of a claim record (AssignCtx::rec)

struct AssignCtx {
    // predictions
    int score_mode;              // 0: pd_scores tensor of probabilities; 1: head-level logits (sigmoid on the fly)
    const float *pd_scores;      // mode 0
    long long ssB, ssA, ssC;     // mode 0 element strides
    int cls_ch0;                 // mode 1: first class channel in the head tensor (4*reg_max)
    const float *pd_bboxes;      // [B,A,4] (box_soa == 0) or [B,4,A] (box_soa == 1); px when box_grid_units == 0,
                                 // grid units (multiplied by stride here) otherwise
    int box_grid_units;
    int box_soa;
    const float *anc;            // [A,2] px, or nullptr when use_grid
    LevelTable t;                // geometry (use_grid) and head pointers (score_mode 1)
    int use_grid;
    // ground truth, generic strides so that packed [B,M,5] / [B,M,17] rows work too
    const float *gt_labels; long long gl_stride;   // label of (b,m) at gt_labels[(b*M+m)*gl_stride]
    const float *gt_bboxes; long long gb_stride;   // box of (b,m) at gt_bboxes[(b*M+m)*gb_stride .. +4]
    const float *mask_gt;                          // [B,M] or nullptr => valid iff sum(box) > 0 (loss.py:226)
    int B, A, nc, M, k;
    float alpha, beta, gamma, eps;
    // 3D extras (TaskAlignedAssigner3d)
    int use_2d, use_3d, kps_l2, constrain;
    const float *pd_kps;  // [B,A,24]
    const float *gt_kps;  // [B,M,24]
    // workspace
    unsigned long long *claim;  // [B,A] zero-initialised; += (1<<32 | m) per claim: count in the high word and,
                                // when the count is 1, the claiming GT in the low word
    int *pos_align;    // [B,M] float bits, zero-initialised
    int *pos_ov;       // [B,M] float bits, zero-initialised
    int *tgi;          // [B,A] out of resolve: assigned GT or -1
    float *alignv;     // [B,A] out of resolve: align_metric of the assigned pair
    // optional (fused loss path): the first GT to claim an anchor appends it to the image's list, so that the
    // finishing kernel touches only claimed anchors.  list_count [B] zero-initialised, list_a [B, list_cap].
    int *list_count;
    int *list_a;
    int list_cap;
    // optional (fused loss, anchor-parallel finish): instead of the list, EVERY claim leaves a record of kRecF4 float4
    // at rec[(b * rec_cap + slot) * kRecF4] (slot from list_count[b]) with everything the finishing kernel needs for
    // this (anchor, GT) pair, gathered and evaluated here -- spread over this kernel's run time instead of one burst of
    // scattered DRAM reads and a latency chain at the very end of the step:
    //   [0] anchor | first-claimer bit << 31, GT index, alignment metric bits, logit of the GT's label
    //   [1] clamped exact CIoU (overlap), 1 - CIoU loss term, DFL loss term, 0     (claim_terms below)
    //   [2] predicted box xyxy (grid units)
    float4 *rec;
    int rec_cap;
    const float *lse;  // [B,A,4] written by the streaming kernel
    // optional, with rec: [B] zero-initialised; +1 (release) after a valid GT's records are written, so that the
    // finishing kernel can take an image as soon as all its GTs are through instead of waiting for the whole grid
    unsigned *topk_done;
};

// bbox2dist (tal.py:328-331) of anchor (gx, gy) in grid units against GT box gbox (px) at stride st, clamped like
// BboxLoss.forward (loss.py:93): the DFL target distance of every side; lower bin = (int)tt, its weight (lower + 1) - tt
__device__ __forceinline__ void dfl_target(float4 gbox, float st, float gx, float gy, float4 &tb, float (&tt)[4]) {
    tb = make_float4(dm::div(gbox.x, st), dm::div(gbox.y, st), dm::div(gbox.z, st), dm::div(gbox.w, st));  // loss.py:248
    const float ltrb[4] = {gx - tb.x, gy - tb.y, tb.z - gx, tb.w - gy};
#pragma unroll
    for (int side = 0; side < 4; ++side) tt[side] = fminf(fmaxf(ltrb[side], 0.0f), 15.0f - 0.01f);
}

struct GtRec {
    float4 box;
    float at1;
    int label;
    int valid;
};

// validity of a GT row alone (padded rows are skipped before any other work)
__device__ __forceinline__ bool gt_valid(const AssignCtx &c, int b, int m) {
    const long long i = (long long)b * c.M + m;
    if (c.mask_gt) return c.mask_gt[i] != 0.0f;
    const float *pb = c.gt_bboxes + i * c.gb_stride;
    return dm::add(dm::add(dm::add(pb[0], pb[1]), pb[2]), pb[3]) > 0.0f;
}

__device__ __forceinline__ GtRec load_gt(const AssignCtx &c, int b, int m) {
    GtRec g;
    long long i = (long long)b * c.M + m;
    const float *pb = c.gt_bboxes + i * c.gb_stride;
    g.box = make_float4(pb[0], pb[1], pb[2], pb[3]);
    g.label = (int)c.gt_labels[i * c.gl_stride];
    if (c.mask_gt)
        g.valid = c.mask_gt[i] != 0.0f;
    else
        g.valid = dm::add(dm::add(dm::add(g.box.x, g.box.y), g.box.z), g.box.w) > 0.0f;
    g.at1 = dm::box1_atan(g.box);
    return g;
}

__device__ __forceinline__ void anchor_px(const AssignCtx &c, int a, float &ax, float &ay, float &st) {
    if (c.use_grid) {
        int l = level_of(c.t, a);
        int cell = a - c.t.start[l];
        int w = c.t.w[l];
        st = c.t.stride[l];
        ax = dm::mul((float)(cell % w) + 0.5f, st);
        ay = dm::mul((float)(cell / w) + 0.5f, st);
    } else {
        float2 p = *reinterpret_cast<const float2 *>(c.anc + 2 * (long long)a);
        ax = p.x;
        ay = p.y;
        st = 1.0f;
    }
}

// The raw loads of one (GT, anchor) pair are split from the arithmetic so that callers can issue the loads of
// several pairs before evaluating any of them (memory-level parallelism: these are latency-bound gathers).
struct PairRaw {
    float4 box;  // as stored (grid units when c.box_grid_units)
    float s;     // probability (score_mode 0) or logit (score_mode 1)
};

__device__ __forceinline__ float pair_load_score(const AssignCtx &c, int b, int a, int label) {
    if (c.score_mode == 0) return c.pd_scores[b * c.ssB + a * c.ssA + (long long)label * c.ssC];
    const int l = level_of(c.t, a);
    return c.t.ptr[l][(long long)b * c.t.sB[l] + (long long)(c.cls_ch0 + label) * c.t.sC[l] + (a - c.t.start[l])];
}

__device__ __forceinline__ PairRaw pair_load_box(const AssignCtx &c, int b, int a) {
    PairRaw r;
    r.s = 0.0f;
    if (c.box_soa) {
        const float *p = c.pd_bboxes + (long long)b * 4 * c.A + a;
        r.box = make_float4(p[0], p[c.A], p[2 * (long long)c.A], p[3 * (long long)c.A]);
    } else {
        r.box = *reinterpret_cast<const float4 *>(c.pd_bboxes + ((long long)b * c.A + a) * 4);
    }
    return r;
}

__device__ __forceinline__ PairRaw pair_load(const AssignCtx &c, int b, int a, int label) {
    PairRaw r = pair_load_box(c, b, a);
    r.s = pair_load_score(c, b, a, label);
    return r;
}

__device__ __forceinline__ float pair_score(const AssignCtx &c, float raw) {
    if (c.score_mode == 0) return raw;
    return 1.0f / (1.0f + expf(-raw));  // pred_scores.detach().sigmoid() loss.py:232
}

__device__ __forceinline__ float4 pair_box(const AssignCtx &c, const PairRaw &r, int a) {
    float4 p = r.box;
    if (c.box_grid_units) {  // pred_bboxes.detach() * stride_tensor loss.py:233
        float st = c.t.stride[level_of(c.t, a)];
        p.x = dm::mul(p.x, st); p.y = dm::mul(p.y, st); p.z = dm::mul(p.z, st); p.w = dm::mul(p.w, st);
    }
    return p;
}

// keypoint similarity of one (prediction, GT) pair: keypoint_distance_3d tal.py:464-470, 1 / exp(mean distance)
// (flags bit2: squared distances, halved).  pk: the anchor's 24 keypoint coordinates, gk: the GT's.
__device__ __forceinline__ float kps_sim(const float4 *pk, const float4 *gk, int flags) {
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        float4 p = pk[j], q = __ldg(gk + j);
        float d0 = dm::sub(p.x, q.x), d1 = dm::sub(p.y, q.y), d2 = dm::sub(p.z, q.z), d3 = dm::sub(p.w, q.w);
        if (flags & 4) {
            acc = dm::add(acc, dm::mul(d0, d0)); acc = dm::add(acc, dm::mul(d1, d1));
            acc = dm::add(acc, dm::mul(d2, d2)); acc = dm::add(acc, dm::mul(d3, d3));
        } else {
            acc = dm::add(acc, fabsf(d0)); acc = dm::add(acc, fabsf(d1));
            acc = dm::add(acc, fabsf(d2)); acc = dm::add(acc, fabsf(d3));
        }
    }
    const float dist = dm::div(acc, 24.0f);
    return dm::div(1.0f, dm::exp_((flags & 4) ? dm::mul(0.5f, dist) : dist));
}

// get_box_metrics (tal.py:108-127) / get_box_kp_metrics / get_keypoint_metrics (tal.py:553-603) for one in-mask pair,
// given sb = score^alpha.  ovl is what the reference calls `overlaps` downstream: CIoU for the 2D assigner, the 3D
// similarity when use_3d.
// Out of line on purpose, with scalar arguments only (a reference to the kernel parameters would force a local
// copy): one copy of the CIoU / atan / pow / keypoint code per kernel keeps the instruction-cache footprint small.
// flags: bit0 use_2d, bit1 use_3d, bit2 kps_l2.
static __device__ __noinline__ float pair_metric_core(float4 gbox, float gat1, float4 pbox, float sb, float beta,
                                                      float gamma, int flags, const float4 *pk, const float4 *gk,
                                                      float *ovl_out) {
    float metric = sb;
    float ovl = 0.0f;
    if (flags & 1) {
        float o = dm::ciou(gbox, pbox, gat1);
        o = o < 0.0f ? 0.0f : o;  // clamp_(0) tal.py:131
        metric = dm::mul(metric, dm::pow_(o, beta));
        ovl = o;
    }
    if (flags & 2) {
        const float sim = kps_sim(pk, gk, flags);
        metric = dm::mul(metric, dm::pow_(sim, gamma));
        ovl = sim;  // tal.py:602-603
    }
    *ovl_out = ovl;
    return metric;
}

__device__ __forceinline__ float pair_metric(const AssignCtx &c, int b, int m, const GtRec &g, int a,
                                             const PairRaw &raw, float sb, float &ovl) {
    const int flags = (c.use_2d ? 1 : 0) | (c.use_3d ? 2 : 0) | (c.kps_l2 ? 4 : 0);
    const float4 *pk = nullptr, *gk = nullptr;
    if (c.use_3d) {
        pk = reinterpret_cast<const float4 *>(c.pd_kps + ((long long)b * c.A + a) * 24);
        gk = reinterpret_cast<const float4 *>(c.gt_kps + ((long long)b * c.M + m) * 24);
    }
    return pair_metric_core(g.box, g.at1, pair_box(c, raw, a), sb, c.beta, c.gamma, flags, pk, gk, &ovl);
}

__device__ __forceinline__ void pair_eval(const AssignCtx &c, int b, int m, const GtRec &g, int a, const PairRaw &raw,
                                          float &metric, float &ovl) {
    metric = pair_metric(c, b, m, g, a, raw, dm::pow_(pair_score(c, raw.s), c.alpha), ovl);
}

__device__ __forceinline__ void pair_eval(const AssignCtx &c, int b, int m, const GtRec &g, int a, float &metric,
                                          float &ovl) {
    pair_eval(c, b, m, g, a, pair_load(c, b, a, g.label), metric, ovl);
}

// norm_align_metric of an assigned anchor (tal.py:89-92)
__device__ __forceinline__ float assigned_norm(const AssignCtx &c, int b, int gi, float alignv) {
    float pa = __int_as_float(c.pos_align[(long long)b * c.M + gi]);
    float po = __int_as_float(c.pos_ov[(long long)b * c.M + gi]);
    return dm::div(dm::mul(alignv, po), dm::add(pa, c.eps));
}

// get_3d_keypoints (keypoint_utils.py:11-118) for one box; op order identical to oracle/y3d_oracle.c::y3d_o_keypoints
__device__ __forceinline__ void keypoints24(float c3x, float c3y, float dep, float s_h, float s_w, float s_l, int hbin,
                                            float hres, const float *cal, float *out) {
    using namespace dm;
    const float kPi = 3.14159265358979323846f, k2Pi = 6.283185307179586f;
    const float cu = cal[0], cv = cal[1], fu = cal[2], fv = cal[3], tx = cal[4], ty = cal[5];
    const float lx = add(div(mul(sub(c3x, cu), dep), fu), tx);  // img_to_rect :113-119
    const float ly = add(div(mul(sub(c3y, cv), dep), fv), ty);
    const float lz = dep;
    float alpha = add(mul((float)hbin, 0.5235987755982988f), hres);  // class2angle :42-47
    if (alpha > kPi) alpha = sub(alpha, k2Pi);
    float ry = add(alpha, atan2_(sub(c3x, cu), fu));  // alpha2ry :94-101
    if (ry > kPi) ry = sub(ry, k2Pi);
    if (ry < -kPi) ry = add(ry, k2Pi);
    float sx, cx, sy, cy;  // to_egoc_rot_mat :87-91  R = Rx(pi/2) @ Ry(-ry)
    sincos_(1.5707963267948966f, &sx, &cx);
    sincos_(-ry, &sy, &cy);
    const float R00 = cy, R01 = 0.0f, R02 = sy;
    const float R10 = mul(sx, sy), R11 = cx, R12 = -mul(sx, cy);
    const float R20 = -mul(cx, sy), R21 = sx, R22 = mul(cx, cy);
    const float hl = div(s_l, 2.0f), hw = div(s_w, 2.0f), hh = div(s_h, 2.0f);  // get_box_corners :20-26
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float px = (k & 2) ? -hl : hl;
        const float py = (k & 1) ? -hw : hw;
        const float pz = (k & 4) ? hh : -hh;
        out[3 * k + 0] = add(add(add(mul(px, R00), mul(py, R10)), mul(pz, R20)), lx);  // transform_to_camera :104-110
        out[3 * k + 1] = add(add(add(mul(px, R01), mul(py, R11)), mul(pz, R21)), ly);
        out[3 * k + 2] = add(add(add(mul(px, R02), mul(py, R12)), mul(pz, R22)), lz);
    }
}


// up to two branches (one2many / one2one of v10DetectLoss) run in the same launches, selected by blockIdx.z
struct AssignCtx2 {
    AssignCtx c[2];
    int *work_counter;  // zero-initialised: dynamic (branch, image, GT) work distribution of tal_topk_kernel
    // optional (fused loss): the valid GTs of every image sorted into kOrdClasses size classes, biggest first --
    // ord_cnt [B][4] = GTs per class, ord_list [B][M] = 32-byte records (GT index, label, box).  The dynamic distribution then hands out the
    // expensive GTs first (longest processing time first) and never sees a padded row.
    const int *ord_cnt, *ord_list;
};
constexpr int kOrdClasses = 4;
constexpr int kOrdMaxSeg = 2047;  // segments = classes x branches x images the top-k kernel's prefix table holds

// ----------------------------------------------------------------------------------------------------------------
// Top-k entries are 64-bit keys: metric bits (>= 0, so integer order == float order) | 0x7fffffff - anchor | in-GT bit.
// A larger key is a better entry (value desc, index asc); 0 is the empty slot.
__device__ __forceinline__ unsigned long long tk_key(float metric, int a, int in) {
    return ((unsigned long long)__float_as_uint(metric) << 32) | ((unsigned long long)(0x7fffffff - a) << 1) |
           (unsigned long long)(in & 1);
}
__device__ __forceinline__ int tk_anchor(unsigned long long key) { return 0x7fffffff - (int)((key & 0xffffffffull) >> 1); }

// bbox_iou(box1, box2, xywh=False, CIoU=True) (metrics.py:96-131) with fast division: for VALUES (loss terms, alignment
// weights), never for anything that decides an index -- those go through the exactly rounded dm::ciou.
__device__ __forceinline__ float ciou_fast(float4 b1, float4 b2) {
    const float eps = 1e-7f;
    const float w1 = b1.z - b1.x, h1 = b1.w - b1.y + eps, w2 = b2.z - b2.x, h2 = b2.w - b2.y + eps;
    const float iw = fmaxf(fminf(b1.z, b2.z) - fmaxf(b1.x, b2.x), 0.f), ih = fmaxf(fminf(b1.w, b2.w) - fmaxf(b1.y, b2.y), 0.f);
    const float inter = iw * ih;
    const float uni = w1 * h1 + w2 * h2 - inter + eps;
    const float iou = __fdividef(inter, uni);
    const float cw = fmaxf(b1.z, b2.z) - fminf(b1.x, b2.x), ch = fmaxf(b1.w, b2.w) - fminf(b1.y, b2.y);
    const float c2 = cw * cw + ch * ch + eps;
    const float dx = b2.x + b2.z - b1.x - b1.z, dy = b2.y + b2.w - b1.y - b1.w;
    const float rho2 = (dx * dx + dy * dy) * 0.25f;
    const float da = atanf(__fdividef(w2, h2)) - atanf(__fdividef(w1, h1));
    const float v = 0.4052847345693511f * da * da;
    const float alpha = __fdividef(v, v - iou + (1.0f + eps));
    return iou - (__fdividef(rho2, c2) + v * alpha);
}

// What a claim record carries besides the pair's identity (assign.cuh, AssignCtx::rec): the loss inputs of (anchor a of
// image b, GT g), gathered from the head tensor and the streaming kernel's planes, and the terms that depend on nothing
// else: the clamped exact CIoU (the pair's `overlaps` entry, tal.py:131), 1 - CIoU(pred, target) in grid units
// (BboxLoss.forward loss.py:85-86) and the DFL cross-entropy averaged over the four sides (_df_loss loss.py:99-113).
struct ClaimTerms {
    float4 terms;  // overlap, 1 - CIoU, DFL, 0
    float4 box;    // predicted box, grid units
    float xlab;    // logit of the GT's class
};
__device__ __forceinline__ ClaimTerms claim_terms(const AssignCtx &c, int b, int a, const GtRec &g) {
    ClaimTerms r;
    const int lv = level_of(c.t, a);
    const int cell = a - c.t.start[lv];
    const float st = c.t.stride[lv];
    const float gx = (float)(cell % c.t.w[lv]) + 0.5f, gy = (float)(cell / c.t.w[lv]) + 0.5f;
    float4 tb;
    float tt[4];
    dfl_target(g.box, st, gx, gy, tb, tt);
    const float *hp = c.t.ptr[lv] + (long long)b * c.t.sB[lv] + cell;
    const long long cs = c.t.sC[lv];
    const int t0 = (int)tt[0], t1 = (int)tt[1], t2 = (int)tt[2], t3 = (int)tt[3];
    const float xl[4] = {hp[(long long)t0 * cs], hp[(long long)(16 + t1) * cs], hp[(long long)(32 + t2) * cs],
                         hp[(long long)(48 + t3) * cs]};
    const float xr[4] = {hp[(long long)(t0 + 1) * cs], hp[(long long)(17 + t1) * cs], hp[(long long)(33 + t2) * cs],
                         hp[(long long)(49 + t3) * cs]};
    r.xlab = hp[(long long)(c.cls_ch0 + (g.label < 0 ? 0 : g.label)) * cs];
    r.box = reinterpret_cast<const float4 *>(c.pd_bboxes)[(long long)b * c.A + a];
    const float4 l4 = reinterpret_cast<const float4 *>(c.lse)[(long long)b * c.A + a];
    const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
    float o = dm::ciou(g.box, make_float4(dm::mul(r.box.x, st), dm::mul(r.box.y, st), dm::mul(r.box.z, st), dm::mul(r.box.w, st)), g.at1);
    o = o < 0.0f ? 0.0f : o;
    float dfl = 0.f;
#pragma unroll
    for (int side = 0; side < 4; ++side) {
        const float wl = (float)((int)tt[side] + 1) - tt[side];
        dfl += (ls[side] - xl[side]) * wl + (ls[side] - xr[side]) * (1.0f - wl);
    }
    r.terms = make_float4(o, 1.0f - ciou_fast(r.box, tb), dfl * 0.25f, 0.f);
    return r;
}

__device__ __forceinline__ void red_release_add1(unsigned *p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// host helpers --------------------------------------------------------------------------------------------------
struct AssignWs {
    size_t off_cnt, off_pa, off_po, off_work, off_tgi, off_align, off_norm, off_lab, total, zero_bytes;
};
inline size_t a256(size_t v) { return (v + 255) & ~(size_t)255; }
inline AssignWs assign_ws_layout(int B, int A, int M) {
    AssignWs w;
    size_t ba = a256(sizeof(int) * (size_t)B * A), bm = a256(sizeof(int) * (size_t)B * (M > 0 ? M : 1));
    w.off_cnt = 0;
    w.off_pa = 2 * ba;
    w.off_po = 2 * ba + bm;
    w.zero_bytes = 2 * ba + 2 * bm + 256;  // claim | pos_align | pos_ov | work counter: zero-filled with one memset
    w.off_work = 2 * ba + 2 * bm;
    w.off_tgi = w.zero_bytes;
    w.off_align = w.off_tgi + ba;
    w.off_norm = w.off_align + ba;
    w.off_lab = w.off_norm + ba;
    w.total = w.off_lab + ba;
    return w;
}
inline void assign_bind_ws(AssignCtx &c, void *ws, const AssignWs &w) {
    char *p = (char *)ws;
    c.claim = (unsigned long long *)(p + w.off_cnt);
    c.pos_align = (int *)(p + w.off_pa);
    c.pos_ov = (int *)(p + w.off_po);
    c.tgi = (int *)(p + w.off_tgi);
    c.alignv = (float *)(p + w.off_align);
}
// enqueue top-k + resolve for n (1 or 2) branches of identical B, A, M (defined in assign.cu).  The caller has
// zero-filled [off_cnt, off_cnt + zero_bytes) of every branch's workspace.  After this c.tgi / c.alignv / c.pos_*
// are final.
int assign_run_core(const AssignCtx2 &cc, int n, cudaStream_t s, cudaEvent_t after_topk = nullptr, bool pdl = false);
// GT keypoints [B,M,24] from packed [B,M,17] rows (add_cls_mean_size tal.py:605-609 + get_3d_keypoints)
int launch_kps_gt(const float *gts, const float *calibs, const float *mean_sizes, int B, int M, int nc, float *gt_kps,
                  cudaStream_t s);
// top-k + claims only (the fused loss path finishes with its own kernel)
int assign_run_topk(const AssignCtx2 &cc, int n, cudaStream_t s, bool pdl = false);
// the same for the fused loss' configuration (logits from the head, alpha 0.5, beta 6, in-GT constraint, grid, [B,A,4]
// boxes in grid units): specialised kernel in topk_fused.cu.  Returns Y3D_EUNSUPPORTED when the context does not match.
int assign_run_topk_fused(const AssignCtx2 &cc, int n, cudaStream_t s, bool pdl);

}  // namespace y3d
