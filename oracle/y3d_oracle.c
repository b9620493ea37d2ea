/*
 * y3d_oracle.c -- CPU restatement of the YOLOv10(-3D) head hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker*: it is compiled by
 * oracle/build_oracle.py (gcc, -ffp-contract=off) into oracle/_build/liby3d_oracle.so and is
 * called only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  Nothing in the product package (yolov10-3d_b200/) links, loads or calls it.
 *
 * It restates, in plain scalar C with dense [M x A] loops exactly like the reference's dense
 * tensors, the algorithm of (all paths relative to the reference repo baldhat/yolov10-3D):
 *   make_anchors                     ultralytics/utils/tal.py:300-312
 *   Detect.inference / DFL / dist2bbox  ultralytics/nn/modules/head.py:53-79, block.py:59-62, tal.py:315-325
 *   ops.v10postprocess / v10_3Dpostprocess   ultralytics/utils/ops.py:852-880
 *   bbox_iou (CIoU)                  ultralytics/utils/metrics.py:78-134
 *   TaskAlignedAssigner.forward      ultralytics/utils/tal.py:44-264
 *   v8DetectionLoss.__call__ / BboxLoss   ultralytics/utils/loss.py:82-113,197-257
 *   v10Detect3d.decode               ultralytics/nn/modules/head.py:755-764
 *   TaskAlignedAssigner3d.forward    ultralytics/utils/tal.py:391-700 + utils/keypoint_utils.py:11-118
 *   KITTIDataset.decode_preds        ultralytics/data/datasets/kitti.py:519-576
 *
 * Parity pinning: the reference has no golden vectors for this path, so this
 * restatement is pinned against outputs of the reference itself, generated in the build container
 * by tests/golden/make_golden.py and committed under tests/golden/ (see tests/test_oracle_vs_golden.py).
 *
 * Arithmetic contract.  Index-deciding arithmetic (in-GT test, CIoU, alignment metric, 3D keypoint
 * similarity) is a fixed sequence of individually rounded IEEE-754 binary32 operations
 * (+ - * / sqrt), never contracted into FMAs, so that a GPU implementation issuing the same
 * sequence with round-to-nearest intrinsics is *bit-identical*.  The transcendental functions the
 * reference takes from torch (atan, sin, cos, atan2, exp, pow) are restated here as explicit
 * polynomial sequences (y3d_atanf etc., classic Cephes single-precision forms, <= 2 ulp), which
 * is within the reference's own CPU-vs-CUDA spread and far inside the 1e-5 relative value bar.
 * Tie-breaking everywhere: lowest index first (stable descending order), per BASELINE.json.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__FP_FAST_FMAF) && !defined(Y3D_ALLOW_FMA)
/* the build recipe passes -ffp-contract=off; this is a belt-and-braces reminder */
#endif

#define Y3D_PI_F 3.14159265358979323846f

/* ------------------------------------------------------------------------------------------ */
/* deterministic single-precision math (spec shared, by documentation, with csrc/y3d_math.cuh) */
/* ------------------------------------------------------------------------------------------ */

/* atan: Cephes atanf form. 2 range reductions + degree-4 odd polynomial; every op rounded. */
float y3d_atanf(float x) {
    float ax = fabsf(x);
    float y0, z;
    if (ax > 2.414213562373095f) {         /* tan(3pi/8) */
        y0 = 1.5707963267948966f;
        z = -(1.0f / ax);
    } else if (ax > 0.4142135623730950f) { /* tan(pi/8) */
        y0 = 0.7853981633974483f;
        z = (ax - 1.0f) / (ax + 1.0f);
    } else {
        y0 = 0.0f;
        z = ax;
    }
    float zz = z * z;
    float p = 8.05374449538e-2f;
    p = p * zz;
    p = p - 1.38776856032e-1f;
    p = p * zz;
    p = p + 1.99777106478e-1f;
    p = p * zz;
    p = p - 3.33329491539e-1f;
    p = p * zz;
    p = p * z;
    float r = y0 + (p + z);
    return copysignf(r, x);
}

/* atan2 built on y3d_atanf (only finite, non-both-zero inputs occur on the path). */
float y3d_atan2f(float y, float x) {
    if (x == 0.0f) {
        if (y > 0.0f) return 1.5707963267948966f;
        if (y < 0.0f) return -1.5707963267948966f;
        return 0.0f;
    }
    float t = y3d_atanf(y / x);
    if (x > 0.0f) return t;
    if (y >= 0.0f) return t + Y3D_PI_F;
    return t - Y3D_PI_F;
}

/* sin / cos for |x| <= ~2*pi+: Cephes sinf/cosf form (octant reduction with 3-part pi/4). */
static void y3d_sincos_core(float x, float *s_out, float *c_out) {
    float ax = fabsf(x);
    int sign_s = x < 0.0f ? -1 : 1;
    int sign_c = 1;
    /* j = integer part of ax/(pi/4), made even */
    int j = (int)(ax * 1.27323954473516f);
    float y = (float)j;
    if (j & 1) {
        j += 1;
        y = y + 1.0f;
    }
    j &= 7;
    if (j > 3) {
        sign_s = -sign_s;
        sign_c = -sign_c;
        j -= 4;
    }
    if (j > 1) sign_c = -sign_c;
    /* extended precision modular arithmetic */
    float r = ax - y * 0.78515625f;
    r = r - y * 2.4187564849853515625e-4f;
    r = r - y * 3.77489497744594108e-8f;
    float z = r * r;
    /* sin poly */
    float ps = -1.9515295891e-4f;
    ps = ps * z;
    ps = ps + 8.3321608736e-3f;
    ps = ps * z;
    ps = ps - 1.6666654611e-1f;
    ps = ps * z;
    ps = ps * r;
    ps = ps + r;
    /* cos poly */
    float pc = 2.443315711809948e-5f;
    pc = pc * z;
    pc = pc - 1.388731625493765e-3f;
    pc = pc * z;
    pc = pc + 4.166664568298827e-2f;
    pc = pc * z;
    pc = pc * z;
    pc = pc - 0.5f * z;
    pc = pc + 1.0f;
    float s, c;
    if (j == 1 || j == 2) {
        s = pc;
        c = ps;
    } else {
        s = ps;
        c = pc;
    }
    *s_out = sign_s < 0 ? -s : s;
    *c_out = sign_c < 0 ? -c : c;
}
float y3d_sinf(float x) { float s, c; y3d_sincos_core(x, &s, &c); return s; }
float y3d_cosf(float x) { float s, c; y3d_sincos_core(x, &s, &c); return c; }

/* exp: Cephes expf form; clamps to +inf / 0 outside the binary32 range. */
float y3d_expf(float x) {
    if (x > 88.72283905206835f) return INFINITY;
    if (x < -103.278929903431851103f) return 0.0f;
    if (x != x) return x;
    float fl = floorf(x * 1.44269504088896341f + 0.5f);
    int n = (int)fl;
    float r = x - fl * 0.693359375f;
    r = r - fl * -2.12194440e-4f;
    float z = r * r;
    float p = 1.9875691500e-4f;
    p = p * r;
    p = p + 1.3981999507e-3f;
    p = p * r;
    p = p + 8.3334519073e-3f;
    p = p * r;
    p = p + 4.1665795894e-2f;
    p = p * r;
    p = p + 1.6666665459e-1f;
    p = p * r;
    p = p + 5.0000001201e-1f;
    p = p * z;
    p = p + r;
    p = p + 1.0f;
    /* scale by 2^n in two exact steps so that subnormal results round once */
    int n1 = n / 2, n2 = n - n1;
    union { uint32_t u; float f; } a, b;
    a.u = (uint32_t)(n1 + 127) << 23;
    b.u = (uint32_t)(n2 + 127) << 23;
    return (p * a.f) * b.f;
}

/* pow with the exponents the path uses made explicit (torch special-cases 0.5/1/2/3 the same way,
 * aten/src/ATen/native/cpu/PowKernel.cpp); 4 and 6 are fixed multiplication trees; anything else
 * falls back to libm powf (not bit-reproducible on the GPU; tolerance-tested only). */
/* sigmoid as the fixed sequence 1 / (1 + y3d_expf(-x)): what the GPU's fused decode + top-k ranks by (csrc
 * y3d_common.cuh dm::sigmoid_).  Weakly monotone over all of binary32 -- oracle/check_sigmoid_monotone.c walks every
 * float -- so per-anchor maxima can be taken on the logits. */
float y3d_sigmoidf(float x) { return 1.0f / (1.0f + y3d_expf(-x)); }

float y3d_powf(float x, float e) {
    if (e == 0.5f) return sqrtf(x);
    if (e == 1.0f) return x;
    if (e == 2.0f) return x * x;
    if (e == 3.0f) return (x * x) * x;
    if (e == 4.0f) { float x2 = x * x; return x2 * x2; }
    if (e == 6.0f) { float x2 = x * x; float x4 = x2 * x2; return x4 * x2; }
    if (e == 0.0f) return 1.0f;
    return powf(x, e);
}

/* ------------------------------------------------------------------------------------------ */
/* make_anchors  (tal.py:300-312)                                                              */
/* ------------------------------------------------------------------------------------------ */
/* lvl_hw: [nl][2] = (h, w).  anc: [A][2] (x, y) in grid units, stride_out: [A]. Returns A. */
int y3d_o_make_anchors(int nl, const int *lvl_hw, const float *lvl_stride, float *anc, float *stride_out) {
    int a = 0;
    for (int l = 0; l < nl; ++l) {
        int h = lvl_hw[2 * l], w = lvl_hw[2 * l + 1];
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                if (anc) {
                    anc[2 * a] = (float)x + 0.5f;
                    anc[2 * a + 1] = (float)y + 0.5f;
                }
                if (stride_out) stride_out[a] = lvl_stride[l];
                ++a;
            }
    }
    return a;
}

/* ------------------------------------------------------------------------------------------ */
/* Detect.inference (head.py:53-79): x_cat [B, 4R+nc, A] -> y [B, 4+nc, A]                      */
/* ------------------------------------------------------------------------------------------ */
static float dfl_expect(const float *x, long cs, int R) { /* block.py:59-62 softmax over R bins, dot arange */
    float m = x[0];
    for (int j = 1; j < R; ++j) m = fmaxf(m, x[j * cs]);
    float e[64];
    float s = 0.0f;
    for (int j = 0; j < R; ++j) {
        e[j] = expf(x[j * cs] - m);
        s += e[j];
    }
    float acc = 0.0f;
    for (int j = 0; j < R; ++j) acc += (float)j * (e[j] / s);
    return acc;
}

void y3d_o_decode2d(const float *xcat, int B, int nc, int R, int A, const float *anc, const float *stride,
                    int xywh, float *y) {
    int C = 4 * R + nc;
    for (int b = 0; b < B; ++b) {
        const float *xb = xcat + (long)b * C * A;
        float *yb = y + (long)b * (4 + nc) * A;
        for (int a = 0; a < A; ++a) {
            float d[4];
            for (int s = 0; s < 4; ++s) d[s] = dfl_expect(xb + (long)(s * R) * A + a, A, R);
            float ax = anc[2 * a], ay = anc[2 * a + 1];
            float x1 = ax - d[0], y1 = ay - d[1], x2 = ax + d[2], y2 = ay + d[3]; /* tal.py:319-320 */
            float o0, o1, o2, o3;
            if (xywh) { /* tal.py:322-324 */
                o0 = (x1 + x2) / 2.0f;
                o1 = (y1 + y2) / 2.0f;
                o2 = x2 - x1;
                o3 = y2 - y1;
            } else {
                o0 = x1; o1 = y1; o2 = x2; o3 = y2;
            }
            float st = stride[a];
            yb[0 * (long)A + a] = o0 * st;
            yb[1 * (long)A + a] = o1 * st;
            yb[2 * (long)A + a] = o2 * st;
            yb[3 * (long)A + a] = o3 * st;
            for (int c = 0; c < nc; ++c) {
                float v = xb[(long)(4 * R + c) * A + a];
                yb[(long)(4 + c) * A + a] = y3d_sigmoidf(v); /* head.py:78 sigmoid */
            }
        }
    }
}

/* training-side decode (loss.py:197-204): pred_dist [B,A,4R] -> xyxy boxes in grid units [B,A,4] */
void y3d_o_bbox_decode(const float *pred_dist, int B, int A, int R, const float *anc, float *boxes) {
    for (long i = 0; i < (long)B * A; ++i) {
        int a = (int)(i % A);
        const float *p = pred_dist + i * 4 * R;
        float d[4];
        for (int s = 0; s < 4; ++s) d[s] = dfl_expect(p + s * R, 1, R);
        boxes[4 * i + 0] = anc[2 * a] - d[0];
        boxes[4 * i + 1] = anc[2 * a + 1] - d[1];
        boxes[4 * i + 2] = anc[2 * a] + d[2];
        boxes[4 * i + 3] = anc[2 * a + 1] + d[3];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* stable descending top-k (lowest index wins ties)                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float v; int i; } kv_t;
static int kv_cmp(const void *pa, const void *pb) {
    const kv_t *a = (const kv_t *)pa, *b = (const kv_t *)pb;
    if (a->v > b->v) return -1;
    if (a->v < b->v) return 1;
    return (a->i > b->i) - (a->i < b->i);
}
/* selects top k of v[0..n) into out_idx/out_val (sorted); O(n*k) insertion for small k, sort otherwise */
static void stable_topk(const float *v, int n, int k, int *out_idx, float *out_val, kv_t *scratch) {
    if (k <= 16) {
        int cnt = 0;
        for (int i = 0; i < n; ++i) {
            float x = v[i];
            if (cnt == k && !(x > out_val[k - 1])) continue;
            int p = cnt < k ? cnt : k - 1;
            while (p > 0 && x > out_val[p - 1]) { /* strict >: earlier index stays ahead on ties */
                out_val[p] = out_val[p - 1];
                out_idx[p] = out_idx[p - 1];
                --p;
            }
            out_val[p] = x;
            out_idx[p] = i;
            if (cnt < k) ++cnt;
        }
        return;
    }
    for (int i = 0; i < n; ++i) { scratch[i].v = v[i]; scratch[i].i = i; }
    qsort(scratch, (size_t)n, sizeof(kv_t), kv_cmp);
    for (int j = 0; j < k; ++j) { out_idx[j] = scratch[j].i; out_val[j] = scratch[j].v; }
}

/* ------------------------------------------------------------------------------------------ */
/* ops.v10postprocess / v10_3Dpostprocess (ops.py:852-880)                                      */
/* preds element (b,a,ch) at preds[b*sB + a*sA + ch*sC]; channel layout:                        */
/*   2D: [reg(4) | scores(nc)]  (scores_first = 0);  3D: [scores(nc) | reg(nreg)] (scores_first=1) */
/* out: reg [B,D,nreg], scores [B,D], labels [B,D] int64, anchor index [B,D] int32 (extra)      */
/* ------------------------------------------------------------------------------------------ */
int y3d_o_postprocess(const float *preds, long sB, long sA, long sC, int B, int A, int nc, int nreg,
                      int scores_first, int D, float *reg, float *scores, int64_t *labels, int32_t *anchor_idx) {
    if (D > A || D <= 0) return -1;
    int soff = scores_first ? 0 : nreg, roff = scores_first ? nc : 0;
    int nmax = A > D * nc ? A : D * nc;
    float *buf = (float *)malloc(sizeof(float) * (size_t)nmax);
    kv_t *scr = (kv_t *)malloc(sizeof(kv_t) * (size_t)nmax);
    int *idx1 = (int *)malloc(sizeof(int) * (size_t)D);
    float *val1 = (float *)malloc(sizeof(float) * (size_t)D);
    int *idx2 = (int *)malloc(sizeof(int) * (size_t)D);
    float *val2 = (float *)malloc(sizeof(float) * (size_t)D);
    for (int b = 0; b < B; ++b) {
        const float *pb = preds + (long)b * sB;
        for (int a = 0; a < A; ++a) { /* scores.amax(-1) */
            float m = pb[a * sA + (long)soff * sC];
            for (int c = 1; c < nc; ++c) m = fmaxf(m, pb[a * sA + (long)(soff + c) * sC]);
            buf[a] = m;
        }
        stable_topk(buf, A, D, idx1, val1, scr); /* first topk over anchors */
        for (int i = 0; i < D; ++i)
            for (int c = 0; c < nc; ++c) buf[i * nc + c] = pb[idx1[i] * sA + (long)(soff + c) * sC];
        stable_topk(buf, D * nc, D, idx2, val2, scr); /* second topk over D*nc */
        for (int j = 0; j < D; ++j) {
            int lab = idx2[j] % nc, i = idx2[j] / nc, a = idx1[i];
            scores[(long)b * D + j] = val2[j];
            labels[(long)b * D + j] = lab;
            if (anchor_idx) anchor_idx[(long)b * D + j] = a;
            for (int r = 0; r < nreg; ++r) reg[((long)b * D + j) * nreg + r] = pb[a * sA + (long)(roff + r) * sC];
        }
    }
    free(buf); free(scr); free(idx1); free(val1); free(idx2); free(val2);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* bbox_iou(box1, box2, xywh=False, CIoU=True)  (metrics.py:96-131), op order preserved         */
/* ------------------------------------------------------------------------------------------ */
float y3d_o_ciou(const float *b1, const float *b2) {
    const float eps = 1e-7f;
    float b1x1 = b1[0], b1y1 = b1[1], b1x2 = b1[2], b1y2 = b1[3];
    float b2x1 = b2[0], b2y1 = b2[1], b2x2 = b2[2], b2y2 = b2[3];
    float w1 = b1x2 - b1x1, h1 = (b1y2 - b1y1) + eps;
    float w2 = b2x2 - b2x1, h2 = (b2y2 - b2y1) + eps;
    float iw = fminf(b1x2, b2x2) - fmaxf(b1x1, b2x1);
    float ih = fminf(b1y2, b2y2) - fmaxf(b1y1, b2y1);
    if (iw < 0.0f) iw = 0.0f;
    if (ih < 0.0f) ih = 0.0f;
    float inter = iw * ih;
    float a1 = w1 * h1, a2 = w2 * h2;
    float uni = ((a1 + a2) - inter) + eps;
    float iou = inter / uni;
    float cw = fmaxf(b1x2, b2x2) - fminf(b1x1, b2x1);
    float ch = fmaxf(b1y2, b2y2) - fminf(b1y1, b2y1);
    float cw2 = cw * cw, ch2 = ch * ch;
    float c2 = (cw2 + ch2) + eps;
    float dx = ((b2x1 + b2x2) - b1x1) - b1x2;
    float dy = ((b2y1 + b2y2) - b1y1) - b1y2;
    float dx2 = dx * dx, dy2 = dy * dy;
    float rho2 = (dx2 + dy2) / 4.0f;
    float da = y3d_atanf(w2 / h2) - y3d_atanf(w1 / h1);
    float da2 = da * da;
    float v = 0.4052847345693511f * da2; /* 4 / pi**2 */
    float alpha = v / ((v - iou) + 1.0000001f); /* 1 + eps rounds to 1 + 2^-23 */
    float pen = (rho2 / c2) + (v * alpha);
    return iou - pen;
}

/* ------------------------------------------------------------------------------------------ */
/* TaskAlignedAssigner.forward (tal.py:44-264), dense, one image at a time                      */
/* ------------------------------------------------------------------------------------------ */
/* inputs : pd_scores [B,A,nc] (already sigmoid), pd_bboxes [B,A,4] xyxy px, anc [A,2] px,
 *          gt_labels [B,M], gt_bboxes [B,M,4], mask_gt [B,M] (0/1)
 * outputs: target_labels [B,A] i64, target_bboxes [B,A,4], target_scores [B,A,nc], fg_mask [B,A] u8,
 *          target_gt_idx [B,A] i64.  Optional debug (may be NULL): mask_pos_out [B,M,A] u8 (final mask_pos),
 *          align_out / overlaps_out [B,M,A] (pre-normalisation align_metric / overlaps).
 * M must be > 0 (the M == 0 early-out of tal.py:68-76 is host-side glue).                         */
static void assign_core(int A, int nc, int M, int k, float alpha, float beta, float eps, const float *sc,
                        const float *pb, const float *anc, const float *gl, const float *gb, const float *mg,
                        int64_t *t_lab, float *t_box, float *t_sc, uint8_t *fg, int64_t *t_gi,
                        const float *extra_sim /* [M,A] or NULL: 3D similarity factor */, float gamma,
                        int use_box, int constrain, uint8_t *mask_pos_out, float *align_out, float *ov_out) {
    float *align = (float *)calloc((size_t)M * A, sizeof(float));
    float *ov = (float *)calloc((size_t)M * A, sizeof(float));   /* "overlaps" as returned by get_*_metrics */
    uint8_t *in_gts = (uint8_t *)calloc((size_t)M * A, 1);
    uint8_t *mpos = (uint8_t *)calloc((size_t)M * A, 1);
    int *tk_idx = (int *)malloc(sizeof(int) * (size_t)k);
    float *tk_val = (float *)malloc(sizeof(float) * (size_t)k);
    int8_t *count = (int8_t *)malloc((size_t)A);
    kv_t *scratch = k > 16 ? (kv_t *)malloc(sizeof(kv_t) * (size_t)A) : NULL;
    for (int m = 0; m < M; ++m) {
        const float *g = gb + 4 * m;
        int lab = (int)(int64_t)gl[m];
        int valid = mg[m] != 0.0f;
        for (int a = 0; a < A; ++a) {
            float ax = anc[2 * a], ay = anc[2 * a + 1];
            /* select_candidates_in_gts tal.py:218-235 */
            float d0 = ax - g[0], d1 = ay - g[1], d2 = g[2] - ax, d3 = g[3] - ay;
            float dm = fminf(fminf(d0, d1), fminf(d2, d3));
            uint8_t ig = dm > 1e-9f;
            in_gts[(long)m * A + a] = ig;
            int sel = constrain ? (ig && valid) : valid;
            if (!sel) continue;
            /* get_box_metrics tal.py:108-127 (3D: get_box_kp_metrics tal.py:578-603) */
            float s = sc[(long)a * nc + lab];
            float metric = y3d_powf(s, alpha);
            float o = 0.0f;
            if (use_box) {
                o = y3d_o_ciou(g, pb + 4 * a);
                if (o < 0.0f) o = 0.0f; /* clamp_(0) tal.py:131 */
                metric = metric * y3d_powf(o, beta);
            }
            if (extra_sim) {
                float sim = extra_sim[(long)m * A + a];
                metric = metric * y3d_powf(sim, gamma);
                o = sim; /* tal.py:602-603: similarities returned in place of overlaps */
            }
            align[(long)m * A + a] = metric;
            ov[(long)m * A + a] = o;
        }
        /* select_topk_candidates tal.py:133-167 */
        memset(count, 0, (size_t)A);
        if (valid) {
            stable_topk(align + (long)m * A, A, k, tk_idx, tk_val, scratch);
            for (int j = 0; j < k; ++j) count[tk_idx[j]] += 1;
        } else {
            count[0] = (int8_t)k; /* indices forced to 0 */
        }
        for (int a = 0; a < A; ++a) {
            int c = count[a] > 1 ? 0 : count[a];
            int gate = constrain ? in_gts[(long)m * A + a] : 1;
            mpos[(long)m * A + a] = (uint8_t)(c && gate && valid); /* get_pos_mask tal.py:104 */
        }
    }
    if (align_out) memcpy(align_out, align, sizeof(float) * (size_t)M * A);
    if (ov_out) memcpy(ov_out, ov, sizeof(float) * (size_t)M * A);
    /* select_highest_overlaps tal.py:237-264 */
    for (int a = 0; a < A; ++a) {
        int s = 0;
        for (int m = 0; m < M; ++m) s += mpos[(long)m * A + a];
        if (s > 1) {
            int best = 0;
            float bv = ov[a];
            for (int m = 1; m < M; ++m)
                if (ov[(long)m * A + a] > bv) { bv = ov[(long)m * A + a]; best = m; }
            for (int m = 0; m < M; ++m) mpos[(long)m * A + a] = (uint8_t)(m == best);
            s = 1;
        }
        int gi = 0;
        for (int m = 0; m < M; ++m)
            if (mpos[(long)m * A + a]) { gi = m; break; }
        fg[a] = (uint8_t)(s > 0);
        t_gi[a] = gi;
    }
    if (mask_pos_out) memcpy(mask_pos_out, mpos, (size_t)M * A);
    /* normalisation tal.py:88-92 */
    float *pos_align = (float *)calloc((size_t)M, sizeof(float));
    float *pos_ov = (float *)calloc((size_t)M, sizeof(float));
    for (int m = 0; m < M; ++m)
        for (int a = 0; a < A; ++a)
            if (mpos[(long)m * A + a]) {
                pos_align[m] = fmaxf(pos_align[m], align[(long)m * A + a]);
                pos_ov[m] = fmaxf(pos_ov[m], ov[(long)m * A + a]);
            }
    /* get_targets tal.py:169-216 */
    for (int a = 0; a < A; ++a) {
        int gi = (int)t_gi[a];
        int64_t lab = (int64_t)gl[gi];
        if (lab < 0) lab = 0;
        if (t_lab) t_lab[a] = lab;
        if (t_box) for (int j = 0; j < 4; ++j) t_box[4 * a + j] = gb[4 * gi + j];
        float norm = 0.0f;
        for (int m = 0; m < M; ++m)
            if (mpos[(long)m * A + a]) {
                float v = (align[(long)m * A + a] * pos_ov[m]) / (pos_align[m] + eps);
                norm = fmaxf(norm, v);
            }
        for (int c = 0; c < nc; ++c) t_sc[(long)a * nc + c] = 0.0f;
        if (fg[a]) t_sc[(long)a * nc + lab] = norm;
    }
    free(align); free(ov); free(in_gts); free(mpos); free(tk_idx); free(tk_val); free(count);
    free(pos_align); free(pos_ov); free(scratch);
}

int y3d_o_tal_assign(const float *pd_scores, const float *pd_bboxes, const float *anc, const float *gt_labels,
                     const float *gt_bboxes, const float *mask_gt, int B, int A, int nc, int M, int k, float alpha,
                     float beta, float eps, int64_t *t_lab, float *t_box, float *t_sc, uint8_t *fg, int64_t *t_gi,
                     uint8_t *mask_pos_out, float *align_out, float *ov_out) {
    if (M <= 0 || k <= 0 || k > A) return -1;
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b)
        assign_core(A, nc, M, k, alpha, beta, eps, pd_scores + (long)b * A * nc, pd_bboxes + (long)b * A * 4, anc,
                    gt_labels + (long)b * M, gt_bboxes + (long)b * M * 4, mask_gt + (long)b * M,
                    t_lab ? t_lab + (long)b * A : NULL, t_box ? t_box + (long)b * A * 4 : NULL,
                    t_sc + (long)b * A * nc, fg + (long)b * A, t_gi + (long)b * A, NULL, 1.0f, 1, 1,
                    mask_pos_out ? mask_pos_out + (long)b * M * A : NULL, align_out ? align_out + (long)b * M * A : NULL,
                    ov_out ? ov_out + (long)b * M * A : NULL);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* v8DetectionLoss.__call__ (loss.py:206-257) for one branch, given packed GT [B,M,5]           */
/* feats as x_cat [B, 4R+nc, A]; anc (grid units) [A,2]; stride [A];                             */
/* out: loss3 = (box, cls, dfl) after gains; also tss (target_scores_sum) and n_fg              */
/* ------------------------------------------------------------------------------------------ */
static double bce_logits(double x, double t) { /* BCEWithLogits, reduction none */
    double ax = fabs(x);
    return (x > 0 ? x : 0) - x * t + log1p(exp(-ax));
}

/* fg_out [B,A] u8 / tgi_out [B,A] i64 (optional): the assignment the loss was computed from */
int y3d_o_v8_loss_dbg(const float *xcat, int B, int nc, int R, int A, const float *anc, const float *stride,
                      const float *gt /* [B,M,5] cls,xyxy px */, int M, int k, float gain_box, float gain_cls,
                      float gain_dfl, double *loss3, double *tss_out, int64_t *nfg_out, uint8_t *fg_out,
                      int64_t *tgi_out) {
    int C = 4 * R + nc;
    long BA = (long)B * A;
    float *pd_scores = (float *)malloc(sizeof(float) * (size_t)BA * nc);
    float *pd_dist = (float *)malloc(sizeof(float) * (size_t)BA * 4 * R);
    float *logit = (float *)malloc(sizeof(float) * (size_t)BA * nc);
    float *pbox = (float *)malloc(sizeof(float) * (size_t)BA * 4);
    float *pbox_px = (float *)malloc(sizeof(float) * (size_t)BA * 4);
    float *anc_px = (float *)malloc(sizeof(float) * (size_t)A * 2);
    for (int b = 0; b < B; ++b) /* permute(0,2,1) loss.py:214-215 */
        for (int a = 0; a < A; ++a) {
            for (int c = 0; c < 4 * R; ++c) pd_dist[((long)b * A + a) * 4 * R + c] = xcat[((long)b * C + c) * A + a];
            for (int c = 0; c < nc; ++c) {
                float v = xcat[((long)b * C + 4 * R + c) * A + a];
                logit[((long)b * A + a) * nc + c] = v;
                pd_scores[((long)b * A + a) * nc + c] = 1.0f / (1.0f + expf(-v));
            }
        }
    y3d_o_bbox_decode(pd_dist, B, A, R, anc, pbox);
    for (long i = 0; i < BA; ++i)
        for (int j = 0; j < 4; ++j) pbox_px[4 * i + j] = pbox[4 * i + j] * stride[i % A];
    for (int a = 0; a < A; ++a) {
        anc_px[2 * a] = anc[2 * a] * stride[a];
        anc_px[2 * a + 1] = anc[2 * a + 1] * stride[a];
    }
    double sum_bce = 0, sum_ts = 0, sum_iou = 0, sum_dfl = 0;
    int64_t nfg = 0;
    float *t_sc = NULL, *t_box = NULL;
    uint8_t *fg = NULL;
    if (M > 0) {
        float *gl = (float *)malloc(sizeof(float) * (size_t)B * M);
        float *gbx = (float *)malloc(sizeof(float) * (size_t)B * M * 4);
        float *mg = (float *)malloc(sizeof(float) * (size_t)B * M);
        for (long i = 0; i < (long)B * M; ++i) {
            gl[i] = gt[5 * i];
            float s = 0.0f;
            for (int j = 0; j < 4; ++j) { gbx[4 * i + j] = gt[5 * i + 1 + j]; }
            s = ((gbx[4 * i] + gbx[4 * i + 1]) + gbx[4 * i + 2]) + gbx[4 * i + 3]; /* sum(2) > 0 loss.py:226 */
            mg[i] = s > 0.0f ? 1.0f : 0.0f;
        }
        t_sc = (float *)malloc(sizeof(float) * (size_t)BA * nc);
        t_box = (float *)malloc(sizeof(float) * (size_t)BA * 4);
        fg = (uint8_t *)malloc((size_t)BA);
        int64_t *t_gi = (int64_t *)malloc(sizeof(int64_t) * (size_t)BA);
        y3d_o_tal_assign(pd_scores, pbox_px, anc_px, gl, gbx, mg, B, A, nc, M, k, 0.5f, 6.0f, 1e-9f, NULL, t_box, t_sc,
                         fg, t_gi, NULL, NULL, NULL);
        if (fg_out) memcpy(fg_out, fg, (size_t)BA);
        if (tgi_out) memcpy(tgi_out, t_gi, sizeof(int64_t) * (size_t)BA);
        free(gl); free(gbx); free(mg); free(t_gi);
    } else {
        if (fg_out) memset(fg_out, 0, (size_t)BA);
        if (tgi_out) memset(tgi_out, 0, sizeof(int64_t) * (size_t)BA);
    }
    for (long i = 0; i < BA; ++i) {
        for (int c = 0; c < nc; ++c) {
            double t = t_sc ? (double)t_sc[i * nc + c] : 0.0;
            sum_ts += t;
            sum_bce += bce_logits((double)logit[i * nc + c], t);
        }
    }
    double tss = sum_ts > 1.0 ? sum_ts : 1.0; /* max(target_scores.sum(), 1) loss.py:240 */
    if (t_sc) {
        for (long i = 0; i < BA; ++i) {
            if (!fg[i]) continue;
            ++nfg;
            int a = (int)(i % A);
            double w = 0;
            for (int c = 0; c < nc; ++c) w += t_sc[i * nc + c];
            float tb[4];
            for (int j = 0; j < 4; ++j) tb[j] = t_box[4 * i + j] / stride[a]; /* loss.py:248 */
            float iou = y3d_o_ciou(pbox + 4 * i, tb); /* BboxLoss.forward loss.py:85 (box1 = pred) */
            sum_iou += (1.0 - (double)iou) * w;
            /* DFL loss.py:90-113; bbox2dist tal.py:328-331 */
            float ltrb[4] = {anc[2 * a] - tb[0], anc[2 * a + 1] - tb[1], tb[2] - anc[2 * a], tb[3] - anc[2 * a + 1]};
            double dfl = 0;
            for (int s = 0; s < 4; ++s) {
                float t = ltrb[s];
                float hi = (float)(R - 1) - 0.01f;
                if (t < 0.0f) t = 0.0f;
                if (t > hi) t = hi;
                int tl = (int)t, tr = tl + 1;
                double wl = (double)((float)tr - t), wr = 1.0 - wl;
                const float *p = pd_dist + i * 4 * R + s * R;
                double mx = p[0];
                for (int j = 1; j < R; ++j) mx = p[j] > mx ? p[j] : mx;
                double se = 0;
                for (int j = 0; j < R; ++j) se += exp((double)p[j] - mx);
                double lse = mx + log(se);
                dfl += (lse - p[tl]) * wl + (lse - p[tr]) * wr;
            }
            sum_dfl += (dfl / 4.0) * w;
        }
    }
    loss3[0] = (nfg ? sum_iou / tss : 0.0) * gain_box;
    loss3[1] = sum_bce / tss * gain_cls;
    loss3[2] = (nfg ? sum_dfl / tss : 0.0) * gain_dfl;
    if (tss_out) *tss_out = tss;
    if (nfg_out) *nfg_out = nfg;
    free(pd_scores); free(pd_dist); free(logit); free(pbox); free(pbox_px); free(anc_px);
    free(t_sc); free(t_box); free(fg);
    return 0;
}

int y3d_o_v8_loss(const float *xcat, int B, int nc, int R, int A, const float *anc, const float *stride,
                  const float *gt, int M, int k, float gain_box, float gain_cls, float gain_dfl, double *loss3,
                  double *tss_out, int64_t *nfg_out) {
    return y3d_o_v8_loss_dbg(xcat, B, nc, R, A, anc, stride, gt, M, k, gain_box, gain_cls, gain_dfl, loss3, tss_out,
                             nfg_out, NULL, NULL);
}

/* ------------------------------------------------------------------------------------------ */
/* v10Detect3d.decode (head.py:755-764): x_cat [B, nc+35, A] -> [B, nc+35, A]                    */
/* channels in : cls(nc) o2d(2) s2d(2) o3d(2) s3d(3) hd(24) dep(1) dep_un(1)                      */
/* channels out: cls(nc) bbox xyxy(4) center3d(2) s3d(3) hd(24) dep(1) dep_un(1)                  */
/* ------------------------------------------------------------------------------------------ */
void y3d_o_decode3d(const float *xcat, int B, int nc, int A, const float *anc, const float *stride, float *y) {
    int C = nc + 35;
    for (int b = 0; b < B; ++b)
        for (int a = 0; a < A; ++a) {
            const float *x = xcat + (long)b * C * A + a;
            float *o = y + (long)b * C * A + a;
#define X(c) x[(long)(c)*A]
#define O(c) o[(long)(c)*A]
            for (int c = 0; c < nc; ++c) O(c) = X(c); /* raw logits pass through */
            float st = stride[a], ax = anc[2 * a], ay = anc[2 * a + 1];
            float s2x = X(nc + 2) * st, s2y = X(nc + 3) * st;
            float ox = (X(nc + 0) + ax) * st, oy = (X(nc + 1) + ay) * st;
            O(nc + 0) = ox - s2x / 2.0f;
            O(nc + 1) = oy - s2y / 2.0f;
            O(nc + 2) = ox + s2x / 2.0f;
            O(nc + 3) = oy + s2y / 2.0f;
            O(nc + 4) = (X(nc + 4) + ax) * st;
            O(nc + 5) = (X(nc + 5) + ay) * st;
            for (int c = nc + 6; c < C; ++c) O(c) = X(c);
#undef X
#undef O
        }
}

/* ------------------------------------------------------------------------------------------ */
/* 3D keypoints (keypoint_utils.py:11-118): one box -> 8 corners in camera frame [8][3]          */
/* c3d (u,v) px, dep, size3d (h,w,l), alpha-bin index + residual, calib (cu,cv,fu,fv,tx,ty)      */
/* ------------------------------------------------------------------------------------------ */
void y3d_o_keypoints(const float *c3d, float dep, const float *size3d, int hbin, float hres, const float *calib,
                     float *kps /* [8][3] */) {
    float cu = calib[0], cv = calib[1], fu = calib[2], fv = calib[3], tx = calib[4], ty = calib[5];
    /* img_to_rect keypoint_utils.py:113-119 */
    float lx = ((c3d[0] - cu) * dep) / fu + tx;
    float ly = ((c3d[1] - cv) * dep) / fv + ty;
    float lz = dep;
    /* class2angle :42-47 (2*pi/12 evaluated in double then cast, as numpy float * torch float32) */
    float apc = 0.5235987755982988f;
    float alpha = (float)hbin * apc + hres;
    if (alpha > Y3D_PI_F) alpha = alpha - 6.283185307179586f;
    /* alpha2ry :94-101 */
    float ry = alpha + y3d_atan2f(c3d[0] - cu, fu);
    if (ry > Y3D_PI_F) ry = ry - 6.283185307179586f;
    if (ry < -Y3D_PI_F) ry = ry + 6.283185307179586f;
    /* to_egoc_rot_mat :87-91: R = Rx(pi/2) @ Ry(-ry) @ Rz(0); boxes = einsum("ji,kj->ki", R, corners) = corners @ R */
    float cx = y3d_cosf(1.5707963267948966f), sx = y3d_sinf(1.5707963267948966f);
    float cy = y3d_cosf(-ry), sy = y3d_sinf(-ry);
    /* Rx = [[1,0,0],[0,cx,-sx],[0,sx,cx]], Ry = [[cy,0,sy],[0,1,0],[-sy,0,cy]]
     * R = Rx @ Ry = [[cy, 0, sy], [sx*sy, cx, -sx*cy], [-cx*sy, sx, cx*cy]]   (Rz = I) */
    float R00 = cy, R01 = 0.0f, R02 = sy;
    float R10 = sx * sy, R11 = cx, R12 = -(sx * cy);
    float R20 = -(cx * sy), R21 = sx, R22 = cx * cy;
    float hl = size3d[2] / 2.0f, hw = size3d[1] / 2.0f, hh = size3d[0] / 2.0f;
    const float sgx[8] = {1, 1, -1, -1, 1, 1, -1, -1};
    const float sgy[8] = {1, -1, 1, -1, 1, -1, 1, -1};
    const float sgz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
    for (int k = 0; k < 8; ++k) {
        float px = sgx[k] * hl, py = sgy[k] * hw, pz = sgz[k] * hh;
        /* out_i = sum_j R[j][i] * p[j], accumulated j = 0,1,2 */
        kps[3 * k + 0] = ((px * R00 + py * R10) + pz * R20) + lx;
        kps[3 * k + 1] = ((px * R01 + py * R11) + pz * R21) + ly;
        kps[3 * k + 2] = ((px * R02 + py * R12) + pz * R22) + lz;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* TaskAlignedAssigner3d.forward (tal.py:391-452)                                               */
/* pd_3d [B,A,31] = o3d(2) s3d(3) hd(24) dep(1) un(1); gts packed [B,M,17] =
 *   label(1) bbox xyxy(4) center_2d(2) size_2d(2) center_3d(2) size_3d(3) depth(1) hbin(1) hres(1)
 * outputs: t_lab [B,A] i64, t_sc [B,A,nc], t_vals [B,A,12] = c2d(2) s2d(2) c3d(2) s3d(3) dep hbin hres,
 *          fg [B,A], t_gi [B,A], pd_kps [B,A,24], gt_kps [B,M,24]                                 */
/* ------------------------------------------------------------------------------------------ */
int y3d_o_tal_assign3d(const float *pd_scores, const float *pd_bboxes, const float *pd_3d, const float *anc,
                       const float *stride, const float *gts, const float *mask_gt, const float *calibs,
                       const float *mean_sizes, int B, int A, int nc, int M, int k, float alpha, float beta,
                       float gamma, float eps, int use_2d, int use_3d, int kps_l2, int constrain, int64_t *t_lab,
                       float *t_sc, float *t_vals, uint8_t *fg, int64_t *t_gi, float *pd_kps_out, float *gt_kps_out,
                       uint8_t *mask_pos_out, float *align_out, float *ov_out) {
    if (M <= 0 || k <= 0 || k > A) return -1;
    if (!use_2d && !use_3d) return -2;
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b) {
        const float *cal = calibs + 6 * b;
        float *pk = (float *)malloc(sizeof(float) * (size_t)A * 24);
        float *gk = (float *)malloc(sizeof(float) * (size_t)M * 24);
        float *gl = (float *)malloc(sizeof(float) * (size_t)M);
        float *gb = (float *)malloc(sizeof(float) * (size_t)M * 4);
        float *sim = use_3d ? (float *)calloc((size_t)M * A, sizeof(float)) : NULL;
        for (int a = 0; a < A; ++a) {
            const float *p = pd_3d + ((long)b * A + a) * 31;
            const float *s = pd_scores + ((long)b * A + a) * nc;
            float c3d[2] = {anc[2 * a] + p[0] * stride[a], anc[2 * a + 1] + p[1] * stride[a]}; /* tal.py:454-456 */
            int cls = 0; /* argmax, first max tal.py:459 */
            for (int c = 1; c < nc; ++c) if (s[c] > s[cls]) cls = c;
            float sz[3] = {mean_sizes[3 * cls] + p[2], mean_sizes[3 * cls + 1] + p[3], mean_sizes[3 * cls + 2] + p[4]};
            int hb = 0;
            for (int j = 1; j < 12; ++j) if (p[5 + j] > p[5 + hb]) hb = j;
            y3d_o_keypoints(c3d, p[29], sz, hb, p[5 + 12 + hb], cal, pk + 24 * a);
        }
        for (int m = 0; m < M; ++m) {
            const float *g = gts + ((long)b * M + m) * 17;
            gl[m] = g[0];
            for (int j = 0; j < 4; ++j) gb[4 * m + j] = g[1 + j];
            int lab = (int)(int64_t)g[0];
            if (lab < 0) lab = 0;
            if (lab >= nc) lab = nc - 1;
            float sz[3] = {mean_sizes[3 * lab] + g[11], mean_sizes[3 * lab + 1] + g[12], mean_sizes[3 * lab + 2] + g[13]};
            y3d_o_keypoints(g + 9, g[14], sz, (int)(int64_t)g[15], g[16], cal, gk + 24 * m);
        }
        if (use_3d)
            for (int m = 0; m < M; ++m)
                for (int a = 0; a < A; ++a) { /* keypoint_distance_3d tal.py:464-470 */
                    float acc = 0.0f;
                    for (int j = 0; j < 24; ++j) {
                        float d = pk[24 * a + j] - gk[24 * m + j];
                        acc = acc + (kps_l2 ? d * d : fabsf(d));
                    }
                    float dist = acc / 24.0f;
                    sim[(long)m * A + a] = 1.0f / y3d_expf(kps_l2 ? 0.5f * dist : dist);
                }
        assign_core(A, nc, M, k, alpha, beta, eps, pd_scores + (long)b * A * nc, pd_bboxes + (long)b * A * 4, anc, gl,
                    gb, mask_gt + (long)b * M, t_lab + (long)b * A, NULL, t_sc + (long)b * A * nc, fg + (long)b * A,
                    t_gi + (long)b * A, sim, gamma, use_2d, constrain,
                    mask_pos_out ? mask_pos_out + (long)b * M * A : NULL, align_out ? align_out + (long)b * M * A : NULL,
                    ov_out ? ov_out + (long)b * M * A : NULL);
        for (int a = 0; a < A; ++a) { /* get_targets tal.py:651-700 */
            const float *g = gts + ((long)b * M + t_gi[(long)b * A + a]) * 17;
            for (int j = 0; j < 12; ++j) t_vals[((long)b * A + a) * 12 + j] = g[5 + j];
        }
        if (pd_kps_out) memcpy(pd_kps_out + (long)b * A * 24, pk, sizeof(float) * (size_t)A * 24);
        if (gt_kps_out) memcpy(gt_kps_out + (long)b * M * 24, gk, sizeof(float) * (size_t)M * 24);
        free(pk); free(gk); free(gl); free(gb); free(sim);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* KITTIDataset.decode_preds (kitti.py:519-576), undo_augment=True, use_camera_dis=False         */
/* dets [B,D,37] = bbox(4) c3d(2) s3d(3) hd(24) dep un score label                               */
/* per image: calib [6], inv_affine [2][3], ratio [2] (ratio_pad[i][0] = (rx, ry)); cls_mean_size [nc][3]
 * rows [B,D,14] = cls alpha x1 y1 x2 y2 h w l x y z ry score ; valid [B,D] u8 (score >= thr)      */
/* the reference runs this in float64 numpy on float32 inputs; so does this restatement.          */
/* ------------------------------------------------------------------------------------------ */
void y3d_o_decode_preds(const float *dets, int B, int D, const double *calib, const double *inv_affine,
                        const double *ratio, const double *cls_mean_size, double thr, double *rows,
                        uint8_t *valid) {
    for (int b = 0; b < B; ++b)
        for (int j = 0; j < D; ++j) {
            const float *p = dets + ((long)b * D + j) * 37;
            double *r = rows + ((long)b * D + j) * 14;
            const double *cal = calib + 6 * b, *T = inv_affine + 6 * b;
            int hb = 0;
            for (int q = 1; q < 12; ++q) if (p[9 + q] > p[9 + hb]) hb = q;
            /* bin2angle decode_helper.py:12-18 on float32 tensors */
            float alpha_f = (float)hb * 0.5235987755982988f + p[9 + 12 + hb];
            if (alpha_f > Y3D_PI_F) alpha_f = alpha_f - 6.283185307179586f;
            double alpha = (double)alpha_f;
            int cls = (int)p[36];
            /* bbox / ratio_pad[i][0][[0,1,0,1]] : float32 / float64 -> float64 */
            double bx[4] = {p[0] / ratio[2 * b], p[1] / ratio[2 * b + 1], p[2] / ratio[2 * b], p[3] / ratio[2 * b + 1]};
            double x = (bx[0] + bx[2]) / 2;
            /* dimensions += cls_mean_size (float32 array += float64 -> stays float32) */
            float dim[3];
            for (int q = 0; q < 3; ++q) dim[q] = (float)((double)p[6 + q] + cls_mean_size[3 * cls + q]);
            double depth = p[33];
            float sigma_f = expf(-p[34]); /* torch.exp on float32 */
            /* affine_transform kitti_utils.py:467-471: float32 [x,y,1] dot float64 2x3 */
            double cx = T[0] * (double)p[4] + T[1] * (double)p[5] + T[2];
            double cy = T[3] * (double)p[4] + T[4] * (double)p[5] + T[5];
            double lx = ((cx - cal[0]) * depth) / cal[2] + cal[4];
            double ly = ((cy - cal[1]) * depth) / cal[3] + cal[5];
            ly += (double)dim[0] / 2;
            double ry = alpha + atan2(x - cal[0], cal[2]);
            if (ry > M_PI) ry -= 2 * M_PI;
            if (ry < -M_PI) ry += 2 * M_PI;
            float sig = 1.0f / (1.0f + expf(-p[35])); /* scores.sigmoid() float32 */
            double score = (double)sig * (double)sigma_f;
            r[0] = cls; r[1] = alpha;
            r[2] = bx[0]; r[3] = bx[1]; r[4] = bx[2]; r[5] = bx[3];
            r[6] = dim[0]; r[7] = dim[1]; r[8] = dim[2];
            r[9] = lx; r[10] = ly; r[11] = depth; r[12] = ry; r[13] = score;
            valid[(long)b * D + j] = (uint8_t)!(score < thr);
        }
}

/* ------------------------------------------------------------------------------------------ */
/* rotate_iou_gpu_eval: ultralytics/data/datasets/kitti_eval.py:60-345 (BEV overlap of the KITTI  */
/* evaluator).  Boxes are (cx, cy, dx, dy, angle); the overlap polygon of two rotated rectangles  */
/* is collected as corners-inside + edge intersections, ordered around its centroid and measured */
/* as a triangle fan.  float32 throughout, like the reference's numba kernel.                     */
/* ------------------------------------------------------------------------------------------ */
static void riou_corners(const float *rb, float *c) { /* rbbox_to_corners :149-172 */
    const float a_cos = cosf(rb[4]), a_sin = sinf(rb[4]);
    const float xd = rb[2], yd = rb[3];
    const float cx[4] = {-xd / 2, -xd / 2, xd / 2, xd / 2};
    const float cy[4] = {-yd / 2, yd / 2, yd / 2, -yd / 2};
    for (int i = 0; i < 4; ++i) {
        c[2 * i] = a_cos * cx[i] + a_sin * cy[i] + rb[0];
        c[2 * i + 1] = -a_sin * cx[i] + a_cos * cy[i] + rb[1];
    }
}
static int riou_point_in_quad(float px, float py, const float *c) { /* :105-122 */
    const float ab0 = c[2] - c[0], ab1 = c[3] - c[1], ad0 = c[6] - c[0], ad1 = c[7] - c[1];
    const float ap0 = px - c[0], ap1 = py - c[1];
    const float abab = ab0 * ab0 + ab1 * ab1, abap = ab0 * ap0 + ab1 * ap1;
    const float adad = ad0 * ad0 + ad1 * ad1, adap = ad0 * ap0 + ad1 * ap1;
    const float eps = -1e-6f;
    return abab - abap >= eps && abap >= eps && adad - adap >= eps && adap >= eps;
}
static int riou_seg_intersect(const float *p1, const float *p2, int i, int j, float *t) { /* :60-102 */
    const float A0 = p1[2 * i], A1 = p1[2 * i + 1], B0 = p1[2 * ((i + 1) % 4)], B1 = p1[2 * ((i + 1) % 4) + 1];
    const float C0 = p2[2 * j], C1 = p2[2 * j + 1], D0 = p2[2 * ((j + 1) % 4)], D1 = p2[2 * ((j + 1) % 4) + 1];
    const float BA0 = B0 - A0, BA1 = B1 - A1, DA0 = D0 - A0, CA0 = C0 - A0, DA1 = D1 - A1, CA1 = C1 - A1;
    const int acd = DA1 * CA0 > CA1 * DA0;
    const int bcd = (D1 - B1) * (C0 - B0) > (C1 - B1) * (D0 - B0);
    if (acd != bcd) {
        const int abc = CA1 * BA0 > BA1 * CA0, abd = DA1 * BA0 > BA1 * DA0;
        if (abc != abd) {
            const float DC0 = D0 - C0, DC1 = D1 - C1;
            const float ABBA = A0 * B1 - B0 * A1, CDDC = C0 * D1 - D0 * C1;
            const float DH = BA1 * DC0 - BA0 * DC1;
            t[0] = (ABBA * DC0 - BA0 * CDDC) / DH;
            t[1] = (ABBA * DC1 - BA1 * CDDC) / DH;
            return 1;
        }
    }
    return 0;
}
static float riou_inter(const float *r1, const float *r2) { /* inter :231-245 */
    float c1[8], c2[8], ip[16 + 32], t[2];
    riou_corners(r1, c1);
    riou_corners(r2, c2);
    int n = 0;
    for (int i = 0; i < 4; ++i) { /* quadrilateral_intersection :125-146 */
        if (riou_point_in_quad(c1[2 * i], c1[2 * i + 1], c2)) { ip[2 * n] = c1[2 * i]; ip[2 * n + 1] = c1[2 * i + 1]; ++n; }
        if (riou_point_in_quad(c2[2 * i], c2[2 * i + 1], c1)) { ip[2 * n] = c2[2 * i]; ip[2 * n + 1] = c2[2 * i + 1]; ++n; }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (riou_seg_intersect(c1, c2, i, j, t)) { ip[2 * n] = t[0]; ip[2 * n + 1] = t[1]; ++n; }
    if (n > 0) { /* sort_vertex_in_convex_polygon :175-212 */
        float cx = 0.0f, cy = 0.0f, vs[24];
        for (int i = 0; i < n; ++i) { cx += ip[2 * i]; cy += ip[2 * i + 1]; }
        cx /= n; cy /= n;
        for (int i = 0; i < n; ++i) {
            float v0 = ip[2 * i] - cx, v1 = ip[2 * i + 1] - cy;
            const float d = sqrtf(v0 * v0 + v1 * v1);
            v0 = v0 / d; v1 = v1 / d;
            if (v1 < 0) v0 = -2 - v0;
            vs[i] = v0;
        }
        for (int i = 1; i < n; ++i)
            if (vs[i - 1] > vs[i]) {
                const float temp = vs[i], tx = ip[2 * i], ty = ip[2 * i + 1];
                int j = i;
                while (j > 0 && vs[j - 1] > temp) {
                    vs[j] = vs[j - 1]; ip[2 * j] = ip[2 * j - 2]; ip[2 * j + 1] = ip[2 * j - 1];
                    --j;
                }
                vs[j] = temp; ip[2 * j] = tx; ip[2 * j + 1] = ty;
            }
    }
    float area = 0.0f; /* area :221-228, trangle_area :215-218 */
    for (int i = 0; i < n - 2; ++i) {
        const float *a = ip, *b = ip + 2 * i + 2, *c = ip + 2 * i + 4;
        area += fabsf(((a[0] - c[0]) * (b[1] - c[1]) - (a[1] - c[1]) * (b[0] - c[0])) / 2.0f);
    }
    return area;
}
/* iou [N,K]: iou[n,k] = devRotateIoUEval(query[k], boxes[n], criterion) (:248-260, call site :299-301) */
void y3d_o_rotate_iou_eval(const float *boxes, int N, const float *query, int K, int criterion, float *iou) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const float *r1 = query + 5 * k, *r2 = boxes + 5 * n;
            const float a1 = r1[2] * r1[3], a2 = r2[2] * r2[3];
            const float ai = riou_inter(r1, r2);
            float v;
            if (criterion == -1) v = ai / (a1 + a2 - ai);
            else if (criterion == 0) v = ai / a1;
            else if (criterion == 1) v = ai / a2;
            else v = ai;
            iou[(long)n * K + k] = v;
        }
}
