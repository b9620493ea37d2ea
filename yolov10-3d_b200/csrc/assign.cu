// assign.cu -- task-aligned assigners (sm_100a), API-faithful entry points.
//   y3d_tal_assign    TaskAlignedAssigner.forward    reference ultralytics/utils/tal.py:44-264
//   y3d_tal_assign3d  TaskAlignedAssigner3d.forward  reference ultralytics/utils/tal.py:391-700
//                     get_3d_keypoints               reference ultralytics/utils/keypoint_utils.py:11-118
// The core (candidate walk, per-GT top-k, conflict resolution) is documented in assign.cuh.
// Algorithmic bytes of the API-faithful 2D assigner per image and branch:
//   read 4*A*(nc+6) + 24*M, write A*(8 + 16 + 4*nc + 1 + 8); the dense target_scores write dominates.
#include "assign.cuh"

namespace y3d {

// grid: persistent CTAs (rectangle walk) or one CTA per (branch, image, GT) (all-anchor scan); block kTopkWarps*32.
// wpg = warps per GT: 1 (rectangle walk) or kTopkWarps (all-anchor scan; the warps' lists are merged through shared
// memory).
//
// The kernel is instruction-issue bound (about 1.6 M (GT, anchor) pairs at cfg2, each an exactly rounded CIoU), so it is
// organised to evaluate as few pairs exactly as possible, on full lanes, with a small code footprint (one call site
// per stage, heavy arithmetic out of line: hundreds of divergent warps share the instruction cache):
//  * the in-GT test (tal.py:218-235) is separable and monotone in the column / row index, so the exact set of in-GT
//    cells of a level is a rectangle whose four edges are found with the reference's own fp32 comparisons -- no
//    per-cell test, no wasted lanes; the rectangles of all levels form one flat index space, walked centre-out
//    128 cells at a time (4 per lane, their score and box gathers issued together);
//  * two upper bounds drop a candidate before any exact arithmetic: sb = score^alpha (CIoU^beta, sim^gamma <= 1) and
//    sb * IoU^beta with a fast IoU (CIoU <= IoU); a candidate whose bound is below the current k-th metric can never
//    enter the list.  Survivors are compacted into a per-warp queue;
//  * stage 2 pops 32 at a time (full lanes, no loads): exact CIoU / keypoint similarity and the sorted-list update.
constexpr int kTopkU = 4;             // cells per lane and round trip
constexpr int kTopkQ = 32 * kTopkU + 32;  // queue slots per warp
constexpr int kTopkSortMerge = 6;  // keys above the threshold from which the sorted list is merged, not inserted into

// score^alpha of a candidate and an upper bound of its metric: sb * IoU^beta with a fast IoU (CIoU <= IoU; the factor
// 1.0001 covers the fast arithmetic and the eps terms of the exact expression).  beta < 0: no IoU bound.  Out of line:
// one copy for the four candidates of a round trip.
static __device__ __noinline__ float2 cand_bounds(float x, int score_mode, float alpha, float4 box, float st,
                                                  float4 gbox, float g_area, float beta) {
    const float s = score_mode == 0 ? x : 1.0f / (1.0f + expf(-x));  // pred_scores.detach().sigmoid() loss.py:232
    const float sb = dm::pow_(s, alpha);
    float ub = sb;
    if (beta >= 0.0f) {
        const float4 p = make_float4(box.x * st, box.y * st, box.z * st, box.w * st);
        const float iw = fmaxf(fminf(gbox.z, p.z) - fmaxf(gbox.x, p.x), 0.f);
        const float ih = fmaxf(fminf(gbox.w, p.w) - fmaxf(gbox.y, p.y), 0.f);
        const float inter = iw * ih;
        const float uni = g_area + (p.z - p.x) * (p.w - p.y) - inter;
        const float iou = fminf(__fdividef(inter, fmaxf(uni, 1e-30f)) * 1.0001f, 1.0f);
        ub = sb * dm::pow_(iou, beta) * 1.0001f;
    }
    return make_float2(sb, ub);
}

#ifdef Y3D_TIMING
// developer instrumentation (tools/phase_timing.py): cycles and event counts per phase, summed over all warps
__device__ unsigned long long g_topk_prof[16];
#define TK_T0() long long tk_t_ = clock64()
#define TK_ACC(i)                                                                          \
    do {                                                                                   \
        const long long n_ = clock64();                                                    \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_topk_prof[i], (unsigned long long)(n_ - tk_t_)); \
        tk_t_ = n_;                                                                        \
    } while (0)
#define TK_CNT(i, v)                                                                       \
    do {                                                                                   \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_topk_prof[i], (unsigned long long)(v)); \
    } while (0)
extern "C" int y3d_debug_read_topk_prof(unsigned long long *host, int reset) {
    int rc = (int)cudaMemcpyFromSymbol(host, g_topk_prof, sizeof(unsigned long long) * 16);
    if (reset) {
        unsigned long long z[16] = {};
        cudaMemcpyToSymbol(g_topk_prof, z, sizeof(z));
    }
    return rc;
}
#else
#define TK_T0()
#define TK_ACC(i)
#define TK_CNT(i, v)
#endif
#ifdef Y3D_TAILTIME
// developer instrumentation (tools/tail_timing.py): start / exit time and item counts of every persistent warp
__device__ unsigned long long g_topk_tail[4096 * 4];
__device__ __forceinline__ unsigned long long tk_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
extern "C" int y3d_debug_read_topk_tail(unsigned long long *host, int n) {
    return (int)cudaMemcpyFromSymbol(host, g_topk_tail, sizeof(unsigned long long) * n);
}
#endif

struct LvlWalk {  // one level's exact in-GT rectangle (per warp, shared memory)
    int off, ncols, c0, r0;        // first flat index, columns, first column / row
    int w, start, cmid, rmid;      // grid width, first anchor of the level, centre-out pivots
    float inv, st;                 // 1 / ncols, stride
    const float *srow;             // score_mode 1: label-channel row of this image and level, minus `start`
};

__global__ void __launch_bounds__(kTopkWarps * 32, 6) tal_topk_kernel(AssignCtx2 cc, int wpg, int n_branch) {
    __shared__ int q_a[kTopkWarps][kTopkQ];
    __shared__ float q_s[kTopkWarps][kTopkQ];
    __shared__ float q_u[kTopkWarps][kTopkQ];
    __shared__ float4 q_b[kTopkWarps][kTopkQ];
    __shared__ LvlWalk walk[kTopkWarps][Y3D_MAX_LEVELS];
    __shared__ unsigned long long mrg[kTopkWarps][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // programmatic dependent launch, both ways: in the fused loss this grid is scheduled while the streaming kernel
    // drains and must wait for its completion before reading anything (a no-op after an ordinary launch); then it lets
    // its own dependent (the finishing kernel) be scheduled as this grid drains -- that one waits for this grid itself
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    const int per_branch = cc.c[0].B * cc.c[0].M;  // host checks n_branch * B * M < 2^31
    const int total = per_branch * n_branch;
    const int wsub = wpg == 1 ? 0 : wid;  // this warp's share of the GT's chunks
    const unsigned lt_mask = (1u << lane) - 1u;
    // Work items = (branch, image, GT).  One warp per GT: persistent warps pull items from a global counter (padded
    // GTs cost one load, big GTs do not stall a whole wave).  Several warps per GT: one item per CTA, static.
    bool static_done = false;
    TK_T0();
#ifdef Y3D_TAILTIME
    const unsigned long long tail_t0 = tk_now();
    unsigned tail_items = 0, tail_valid = 0;
#endif
    // Longest-first order (fused loss): the streaming kernel has sorted every image's valid GTs into size classes.
    // Segment s = (class, branch, image) holds ord_cnt[image][class] items; s_pref = exclusive prefix over the segments.
    // (Handing the very biggest GTs to whole CTAs, merged as in the static mode, was measured too: the four warps of
    // such an item cost more warp-time than the shorter tail gives back.)
    __shared__ int s_pref[kOrdMaxSeg + 1];
    const bool ordered = wpg == 1 && cc.ord_cnt != nullptr;
    const int n_img = cc.c[0].B;
    const int n_seg = kOrdClasses * n_branch * n_img;
    if (ordered) {
        for (int s = threadIdx.x; s < n_seg; s += blockDim.x)
            s_pref[s] = __ldcg(cc.ord_cnt + (s % n_img) * 4 + s / (n_branch * n_img));  // written by the primary grid: not .nc
        __syncthreads();
        if (wid == 0) {  // in-place exclusive scan: lane-contiguous chunks, warp scan of the chunk sums
            const int chunk = (n_seg + 31) / 32;
            const int lo = min(lane * chunk, n_seg), hi = min(lo + chunk, n_seg);
            int sum = 0;
            for (int s = lo; s < hi; ++s) sum += s_pref[s];
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            int run = inc - sum;
            for (int s = lo; s < hi; ++s) {
                const int v = s_pref[s];
                s_pref[s] = run;
                run += v;
            }
            if (lane == 31) s_pref[n_seg] = inc;
        }
        __syncthreads();
    }
    const int seg_stride = (n_seg + 32) / 32;  // 32 * seg_stride >= n_seg + 1; <= 64
    for (int item = (int)blockIdx.x;;) {
    if (wpg == 1) {
        int it = 0;
        if (lane == 0) it = atomicAdd(cc.work_counter, 1);
        item = __shfl_sync(0xffffffffu, it, 0);
    } else {
        if (static_done) break;  // static mapping: a single pass
        static_done = true;
    }
    int z, b, m;
    GtRec og;
    if (ordered) {
        if (item >= s_pref[n_seg]) break;
        // last segment whose prefix is <= item: a 32-way step, then up to 64 entries
        const int i1 = min(lane * seg_stride, n_seg);
        const unsigned m1 = __ballot_sync(0xffffffffu, s_pref[i1] <= item);
        const int base = (31 - __clz(m1)) * seg_stride;
        const int i2 = base + lane;
        const unsigned m2 = __ballot_sync(0xffffffffu, lane < seg_stride && i2 < n_seg && s_pref[i2] <= item);
        const unsigned m3 = __ballot_sync(0xffffffffu, lane + 32 < seg_stride && i2 + 32 < n_seg && s_pref[min(i2 + 32, n_seg)] <= item);
        const int seg = base + (m3 ? 63 - __clz(m3) : 31 - __clz(m2));
        const int cls = seg / (n_branch * n_img);
        z = (seg / n_img) % n_branch;
        b = seg % n_img;
        int off = item - s_pref[seg];  // position in the image's class-sorted list
        for (int c2 = 0; c2 < cls; ++c2) {
            const int s2 = (c2 * n_branch + z) * n_img + b;
            off += s_pref[s2 + 1] - s_pref[s2];
        }
        // the sorted list carries the GT record itself (index, label, box): one load, no second round trip
        const int4 *rec = reinterpret_cast<const int4 *>(cc.ord_list) + ((long long)b * cc.c[0].M + off) * 2;
        const int4 r0 = __ldcg(rec), r1 = __ldcg(rec + 1);
        m = r0.x;
        og.label = r0.y;
        og.box = make_float4(__int_as_float(r0.z), __int_as_float(r0.w), __int_as_float(r1.x), __int_as_float(r1.y));
    } else {
        if (item >= total) break;
        z = item >= per_branch ? 1 : 0;
        const int gt_id = item - z * per_branch;
        b = gt_id / cc.c[0].M;
        m = gt_id - b * cc.c[0].M;
    }
#ifdef Y3D_TAILTIME
    ++tail_items;
#endif
    const AssignCtx &c = cc.c[z];
    if (!ordered && !gt_valid(c, b, m)) {  // padded GT: top-k indices forced to 0 and masked out (tal.py:155,104)
        if (wpg == 1) continue;
        break;
    }
    TK_ACC(0);  // item fetch + validity (incl. padded GTs)
#ifdef Y3D_TAILTIME
    ++tail_valid;
#endif
    if (ordered) {
        og.valid = true;
        og.at1 = dm::box1_atan(og.box);
    } else {
        og = load_gt(c, b, m);
    }
    const GtRec g = og;
    const int k = c.k;
    const bool rect = c.use_grid && c.constrain;
    const bool prune = c.beta >= 0.0f && c.gamma >= 0.0f;  // the upper bounds need non-negative exponents
    const bool iou_bound = prune && c.use_2d &&
                           (c.beta == 1.0f || c.beta == 2.0f || c.beta == 3.0f || c.beta == 4.0f || c.beta == 6.0f);
    const float g_area = (g.box.z - g.box.x) * (g.box.w - g.box.y);
    const bool fast_sq = iou_bound && c.score_mode == 1 && c.alpha == 0.5f && c.beta == 6.0f;

    int cells = c.A;  // candidates in the flat index space
    int off1 = 0x7fffffff, off2 = 0x7fffffff, off3 = 0x7fffffff;
    if (rect) {
        // exact in-GT rectangle of every level: in_gt(ax, ay) = min(ax-x1, ay-y1, x2-ax, y2-ay) > 1e-9 splits into four
        // monotone edge tests; each edge is located with the very comparison the dense test uses
        // The 4 edges x nl levels are independent: lane 4*l + e finds edge e (x lo, x hi, y lo, y hi) of level l.
        // lo edge: first index i with (i + 0.5) * st - lo > 1e-9; hi edge: last index with hi - (i + 0.5) * st > 1e-9.
        // The real-valued estimate is off by at most one cell, so a window of four cells around it brackets the flip
        // of the (monotone) comparison; anything else falls back to a linear search.
        int my_edge = 0;
        if (lane < 4 * c.t.nl) {
            const int lv = lane >> 2, e = lane & 3;
            const float st = c.t.stride[lv];
            const int n = e < 2 ? c.t.w[lv] : c.t.h[lv];
            const float v = e == 0 ? g.box.x : e == 1 ? g.box.z : e == 2 ? g.box.y : g.box.w;
            const bool hi = e & 1;
            auto ok = [&](int i) {
                const float p = dm::mul((float)i + 0.5f, st);
                return hi ? dm::sub(v, p) > 1e-9f : dm::sub(p, v) > 1e-9f;
            };
            const float r = v / st - 0.5f;
            if (!hi) {
                const int f = (int)floorf(r);
                int ed = f - 1 + (ok(f - 1) ? 0 : 1) + (ok(f) ? 0 : 1) + (ok(f + 1) ? 0 : 1) + (ok(f + 2) ? 0 : 1);
                if (ok(f - 1) || !ok(f + 2)) {  // flip not inside the window
                    ed = f - 1;
                    while (ed > 0 && ok(ed - 1)) --ed;
                    while (ed < n && !ok(ed)) ++ed;
                }
                my_edge = max(ed, 0);
            } else {
                const int f = (int)ceilf(r);
                int ed = f + 1 - (ok(f + 1) ? 0 : 1) - (ok(f) ? 0 : 1) - (ok(f - 1) ? 0 : 1) - (ok(f - 2) ? 0 : 1);
                if (ok(f + 1) || !ok(f - 2)) {
                    ed = f + 1;
                    while (ed < n - 1 && ok(ed + 1)) ++ed;
                    while (ed >= 0 && !ok(ed)) --ed;
                }
                my_edge = min(ed, n - 1);
            }
        }
        cells = 0;
        for (int l = 0; l < c.t.nl; ++l) {
            const float st = c.t.stride[l];
            const int w = c.t.w[l];
            const int cA = __shfl_sync(0xffffffffu, my_edge, 4 * l), cB = __shfl_sync(0xffffffffu, my_edge, 4 * l + 1);
            const int rA = __shfl_sync(0xffffffffu, my_edge, 4 * l + 2), rB = __shfl_sync(0xffffffffu, my_edge, 4 * l + 3);
            const int ncols = cB >= cA ? cB - cA + 1 : 0, nrows = rB >= rA ? rB - rA + 1 : 0;
            if (lane == 0) {
                LvlWalk &L = walk[wid][l];
                L.off = cells; L.ncols = ncols > 0 ? ncols : 1; L.c0 = cA; L.r0 = rA;
                L.w = w; L.start = c.t.start[l]; L.cmid = (ncols - 1) >> 1; L.rmid = (nrows - 1) >> 1;
                L.inv = __frcp_rn((float)(ncols > 0 ? ncols : 1)); L.st = st;
                L.srow = c.score_mode == 1 ? c.t.ptr[l] + (long long)b * c.t.sB[l] +
                                                 (long long)(c.cls_ch0 + g.label) * c.t.sC[l] - c.t.start[l]
                                           : nullptr;
            }
            cells += ncols * nrows;
            if (l == 0) off1 = cells;
            else if (l == 1) off2 = cells;
            else if (l == 2) off3 = cells;
        }
        __syncwarp();
    }

    TK_ACC(1);  // GT load + exact rectangles
    TK_CNT(8, 1);
    TK_CNT(9, cells);
    unsigned long long tk = 0ull;   // lane-distributed sorted list (descending), lanes >= k unused
    unsigned long long thr = 0ull;  // key of the k-th entry (warp-uniform)

    // phase 0 (first warp of the GT): the first k anchors enter the list even when outside the GT or at metric 0
    // (they are what a dense stable top-k picks among zeros), see assign.cuh
    unsigned long long key0 = 0ull;
    if (wsub == 0 && lane < k && lane < c.A) {
        float ax, ay, st;
        anchor_px(c, lane, ax, ay, st);
        const int cin = c.constrain ? (int)dm::in_gt(ax, ay, g.box) : 1;
        float metric = 0.0f, ovl;
        if (cin) pair_eval(c, b, m, g, lane, metric, ovl);
        key0 = tk_key(metric, lane, cin);
    }

    // phase 1: candidates = anchors >= k inside the GT.  One loop body: first entry, then trips while the queue cannot
    // fill a warp, pops otherwise, the remainder at the end.
    TK_ACC(2);  // phase 0 (first k anchors)
    int qn = 0, qh = 0;  // FIFO ring: candidates are evaluated in walk order (centre first)
    int i0 = wsub * (32 * kTopkU);
    // the first round trip of a one-warp walk takes only the 32 most central cells: their metrics set the k-th
    // threshold before the bulk of the rectangle is tested against it
    bool seed_trip = wpg == 1 && rect, flush_seed = false;
    bool done = false;
    for (bool first = true;; first = false) {
        unsigned long long key = 0ull;
        if (first) {
            key = key0;
        } else if (qn < 32 && !done && !(flush_seed && qn > 0)) {
            if (i0 >= cells) { done = true; continue; }
            // ---- stage 1: locate, one round trip for the gathers of up to kTopkU candidates per lane, bounds
            bool h[kTopkU];
            int a[kTopkU];
            float x[kTopkU], sv[kTopkU];
            float4 bx[kTopkU];
#pragma unroll
            for (int u = 0; u < kTopkU; ++u) {
                const int j = i0 + u * 32 + lane;
                const bool live = !seed_trip || u == 0;
                h[u] = false;
                a[u] = 0;
                x[u] = 0.f;
                sv[u] = 1.0f;
                bx[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live && j < cells) {
                    if (rect) {
                        const int l = (j >= off1 ? 1 : 0) + (j >= off2 ? 1 : 0) + (j >= off3 ? 1 : 0);
                        const LvlWalk &L = walk[wid][l];
                        const int jj = j - L.off;
                        int r = (int)(((float)jj + 0.5f) * L.inv);  // jj / ncols; fixed up if the product rounded across
                        int cc_ = jj - r * L.ncols;
                        if (cc_ < 0) { --r; cc_ += L.ncols; }
                        else if (cc_ >= L.ncols) { ++r; cc_ -= L.ncols; }
                        // centre-out inside the rectangle: good candidates first raise the k-th metric early
                        const int col = L.c0 + ((cc_ & 1) ? L.cmid + ((cc_ + 1) >> 1) : L.cmid - (cc_ >> 1));
                        const int row = L.r0 + ((r & 1) ? L.rmid + ((r + 1) >> 1) : L.rmid - (r >> 1));
                        a[u] = L.start + row * L.w + col;
                        sv[u] = L.st;
                        h[u] = a[u] >= k;
                        if (h[u]) x[u] = L.srow ? L.srow[a[u]] : pair_load_score(c, b, a[u], g.label);
                    } else {
                        a[u] = j;
                        h[u] = j >= k;
                        if (h[u] && c.constrain) {
                            float ax, ay, s_;
                            anchor_px(c, j, ax, ay, s_);
                            h[u] = dm::in_gt(ax, ay, g.box);
                        }
                        if (h[u]) x[u] = pair_load_score(c, b, j, g.label);
                        if (h[u] && c.box_grid_units) sv[u] = c.t.stride[level_of(c.t, j)];
                    }
                    if (h[u]) bx[u] = pair_load_box(c, b, a[u]).box;
                }
            }
            i0 += seed_trip ? 32 : wpg * (32 * kTopkU);
            flush_seed = seed_trip;  // evaluate the seed candidates before walking on
            seed_trip = false;
            const unsigned tm = prune ? (unsigned)(thr >> 32) : 0u;
            const float thr_m = __uint_as_float(tm);
            const float tq = thr_m * thr_m * 0.999f;  // threshold of the squared-domain bound (conservative)
#pragma unroll
            for (int u = 0; u < kTopkU; ++u) {
                float payload = 0.f, ub = 0.f;
                bool pass = false;
                if (fast_sq) {
                    // the common configuration (logits in, alpha = 0.5, beta = 6): bound metric^2 = sigmoid(x) * CIoU^12
                    // from above with SFU approximations and safety factors -- no division, no square root; the exact
                    // score^alpha is computed in stage 2, for the survivors only
                    const float e = im::ex2a(-x[u] * im::kLog2e);
                    const float s_ub = __fdividef(1.0002f, __fmaf_rn(0.9998f, e, 1.0f));
                    const float stv = c.box_grid_units ? sv[u] : 1.0f;
                    const float px1 = bx[u].x * stv, py1 = bx[u].y * stv, px2 = bx[u].z * stv, py2 = bx[u].w * stv;
                    const float iw = fmaxf(fminf(g.box.z, px2) - fmaxf(g.box.x, px1), 0.f);
                    const float ih = fmaxf(fminf(g.box.w, py2) - fmaxf(g.box.y, py1), 0.f);
                    const float inter = iw * ih;
                    const float uni = g_area + (px2 - px1) * (py2 - py1) - inter;
                    const float iou = fminf(__fdividef(inter, fmaxf(uni, 1e-30f)) * 1.0001f, 1.0f);  // CIoU <= IoU
                    const float i2 = iou * iou, i4 = i2 * i2;
                    ub = s_ub * (i4 * i4 * i4) * 1.0001f;
                    payload = x[u];
                    pass = h[u] && ub >= tq;
                } else {
                    if (__any_sync(0xffffffffu, h[u])) {
                        const float2 r2 = cand_bounds(x[u], c.score_mode, c.alpha, bx[u],
                                                      c.box_grid_units ? sv[u] : 1.0f, g.box, g_area,
                                                      iou_bound ? c.beta : -1.0f);
                        payload = h[u] ? r2.x : 0.f;  // score^alpha
                        ub = r2.y;
                    }
                    pass = h[u] && payload > 0.0f && __float_as_uint(ub) >= tm;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, pass);
                if (pass) {
                    int pos = qh + qn + __popc(bal & lt_mask);
                    if (pos >= kTopkQ) pos -= kTopkQ;
                    q_a[wid][pos] = a[u];
                    q_s[wid][pos] = payload;
                    q_u[wid][pos] = ub;
                    q_b[wid][pos] = bx[u];
                }
                qn += __popc(bal);
            }
            __syncwarp();
            TK_ACC(3);  // stage 1 trips
            TK_CNT(10, 1);
            continue;
        } else if (qn > 0) {
            // ---- stage 2: pop full lanes while trips remain, the rest at the end (no loads: pure arithmetic)
            flush_seed = false;
            const int take = qn < 32 ? qn : 32;
            int qi = qh + lane;
            if (qi >= kTopkQ) qi -= kTopkQ;
            if (lane < take) {
                const int a2 = q_a[wid][qi];
                const float thr_now = __uint_as_float((unsigned)(thr >> 32));
                const bool alive = !prune || (fast_sq ? q_u[wid][qi] >= thr_now * thr_now * 0.999f
                                                      : __float_as_uint(q_u[wid][qi]) >= (unsigned)(thr >> 32));
                if (alive) {
                    PairRaw raw;
                    raw.box = q_b[wid][qi];
                    raw.s = 0.0f;
                    float ovl;
                    const float sb = fast_sq ? dm::pow_(pair_score(c, q_s[wid][qi]), c.alpha) : q_s[wid][qi];
                    const float metric = pair_metric(c, b, m, g, a2, raw, sb, ovl);
                    if (metric > 0.0f) key = tk_key(metric, a2, 1);
                }
            }
            qn -= take;
            qh += take;
            if (qh >= kTopkQ) qh -= kTopkQ;
            __syncwarp();
            TK_ACC(4);  // stage 2 pops
            TK_CNT(11, 1);
            TK_CNT(12, take);
        } else {
            break;
        }
        // ---- sorted-list update with this lane's candidate key (0 = none).  Many keys at once (a GT's first pops,
        //      when the list is still short): sort the keys across the lanes (bitonic network), reverse them against
        //      the sorted list -- the lane-wise maximum holds the 32 largest of both and is bitonic -- and sort that
        //      with five more exchanges.  A few keys: insert them one by one.
        if (__popc(__ballot_sync(0xffffffffu, key > thr)) >= kTopkSortMerge) {
            unsigned long long v = key > thr ? key : 0ull;
#pragma unroll
            for (int ksz = 2; ksz <= 32; ksz <<= 1) {
#pragma unroll
                for (int j = ksz >> 1; j > 0; j >>= 1) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, j);
                    const bool keep_max = ((lane & ksz) == 0) == ((lane & j) == 0);  // descending: lane 0 ends largest
                    v = (o > v) == keep_max ? o : v;
                }
            }
            const unsigned long long vr = __shfl_sync(0xffffffffu, v, 31 - lane);
            tk = vr > tk ? vr : tk;
#pragma unroll
            for (int j = 16; j > 0; j >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, tk, j);
                tk = (o > tk) == ((lane & j) == 0) ? o : tk;
            }
            thr = __shfl_sync(0xffffffffu, tk, k - 1);
            key = 0ull;
            TK_CNT(13, 1);
        }
        for (;;) {
            const unsigned mk = __ballot_sync(0xffffffffu, key > thr);
            if (!mk) break;
            const int src = __ffs(mk) - 1;
            const unsigned long long xk = __shfl_sync(0xffffffffu, key, src);
            const int pos = __popc(__ballot_sync(0xffffffffu, tk > xk && lane < k));
            const unsigned long long up = __shfl_up_sync(0xffffffffu, tk, 1);
            if (lane == pos) tk = xk;
            else if (lane > pos) tk = up;
            thr = __shfl_sync(0xffffffffu, tk, k - 1);
            if (lane == src) key = 0ull;
            TK_CNT(13, 1);
        }
        TK_ACC(5);  // list updates
    }

    if (wpg > 1) {  // merge the per-warp lists into the first warp's
        mrg[wid][lane] = lane < k ? tk : 0ull;
        __syncthreads();
        if (wid != 0) break;
        for (int w2 = 1; w2 < wpg; ++w2) {
            unsigned long long km = mrg[w2][lane];
            for (;;) {
                const unsigned mk = __ballot_sync(0xffffffffu, km > thr);
                if (!mk) break;
                const int src = __ffs(mk) - 1;
                const unsigned long long xk = __shfl_sync(0xffffffffu, km, src);
                const int pos = __popc(__ballot_sync(0xffffffffu, tk > xk && lane < k));
                const unsigned long long up = __shfl_up_sync(0xffffffffu, tk, 1);
                if (lane == pos) tk = xk;
                else if (lane > pos) tk = up;
                thr = __shfl_sync(0xffffffffu, tk, k - 1);
                if (lane == src) km = 0ull;
            }
        }
    }
    // claims: mask_pos = mask_topk * mask_in_gts * mask_gt (tal.py:104)
    if (c.rec) {
        // record path: the claim atomic, the slot allocation and the gathers of the pair's loss inputs are all in flight
        // together (one round trip), then the record is written
        const bool claiming = lane < k && tk != 0ull && (tk & 1ull);
        const unsigned cmask = __ballot_sync(0xffffffffu, claiming);
        if (cmask) {
            int a = 0;
            unsigned long long old = 0ull;
            ClaimTerms ct;
            if (claiming) {
                a = tk_anchor(tk);
                old = atomicAdd(c.claim + (long long)b * c.A + a, (1ull << 32) | (unsigned long long)m);
            }
            const int leader = __ffs(cmask) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(c.list_count + b, __popc(cmask));
            if (claiming) ct = claim_terms(c, b, a, g);
            base = __shfl_sync(0xffffffffu, base, leader);
            const int slot = base + __popc(cmask & lt_mask);
            if (claiming && slot < c.rec_cap) {
                float4 *r = c.rec + ((long long)b * c.rec_cap + slot) * kRecF4;
                const int first = (old >> 32) == 0 ? (int)0x80000000 : 0;  // the first claimer owns the anchor downstream
                r[0] = make_float4(__int_as_float(a | first), __int_as_float(m), __uint_as_float((unsigned)(tk >> 32)), ct.xlab);
                r[1] = ct.terms;
                r[2] = ct.box;
            }
        }
        __syncwarp();
        if (lane == 0 && c.topk_done) red_release_add1(c.topk_done + b);  // this GT is through (see assign.cuh)
    } else if (lane < k && tk != 0ull && (tk & 1ull)) {
        const int a = tk_anchor(tk);
        const unsigned long long old = atomicAdd(c.claim + (long long)b * c.A + a, (1ull << 32) | (unsigned long long)m);
        if (c.list_a) {
            if ((old >> 32) == 0) {  // first claim of this anchor: publish it
                const int pos = atomicAdd(c.list_count + b, 1);
                if (pos < c.list_cap) c.list_a[(long long)b * c.list_cap + pos] = a;
            }
            if (c.score_mode == 1 && c.cls_ch0 == 64) {
                // the finishing kernel will gather the two DFL bins around each target distance (loss.py:99-113):
                // start pulling them into L2 now
                const int lv = level_of(c.t, a);
                const int cell = a - c.t.start[lv];
                const float sv = c.t.stride[lv];
                const float ax = (float)(cell % c.t.w[lv]) + 0.5f, ay = (float)(cell / c.t.w[lv]) + 0.5f;
                const float ltrb[4] = {ax - g.box.x / sv, ay - g.box.y / sv, g.box.z / sv - ax, g.box.w / sv - ay};
                const float *hp = c.t.ptr[lv] + (long long)b * c.t.sB[lv] + cell;
#pragma unroll
                for (int side = 0; side < 4; ++side) {
                    const int tl = (int)fminf(fmaxf(ltrb[side], 0.0f), 14.99f);
                    prefetch_l2(hp + (long long)(side * 16 + tl) * c.t.sC[lv]);
                    prefetch_l2(hp + (long long)(side * 16 + tl + 1) * c.t.sC[lv]);
                }
            }
        }
    }
    __syncwarp();
    TK_ACC(6);  // claims + prefetch
    }  // item loop
#ifdef Y3D_TAILTIME
    if (lane == 0 && blockIdx.x * kTopkWarps + wid < 4096) {
        unsigned long long *o = g_topk_tail + (blockIdx.x * kTopkWarps + wid) * 4;
        o[0] = tail_t0; o[1] = tk_now(); o[2] = tail_items; o[3] = tail_valid;
    }
#endif
}

// grid (ceil(A/256), B, n_branch); dynamic smem: M GtRec
__global__ void __launch_bounds__(256) tal_resolve_kernel(AssignCtx2 cc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GtRec *gts = reinterpret_cast<GtRec *>(smem_raw);
    const AssignCtx &c = cc.c[blockIdx.z];
    const int b = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const long long o = (long long)b * c.A + a;
    // (a no-op after an ordinary launch; the fused 3D loss launches this grid programmatically dependent on the top-k grid)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    const unsigned long long cl = a < c.A ? c.claim[o] : 0ull;
    const int cnt = (int)(cl >> 32);
    const int any_multi = __syncthreads_or(cnt > 1);  // does any anchor of this block need the all-GT scan?
    if (any_multi) {
        for (int m = threadIdx.x; m < c.M; m += blockDim.x) gts[m] = load_gt(c, b, m);
        __syncthreads();
    }
    // multiply-claimed anchors (select_highest_overlaps tal.py:252-263: argmax over all GTs of the overlap, first
    // maximum): the warp takes them one at a time, lanes over the GTs; only the overlap itself is evaluated (the CIoU
    // for the 2D assigner, the keypoint similarity when use_3d)
    int gi = cnt == 1 ? (int)(cl & 0xffffffffull) : -1;
    {
        const int lane = threadIdx.x & 31;
        const int kflags = c.kps_l2 ? 4 : 0;
        unsigned cm = __ballot_sync(0xffffffffu, cnt > 1);
        while (cm) {
            const int src = __ffs(cm) - 1;
            cm &= cm - 1;
            const int a_s = __shfl_sync(0xffffffffu, a, src);
            float ax, ay, st;
            anchor_px(c, a_s, ax, ay, st);
            float4 pbox = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 *pk = nullptr;
            if (c.use_3d) pk = reinterpret_cast<const float4 *>(c.pd_kps + ((long long)b * c.A + a_s) * 24);
            else pbox = pair_box(c, pair_load_box(c, b, a_s), a_s);
            unsigned long long best = 0xffffffffull;  // overlap 0 at GT 0: what the reference's argmax of zeros gives
            for (int m = lane; m < c.M; m += 32) {
                const GtRec g = gts[m];
                if (g.valid && (!c.constrain || dm::in_gt(ax, ay, g.box))) {
                    float ovl;
                    if (c.use_3d) {
                        ovl = kps_sim(pk, reinterpret_cast<const float4 *>(c.gt_kps + ((long long)b * c.M + m) * 24), kflags);
                    } else {
                        ovl = dm::ciou(g.box, pbox, g.at1);
                        ovl = ovl < 0.0f ? 0.0f : ovl;
                    }
                    const unsigned long long key =
                        ((unsigned long long)__float_as_uint(ovl) << 32) | (unsigned long long)(0xffffffffu - (unsigned)m);
                    best = key > best ? key : best;
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, best, off);
                best = o2 > best ? o2 : best;
            }
            if (lane == src) gi = (int)(0xffffffffu - (unsigned)(best & 0xffffffffull));
        }
    }
    if (a >= c.A) return;
    float alignv = 0.0f;
    if (cnt > 0) {
        float ax, ay, st;
        anchor_px(c, a, ax, ay, st);
        const GtRec g = cnt > 1 ? gts[gi] : load_gt(c, b, gi);
        float metric = 0.0f, ovl = 0.0f;
        bool sel = g.valid && (!c.constrain || dm::in_gt(ax, ay, g.box));
        if (sel) pair_eval(c, b, gi, g, a, metric, ovl);
        alignv = metric;
        atomicMax(c.pos_align + (long long)b * c.M + gi, __float_as_int(metric));  // values >= 0: int order == float order
        atomicMax(c.pos_ov + (long long)b * c.M + gi, __float_as_int(ovl));
    }
    c.tgi[o] = gi;
    c.alignv[o] = alignv;
}

int assign_run_topk(const AssignCtx2 &cc, int n, cudaStream_t s, bool pdl) {
    const AssignCtx &c = cc.c[0];
    const long long items = (long long)c.B * c.M * n;
    // one persistent warp per GT fills the machine only when there are many GTs; with few (e.g. KITTI: 32 x 50) the
    // kernel's duration is one GT's latency, so each GT is split over kTopkWarps warps instead
    const int sms = device_sm_count();
    // (3D similarity: every candidate costs a 96-byte gather and 24 coordinate terms, so the split pays for longer)
    const bool few = items < (c.use_3d ? 32LL : 16LL) * sms;
    const int wpg = (c.use_grid && c.constrain && cc.work_counter && !few) ? 1 : kTopkWarps;
    if (items >= 0x7fffffffLL) return Y3D_EUNSUPPORTED;
    long long blocks = wpg == 1 ? (items + kTopkWarps - 1) / kTopkWarps : items;
    if (wpg == 1 && blocks > 6LL * sms) blocks = 6LL * sms;  // persistent: no more CTAs than fit at once
    if (pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)blocks);
        cfg.blockDim = dim3(kTopkWarps * 32);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t le = cudaLaunchKernelEx(&cfg, tal_topk_kernel, cc, wpg, n);
        if (le != cudaSuccess) return (int)le;
    } else {
        tal_topk_kernel<<<(unsigned)blocks, kTopkWarps * 32, 0, s>>>(cc, wpg, n);
    }
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

int assign_run_core(const AssignCtx2 &cc, int n, cudaStream_t s, cudaEvent_t after_topk, bool pdl) {
    const AssignCtx &c = cc.c[0];
    int rc0 = assign_run_topk(cc, n, s, pdl);
    if (rc0) return rc0;
    if (after_topk) cudaEventRecord(after_topk, s);
    size_t smem = sizeof(GtRec) * (size_t)c.M;
    static size_t smem_limit = 48 * 1024;  // raised once per process to the largest size asked for
    if (smem > smem_limit) {
        cudaError_t e = cudaFuncSetAttribute(tal_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        smem_limit = smem;
    }
    dim3 grid((c.A + 255) / 256, c.B, n);
    if (pdl) {  // scheduled as the top-k grid drains; waits for it before it reads
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t le = cudaLaunchKernelEx(&cfg, tal_resolve_kernel, cc);
        if (le != cudaSuccess) return (int)le;
    } else {
        tal_resolve_kernel<<<grid, 256, smem, s>>>(cc);
    }
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

// ------------------------------------------------------------------------------------------------ emit kernels
// get_targets (tal.py:169-216 / 651-700) + normalisation (tal.py:88-92): per-anchor small outputs
__global__ void __launch_bounds__(256) tal_emit_small_kernel(AssignCtx c, int64_t *__restrict__ t_lab,
                                                             float *__restrict__ t_box, uint8_t *__restrict__ fg,
                                                             int64_t *__restrict__ t_gi, float *__restrict__ norm_ws,
                                                             int *__restrict__ lab_ws, const float *__restrict__ gts17,
                                                             float *__restrict__ t_vals) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)c.B * c.A) return;
    const int b = (int)(i / c.A);
    const int gi = c.tgi[i];
    const bool isfg = gi >= 0;
    const long long gm = (long long)b * c.M + (isfg ? gi : 0);  // background anchors index GT 0 (argmax of zeros)
    long long lab = (long long)c.gt_labels[gm * c.gl_stride];
    lab = lab < 0 ? 0 : lab;  // target_labels.clamp_(0)
    const float norm = isfg ? assigned_norm(c, b, gi, c.alignv[i]) : 0.0f;
    norm_ws[i] = norm;
    lab_ws[i] = isfg ? (int)lab : -1;
    if (t_lab) t_lab[i] = lab;
    if (t_box) {
        const float *g = c.gt_bboxes + gm * c.gb_stride;
        *reinterpret_cast<float4 *>(t_box + 4 * i) = make_float4(g[0], g[1], g[2], g[3]);
    }
    if (t_vals) {
        const float *g = gts17 + gm * 17 + 5;
#pragma unroll
        for (int j = 0; j < 12; ++j) t_vals[i * 12 + j] = g[j];
    }
    fg[i] = (uint8_t)isfg;
    t_gi[i] = isfg ? gi : 0;
}

// dense target_scores [B,A,nc] = one_hot(label) * fg * norm: 128-bit streaming stores
template <int VEC>
__global__ void __launch_bounds__(256) tal_emit_scores_kernel(const float *__restrict__ norm_ws,
                                                              const int *__restrict__ lab_ws, int nc, long long total,
                                                              float *__restrict__ t_sc) {
    long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long e0 = f * VEC;
    if (e0 >= total) return;
    long long i = e0 / nc;
    int c0 = (int)(e0 - i * nc);
    if constexpr (VEC == 4) {  // nc % 4 == 0: the four elements belong to one anchor
        int lab = lab_ws[i];
        float nv = norm_ws[i];
        int d = lab - c0;
        float4 v = make_float4(d == 0 ? nv : 0.f, d == 1 ? nv : 0.f, d == 2 ? nv : 0.f, d == 3 ? nv : 0.f);
        stg_stream4(t_sc + e0, v);
    } else {
        t_sc[e0] = (lab_ws[i] == c0) ? norm_ws[i] : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------ 3D keypoints
// keypoints24(): assign.cuh
__global__ void __launch_bounds__(128) kps_pred_kernel(const float *__restrict__ pd_scores,
                                                       const float *__restrict__ pd_3d, const float *__restrict__ anc,
                                                       const float *__restrict__ stride,
                                                       const float *__restrict__ calibs,
                                                       const float *__restrict__ mean_sizes, int B, int A, int nc,
                                                       float *__restrict__ pd_kps) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * A) return;
    const int b = (int)(i / A), a = (int)(i % A);
    const float *p = pd_3d + i * 31;
    const float *s = pd_scores + i * nc;
    const float st = stride[a];
    const float c3x = dm::add(anc[2 * a], dm::mul(p[0], st));  // decode_3d_center tal.py:454-456
    const float c3y = dm::add(anc[2 * a + 1], dm::mul(p[1], st));
    int cls = 0;  // decode_3d_size tal.py:458-462 (argmax, first max)
    float best = s[0];
    for (int c = 1; c < nc; ++c) {
        float v = s[c];
        if (v > best) { best = v; cls = c; }
    }
    const float sh_ = dm::add(mean_sizes[3 * cls], p[2]), sw_ = dm::add(mean_sizes[3 * cls + 1], p[3]),
                sl_ = dm::add(mean_sizes[3 * cls + 2], p[4]);
    int hb = 0;
    float hbv = p[5];
    for (int j = 1; j < 12; ++j) {
        float v = p[5 + j];
        if (v > hbv) { hbv = v; hb = j; }
    }
    float out[24];
    keypoints24(c3x, c3y, p[29], sh_, sw_, sl_, hb, p[5 + 12 + hb], calibs + 6 * b, out);
    float4 *o = reinterpret_cast<float4 *>(pd_kps + i * 24);
#pragma unroll
    for (int j = 0; j < 6; ++j) o[j] = make_float4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
}

__global__ void kps_gt_kernel(const float *__restrict__ gts, const float *__restrict__ calibs,
                              const float *__restrict__ mean_sizes, int B, int M, int nc,
                              float *__restrict__ gt_kps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * M) return;
    const int b = i / M;
    const float *g = gts + (long long)i * 17;
    int lab = (int)g[0];
    lab = lab < 0 ? 0 : (lab >= nc ? nc - 1 : lab);
    const float sh_ = dm::add(mean_sizes[3 * lab], g[11]), sw_ = dm::add(mean_sizes[3 * lab + 1], g[12]),
                sl_ = dm::add(mean_sizes[3 * lab + 2], g[13]);  // add_cls_mean_size tal.py:605-609
    float out[24];
    keypoints24(g[9], g[10], g[14], sh_, sw_, sl_, (int)g[15], g[16], calibs + 6 * b, out);
    for (int j = 0; j < 24; ++j) gt_kps[(long long)i * 24 + j] = out[j];
}

static int run_emit(const AssignCtx &c, void *ws, const AssignWs &w, int64_t *t_lab, float *t_box, float *t_sc,
                    uint8_t *fg, int64_t *t_gi, const float *gts17, float *t_vals, cudaStream_t s) {
    float *norm_ws = (float *)((char *)ws + w.off_norm);
    int *lab_ws = (int *)((char *)ws + w.off_lab);
    long long n = (long long)c.B * c.A;
    tal_emit_small_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(c, t_lab, t_box, fg, t_gi, norm_ws, lab_ws, gts17,
                                                                      t_vals);
    Y3D_CHECK_LAUNCH();
    if (t_sc) {
        long long total = n * c.nc;
        if (c.nc % 4 == 0 && ((uintptr_t)t_sc) % 16 == 0) {
            long long th = total / 4;
            tal_emit_scores_kernel<4><<<(unsigned)((th + 255) / 256), 256, 0, s>>>(norm_ws, lab_ws, c.nc, total, t_sc);
        } else {
            tal_emit_scores_kernel<1><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(norm_ws, lab_ws, c.nc, total, t_sc);
        }
        Y3D_CHECK_LAUNCH();
    }
    return Y3D_OK;
}

size_t assign_workspace_bytes(int B, int A, int M) { return assign_ws_layout(B, A, M).total; }

int launch_kps_gt(const float *gts, const float *calibs, const float *mean_sizes, int B, int M, int nc, float *gt_kps,
                  cudaStream_t s) {
    kps_gt_kernel<<<(B * M + 127) / 128, 128, 0, s>>>(gts, calibs, mean_sizes, B, M, nc, gt_kps);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

}  // namespace y3d

using namespace y3d;

extern "C" int y3d_tal_assign(const float *pd_scores, int64_t ss_B, int64_t ss_A, int64_t ss_C,
                              const float *pd_bboxes, const float *anc_points, const float *gt_labels,
                              const float *gt_bboxes, const float *mask_gt, int B, int A, int nc, int M, int topk,
                              float alpha, float beta, float eps, const int *lvl_hw, const float *lvl_stride, int nl,
                              int64_t *target_labels, float *target_bboxes, float *target_scores, uint8_t *fg_mask,
                              int64_t *target_gt_idx, void *ws, size_t ws_bytes, void *stream) {
    if (!pd_scores || !pd_bboxes || !gt_labels || !gt_bboxes || !mask_gt || !fg_mask || !target_gt_idx)
        return Y3D_EINVAL;
    if (B < 0 || A < 1 || nc < 1 || M < 1 || topk < 1 || topk > A) return Y3D_EINVAL;
    if (topk > Y3D_MAX_TOPK) return Y3D_EUNSUPPORTED;
    if (((uintptr_t)pd_bboxes) % 16 || (target_bboxes && ((uintptr_t)target_bboxes) % 16)) return Y3D_EALIGN;
    AssignCtx c{};
    if (lvl_hw && lvl_stride && nl > 0) {
        int a = make_level_table(c.t, nullptr, nullptr, nullptr, lvl_hw, lvl_stride, nl);
        if (a != A) return Y3D_EINVAL;
        c.use_grid = 1;
    } else {
        if (!anc_points) return Y3D_EINVAL;
        if (((uintptr_t)anc_points) % 8) return Y3D_EALIGN;
        c.use_grid = 0;
    }
    AssignWs w = assign_ws_layout(B, A, M);
    if (!ws || ws_bytes < w.total) return Y3D_EWORKSPACE;
    if (((uintptr_t)ws) % 256) return Y3D_EALIGN;
    if (B == 0) return Y3D_OK;
    c.score_mode = 0;
    c.pd_scores = pd_scores; c.ssB = ss_B; c.ssA = ss_A; c.ssC = ss_C;
    c.pd_bboxes = pd_bboxes; c.box_grid_units = 0;
    c.anc = anc_points;
    c.gt_labels = gt_labels; c.gl_stride = 1;
    c.gt_bboxes = gt_bboxes; c.gb_stride = 4;
    c.mask_gt = mask_gt;
    c.B = B; c.A = A; c.nc = nc; c.M = M; c.k = topk;
    c.alpha = alpha; c.beta = beta; c.gamma = 1.0f; c.eps = eps;
    c.use_2d = 1; c.use_3d = 0; c.kps_l2 = 0; c.constrain = 1;
    assign_bind_ws(c, ws, w);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync((char *)ws + w.off_cnt, 0, w.zero_bytes, s);
    if (e != cudaSuccess) return (int)e;
    AssignCtx2 cc{};
    cc.c[0] = c;
    cc.work_counter = (int *)((char *)ws + w.off_work);
    int rc = assign_run_core(cc, 1, s);
    if (rc) return rc;
    return run_emit(c, ws, w, target_labels, target_bboxes, target_scores, fg_mask, target_gt_idx, nullptr, nullptr, s);
}

extern "C" int y3d_tal_assign3d(const float *pd_scores, const float *pd_bboxes, const float *pd_3d,
                                const float *anc_points, const float *stride, const float *gts,
                                const float *mask_gt, const float *calibs, const float *mean_sizes, int B, int A,
                                int nc, int M, int topk, float alpha, float beta, float gamma, float eps, int flags,
                                const int *lvl_hw, const float *lvl_stride, int nl, int64_t *target_labels,
                                float *target_scores, float *target_vals, uint8_t *fg_mask, int64_t *target_gt_idx,
                                float *pd_keypoints, float *gt_keypoints, void *ws, size_t ws_bytes, void *stream) {
    if (!pd_scores || !pd_bboxes || !pd_3d || !anc_points || !stride || !gts || !mask_gt || !calibs || !mean_sizes ||
        !fg_mask || !target_gt_idx || !pd_keypoints || !gt_keypoints)
        return Y3D_EINVAL;
    if (B < 0 || A < 1 || nc < 1 || M < 1 || topk < 1 || topk > A) return Y3D_EINVAL;
    if (topk > Y3D_MAX_TOPK) return Y3D_EUNSUPPORTED;
    const int use_2d = flags & 1, use_3d = (flags >> 1) & 1, kps_l2 = (flags >> 2) & 1, constrain = (flags >> 3) & 1;
    if (!use_2d && !use_3d) return Y3D_EINVAL;  // tal.py:486 RuntimeError
    if (((uintptr_t)pd_bboxes) % 16 || ((uintptr_t)pd_keypoints) % 16 || ((uintptr_t)gt_keypoints) % 16 ||
        ((uintptr_t)anc_points) % 8)
        return Y3D_EALIGN;
    AssignCtx c{};
    if (lvl_hw && lvl_stride && nl > 0) {
        int a = make_level_table(c.t, nullptr, nullptr, nullptr, lvl_hw, lvl_stride, nl);
        if (a != A) return Y3D_EINVAL;
        c.use_grid = 1;
    }
    AssignWs w = assign_ws_layout(B, A, M);
    if (!ws || ws_bytes < w.total) return Y3D_EWORKSPACE;
    if (((uintptr_t)ws) % 256) return Y3D_EALIGN;
    if (B == 0) return Y3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    long long n = (long long)B * A;
    kps_pred_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(pd_scores, pd_3d, anc_points, stride, calibs, mean_sizes,
                                                               B, A, nc, pd_keypoints);
    Y3D_CHECK_LAUNCH();
    kps_gt_kernel<<<(B * M + 127) / 128, 128, 0, s>>>(gts, calibs, mean_sizes, B, M, nc, gt_keypoints);
    Y3D_CHECK_LAUNCH();
    c.score_mode = 0;
    c.pd_scores = pd_scores; c.ssB = (long long)A * nc; c.ssA = nc; c.ssC = 1;
    c.pd_bboxes = pd_bboxes; c.box_grid_units = 0;
    c.anc = anc_points;
    c.gt_labels = gts; c.gl_stride = 17;
    c.gt_bboxes = gts + 1; c.gb_stride = 17;
    c.mask_gt = mask_gt;
    c.B = B; c.A = A; c.nc = nc; c.M = M; c.k = topk;
    c.alpha = alpha; c.beta = beta; c.gamma = gamma; c.eps = eps;
    c.use_2d = use_2d; c.use_3d = use_3d; c.kps_l2 = kps_l2; c.constrain = constrain;
    c.pd_kps = pd_keypoints; c.gt_kps = gt_keypoints;
    assign_bind_ws(c, ws, w);
    cudaError_t e = cudaMemsetAsync((char *)ws + w.off_cnt, 0, w.zero_bytes, s);
    if (e != cudaSuccess) return (int)e;
    AssignCtx2 cc{};
    cc.c[0] = c;
    cc.work_counter = (int *)((char *)ws + w.off_work);
    int rc = assign_run_core(cc, 1, s);
    if (rc) return rc;
    return run_emit(c, ws, w, target_labels, nullptr, target_scores, fg_mask, target_gt_idx, gts, target_vals, s);
}
