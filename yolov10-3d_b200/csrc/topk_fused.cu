// topk_fused.cu -- candidate walk + per-GT top-k + claims of the fused loss, specialised (sm_100a).
//
// Same algorithm and the same results, bit for bit, as tal_topk_kernel (assign.cu) -- select_candidates_in_gts,
// get_box_metrics and select_topk_candidates of reference ultralytics/utils/tal.py:108-167,218-235 -- for the ONE
// configuration v8DetectionLoss uses (loss.py:176,231-238): scores are the head's class logits, alpha = 0.5, beta = 6,
// candidates constrained to the GT box, anchors on the level grid, predicted boxes [B,A,4] in grid units.  The generic
// kernel carries every other mode (3D similarity, free anchors, probability inputs, several warps per GT) and had grown
// past the instruction cache (its top stall was `no_instruction`); this one is a fraction of its size and issues about
// half the instructions per candidate cell:
//   * a level's in-GT rectangle is walked in row-packed units: lane = (row within the unit, column), fixed per level, so
//     a cell's index is two adds and a multiply (the generic kernel maps a flat centre-out index through a division for
//     every cell); row groups are still taken centre-out, and the first trip is the most central unit alone (seed);
//   * one 16-byte load fetches a candidate's box, the fast IoU bound works in grid units against the GT box divided
//     by the level's stride (no per-candidate scaling);
//   * the next work ticket is requested before the claims go out, so its round trip overlaps theirs.
#include "assign.cuh"

#ifndef Y3D_FSORTMERGE
#define Y3D_FSORTMERGE 6
#endif

namespace y3d {

constexpr int kFU = 4;                  // units (<= 32 cells each) per round trip
constexpr int kFQ = 32 * kFU + 32;      // candidate queue slots per warp
constexpr int kFSortMerge = Y3D_FSORTMERGE;         // keys above the threshold from which the sorted list is merged, not inserted into
constexpr int kFCtasPerSM = 8;

__device__ __forceinline__ float f_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float f_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// exact alignment metric of a pair: sigmoid(x)^0.5 * clamp(CIoU, 0)^6 -- the operation sequence of pair_eval /
// pair_metric_core (assign.cuh) for score_mode 1, alpha 0.5, beta 6, use_2d only
__device__ __forceinline__ float fused_metric(float4 gbox, float gat1, float4 pbox_grid, float st, float x) {
    const float sb = dm::sqrt_(1.0f / (1.0f + expf(-x)));
    const float4 p = make_float4(dm::mul(pbox_grid.x, st), dm::mul(pbox_grid.y, st), dm::mul(pbox_grid.z, st),
                                 dm::mul(pbox_grid.w, st));
    float o = dm::ciou(gbox, p, gat1);
    o = o < 0.0f ? 0.0f : o;
    const float o2 = dm::mul(o, o);
    const float o4 = dm::mul(o2, o2);
    return dm::mul(sb, dm::mul(o4, o2));
}

#ifdef Y3D_TIMING
// developer builds: [0] ~first warp past the dependency wait [1] last warp exit [2] ~first warp exit [3] last prologue end
__device__ unsigned long long g_tkf_tl[16 * 4];
__device__ unsigned g_tkf_step;  // bumped by the warp that draws the first ticket past the end
__device__ __forceinline__ unsigned long long tkf_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
extern "C" int y3d_debug_read_topk_timeline(unsigned long long *host, int reset) {
    int rc = (int)cudaMemcpyFromSymbol(host, g_tkf_tl, sizeof(unsigned long long) * 16 * 4);
    if (reset) {
        unsigned long long z[16 * 4] = {};
        cudaMemcpyToSymbol(g_tkf_tl, z, sizeof(z));
        unsigned zero = 0;
        cudaMemcpyToSymbol(g_tkf_step, &zero, sizeof(zero));
    }
    return rc;
}
#define TKF_STEP0() const unsigned tkf_step = *(volatile unsigned *)&g_tkf_step
#define TKF_MIN(i) do { if ((threadIdx.x & 31) == 0) atomicMax(&g_tkf_tl[(tkf_step & 15) * 4 + (i)], ~tkf_now()); } while (0)
#define TKF_MAX(i) do { if ((threadIdx.x & 31) == 0) atomicMax(&g_tkf_tl[(tkf_step & 15) * 4 + (i)], tkf_now()); } while (0)
#define TKF_BUMP(cond) do { if ((cond) && (threadIdx.x & 31) == 0) atomicAdd(&g_tkf_step, 1u); } while (0)
#else
#define TKF_STEP0()
#define TKF_MIN(i)
#define TKF_MAX(i)
#define TKF_BUMP(cond)
#endif

// grid: persistent CTAs of kTopkWarps warps; a warp pulls (branch, image, GT) items from cc.work_counter.
// dynamic shared memory: ORDERED ? (2 * kOrdClasses * n_branch * B + 1) ints : 0
template <bool ORDERED, bool REC>
__global__ void __launch_bounds__(kTopkWarps * 32, kFCtasPerSM)
tal_topk_fused_kernel(const __grid_constant__ AssignCtx2 cc, const int n_branch) {
    extern __shared__ int s_pref[];
    __shared__ float4 q_k[kTopkWarps][kFQ];  // (anchor index bits, class logit, metric^2 bound, -)
    __shared__ float4 q_b[kTopkWarps][kFQ];  // predicted box, grid units
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // programmatic dependent launch, both ways (see tal_topk_kernel)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    TKF_STEP0();
    TKF_MIN(0);
    const unsigned lt_mask = (1u << lane) - 1u;
    const int n_img = cc.c[0].B, M = cc.c[0].M, A = cc.c[0].A;
    const int per_branch = n_img * M, total = per_branch * n_branch;
    const int n_seg = kOrdClasses * n_branch * n_img;
    // ORDERED: s_pref = exclusive prefix of the item counts over the (class, branch, image) segments of the longest-first
    // order; s_info[seg] = image | branch << 12 | (position of the class in the image's sorted list) << 16
    int *s_info = s_pref + (n_seg + 1);
    if (ORDERED) {
        for (int bb = threadIdx.x; bb < n_img; bb += blockDim.x) {
            // (plain load: the CTAs of an SM share the line through L1 instead of all 1184 CTAs queueing at one L2 slice)
            const int4 v = reinterpret_cast<const int4 *>(cc.ord_cnt)[bb];
            const int cnt[4] = {v.x, v.y, v.z, v.w};
            int basec = 0;
#pragma unroll
            for (int cl = 0; cl < kOrdClasses; ++cl) {
                for (int zz = 0; zz < n_branch; ++zz) {
                    const int sg = (cl * n_branch + zz) * n_img + bb;
                    s_pref[sg] = cnt[cl];
                    s_info[sg] = bb | (zz << 12) | (basec << 16);
                }
                basec += cnt[cl];
            }
        }
        __syncthreads();
        {   // exclusive scan by the whole CTA: thread-contiguous chunks, warp scans, warp totals through shared memory
            __shared__ int s_wtot[kTopkWarps];
            const int nth = kTopkWarps * 32, tid = threadIdx.x;
            const int chunk = (n_seg + nth - 1) / nth;
            const int lo = min(tid * chunk, n_seg), hi = min(lo + chunk, n_seg);
            int sum = 0;
            for (int sg = lo; sg < hi; ++sg) sum += s_pref[sg];
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane == 31) s_wtot[wid] = inc;
            __syncthreads();
            int wbase = 0, all = 0;
#pragma unroll
            for (int w2 = 0; w2 < kTopkWarps; ++w2) {
                const int v = s_wtot[w2];
                wbase += w2 < wid ? v : 0;
                all += v;
            }
            int run = wbase + inc - sum;
            for (int sg = lo; sg < hi; ++sg) {
                const int v = s_pref[sg];
                s_pref[sg] = run;
                run += v;
            }
            if (tid == 0) s_pref[n_seg] = all;
        }
        __syncthreads();
    }
    const int seg_stride = (n_seg + 32) / 32;  // 32 * seg_stride >= n_seg + 1; <= 64
    TKF_MAX(3);
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(cc.work_counter, 1);
    for (;;) {
        const int item = __shfl_sync(0xffffffffu, ticket, 0);
        int z, b, m;
        GtRec g;
        if (ORDERED) {
            TKF_BUMP(item == s_pref[n_seg]);
            if (item >= s_pref[n_seg]) break;
            // last segment whose prefix is <= item: a 32-way step, then up to 64 entries
            const int i1 = min(lane * seg_stride, n_seg);
            const unsigned m1 = __ballot_sync(0xffffffffu, s_pref[i1] <= item);
            const int base = (31 - __clz(m1)) * seg_stride;
            const int i2 = base + lane;
            const unsigned m2 = __ballot_sync(0xffffffffu, lane < seg_stride && i2 < n_seg && s_pref[i2] <= item);
            const unsigned m3 = __ballot_sync(0xffffffffu, lane + 32 < seg_stride && i2 + 32 < n_seg && s_pref[min(i2 + 32, n_seg)] <= item);
            const int seg = base + (m3 ? 63 - __clz(m3) : 31 - __clz(m2));
            const int info = s_info[seg];
            b = info & 0xfff;
            z = (info >> 12) & 1;
            const int off = item - s_pref[seg] + (info >> 16);  // position in the image's class-sorted list
            const int4 *rec = reinterpret_cast<const int4 *>(cc.ord_list) + ((long long)b * M + off) * 2;
            const int4 r0 = __ldcg(rec), r1 = __ldcg(rec + 1);
            m = r0.x;
            g.label = r0.y;
            g.box = make_float4(__int_as_float(r0.z), __int_as_float(r0.w), __int_as_float(r1.x), __int_as_float(r1.y));
            g.valid = 1;
            g.at1 = dm::box1_atan(g.box);
        } else {
            if (item >= total) break;
            z = item >= per_branch ? 1 : 0;
            const int gt_id = item - z * per_branch;
            b = gt_id / M;
            m = gt_id - b * M;
            if (!gt_valid(cc.c[z], b, m)) {  // padded GT: top-k indices forced to 0 and masked out (tal.py:155,104)
                if (lane == 0) ticket = atomicAdd(cc.work_counter, 1);
                continue;
            }
            g = load_gt(cc.c[z], b, m);
        }
        const AssignCtx &c = cc.c[z];
        const int k = c.k;
        const int nl = c.t.nl;
        // exact in-GT rectangle of every level: lane 4*l + e finds edge e (x lo, x hi, y lo, y hi) of level l with the
        // very comparison of select_candidates_in_gts (see tal_topk_kernel)
        int my_edge = 0;
        if (lane < 4 * nl) {
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
        unsigned long long tk = 0ull;   // lane-distributed sorted list (descending), lanes >= k unused
        unsigned long long thr = 0ull;  // key of the k-th entry (warp-uniform)
        // phase 0: the first k anchors enter the list even when outside the GT or at metric 0 (they are what a dense
        // stable top-k picks among zeros), see assign.cuh
        unsigned long long key0 = 0ull;
        if (lane < k && lane < A) {
            float ax, ay, st;
            anchor_px(c, lane, ax, ay, st);
            const int cin = (int)dm::in_gt(ax, ay, g.box);
            float metric = 0.0f;
            if (cin) {
                const int lv = level_of(c.t, lane);
                const float x0 = c.t.ptr[lv][(long long)b * c.t.sB[lv] + (long long)(64 + g.label) * c.t.sC[lv] + (lane - c.t.start[lv])];
                const float4 b0 = reinterpret_cast<const float4 *>(c.pd_bboxes)[(long long)b * A + lane];
                metric = fused_metric(g.box, g.at1, b0, st, x0);
            }
            key0 = tk_key(metric, lane, cin);
        }
        // ---- candidate walk.  One loop body: first entry, then trips while the queue cannot fill a warp, pops otherwise,
        //      the remainder at the end.
        int qn = 0, qh = 0;  // FIFO ring
        int l = -1, t0 = 0, n_units = 0;
        int L_w = 0, L_start = 0, L_cA = 0, L_cB = 0, L_rA = 0, L_rB = 0, L_R = 1, L_nseg = 1, L_wseg = 1, L_mid = 0;
        int rsub = 0, csub = 0;
        bool lane_ok = false;
        float L_st = 1.0f, gs_area = 0.f;
        float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *L_srow = nullptr;
        const float4 *L_box = nullptr;
        bool more = true, seed_trip = true, flush_seed = false;
        for (bool first = true;; first = false) {
            unsigned long long key = 0ull;
            if (first) {
                key = key0;
            } else if (qn < 32 && more && !(flush_seed && qn > 0)) {
                while (t0 >= n_units) {  // next level with a non-empty rectangle
                    if (++l >= nl) { more = false; break; }
                    L_cA = __shfl_sync(0xffffffffu, my_edge, 4 * l);
                    L_cB = __shfl_sync(0xffffffffu, my_edge, 4 * l + 1);
                    L_rA = __shfl_sync(0xffffffffu, my_edge, 4 * l + 2);
                    L_rB = __shfl_sync(0xffffffffu, my_edge, 4 * l + 3);
                    const int ncols = L_cB - L_cA + 1, nrows = L_rB - L_rA + 1;
                    t0 = 0;
                    n_units = 0;
                    if (ncols <= 0 || nrows <= 0) continue;
                    // (small non-negative integers: quotients through the SFU reciprocal; a true quotient's fraction is a
                    // multiple of 1/32 at least, far above the approximation's error, and the bias keeps exact ones exact)
                    L_nseg = (ncols + 31) >> 5;
                    L_wseg = L_nseg == 1 ? ncols : (ncols + L_nseg - 1) / L_nseg;
                    const float rw = f_rcp((float)L_wseg);
                    L_R = (int)(32.0f * rw + 0.01f);
                    rsub = (int)((float)lane * rw + 0.01f);
                    csub = lane - rsub * L_wseg;
                    lane_ok = rsub < L_R;
                    const int nrg = (int)((float)(nrows + L_R - 1) * f_rcp((float)L_R) + 0.01f);
                    L_mid = (nrg - 1) >> 1;
                    n_units = nrg * L_nseg;
                    L_w = c.t.w[l];
                    L_start = c.t.start[l];
                    L_st = c.t.stride[l];
                    L_srow = c.t.ptr[l] + (long long)b * c.t.sB[l] + (long long)(64 + g.label) * c.t.sC[l];
                    L_box = reinterpret_cast<const float4 *>(c.pd_bboxes) + (long long)b * A + L_start;
                    const float inv = __frcp_rn(L_st);
                    gs = make_float4(g.box.x * inv, g.box.y * inv, g.box.z * inv, g.box.w * inv);
                    gs_area = (gs.z - gs.x) * (gs.w - gs.y);
                }
                if (!more) continue;
                // ---- stage 1: one round trip for the gathers of up to kFU cells per lane, then the bounds
                const int nu = min(seed_trip ? 1 : kFU, n_units - t0);  // units of this trip (warp-uniform)
                bool h[kFU];
                int ca[kFU];
                float x[kFU];
                float4 bx[kFU];
#pragma unroll
                for (int u = 0; u < kFU; ++u) {
                    h[u] = false;
                    ca[u] = 0;
                    x[u] = 0.f;
                    bx[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int t = t0 + u;
                    if (u < nu) {
                        int rg = t, sg = 0;
                        if (L_nseg > 1) { rg = t / L_nseg; sg = t - rg * L_nseg; }
                        // centre-out over the row groups: good candidates first raise the k-th metric early
                        const int rgo = (rg & 1) ? L_mid + ((rg + 1) >> 1) : L_mid - (rg >> 1);
                        const int row = L_rA + rgo * L_R + rsub, col = L_cA + sg * L_wseg + csub;
                        const int cell = row * L_w + col;
                        ca[u] = L_start + cell;
                        h[u] = lane_ok && row <= L_rB && col <= L_cB && ca[u] >= k;
                        if (h[u]) {
                            x[u] = L_srow[cell];
                            bx[u] = L_box[cell];
                        }
                    }
                }
                const int nlive = nu;
                t0 += nu;
                flush_seed = seed_trip;  // evaluate the seed candidates before walking on
                seed_trip = false;
                const float thr_m = __uint_as_float((unsigned)(thr >> 32));
                const float tq = thr_m * thr_m * 0.999f;  // threshold of the squared-domain bound (conservative)
#pragma unroll
                for (int u = 0; u < kFU; ++u) {
                    if (u >= nlive) break;  // warp-uniform
                    // metric^2 = sigmoid(x) * CIoU^12 bounded from above with SFU approximations and safety factors (no
                    // division, no square root, CIoU <= IoU); the exact metric is computed in stage 2, for the survivors
                    const float e = f_ex2(-x[u] * 1.4426950408889634f);
                    const float s_ub = 1.0002f * f_rcp(__fmaf_rn(0.9998f, e, 1.0f));
                    const float iw = fmaxf(fminf(gs.z, bx[u].z) - fmaxf(gs.x, bx[u].x), 0.f);
                    const float ih = fmaxf(fminf(gs.w, bx[u].w) - fmaxf(gs.y, bx[u].y), 0.f);
                    const float inter = iw * ih;
                    const float uni = gs_area + (bx[u].z - bx[u].x) * (bx[u].w - bx[u].y) - inter;
                    const float iou = fminf(inter * f_rcp(fmaxf(uni, 1e-30f)) * 1.0002f, 1.0f);
                    const float i2 = iou * iou, i4 = i2 * i2;
                    const float ub = s_ub * (i4 * i4 * i4) * 1.0001f;
                    const bool pass = h[u] && ub >= tq;
                    const unsigned bal = __ballot_sync(0xffffffffu, pass);
                    if (pass) {
                        int pos = qh + qn + __popc(bal & lt_mask);
                        if (pos >= kFQ) pos -= kFQ;
                        q_k[wid][pos] = make_float4(__int_as_float(ca[u]), x[u], ub, L_st);
                        q_b[wid][pos] = bx[u];
                    }
                    qn += __popc(bal);
                }
                __syncwarp();
                continue;
            } else if (qn > 0) {
                // ---- stage 2: pop full lanes while trips remain, the rest at the end (no loads: pure arithmetic)
                flush_seed = false;
                const int take = qn < 32 ? qn : 32;
                int qi = qh + lane;
                if (qi >= kFQ) qi -= kFQ;
                if (lane < take) {
                    const float4 kk = q_k[wid][qi];
                    const float thr_now = __uint_as_float((unsigned)(thr >> 32));
                    if (kk.z >= thr_now * thr_now * 0.999f) {
                        const float metric = fused_metric(g.box, g.at1, q_b[wid][qi], kk.w, kk.y);
                        if (metric > 0.0f) key = tk_key(metric, __float_as_int(kk.x), 1);
                    }
                }
                qn -= take;
                qh += take;
                if (qh >= kFQ) qh -= kFQ;
                __syncwarp();
            } else {
                break;
            }
            // ---- sorted-list update with this lane's candidate key (0 = none), as in tal_topk_kernel: many keys -> sort
            //      them across the lanes and merge with the list; a few -> insert one by one
            if (__popc(__ballot_sync(0xffffffffu, key > thr)) >= kFSortMerge) {
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
            }
        }
        // the next ticket's round trip overlaps the claims'
        if (lane == 0) ticket = atomicAdd(cc.work_counter, 1);
        // ---- claims: mask_pos = mask_topk * mask_in_gts * mask_gt (tal.py:104)
        const bool claiming = lane < k && tk != 0ull && (tk & 1ull);
        if (REC) {
            // the claim atomic, the slot allocation and the gathers of the pair's loss inputs are in flight together
            const unsigned cmask = __ballot_sync(0xffffffffu, claiming);
            if (cmask) {
                int a = 0;
                unsigned long long old = 0ull;
                ClaimTerms ct;
                if (claiming) {
                    a = tk_anchor(tk);
                    old = atomicAdd(c.claim + (long long)b * A + a, (1ull << 32) | (unsigned long long)m);
                }
                const int leader = __ffs(cmask) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(c.list_count + b, __popc(cmask));
                if (claiming) ct = claim_terms(c, b, a, g);
                base = __shfl_sync(0xffffffffu, base, leader);
                const int slot = base + __popc(cmask & lt_mask);
                if (claiming && slot < c.rec_cap) {
                    float4 *r = c.rec + ((long long)b * c.rec_cap + slot) * kRecF4;
                    const int first_bit = (old >> 32) == 0 ? (int)0x80000000 : 0;  // the first claimer owns the anchor downstream
                    r[0] = make_float4(__int_as_float(a | first_bit), __int_as_float(m), __uint_as_float((unsigned)(tk >> 32)), ct.xlab);
                    r[1] = ct.terms;
                    r[2] = ct.box;
                }
            }
            // this GT is through: the finishing kernel takes the image when its count is complete
            __syncwarp();
            if (lane == 0 && c.topk_done) red_release_add1(c.topk_done + b);
        } else if (claiming) {
            const int a = tk_anchor(tk);
            const unsigned long long old = atomicAdd(c.claim + (long long)b * A + a, (1ull << 32) | (unsigned long long)m);
            if ((old >> 32) == 0) {  // first claim of this anchor: publish it
                const int pos = atomicAdd(c.list_count + b, 1);
                if (pos < c.list_cap) c.list_a[(long long)b * c.list_cap + pos] = a;
            }
            // the finishing kernel will gather the two DFL bins around each target distance (loss.py:99-113): start
            // pulling them into L2 now
            const int lv = level_of(c.t, a);
            const int cell = a - c.t.start[lv];
            const float gx = (float)(cell % c.t.w[lv]) + 0.5f, gy = (float)(cell / c.t.w[lv]) + 0.5f;
            float4 tb;
            float tt[4];
            dfl_target(g.box, c.t.stride[lv], gx, gy, tb, tt);
            const float *hp = c.t.ptr[lv] + (long long)b * c.t.sB[lv] + cell;
#pragma unroll
            for (int side = 0; side < 4; ++side) {
                const int tl = (int)tt[side];
                prefetch_l2(hp + (long long)(side * 16 + tl) * c.t.sC[lv]);
                prefetch_l2(hp + (long long)(side * 16 + tl + 1) * c.t.sC[lv]);
            }
        }
        __syncwarp();
    }  // item loop
    TKF_MIN(2);
    TKF_MAX(1);
}

int device_sm_count() {
    static int sms = 0;  // one device per process (one process per GPU)
    if (sms == 0) {
        int dev = 0, v = 148;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms = v;
    }
    return sms;
}

template <bool ORDERED, bool REC>
static int launch_fused(const AssignCtx2 &cc, int n, long long items, size_t smem, cudaStream_t s, bool pdl) {
    static int ctas_per_sm = 0;  // per instantiation and shared-memory size (host-side calculation, cached)
    static size_t ctas_smem = ~(size_t)0;
    if (ctas_per_sm == 0 || ctas_smem != smem) {
        int v = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, tal_topk_fused_kernel<ORDERED, REC>, kTopkWarps * 32, smem) !=
                cudaSuccess || v <= 0)
            v = 4;
        ctas_per_sm = v;
        ctas_smem = smem;
    }
    long long blocks = (items + kTopkWarps - 1) / kTopkWarps;
    const long long cap = (long long)ctas_per_sm * device_sm_count();
    if (blocks > cap) blocks = cap;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(kTopkWarps * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t le = cudaLaunchKernelEx(&cfg, tal_topk_fused_kernel<ORDERED, REC>, cc, n);
    if (le != cudaSuccess) return (int)le;
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

int assign_run_topk_fused(const AssignCtx2 &cc, int n, cudaStream_t s, bool pdl) {
    for (int z = 0; z < n; ++z) {
        const AssignCtx &c = cc.c[z];
        if (c.score_mode != 1 || c.cls_ch0 != 64 || c.alpha != 0.5f || c.beta != 6.0f || !c.use_2d || c.use_3d ||
            !c.constrain || !c.use_grid || !c.box_grid_units || c.box_soa || c.mask_gt || !cc.work_counter)
            return Y3D_EUNSUPPORTED;
        if (c.B != cc.c[0].B || c.M != cc.c[0].M || c.A != cc.c[0].A) return Y3D_EUNSUPPORTED;
        if ((c.rec != nullptr) != (cc.c[0].rec != nullptr) || (!c.rec && !c.list_a)) return Y3D_EUNSUPPORTED;
    }
    const AssignCtx &c = cc.c[0];
    const long long items = (long long)c.B * c.M * n;
    if (items >= 0x7fffffffLL) return Y3D_EUNSUPPORTED;
    // with few GTs (e.g. 32 images x 50) the generic kernel splits every GT over several warps instead
    if (items < 16LL * device_sm_count()) return Y3D_EUNSUPPORTED;
    const bool ordered = cc.ord_cnt != nullptr;
    const size_t smem = ordered ? sizeof(int) * (size_t)(2 * kOrdClasses * n * c.B + 1) : 0;  // prefix + info tables
    if (ordered && (kOrdClasses * n * c.B > kOrdMaxSeg || c.B >= 4096 || c.M >= 32768 || n > 2)) return Y3D_EUNSUPPORTED;
    const bool rec = c.rec != nullptr;
    if (ordered) return rec ? launch_fused<true, true>(cc, n, items, smem, s, pdl) : launch_fused<true, false>(cc, n, items, smem, s, pdl);
    return rec ? launch_fused<false, true>(cc, n, items, smem, s, pdl) : launch_fused<false, false>(cc, n, items, smem, s, pdl);
}

}  // namespace y3d
