// loss.cuh -- declarations shared by loss.cu (fused forward) and loss_bwd.cu (backward): constants and the
// workspace layout.  The backward pass reads what the forward pass left in the caller-owned workspace (claimed-anchor
// lists, and the claim words carrying the assigned GT and alignment weight), so both sides must agree on the layout.
#pragma once
#include "assign.cuh"

namespace y3d {

constexpr int kR = 16;  // reg_max (head.py:37)
constexpr int kFinishApMaxM = 128;  // GTs per image up to which the anchor-parallel finishing kernel is used

struct LossWs {  // all offsets 256-byte aligned; per-branch blocks are contiguous
    size_t claim, boxes, lse, list_a, list_gi, list_al, rec, list_count, img_cnt, topk_done, pos, per_branch;
    size_t off_counter, off_pfg, off_pbce, off_ord_cnt, off_ord_list, total;
    int cap, rcap, n_bce;
};
inline int stream_blocks_x(int A) { return (A + 31) / 32; }  // upper bound (the scalar path)
inline LossWs loss_ws_layout(int nb, int B, int A, int M, int k) {
    LossWs w;
    long long cap = (long long)(M > 0 ? M : 1) * (k > 0 ? k : 1);
    w.cap = (int)(cap < A ? cap : A);
    size_t o = 0;
    w.claim = o;      o += a256(sizeof(unsigned long long) * (size_t)B * A);
    w.boxes = o;      o += a256(sizeof(float) * 4 * (size_t)B * A);
    w.lse = o;        o += a256(sizeof(float) * 4 * (size_t)B * A);
    // M > kFinishApMaxM (per-image finishing kernel): list of claimed anchors + two scratch lists, [B, cap] each.
    // M <= kFinishApMaxM (anchor-parallel finishing kernel): one kRecF4 x 16-byte record per CLAIM, [B, rcap]
    const bool ap = M <= kFinishApMaxM;
    w.rcap = ap ? (M > 0 ? M : 1) * (k > 0 ? k : 1) : 0;
    const size_t lcap = ap ? 0 : (size_t)w.cap;
    w.list_a = o;     o += a256(sizeof(int) * (size_t)B * lcap);
    w.list_gi = o;    o += a256(sizeof(int) * (size_t)B * lcap);
    w.list_al = o;    o += a256(sizeof(float) * (size_t)B * lcap);
    w.rec = o;        o += a256(16 * (size_t)kRecF4 * B * w.rcap);
    w.list_count = o; o += a256(sizeof(int) * (size_t)B);
    w.img_cnt = o;    o += a256(sizeof(int) * (size_t)B);
    w.topk_done = o;  o += a256(sizeof(int) * (size_t)B);
    w.pos = o;        o += a256(sizeof(int) * 2 * (size_t)B * (M > 0 ? M : 1));  // per GT: max alignment, max overlap
    w.per_branch = o;
    w.off_counter = (size_t)nb * w.per_branch;
    w.off_pfg = w.off_counter + 256;
    w.off_pbce = w.off_pfg + a256(sizeof(double) * 5 * (size_t)nb * B);
    w.n_bce = stream_blocks_x(A) * B;
    w.off_ord_cnt = w.off_pbce + a256(sizeof(double) * (size_t)nb * w.n_bce);  // [B][4]: valid GTs per size class
    w.off_ord_list = w.off_ord_cnt + a256(sizeof(int) * 4 * (size_t)B);  // [B][M] x 32 B: GT records, big first
    w.total = w.off_ord_list + a256(32 * (size_t)B * (M > 0 ? M : 1));
    return w;
}


}  // namespace y3d
