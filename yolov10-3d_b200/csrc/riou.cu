// riou.cu -- rotated-box overlap of the KITTI evaluator (sm_100a).
//   y3d_rotate_iou_eval   rotate_iou_gpu_eval / rotate_iou_kernel_eval
//                         reference ultralytics/data/datasets/kitti_eval.py:60-345 (the reference's only GPU code: a
//                         numba-CUDA kernel JIT-compiled at import, with a host round trip per call)
// A CTA of 256 threads owns a 64 x 64 tile of the overlap matrix (four threads per box row, 16 queries each), the tile's
// boxes staged in shared memory.  The overlap polygon of two
// rotated rectangles = corners of one inside the other + edge intersections, ordered around the centroid, measured
// as a triangle fan -- the reference's sequence of float32 operations, so degenerate pairs (identical boxes give
// 1/3, not 1) come out exactly as they do there.
#include "y3d_common.cuh"

namespace y3d {

constexpr int kTile = 64;

__device__ __forceinline__ void riou_corners(const float *rb, float *c) {  // rbbox_to_corners :149-172
    const float a_cos = cosf(rb[4]), a_sin = sinf(rb[4]);
    const float xd = rb[2], yd = rb[3];
    const float cx[4] = {-xd / 2, -xd / 2, xd / 2, xd / 2};
    const float cy[4] = {-yd / 2, yd / 2, yd / 2, -yd / 2};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        c[2 * i] = __fadd_rn(__fadd_rn(__fmul_rn(a_cos, cx[i]), __fmul_rn(a_sin, cy[i])), rb[0]);
        c[2 * i + 1] = __fadd_rn(__fadd_rn(__fmul_rn(-a_sin, cx[i]), __fmul_rn(a_cos, cy[i])), rb[1]);
    }
}
__device__ __forceinline__ bool riou_point_in_quad(float px, float py, const float *c) {  // :105-122
    const float ab0 = c[2] - c[0], ab1 = c[3] - c[1], ad0 = c[6] - c[0], ad1 = c[7] - c[1];
    const float ap0 = px - c[0], ap1 = py - c[1];
    const float abab = __fadd_rn(__fmul_rn(ab0, ab0), __fmul_rn(ab1, ab1)), abap = __fadd_rn(__fmul_rn(ab0, ap0), __fmul_rn(ab1, ap1));
    const float adad = __fadd_rn(__fmul_rn(ad0, ad0), __fmul_rn(ad1, ad1)), adap = __fadd_rn(__fmul_rn(ad0, ap0), __fmul_rn(ad1, ap1));
    const float eps = -1e-6f;
    return abab - abap >= eps && abap >= eps && adad - adap >= eps && adap >= eps;
}
__device__ __forceinline__ bool riou_seg_intersect(const float *p1, const float *p2, int i, int j, float *t) {  // :60-102
    const int i1 = (i + 1) & 3, j1 = (j + 1) & 3;
    const float A0 = p1[2 * i], A1 = p1[2 * i + 1], B0 = p1[2 * i1], B1 = p1[2 * i1 + 1];
    const float C0 = p2[2 * j], C1 = p2[2 * j + 1], D0 = p2[2 * j1], D1 = p2[2 * j1 + 1];
    const float BA0 = B0 - A0, BA1 = B1 - A1, DA0 = D0 - A0, CA0 = C0 - A0, DA1 = D1 - A1, CA1 = C1 - A1;
    const bool acd = __fmul_rn(DA1, CA0) > __fmul_rn(CA1, DA0);
    const bool bcd = __fmul_rn(D1 - B1, C0 - B0) > __fmul_rn(C1 - B1, D0 - B0);
    if (acd == bcd) return false;
    const bool abc = __fmul_rn(CA1, BA0) > __fmul_rn(BA1, CA0), abd = __fmul_rn(DA1, BA0) > __fmul_rn(BA1, DA0);
    if (abc == abd) return false;
    const float DC0 = D0 - C0, DC1 = D1 - C1;
    const float ABBA = __fsub_rn(__fmul_rn(A0, B1), __fmul_rn(B0, A1)), CDDC = __fsub_rn(__fmul_rn(C0, D1), __fmul_rn(D0, C1));
    const float DH = __fsub_rn(__fmul_rn(BA1, DC0), __fmul_rn(BA0, DC1));
    t[0] = __fdiv_rn(__fsub_rn(__fmul_rn(ABBA, DC0), __fmul_rn(BA0, CDDC)), DH);
    t[1] = __fdiv_rn(__fsub_rn(__fmul_rn(ABBA, DC1), __fmul_rn(BA1, CDDC)), DH);
    return true;
}
__device__ float riou_inter(const float *r1, const float *r2) {  // inter :231-245
    float c1[8], c2[8], ip[48], t[2];
    riou_corners(r1, c1);
    riou_corners(r2, c2);
    int n = 0;
    for (int i = 0; i < 4; ++i) {  // quadrilateral_intersection :125-146
        if (riou_point_in_quad(c1[2 * i], c1[2 * i + 1], c2)) { ip[2 * n] = c1[2 * i]; ip[2 * n + 1] = c1[2 * i + 1]; ++n; }
        if (riou_point_in_quad(c2[2 * i], c2[2 * i + 1], c1)) { ip[2 * n] = c2[2 * i]; ip[2 * n + 1] = c2[2 * i + 1]; ++n; }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (riou_seg_intersect(c1, c2, i, j, t)) { ip[2 * n] = t[0]; ip[2 * n + 1] = t[1]; ++n; }
    if (n > 0) {  // sort_vertex_in_convex_polygon :175-212
        float cx = 0.0f, cy = 0.0f, vs[24];
        for (int i = 0; i < n; ++i) { cx = __fadd_rn(cx, ip[2 * i]); cy = __fadd_rn(cy, ip[2 * i + 1]); }
        cx = __fdiv_rn(cx, (float)n);
        cy = __fdiv_rn(cy, (float)n);
        for (int i = 0; i < n; ++i) {
            float v0 = ip[2 * i] - cx, v1 = ip[2 * i + 1] - cy;
            const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(v0, v0), __fmul_rn(v1, v1)));
            v0 = __fdiv_rn(v0, d);
            v1 = __fdiv_rn(v1, d);
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
    float area = 0.0f;  // area :221-228, trangle_area :215-218
    for (int i = 0; i < n - 2; ++i) {
        const float *a = ip, *b = ip + 2 * i + 2, *c = ip + 2 * i + 4;
        const float tr = __fdiv_rn(__fsub_rn(__fmul_rn(a[0] - c[0], b[1] - c[1]), __fmul_rn(a[1] - c[1], b[0] - c[0])), 2.0f);
        area = __fadd_rn(area, fabsf(tr));
    }
    return area;
}

// grid (ceil(K/64), ceil(N/64)), block 256: a CTA owns a 64 x 64 tile of the overlap matrix (box rows x queries), both
// box sets of the tile staged in shared memory; four threads share a box row, thread (tid & 3) takes the queries
// (tid & 3) + 4 j, j = 0..15, so the four stores of a row group are adjacent.  Work per pair: two 4 x 4 edge-intersection
// tests, 8 point-in-quadrilateral tests, a <= 24-point convex polygon sort and area (about 1.5 k fp32 operations); the
// kernel is ALU-bound, its memory traffic (20 (N + K) + 4 N K bytes) is negligible.
__global__ void __launch_bounds__(256) rotate_iou_kernel(const float *__restrict__ boxes, int N,
                                                         const float *__restrict__ query, int K, int criterion,
                                                         float *__restrict__ iou) {
    __shared__ float sb[kTile * 5], sq[kTile * 5];
    const int n0 = blockIdx.y * kTile, k0 = blockIdx.x * kTile;
    const int tid = threadIdx.x;
    for (int i = tid; i < kTile * 5; i += blockDim.x) {
        const int r = i / 5;
        sb[i] = (n0 + r < N) ? boxes[(long long)n0 * 5 + i] : 0.f;
        sq[i] = (k0 + r < K) ? query[(long long)k0 * 5 + i] : 0.f;
    }
    __syncthreads();
    const int r = tid >> 2;
    if (n0 + r >= N) return;
    const float *r2 = sb + 5 * r;
    const float a2 = __fmul_rn(r2[2], r2[3]);
    for (int j = 0; j < 16; ++j) {
        const int cidx = (tid & 3) + 4 * j;
        if (k0 + cidx >= K) continue;
        const float *r1 = sq + 5 * cidx;
        const float a1 = __fmul_rn(r1[2], r1[3]);
        const float ai = riou_inter(r1, r2);  // devRotateIoUEval(query, box) :248-260, call site :299-301
        float v;
        if (criterion == -1) v = __fdiv_rn(ai, __fsub_rn(__fadd_rn(a1, a2), ai));
        else if (criterion == 0) v = __fdiv_rn(ai, a1);
        else if (criterion == 1) v = __fdiv_rn(ai, a2);
        else v = ai;
        iou[(long long)(n0 + r) * K + k0 + cidx] = v;
    }
}

}  // namespace y3d

using namespace y3d;

extern "C" int y3d_rotate_iou_eval(const float *boxes, int N, const float *query_boxes, int K, int criterion, float *iou,
                                   void *stream) {
    if (N < 0 || K < 0) return Y3D_EINVAL;
    if (N == 0 || K == 0) return Y3D_OK;
    if (!boxes || !query_boxes || !iou) return Y3D_EINVAL;
    dim3 grid((K + kTile - 1) / kTile, (N + kTile - 1) / kTile);
    rotate_iou_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(boxes, N, query_boxes, K, criterion, iou);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}
