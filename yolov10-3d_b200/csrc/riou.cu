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

// Individually rounded helpers: the reference computes every product and sum in float32, one operation at a time.
__device__ __forceinline__ float det2(float a, float b, float c, float d) { return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d)); }
__device__ __forceinline__ float dot2(float2 u, float2 v) { return __fadd_rn(__fmul_rn(u.x, v.x), __fmul_rn(u.y, v.y)); }
__device__ __forceinline__ float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// the four corners of a rotated rectangle (cx, cy, dx, dy, angle), in the reference's order (rbbox_to_corners :149-172)
__device__ __forceinline__ void rect_corners(const float *rb, float2 (&q)[4]) {
    const float ca = cosf(rb[4]), sa = sinf(rb[4]);
    const float hx = rb[2] / 2, hy = rb[3] / 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float lx = i < 2 ? -hx : hx, ly = (i == 1 || i == 2) ? hy : -hy;  // (-,-) (-,+) (+,+) (+,-)
        q[i].x = __fadd_rn(__fadd_rn(__fmul_rn(ca, lx), __fmul_rn(sa, ly)), rb[0]);
        q[i].y = __fadd_rn(__fadd_rn(__fmul_rn(-sa, lx), __fmul_rn(ca, ly)), rb[1]);
    }
}
// is p inside the rectangle q?  Projections of (p - q0) on the two edges leaving q0 lie within the edges' squared
// lengths, with the reference's tolerance (point_in_quadrilateral :105-122)
__device__ __forceinline__ bool inside_rect(float2 p, const float2 (&q)[4]) {
    const float2 e1 = q[1] - q[0], e3 = q[3] - q[0], w = p - q[0];
    const float l1 = dot2(e1, e1), p1 = dot2(e1, w), l3 = dot2(e3, e3), p3 = dot2(e3, w);
    const float tol = -1e-6f;
    return l1 - p1 >= tol && p1 >= tol && l3 - p3 >= tol && p3 >= tol;
}
// orientation predicate of the reference's segment test: (r - p) x (q - p) compared as two rounded products
__device__ __forceinline__ bool turns(float2 p, float2 q, float2 r) {
    return __fmul_rn(r.y - p.y, q.x - p.x) > __fmul_rn(q.y - p.y, r.x - p.x);
}
// edge i of rectangle u against edge j of rectangle v: the crossing point, if the two segments straddle each other
// (line_segment_intersection :60-102; Cramer's rule on the two line equations)
__device__ __forceinline__ bool edges_cross(const float2 (&u)[4], const float2 (&v)[4], int i, int j, float2 &x) {
    const float2 a = u[i], b = u[(i + 1) & 3], c = v[j], d = v[(j + 1) & 3];
    if (turns(a, c, d) == turns(b, c, d)) return false;
    if (turns(a, b, c) == turns(a, b, d)) return false;
    const float2 ab = b - a, cd = d - c;
    const float ka = det2(a.x, b.y, b.x, a.y), kc = det2(c.x, d.y, d.x, c.y);
    const float den = det2(ab.y, cd.x, ab.x, cd.y);
    x.x = __fdiv_rn(det2(ka, cd.x, ab.x, kc), den);
    x.y = __fdiv_rn(det2(ka, cd.y, ab.y, kc), den);
    return true;
}
// area of the overlap polygon of two rotated rectangles (inter :231-245): its vertices are the corners of one rectangle
// inside the other plus the edge crossings (quadrilateral_intersection :125-146), ordered around their centroid by a
// monotone function of the angle (sort_vertex_in_convex_polygon :175-212), measured as a triangle fan (area :215-228)
__device__ float riou_inter(const float *r1, const float *r2) {
    float2 u[4], v[4], pts[24];
    rect_corners(r1, u);
    rect_corners(r2, v);
    int n = 0;
    for (int i = 0; i < 4; ++i) {
        if (inside_rect(u[i], v)) pts[n++] = u[i];
        if (inside_rect(v[i], u)) pts[n++] = v[i];
    }
    float2 x;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (edges_cross(u, v, i, j, x)) pts[n++] = x;
    if (n > 0) {
        float2 ctr = make_float2(0.0f, 0.0f);
        float key[24];
        for (int i = 0; i < n; ++i) { ctr.x = __fadd_rn(ctr.x, pts[i].x); ctr.y = __fadd_rn(ctr.y, pts[i].y); }
        ctr.x = __fdiv_rn(ctr.x, (float)n);
        ctr.y = __fdiv_rn(ctr.y, (float)n);
        for (int i = 0; i < n; ++i) {  // unit direction from the centroid: x in [-1, 1] above it, -2 - x below
            const float2 dir = pts[i] - ctr;
            const float len = __fsqrt_rn(dot2(dir, dir));
            const float ux = __fdiv_rn(dir.x, len), uy = __fdiv_rn(dir.y, len);
            key[i] = uy < 0 ? -2 - ux : ux;
        }
        for (int i = 1; i < n; ++i) {  // stable insertion sort, ascending keys
            if (!(key[i - 1] > key[i])) continue;
            const float k = key[i];
            const float2 p = pts[i];
            int j = i;
            for (; j > 0 && key[j - 1] > k; --j) { key[j] = key[j - 1]; pts[j] = pts[j - 1]; }
            key[j] = k;
            pts[j] = p;
        }
    }
    float area = 0.0f;
    for (int i = 1; i + 1 < n; ++i) {  // fan around pts[0]: triangles (pts[0], pts[i], pts[i + 1])
        const float2 e = pts[0] - pts[i + 1], f = pts[i] - pts[i + 1];
        area = __fadd_rn(area, fabsf(__fdiv_rn(det2(e.x, f.y, e.y, f.x), 2.0f)));
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
