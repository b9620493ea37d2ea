// decode.cu -- head decode kernels (sm_100a).
//   y3d_decode2d       Detect.inference        reference ultralytics/nn/modules/head.py:53-79
//                      DFL integral            reference ultralytics/nn/modules/block.py:59-62
//                      dist2bbox               reference ultralytics/utils/tal.py:315-325
//   y3d_decode3d       v10Detect3d.decode      reference ultralytics/nn/modules/head.py:755-764
//   y3d_decode_preds3d KITTIDataset.decode_preds reference ultralytics/data/datasets/kitti.py:519-576
//
// HBM-bound streaming kernels: each head element is read once with 128-bit no-allocate loads along the anchor
// axis (the contiguous one), each output element is written once with 128-bit streaming stores; anchors and
// strides are recomputed from the cell index.  Algorithmic bytes: 4*(4R+nc)*A read + 4*(4+nc)*A written per image.
#include "y3d_common.cuh"

namespace y3d {

template <int VEC>
struct VecF {
    float v[VEC];
};

template <int VEC>
__device__ __forceinline__ VecF<VEC> load_vec(const float *p) {
    VecF<VEC> r;
    if constexpr (VEC == 4) {
        float4 t = ldg_stream4(p);
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
        r.v[0] = ldg_stream1(p);
    }
    return r;
}
template <int VEC>
__device__ __forceinline__ void store_vec(float *p, const VecF<VEC> &r) {
    if constexpr (VEC == 4) {
        stg_stream4(p, make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
    } else {
        *p = r.v[0];
    }
}

struct QuadMap {  // quads (groups of VEC consecutive cells of one level) per level
    int qstart[Y3D_MAX_LEVELS + 1];
};

template <int VEC>
__device__ __forceinline__ bool locate(const LevelTable &t, const QuadMap &qm, int q, int &l, int &cell) {
    if (q >= qm.qstart[t.nl]) return false;
    l = 0;
#pragma unroll
    for (int i = 1; i < Y3D_MAX_LEVELS; ++i) l += (i < t.nl && q >= qm.qstart[i]) ? 1 : 0;
    cell = (q - qm.qstart[l]) * VEC;
    return true;
}

// grid (ceil(Q/32), B), block 128 = 32 units of VEC anchors x 4 channel parts (warp = part).  The 64 + nc channel rows
// are split evenly: parts 0 / 1 own the x / y axis of the box (DFL sides l,r / t,b: 32 rows) plus a few class rows,
// parts 2 / 3 the remaining class rows.  One warp instruction touches 32*VEC consecutive anchors of a channel row;
// class rows are loaded 10 at a time before any arithmetic.
template <int VEC>
__global__ void __launch_bounds__(128, 4) decode2d_kernel(LevelTable t, QuadMap qm, int nc, int xywh, int A,
                                                          float *__restrict__ y) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
    const int q = blockIdx.x * 32 + lane;
    int l, cell;
    if (!locate<VEC>(t, qm, q, l, cell)) return;
    const float *base = t.ptr[l] + (long long)b * t.sB[l] + cell;
    const long long cs = t.sC[l];
    const int a0 = t.start[l] + cell;
    float *yb = y + (long long)b * (4 + nc) * A + a0;
    // class rows of this part
    const int rpp = (64 + nc + 3) >> 2;                   // rows per part
    const int n01 = max(0, min(nc / 2, rpp - 32));        // class rows of parts 0 and 1 (each)
    const int rest = nc - 2 * n01, h2 = (rest + 1) >> 1;  // parts 2 and 3 share the rest
    int c0, c1;
    if (part < 2) { c0 = part * n01; c1 = c0 + n01; }
    else if (part == 2) { c0 = 2 * n01; c1 = c0 + h2; }
    else { c0 = 2 * n01 + h2; c1 = nc; }
    if (part < 2) {  // box axis `part`
        VecF<VEC> d_lo, d_hi;
#pragma unroll
        for (int hs = 0; hs < 2; ++hs) {  // l then r (t then b): 16 rows in flight at a time
            VecF<VEC> x[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = load_vec<VEC>(base + (long long)((part + 2 * hs) * 16 + j) * cs);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                float a[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) a[j] = x[j].v[e];
                (hs ? d_hi : d_lo).v[e] = im::dfl16(a);
            }
        }
        const float st = t.stride[l];
        const int w = t.w[l];
        int cx = cell % w, cy = cell / w;
        VecF<VEC> o0, o1;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const float anc = (part == 0 ? (float)cx : (float)cy) + 0.5f;
            im::box_axis(anc, d_lo.v[e], d_hi.v[e], st, xywh, o0.v[e], o1.v[e]);
            if (++cx >= w) { cx = 0; ++cy; }
        }
        store_vec<VEC>(yb + (long long)part * A, o0);
        store_vec<VEC>(yb + (long long)(part + 2) * A, o1);
    }
    const float *p = base + (long long)(64 + c0) * cs;
    float *o = yb + (long long)(4 + c0) * A;
    int c = c0;
    constexpr int CB = 10;
    for (; c + CB <= c1; c += CB, p += CB * cs, o += (long long)CB * A) {
        VecF<VEC> v[CB];
#pragma unroll
        for (int u = 0; u < CB; ++u) v[u] = load_vec<VEC>(p + u * cs);
#pragma unroll
        for (int u = 0; u < CB; ++u) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) v[u].v[e] = im::sigmoid(v[u].v[e]);  // head.py:78
            store_vec<VEC>(o + (long long)u * A, v[u]);
        }
    }
    for (; c < c1; ++c, p += cs, o += A) {
        VecF<VEC> v = load_vec<VEC>(p);
#pragma unroll
        for (int e = 0; e < VEC; ++e) v.v[e] = im::sigmoid(v.v[e]);
        store_vec<VEC>(o, v);
    }
}

// v10Detect3d.decode: flat over (b, channel, quad)
template <int VEC>
__global__ void __launch_bounds__(256) decode3d_kernel(LevelTable t, QuadMap qm, int nc, int A, long long total,
                                                       float *__restrict__ y) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int Q = qm.qstart[t.nl];
    const int C = nc + 35;
    int q = (int)(idx % Q);
    long long r = idx / Q;
    int c = (int)(r % C);
    int b = (int)(r / C);
    int l, cell;
    locate<VEC>(t, qm, q, l, cell);
    const float *base = t.ptr[l] + (long long)b * t.sB[l] + cell;
    const long long cs = t.sC[l];
    float *o = y + ((long long)b * C + c) * A + t.start[l] + cell;
    const int k = c - nc;  // 0,1 o2d | 2,3 s2d | 4,5 o3d
    VecF<VEC> out;
    if (k < 0 || k > 5) {
        out = load_vec<VEC>(base + (long long)c * cs);
    } else {
        const float st = t.stride[l];
        const int w = t.w[l];
        const int axis = k & 1;
        if (k < 4) {  // bbox: (o2d + anchor)*stride -/+ (s2d*stride)/2   head.py:756-759
            VecF<VEC> o2 = load_vec<VEC>(base + (long long)(nc + axis) * cs);
            VecF<VEC> s2 = load_vec<VEC>(base + (long long)(nc + 2 + axis) * cs);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                int cc = cell + e;
                float anc = (axis == 0 ? (float)(cc % w) : (float)(cc / w)) + 0.5f;
                float ctr = __fmul_rn(__fadd_rn(o2.v[e], anc), st);
                float half = __fmul_rn(s2.v[e], st) / 2.0f;
                out.v[e] = k < 2 ? __fsub_rn(ctr, half) : __fadd_rn(ctr, half);
            }
        } else {  // center3d = (o3d + anchor) * stride   head.py:762
            VecF<VEC> o3 = load_vec<VEC>(base + (long long)c * cs);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                int cc = cell + e;
                float anc = (axis == 0 ? (float)(cc % w) : (float)(cc / w)) + 0.5f;
                out.v[e] = __fmul_rn(__fadd_rn(o3.v[e], anc), st);
            }
        }
    }
    store_vec<VEC>(o, out);
}

// KITTIDataset.decode_preds: one thread per detection, float64 like the reference's numpy math.
__global__ void decode_preds3d_kernel(const float *__restrict__ dets, int B, int D, int nc,
                                      const double *__restrict__ calib, const double *__restrict__ inv_affine,
                                      const double *__restrict__ ratio, const double *__restrict__ cms, double thr,
                                      double *__restrict__ rows, uint8_t *__restrict__ valid) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * D) return;
    const int b = i / D;
    const float *p = dets + (long long)i * 37;
    double *r = rows + (long long)i * 14;
    const double *cal = calib + 6 * b, *T = inv_affine + 6 * b;
    int hb = 0;
    for (int q = 1; q < 12; ++q)
        if (p[9 + q] > p[9 + hb]) hb = q;  // argmax, first max (kitti.py:526)
    float alpha_f = __fadd_rn(__fmul_rn((float)hb, 0.5235987755982988f), p[9 + 12 + hb]);  // decode_helper.py:12-18
    if (alpha_f > 3.14159265358979323846f) alpha_f = __fsub_rn(alpha_f, 6.283185307179586f);
    const double alpha = (double)alpha_f;
    int cls = (int)p[36];
    cls = cls < 0 ? 0 : (cls >= nc ? nc - 1 : cls);
    const double rx = ratio[2 * b], ry_ = ratio[2 * b + 1];
    const double bx0 = p[0] / rx, bx1 = p[1] / ry_, bx2 = p[2] / rx, bx3 = p[3] / ry_;  // kitti.py:538
    const double x = (bx0 + bx2) / 2;
    float dim[3];
    for (int q = 0; q < 3; ++q) dim[q] = (float)((double)p[6 + q] + cms[3 * cls + q]);  // kitti.py:541-542
    const double depth = p[33];
    const float sigma_f = expf(-p[34]);  // kitti.py:545
    const double cx = T[0] * (double)p[4] + T[1] * (double)p[5] + T[2];  // affine_transform kitti_utils.py:467
    const double cy = T[3] * (double)p[4] + T[4] * (double)p[5] + T[5];
    const double lx = ((cx - cal[0]) * depth) / cal[2] + cal[4];  // img_to_rect kitti_utils.py:241
    double ly = ((cy - cal[1]) * depth) / cal[3] + cal[5];
    ly += (double)dim[0] / 2;  // kitti.py:564
    const double kPi = 3.14159265358979323846;
    double ry = alpha + atan2(x - cal[0], cal[2]);  // alpha2ry kitti_utils.py:311
    if (ry > kPi) ry -= 2 * kPi;
    if (ry < -kPi) ry += 2 * kPi;
    const float sig = 1.0f / (1.0f + expf(-p[35]));
    const double score = (double)sig * (double)sigma_f;  // kitti.py:569
    r[0] = cls; r[1] = alpha;
    r[2] = bx0; r[3] = bx1; r[4] = bx2; r[5] = bx3;
    r[6] = dim[0]; r[7] = dim[1]; r[8] = dim[2];
    r[9] = lx; r[10] = ly; r[11] = depth; r[12] = ry; r[13] = score;
    valid[i] = (uint8_t)!(score < thr);
}

// can the head be read with 128-bit loads along the anchor axis?
static bool vec4_ok(const LevelTable &t) {
    for (int l = 0; l < t.nl; ++l) {
        if ((t.h[l] * t.w[l]) % 4) return false;
        if (((uintptr_t)t.ptr[l]) % 16) return false;
        if (t.sB[l] % 4 || t.sC[l] % 4) return false;
    }
    return true;
}

template <int VEC>
static QuadMap make_quads(const LevelTable &t) {
    QuadMap qm;
    int q = 0;
    for (int l = 0; l <= Y3D_MAX_LEVELS; ++l) {
        qm.qstart[l] = q;
        if (l < t.nl) q += t.h[l] * t.w[l] / VEC;
    }
    for (int l = t.nl; l <= Y3D_MAX_LEVELS; ++l) qm.qstart[l] = q;
    return qm;
}

}  // namespace y3d

using namespace y3d;

extern "C" int y3d_decode2d(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC,
                            const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, int reg_max, int xywh,
                            float *y, void *stream) {
    if (!lvl_ptr || !lvl_sB || !lvl_sC || !y || B < 0 || nc < 1) return Y3D_EINVAL;
    if (reg_max != 16) return Y3D_EUNSUPPORTED;
    LevelTable t;
    int A = make_level_table(t, lvl_ptr, lvl_sB, lvl_sC, lvl_hw, lvl_stride, nl);
    if (A < 0) return A;
    for (int l = 0; l < nl; ++l)
        if (!lvl_ptr[l]) return Y3D_EINVAL;
    if (B == 0) return Y3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    bool v4 = vec4_ok(t) && (((uintptr_t)y) % 16 == 0) && A % 4 == 0;
    if (v4) {
        QuadMap qm = make_quads<4>(t);
        dim3 grid((qm.qstart[nl] + 31) / 32, B);
        decode2d_kernel<4><<<grid, 128, 0, s>>>(t, qm, nc, xywh, A, y);
    } else {
        QuadMap qm = make_quads<1>(t);
        dim3 grid((qm.qstart[nl] + 31) / 32, B);
        decode2d_kernel<1><<<grid, 128, 0, s>>>(t, qm, nc, xywh, A, y);
    }
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

extern "C" int y3d_decode3d(const float *const *lvl_ptr, const int64_t *lvl_sB, const int64_t *lvl_sC,
                            const int *lvl_hw, const float *lvl_stride, int nl, int B, int nc, float *y,
                            void *stream) {
    if (!lvl_ptr || !lvl_sB || !lvl_sC || !y || B < 0 || nc < 1) return Y3D_EINVAL;
    LevelTable t;
    int A = make_level_table(t, lvl_ptr, lvl_sB, lvl_sC, lvl_hw, lvl_stride, nl);
    if (A < 0) return A;
    for (int l = 0; l < nl; ++l)
        if (!lvl_ptr[l]) return Y3D_EINVAL;
    if (B == 0) return Y3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int C = nc + 35;
    bool v4 = vec4_ok(t) && (((uintptr_t)y) % 16 == 0);
    if (v4) {
        QuadMap qm = make_quads<4>(t);
        long long total = (long long)B * C * qm.qstart[nl];
        decode3d_kernel<4><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(t, qm, nc, A, total, y);
    } else {
        QuadMap qm = make_quads<1>(t);
        long long total = (long long)B * C * qm.qstart[nl];
        decode3d_kernel<1><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(t, qm, nc, A, total, y);
    }
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}

extern "C" int y3d_decode_preds3d(const float *dets, int B, int D, int nc, const double *calib,
                                  const double *inv_affine, const double *ratio, const double *cls_mean_size,
                                  double threshold, double *rows, uint8_t *valid, void *stream) {
    if (!dets || !calib || !inv_affine || !ratio || !cls_mean_size || !rows || !valid || B < 0 || D < 0 || nc < 1)
        return Y3D_EINVAL;
    if (B * D == 0) return Y3D_OK;
    decode_preds3d_kernel<<<(B * D + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dets, B, D, nc, calib, inv_affine,
                                                                               ratio, cls_mean_size, threshold,
                                                                               rows, valid);
    Y3D_CHECK_LAUNCH();
    return Y3D_OK;
}
