// y3d_common.cuh -- shared device helpers for liby3d_b200 (sm_100a).
//
//  * LevelTable: head geometry passed by value to kernels; anchors are recomputed from the cell index
//    (make_anchors, reference ultralytics/utils/tal.py:300-312) so no anchor tensor is ever read.
//  * dm:: deterministic fp32 math.  Everything that decides an index on the path (in-GT test, CIoU, alignment
//    metric, 3D keypoint similarity) is issued as individually rounded IEEE operations through the
//    __f*_rn intrinsics, which nvcc never contracts into FMAs, so the result is independent of -fmad and
//    bit-identical to a scalar CPU evaluation of the same sequence.  atan / sin / cos / exp follow the classic
//    Cephes single-precision forms (<= 2 ulp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/y3d.h"

#define Y3D_CHECK_LAUNCH()                               \
    do {                                                 \
        cudaError_t e__ = cudaPeekAtLastError();         \
        if (e__ != cudaSuccess) return (int)e__;         \
    } while (0)

namespace y3d {

// number of SMs of the current device (148 on B200), queried once per process: one process per GPU (topk_fused.cu)
int device_sm_count();

struct LevelTable {
    const float *ptr[Y3D_MAX_LEVELS];
    long long sB[Y3D_MAX_LEVELS];
    long long sC[Y3D_MAX_LEVELS];
    int h[Y3D_MAX_LEVELS];
    int w[Y3D_MAX_LEVELS];
    int start[Y3D_MAX_LEVELS + 1];  // first anchor index of each level; start[nl] = A
    float stride[Y3D_MAX_LEVELS];
    int nl;
};

// host: fills a LevelTable; returns A or a negative error.  ptrs may be NULL (geometry only).
inline int make_level_table(LevelTable &t, const float *const *lvl_ptr, const int64_t *sB, const int64_t *sC,
                            const int *lvl_hw, const float *lvl_stride, int nl) {
    if (nl < 1 || nl > Y3D_MAX_LEVELS || !lvl_hw || !lvl_stride) return Y3D_EINVAL;
    int a = 0;
    for (int l = 0; l < Y3D_MAX_LEVELS; ++l) {
        bool on = l < nl;
        t.ptr[l] = (on && lvl_ptr) ? lvl_ptr[l] : nullptr;
        t.sB[l] = (on && sB) ? sB[l] : 0;
        t.sC[l] = (on && sC) ? sC[l] : 0;
        t.h[l] = on ? lvl_hw[2 * l] : 0;
        t.w[l] = on ? lvl_hw[2 * l + 1] : 0;
        t.stride[l] = on ? lvl_stride[l] : 0.f;
        t.start[l] = a;
        if (on) {
            if (t.h[l] <= 0 || t.w[l] <= 0) return Y3D_EINVAL;
            a += t.h[l] * t.w[l];
        }
    }
    t.start[Y3D_MAX_LEVELS] = a;
    for (int l = nl; l < Y3D_MAX_LEVELS; ++l) t.start[l] = a;
    t.nl = nl;
    return a;
}

__device__ __forceinline__ int level_of(const LevelTable &t, int a) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < Y3D_MAX_LEVELS; ++i) l += (i < t.nl && a >= t.start[i]) ? 1 : 0;
    return l;
}

// --------------------------------------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 128-bit streaming loads: the head tensor is touched exactly once per pass, so it bypasses L1 and is marked
// evict-first in L2 -- the small intermediates (boxes, claims, lists) written next to it are what later kernels
// gather from and should be the lines that survive in the 126 MB L2.
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ldg_stream4(const float *p, unsigned long long pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float *p, unsigned long long pol) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream4(float *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// fire-and-forget pull of one line into L2 (latency-bound gathers of a later phase / kernel then hit L2)
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// --------------------------------------------------------------------------------------------- im:: math
// Value-only arithmetic of the inference path (sigmoid of the class logits, DFL softmax expectation): fast SFU
// intrinsics, relative error ~1e-6, far inside the 1e-5 bar.  decode.cu and topk.cu share these definitions so that the
// fused decode + top-k path and the unfused one produce identical bits.
namespace im {
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2a(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid(float v) { return __fdividef(1.0f, 1.0f + ex2a(-v * kLog2e)); }
// softmax over 16 bins followed by the expectation sum_j j * p_j (DFL.forward block.py:59-62)
__device__ __forceinline__ float dfl16(const float (&x)[16]) {
    float m = x[0];
#pragma unroll
    for (int j = 1; j < 16; ++j) m = fmaxf(m, x[j]);
    const float mo = -m * kLog2e;
    float s = 0.f, acc = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float e = ex2a(__fmaf_rn(x[j], kLog2e, mo));
        s = __fadd_rn(s, e);
        acc = __fmaf_rn((float)j, e, acc);
    }
    return __fdividef(acc, s);
}
// dist2bbox (tal.py:315-325) * stride (head.py:76) for one axis: lo = anchor - d_lo, hi = anchor + d_hi
__device__ __forceinline__ void box_axis(float anc, float d_lo, float d_hi, float st, int xywh, float &o0, float &o1) {
    const float lo = __fsub_rn(anc, d_lo), hi = __fadd_rn(anc, d_hi);
    if (xywh) {
        o0 = __fmul_rn(__fmul_rn(__fadd_rn(lo, hi), 0.5f), st);
        o1 = __fmul_rn(__fsub_rn(hi, lo), st);
    } else {
        o0 = __fmul_rn(lo, st);
        o1 = __fmul_rn(hi, st);
    }
}
}  // namespace im

// --------------------------------------------------------------------------------------------- dm:: math
namespace dm {
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sqrt_(float a) { return __fsqrt_rn(a); }

__device__ __forceinline__ float atan_(float x) {
    float ax = fabsf(x);
    float y0, z;
    if (ax > 2.414213562373095f) {
        y0 = 1.5707963267948966f;
        z = -div(1.0f, ax);
    } else if (ax > 0.4142135623730950f) {
        y0 = 0.7853981633974483f;
        z = div(sub(ax, 1.0f), add(ax, 1.0f));
    } else {
        y0 = 0.0f;
        z = ax;
    }
    float zz = mul(z, z);
    float p = 8.05374449538e-2f;
    p = mul(p, zz);
    p = sub(p, 1.38776856032e-1f);
    p = mul(p, zz);
    p = add(p, 1.99777106478e-1f);
    p = mul(p, zz);
    p = sub(p, 3.33329491539e-1f);
    p = mul(p, zz);
    p = mul(p, z);
    float r = add(y0, add(p, z));
    return copysignf(r, x);
}

__device__ __forceinline__ float atan2_(float y, float x) {
    const float pi = 3.14159265358979323846f;
    if (x == 0.0f) {
        if (y > 0.0f) return 1.5707963267948966f;
        if (y < 0.0f) return -1.5707963267948966f;
        return 0.0f;
    }
    float t = atan_(div(y, x));
    if (x > 0.0f) return t;
    if (y >= 0.0f) return add(t, pi);
    return sub(t, pi);
}

__device__ __forceinline__ void sincos_(float x, float *s_out, float *c_out) {
    float ax = fabsf(x);
    int sign_s = x < 0.0f ? -1 : 1;
    int sign_c = 1;
    int j = (int)mul(ax, 1.27323954473516f);
    float y = (float)j;
    if (j & 1) {
        j += 1;
        y = add(y, 1.0f);
    }
    j &= 7;
    if (j > 3) {
        sign_s = -sign_s;
        sign_c = -sign_c;
        j -= 4;
    }
    if (j > 1) sign_c = -sign_c;
    float r = sub(ax, mul(y, 0.78515625f));
    r = sub(r, mul(y, 2.4187564849853515625e-4f));
    r = sub(r, mul(y, 3.77489497744594108e-8f));
    float z = mul(r, r);
    float ps = -1.9515295891e-4f;
    ps = mul(ps, z);
    ps = add(ps, 8.3321608736e-3f);
    ps = mul(ps, z);
    ps = sub(ps, 1.6666654611e-1f);
    ps = mul(ps, z);
    ps = mul(ps, r);
    ps = add(ps, r);
    float pc = 2.443315711809948e-5f;
    pc = mul(pc, z);
    pc = sub(pc, 1.388731625493765e-3f);
    pc = mul(pc, z);
    pc = add(pc, 4.166664568298827e-2f);
    pc = mul(pc, z);
    pc = mul(pc, z);
    pc = sub(pc, mul(0.5f, z));
    pc = add(pc, 1.0f);
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

__device__ __forceinline__ float exp_(float x) {
    if (x > 88.72283905206835f) return __int_as_float(0x7f800000);
    if (x < -103.278929903431851103f) return 0.0f;
    if (x != x) return x;
    float fl = floorf(add(mul(x, 1.44269504088896341f), 0.5f));
    int n = (int)fl;
    float r = sub(x, mul(fl, 0.693359375f));
    r = sub(r, mul(fl, -2.12194440e-4f));
    float z = mul(r, r);
    float p = 1.9875691500e-4f;
    p = mul(p, r);
    p = add(p, 1.3981999507e-3f);
    p = mul(p, r);
    p = add(p, 8.3334519073e-3f);
    p = mul(p, r);
    p = add(p, 4.1665795894e-2f);
    p = mul(p, r);
    p = add(p, 1.6666665459e-1f);
    p = mul(p, r);
    p = add(p, 5.0000001201e-1f);
    p = mul(p, z);
    p = add(p, r);
    p = add(p, 1.0f);
    int n1 = n / 2, n2 = n - n1;
    float a = __int_as_float((n1 + 127) << 23);
    float b = __int_as_float((n2 + 127) << 23);
    return mul(mul(p, a), b);
}

// sigmoid as a fixed sequence of rounded operations: 1 / (1 + exp_(-x)).  Exactly the oracle's (oracle/y3d_oracle.c
// y3d_sigmoidf), and weakly monotone over all of binary32 (checked exhaustively: oracle/check_sigmoid_monotone.c), so the
// largest score of an anchor is the sigmoid of its largest logit.  Used where a score decides an index (fused decode +
// top-k); value-only outputs use im::sigmoid.
__device__ __forceinline__ float sigmoid_(float x) { return div(1.0f, add(1.0f, exp_(-x))); }

static __device__ __noinline__ float pow_generic(float x, float e) { return powf(x, e); }

// pow with the path's exponents explicit (see oracle/y3d_oracle.c::y3d_powf); other exponents -> powf.
__device__ __forceinline__ float pow_(float x, float e) {
    if (e == 0.5f) return sqrt_(x);
    if (e == 1.0f) return x;
    if (e == 2.0f) return mul(x, x);
    if (e == 3.0f) return mul(mul(x, x), x);
    if (e == 4.0f) {
        float x2 = mul(x, x);
        return mul(x2, x2);
    }
    if (e == 6.0f) {
        float x2 = mul(x, x);
        float x4 = mul(x2, x2);
        return mul(x4, x2);
    }
    if (e == 0.0f) return 1.0f;
    return pow_generic(x, e);
}

// atan(w/h) term of a box, h already has +eps (metrics.py:103-104,126)
__device__ __forceinline__ float box_atan(float w, float h) { return atan_(div(w, h)); }

// bbox_iou(box1, box2, xywh=False, CIoU=True), metrics.py:96-131, op order preserved.
// at1 = atan(w1/h1) of box1, precomputed by the caller when box1 is reused.
__device__ __forceinline__ float ciou(float4 b1, float4 b2, float at1) {
    const float eps = 1e-7f;
    float w1 = sub(b1.z, b1.x), h1 = add(sub(b1.w, b1.y), eps);
    float w2 = sub(b2.z, b2.x), h2 = add(sub(b2.w, b2.y), eps);
    float iw = sub(fminf(b1.z, b2.z), fmaxf(b1.x, b2.x));
    float ih = sub(fminf(b1.w, b2.w), fmaxf(b1.y, b2.y));
    iw = iw < 0.0f ? 0.0f : iw;
    ih = ih < 0.0f ? 0.0f : ih;
    float inter = mul(iw, ih);
    float uni = add(sub(add(mul(w1, h1), mul(w2, h2)), inter), eps);
    float iou = div(inter, uni);
    float cw = sub(fmaxf(b1.z, b2.z), fminf(b1.x, b2.x));
    float ch = sub(fmaxf(b1.w, b2.w), fminf(b1.y, b2.y));
    float c2 = add(add(mul(cw, cw), mul(ch, ch)), eps);
    float dx = sub(sub(add(b2.x, b2.z), b1.x), b1.z);
    float dy = sub(sub(add(b2.y, b2.w), b1.y), b1.w);
    float rho2 = div(add(mul(dx, dx), mul(dy, dy)), 4.0f);
    float da = sub(box_atan(w2, h2), at1);
    float v = mul(0.4052847345693511f, mul(da, da));
    float alpha = div(v, add(sub(v, iou), 1.0000001f));
    return sub(iou, add(div(rho2, c2), mul(v, alpha)));
}
__device__ __forceinline__ float box1_atan(float4 b1) {
    return box_atan(sub(b1.z, b1.x), add(sub(b1.w, b1.y), 1e-7f));
}

// select_candidates_in_gts, tal.py:218-235
__device__ __forceinline__ bool in_gt(float ax, float ay, float4 g) {
    float d = fminf(fminf(sub(ax, g.x), sub(ay, g.y)), fminf(sub(g.z, ax), sub(g.w, ay)));
    return d > 1e-9f;
}
}  // namespace dm

}  // namespace y3d
