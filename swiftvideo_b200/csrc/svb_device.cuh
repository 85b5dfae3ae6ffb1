// svb_device.cuh -- device-side types and the bit-exact per-pixel arithmetic shared by every kernel.
//
// What is computed is fixed by the reference's OpenCL kernels (the parity target named by north_star):
//   /root/reference/Sources/SwiftVideo/kernels.cl.swift:63-108 (img_nv12_nv12 and its YUV siblings),
//   :283-334,:485-531 (BGRA/RGBA sources), :38-46,:174-185,:257-265 (clear),
// with OpenCL 1.2 image semantics (UNORM8 read c/255.0f, write sat_rte(f*255.0f), linear sampler
// section 8.2).  HOW it is computed is ours.  Every multiply and add must round separately (the
// reference builds its CUDA with --fmad=false, compute.cuda.swift:177): all arithmetic below goes
// through __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn, which the compiler never contracts or reassociates.
#pragma once
#include <stdint.h>

#include "svb_desc.h"

// The drop-in kernels receive the reference's own 236-byte ImageUniforms upload; SvbUniforms has the same
// offsets (the device-side struct pads to 240 upstream as well, kernels.cl.swift:49-59).
typedef SvbUniforms ImageUniforms;

namespace svb {

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ float4 ldrow(const float* m, int row) { return __ldg((const float4*)m + row); }

// dot(float4,float4) summed left to right
__device__ __forceinline__ float dot4(float x, float y, float z, float w, const float4 m) {
    return add(add(add(mul(x, m.x), mul(y, m.y)), mul(z, m.z)), mul(w, m.w));
}

// OpenCL clamp = fmin(fmax(x, lo), hi); CUDA fminf/fmaxf drop a NaN operand the same way
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// UNORM8 read: c / 255.0f, correctly rounded.  Division-free and exact for every byte: with K1 = fl(1/255)
// and K2 = fl(1/255 - K1), fma(c, K1, fl(c*K2)) carries ~48 bits of c/255 into one rounding, and c/255 (an
// 8-bit pattern repeating in binary) is never that close to a rounding boundary.  tests/test_gpu_parity.py
// checks all 256 values against __fdiv_rn on the device.
#define SVB_K1 __uint_as_float(0x3b808081u)
#define SVB_K2 __uint_as_float(0xaf7efeffu)
__device__ __forceinline__ float unorm_f(float c) { return __fmaf_rn(c, SVB_K1, __fmul_rn(c, SVB_K2)); }
// Byte loads that hand the compiler an opaque 32-bit value: knowing the 8-bit range it would convert with
// I2F.U8/U16, which runs on the quarter-rate XU pipe (profiles/: the XU pipe saturated at 149 % before this);
// an unknown u32 converts with the full-rate I2FP.F32.U32.
__device__ __forceinline__ unsigned ldg_u8(const uint8_t* p) {
    unsigned v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned ldg_u16(const uint8_t* p) {
    unsigned v;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned opaque(unsigned v) {
    asm volatile("" : "+r"(v));
    return v;
}
__device__ __forceinline__ float unorm(unsigned c) { return unorm_f(__uint2float_rn(c)); }
__device__ __forceinline__ float unorm_div(unsigned c) { return __fdiv_rn((float)c, 255.0f); }

// UNORM8 write: convert_uchar_sat_rte(f * 255.0f); NaN -> 0
__device__ __forceinline__ unsigned rte8(float f) {
    float v = fminf(fmaxf(mul(f, 255.0f), 0.0f), 255.0f);
    return (unsigned)__float2int_rn(v);
}

// The same write kept as an integer-valued float (no int conversion at all): adding 2^23 rounds to the nearest
// integer, ties to even, exactly like rint.
__device__ __forceinline__ float quantf(float f) {
    const float v = fminf(fmaxf(mul(f, 255.0f), 0.0f), 255.0f);
    return sub(add(v, 8388608.0f), 8388608.0f);
}

__device__ __forceinline__ bool in01(float x, float y) { return x >= 0.f && y >= 0.f && x <= 1.f && y <= 1.f; }

// Linear sampler footprint: OpenCL 1.2 section 8.2, normalised coords, clamp-to-edge.
struct Taps {
    int i0, i1, j0, j1;
    float w00, w10, w01, w11;
};
__device__ __forceinline__ Taps make_taps(float s, float t, int w, int h) {
    Taps k;
    float u = mul(s, (float)w), v = mul(t, (float)h);
    float um = sub(u, 0.5f), vm = sub(v, 0.5f);
    float fu = floorf(um), fv = floorf(vm);
    float a = sub(um, fu), b = sub(vm, fv);
    int i = (int)fu, j = (int)fv;
    k.i0 = min(max(i, 0), w - 1);
    k.i1 = min(max(i + 1, 0), w - 1);
    k.j0 = min(max(j, 0), h - 1);
    k.j1 = min(max(j + 1, 0), h - 1);
    float na = sub(1.0f, a), nb = sub(1.0f, b);
    k.w00 = mul(na, nb);
    k.w10 = mul(a, nb);
    k.w01 = mul(na, b);
    k.w11 = mul(a, b);
    return k;
}
// (1-a)(1-b)T00 + a(1-b)T10 + (1-a)b T01 + ab T11, left to right
__device__ __forceinline__ float filt(const Taps& k, float t00, float t10, float t01, float t11) {
    return add(add(add(mul(k.w00, t00), mul(k.w10, t10)), mul(k.w01, t01)), mul(k.w11, t11));
}
__device__ __forceinline__ float sample1(const uint8_t* __restrict__ p, int stride, int ncomp, int c, const Taps& k) {
    const uint8_t* r0 = p + (size_t)k.j0 * stride + c;
    const uint8_t* r1 = p + (size_t)k.j1 * stride + c;
    return filt(k, unorm(ldg_u8(r0 + k.i0 * ncomp)), unorm(ldg_u8(r0 + k.i1 * ncomp)), unorm(ldg_u8(r1 + k.i0 * ncomp)),
                unorm(ldg_u8(r1 + k.i1 * ncomp)));
}
__device__ __forceinline__ float4 sample4(const uint8_t* __restrict__ p, int stride, const Taps& k) {
    const unsigned a = __ldg((const unsigned*)(p + (size_t)k.j0 * stride) + k.i0);
    const unsigned b = __ldg((const unsigned*)(p + (size_t)k.j0 * stride) + k.i1);
    const unsigned c = __ldg((const unsigned*)(p + (size_t)k.j1 * stride) + k.i0);
    const unsigned d = __ldg((const unsigned*)(p + (size_t)k.j1 * stride) + k.i1);
    float4 r;
#define SVB_CH(w, n) unorm(opaque(((w) >> (8 * (n))) & 0xffu))
    r.x = filt(k, SVB_CH(a, 0), SVB_CH(b, 0), SVB_CH(c, 0), SVB_CH(d, 0));
    r.y = filt(k, SVB_CH(a, 1), SVB_CH(b, 1), SVB_CH(c, 1), SVB_CH(d, 1));
    r.z = filt(k, SVB_CH(a, 2), SVB_CH(b, 2), SVB_CH(c, 2), SVB_CH(d, 2));
    r.w = filt(k, SVB_CH(a, 3), SVB_CH(b, 3), SVB_CH(c, 3), SVB_CH(d, 3));
#undef SVB_CH
    return r;
}

// RGB2YUV rows, kernels.cl.swift:96-99 (0.113 as written upstream); v.w is always 1
__device__ __forceinline__ float3 rgb2yuv(float r, float g, float b) {
    float3 o;
    o.x = dot4(r, g, b, 1.0f, make_float4(0.299f, 0.587f, 0.113f, 0.f));
    o.y = dot4(r, g, b, 1.0f, make_float4(-0.169f, -0.331f, 0.5f, 0.5f));
    o.z = dot4(r, g, b, 1.0f, make_float4(0.5f, -0.419f, -0.081f, 0.5f));
    return o;
}

// Source picture as the kernels see it.
struct Src {
    const uint8_t* p[3];
    int stride[3];
    int w, h;    // luma / RGBA plane size
    int cw, ch;  // chroma plane size
    int format;
};

// img_{bgra,rgba}_{nv12,y420p} for a pixel inside the picture's rectangle (kernels.cl.swift:509-528): the fill colour is
// blended first, always; then, where uv is inside [0,1], the sampled pixel, premultiplied by a = alpha * opacity before the
// colour matrix and blended with weight a again -- as written upstream.  `k` is only read when in_uv.
__device__ __forceinline__ void rgba_pixel(const Src& s, float opacity, const float4 fc, bool in_uv, const Taps& k, float cy, float cu, float cv, float& oy,
                                           float& ou, float& ov) {
    const float a = mul(opacity, fc.w), na = sub(1.f, a);
    const float3 f = rgb2yuv(mul(fc.x, a), mul(fc.y, a), mul(fc.z, a));
    float r0 = add(mul(cy, na), mul(f.x, a));
    float r1 = clampf(add(mul(cu, na), mul(f.y, a)), -1.f, 1.f);
    float r2 = clampf(add(mul(cv, na), mul(f.z, a)), -1.f, 1.f);
    if (in_uv) {
        float4 px = sample4(s.p[0], s.stride[0], k);
        if (s.format == SVB_BGRA) {
            float t = px.x;
            px.x = px.z;
            px.z = t;
        }
        const float a2 = mul(px.w, opacity), n2 = sub(1.f, a2);
        const float3 q = rgb2yuv(mul(px.x, a2), mul(px.y, a2), mul(px.z, a2));
        r0 = add(mul(r0, n2), mul(q.x, a2));
        r1 = add(mul(r1, n2), mul(q.y, a2));
        r2 = add(mul(r2, n2), mul(q.z, a2));
    }
    oy = r0;
    ou = r1;
    ov = r2;
}

// One work-item of img_<src>_<dst> on luma pixel (x,y) of a WxH target.
//   cy/cu/cv : current target values as UNORM floats (cu/cv only meaningful when `chroma`)
//   returns false when the pixel is left untouched; otherwise oy (and ou/ov when `chroma`) hold the
//   floats the reference hands to write_imagef.
__device__ __forceinline__ bool eval_pixel(const ImageUniforms* __restrict__ U, const Src& s, int x, int y, float W,
                                           float H, bool chroma, float cy, float cu, float cv, float& oy, float& ou,
                                           float& ov) {
    const float nx = sub(mul(__fdiv_rn((float)x, W), 2.f), 1.f);
    const float ny = sub(mul(__fdiv_rn((float)y, H), 2.f), 1.f);
    const float bx = dot4(nx, ny, 0.f, 1.f, ldrow(U->borderMatrix, 0));
    const float by = dot4(nx, ny, 0.f, 1.f, ldrow(U->borderMatrix, 1));
    if (!in01(bx, by)) return false;
    const float t0 = dot4(nx, ny, 0.f, 1.f, ldrow(U->transform, 0));
    const float t1 = dot4(nx, ny, 0.f, 1.f, ldrow(U->transform, 1));
    const float t2 = dot4(nx, ny, 0.f, 1.f, ldrow(U->transform, 2));
    const float t3 = dot4(nx, ny, 0.f, 1.f, ldrow(U->transform, 3));
    const float uu = dot4(t0, t1, t2, t3, ldrow(U->textureTx, 0));
    const float vv = dot4(t0, t1, t2, t3, ldrow(U->textureTx, 1));
    const float opacity = __ldg(&U->opacity);
    const float4 fc = ldrow(U->fillColor, 0);
    const bool in_tx = in01(t0, t1), in_uv = in01(uu, vv);

    if (SVB_FORMAT_IS_YUV(s.format)) {
        if (in_tx && in_uv) {
            const float na = sub(1.f, opacity);
            Taps k = make_taps(uu, vv, s.w, s.h);
            oy = add(mul(cy, na), mul(sample1(s.p[0], s.stride[0], 1, 0, k), opacity));
            if (chroma) {
                Taps kc = make_taps(uu, vv, s.cw, s.ch);
                float cb, cr;
                if (SVB_FORMAT_IS_SEMIPLANAR(s.format)) {
                    const int swap = s.format == SVB_NV21 ? 1 : 0;  // NV21 stores (Cr, Cb)
                    cb = sample1(s.p[1], s.stride[1], 2, swap, kc);
                    cr = sample1(s.p[1], s.stride[1], 2, 1 - swap, kc);
                } else {
                    cb = sample1(s.p[1], s.stride[1], 1, 0, kc);
                    cr = sample1(s.p[2], s.stride[2], 1, 0, kc);
                }
                ou = add(mul(cu, na), mul(cb, opacity));
                ov = add(mul(cv, na), mul(cr, opacity));
            }
            return true;
        }
        const float3 f = rgb2yuv(fc.x, fc.y, fc.z);
        const float a = mul(opacity, fc.w), na = sub(1.f, a);
        oy = clampf(add(mul(cy, na), mul(f.x, a)), 0.f, 1.f);
        if (chroma) {
            ou = clampf(add(mul(cu, na), mul(f.y, a)), -1.f, 1.f);
            ov = clampf(add(mul(cv, na), mul(f.z, a)), -1.f, 1.f);
        }
        return true;
    }
    // BGRA / RGBA sources
    if (!in_tx) return false;
    Taps k = {};
    if (in_uv) k = make_taps(uu, vv, s.w, s.h);
    rgba_pixel(s, opacity, fc, in_uv, k, cy, cu, cv, oy, ou, ov);
    return true;
}

}  // namespace svb
