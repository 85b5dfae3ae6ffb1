// cl_shim.hpp -- just enough OpenCL C 1.2 (vector types, image reads/writes, samplers) for the
// reference's image kernels to compile as C++.  TEST INFRASTRUCTURE ONLY (see ../mixer_oracle.h).
//
// The kernel BODIES come verbatim from /root/reference/Sources/SwiftVideo/kernels.cl.swift
// (extracted by extract_kernels.py into oracle/_ref/, never committed).  This header supplies what
// an OpenCL runtime would: the builtins below follow the OpenCL 1.2 specification --
//   section 6.12.14 / 8.2  read_imagef: nearest (unnormalised) and linear (normalised, clamp-to-edge)
//   section 8.3.1.1        CL_UNORM_INT8 conversion rules (read c/255.0f, write sat_rte(f*255.0f))
//   section 6.12.2/6.12.4  dot, clamp
// with the canonical choices SURVEY.md appendix A fixes: fp32, no contraction, dot summed left to right.
#pragma once
#include <cmath>
#include <cstdint>

struct alignas(8) float2 {
    float x, y;
    float2() : x(0), y(0) {}
    float2(float a, float b) : x(a), y(b) {}
};
struct int2 {
    int x, y;
    int2() : x(0), y(0) {}
    int2(int a, int b) : x(a), y(b) {}
};
inline int2 operator/(int2 a, int b) { return int2(a.x / b, a.y / b); }

struct alignas(16) float4 {
    float x, y, z, w;
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float4 yzyz() const { return float4(y, z, y, z); }
};
struct float3 {
    float x, y, z;
    float3() : x(0), y(0), z(0) {}
    float4 yzyz() const { return float4(y, z, y, z); }
};
inline float4 operator*(float4 a, float s) { return float4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline float4 operator+(float4 a, float4 b) { return float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline float4 operator+(float4 a, float s) { return float4(a.x + s, a.y + s, a.z + s, a.w + s); }

inline float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float clamp(float v, float lo, float hi) { return std::fmin(std::fmax(v, lo), hi); }
inline float4 clamp(float4 v, float lo, float hi) {
    return float4(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi), clamp(v.w, lo, hi));
}

// An image2d_t: same field order as svo_plane so the driver can pass planes straight through.
struct ClImage {
    uint8_t* data;
    int32_t width, height, stride, ncomp;
};
typedef ClImage* image2d_t;
typedef unsigned sampler_t;
enum : unsigned {
    CLK_NORMALIZED_COORDS_FALSE = 0, CLK_NORMALIZED_COORDS_TRUE = 1,
    CLK_ADDRESS_NONE = 0, CLK_ADDRESS_CLAMP_TO_EDGE = 2,
    CLK_FILTER_NEAREST = 0, CLK_FILTER_LINEAR = 16
};
#define __kernel
#define __write_only
#define __read_only
#define __read_write
#define __constant const
#define __global

extern thread_local int g_gid[2];
extern thread_local int g_gsize[2];
inline int get_global_id(int d) { return g_gid[d]; }
inline int get_global_size(int d) { return g_gsize[d]; }

inline float cl_texel(const ClImage* im, int i, int j, int c) {
    if (c >= im->ncomp) return c == 3 ? 1.0f : 0.0f;  // CL_R / CL_RG read back (r,0,0,1) / (r,g,0,1)
    return (float)im->data[(int64_t)j * im->stride + i * im->ncomp + c] / 255.0f;
}
// nearest, unnormalised, CLK_ADDRESS_NONE (the "curSampler")
inline float4 read_imagef(image2d_t im, sampler_t, int2 p) {
    return float4(cl_texel(im, p.x, p.y, 0), cl_texel(im, p.x, p.y, 1), cl_texel(im, p.x, p.y, 2), cl_texel(im, p.x, p.y, 3));
}
inline int cl_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// linear, normalised, clamp-to-edge: OpenCL 1.2 section 8.2
inline float4 read_imagef(image2d_t im, sampler_t, float2 st) {
    float u = st.x * (float)im->width, v = st.y * (float)im->height;
    float fu = std::floor(u - 0.5f), fv = std::floor(v - 0.5f);
    float a = (u - 0.5f) - fu, b = (v - 0.5f) - fv;
    int i0 = cl_clampi((int)fu, 0, im->width - 1), i1 = cl_clampi((int)fu + 1, 0, im->width - 1);
    int j0 = cl_clampi((int)fv, 0, im->height - 1), j1 = cl_clampi((int)fv + 1, 0, im->height - 1);
    float r[4];
    for (int c = 0; c < 4; ++c)
        r[c] = (1.0f - a) * (1.0f - b) * cl_texel(im, i0, j0, c) + a * (1.0f - b) * cl_texel(im, i1, j0, c) +
               (1.0f - a) * b * cl_texel(im, i0, j1, c) + a * b * cl_texel(im, i1, j1, c);
    return float4(r[0], r[1], r[2], r[3]);
}
inline uint8_t cl_sat_rte(float f) {
    float v = f * 255.0f;
    if (!(v == v) || v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)std::rint(v);
}
inline void write_imagef(image2d_t im, int2 p, float4 c) {
    uint8_t* t = im->data + (int64_t)p.y * im->stride + p.x * im->ncomp;
    const float v[4] = {c.x, c.y, c.z, c.w};
    for (int k = 0; k < im->ncomp; ++k) t[k] = cl_sat_rte(v[k]);
}
inline void write_imagef(image2d_t im, int2 p, float s) { write_imagef(im, p, float4(s, s, s, s)); }
inline void write_imagef(image2d_t im, int2 p, double s) { write_imagef(im, p, (float)s); }

// swizzles used by the kernels (`x.yzyz`) become method calls
#define yzyz yzyz()
