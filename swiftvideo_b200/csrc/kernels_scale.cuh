// kernels_scale.cuh -- svb_scale_convert: NV12 / P010 -> BGRA with a separable bilinear or Lanczos-3 resize, one launch.
//
// An EXTENSION of the reference's operator set, not a replacement of anything in it: upstream has no high-bit-depth
// format (sample.pict.swift:19 "TODO: Higher bit-depth formats"), no filter but the OpenCL linear sampler
// (kernels.cl.swift:61) and no kernel that writes BGRA on Linux (img_bgra_bgra, compute.swift:54, exists only as a
// half-written Metal body, kernels.metal:51-62).  BASELINE.json's config 5 (3840x2160 P010 -> 1920x1080 BGRA, Lanczos)
// and the NV12 -> BGRA leg of config 2 name the operator with libswscale as comparator; what is computed is defined by
// oracle/scale_oracle.c (swscale's filter construction in floating point, every multiply-add one fused operation) and the
// bytes below equal that definition's exactly; against libswscale 9.1 in its accurate mode they differ by at most one
// code value (tests/test_scale.py).
//
// One CTA = a 64 x 32 tile of the output.  Pass 1 filters horizontally every source row the tile's output rows reach
// (luma rows and chroma rows, U and V together) into shared memory: a thread owns one output COLUMN, keeps that
// column's tap weights in registers and walks down the rows, so neighbouring lanes read neighbouring source samples.
// Pass 2 filters vertically out of shared memory (conflict-free: a warp reads 32 consecutive floats of a row), converts
// BT.601 limited-range YUV to RGB and stores one BGRA pixel per thread -- a warp writes 128 contiguous bytes.  The
// intermediate never touches HBM: algorithmic traffic = source planes once + BGRA once.
#pragma once
#include "svb_device.cuh"

#define SVB_SCALE_TW 64
#define SVB_SCALE_TH 32
#define SVB_SCALE_MAX_TAPS 16  // Lanczos-3 down to 1 : 2.66, bilinear down to 1 : 8

namespace svb {

__device__ __forceinline__ unsigned ldg_u16(const uint8_t* p) {
    unsigned v;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned ldg_u32(const uint8_t* p) {
    unsigned v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// P010: ten bits in the MSBs of a little-endian 16-bit word -> the 8-bit scale (exact: a multiple of 1/4)
__device__ __forceinline__ float p010_f(unsigned w) { return __fmul_rn(__uint2float_rn(w >> 6), 0.25f); }

__device__ __forceinline__ unsigned store8(float v) { return (unsigned)__float2int_rn(fminf(fmaxf(v, 0.f), 255.f)); }

}  // namespace svb

extern "C" __global__ void __launch_bounds__(256) svb_scale_convert(const SvbScaleDesc d) {
    using namespace svb;
    extern __shared__ __align__(16) float sc_smem[];
    float* const hy = sc_smem;                                        // [spanYy][TW]
    float* const hu = hy + (size_t)d.spanYy * SVB_SCALE_TW;           // [spanCy][TW]
    float* const hv = hu + (size_t)d.spanCy * SVB_SCALE_TW;           // [spanCy][TW]
    const int t = threadIdx.x, col = t & (SVB_SCALE_TW - 1), rg = t >> 6;
    const int x0 = blockIdx.x * SVB_SCALE_TW, y0 = blockIdx.y * SVB_SCALE_TH;
    const int x = min(x0 + col, d.dstW - 1);
    const int ylast = min(y0 + SVB_SCALE_TH, d.dstH) - 1;
    const int32_t* __restrict__ fYy = (const int32_t*)d.fYy;
    const int32_t* __restrict__ fCy = (const int32_t*)d.fCy;
    const int ry0 = __ldg(fYy + y0), nry = __ldg(fYy + ylast) + d.nYy - ry0;  // source rows [ry0, ry0 + nry) (before clamping)
    const int cy0 = __ldg(fCy + y0), ncr = __ldg(fCy + ylast) + d.nCy - cy0;
    const bool p010 = d.format == 1;
    const int cw = d.srcW >> 1, ch = d.srcH >> 1;

    // ---- pass 1: horizontal -----------------------------------------------------------------------------------------
    {
        float w[SVB_SCALE_MAX_TAPS];
        const float* __restrict__ wt = (const float*)d.wYx + (size_t)x * d.nYx;
#pragma unroll
        for (int k = 0; k < SVB_SCALE_MAX_TAPS; ++k) w[k] = k < d.nYx ? __ldg(wt + k) : 0.f;
        const int f = __ldg((const int32_t*)d.fYx + x);
        const uint8_t* __restrict__ src = (const uint8_t*)d.srcY;
        for (int i = rg; i < nry; i += 4) {
            const uint8_t* __restrict__ row = src + (size_t)min(max(ry0 + i, 0), d.srcH - 1) * d.strideY;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < SVB_SCALE_MAX_TAPS; ++k)
                if (k < d.nYx) {
                    const int sx = min(max(f + k, 0), d.srcW - 1);
                    const float v = p010 ? p010_f(ldg_u16(row + 2 * sx)) : __uint2float_rn(ldg_u8(row + sx));
                    acc = __fmaf_rn(w[k], v, acc);
                }
            hy[i * SVB_SCALE_TW + col] = acc;
        }
    }
    {
        float w[SVB_SCALE_MAX_TAPS];
        const float* __restrict__ wt = (const float*)d.wCx + (size_t)x * d.nCx;
#pragma unroll
        for (int k = 0; k < SVB_SCALE_MAX_TAPS; ++k) w[k] = k < d.nCx ? __ldg(wt + k) : 0.f;
        const int f = __ldg((const int32_t*)d.fCx + x);
        const uint8_t* __restrict__ src = (const uint8_t*)d.srcC;
        for (int i = rg; i < ncr; i += 4) {
            const uint8_t* __restrict__ row = src + (size_t)min(max(cy0 + i, 0), ch - 1) * d.strideC;
            float au = 0.f, av = 0.f;
#pragma unroll
            for (int k = 0; k < SVB_SCALE_MAX_TAPS; ++k)
                if (k < d.nCx) {
                    const int sx = min(max(f + k, 0), cw - 1);
                    float u, v;
                    if (p010) {
                        const unsigned q = ldg_u32(row + 4 * sx);
                        u = p010_f(q & 0xffffu), v = p010_f(q >> 16);
                    } else {
                        const unsigned q = ldg_u16(row + 2 * sx);
                        u = __uint2float_rn(opaque(q & 0xffu)), v = __uint2float_rn(opaque(q >> 8));
                    }
                    au = __fmaf_rn(w[k], u, au);
                    av = __fmaf_rn(w[k], v, av);
                }
            hu[i * SVB_SCALE_TW + col] = au;
            hv[i * SVB_SCALE_TW + col] = av;
        }
    }
    __syncthreads();

    // ---- pass 2: vertical, colour, store ------------------------------------------------------------------------------
    if (x0 + col >= d.dstW) return;
    const float* __restrict__ wYy = (const float*)d.wYy;
    const float* __restrict__ wCy = (const float*)d.wCy;
    for (int j = rg; j < SVB_SCALE_TH; j += 4) {
        const int y = y0 + j;
        if (y >= d.dstH) break;
        const float* __restrict__ hyc = hy + (__ldg(fYy + y) - ry0) * SVB_SCALE_TW + col;
        const float* __restrict__ wy = wYy + (size_t)y * d.nYy;
        float Y = 0.f;
        for (int k = 0; k < d.nYy; ++k) Y = __fmaf_rn(__ldg(wy + k), hyc[k * SVB_SCALE_TW], Y);
        const int co = (__ldg(fCy + y) - cy0) * SVB_SCALE_TW + col;
        const float* __restrict__ wc = wCy + (size_t)y * d.nCy;
        float U = 0.f, V = 0.f;
        for (int k = 0; k < d.nCy; ++k) {
            const float wk = __ldg(wc + k);
            U = __fmaf_rn(wk, hu[co + k * SVB_SCALE_TW], U);
            V = __fmaf_rn(wk, hv[co + k * SVB_SCALE_TW], V);
        }
        const float yy = __fmul_rn(1.164383f, __fsub_rn(Y, 16.f)), du = __fsub_rn(U, 128.f), dv = __fsub_rn(V, 128.f);
        const float R = __fmaf_rn(1.596027f, dv, yy);
        const float G = __fmaf_rn(-0.812968f, dv, __fmaf_rn(-0.391762f, du, yy));
        const float B = __fmaf_rn(2.017232f, du, yy);
        *(unsigned*)((uint8_t*)d.dst + (size_t)y * d.dstStride + 4 * (size_t)(x0 + col)) = store8(B) | (store8(G) << 8) | (store8(R) << 16) | 0xff000000u;
    }
}
