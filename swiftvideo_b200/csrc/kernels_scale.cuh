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
// One CTA = a 64 x 16 tile of the output (SVB_SCALE_TW x SVB_SCALE_TH).  The tile's source window is staged in shared memory as floats (each source
// sample read from HBM and converted once, coalesced), then filtered horizontally for every source row the tile's output
// rows reach (luma rows, then chroma rows, U and V together): a thread owns one output COLUMN, keeps that column's tap
// weights in registers and walks down the rows.  The last pass filters vertically out of shared memory (conflict-free:
// a warp reads 32 consecutive floats of a row), converts BT.601 limited-range YUV to RGB and stores one BGRA pixel per thread -- a warp writes 128 contiguous bytes.  The
// intermediate never touches HBM: algorithmic traffic = source planes once + BGRA once.
#pragma once
#include "svb_device.cuh"


namespace svb {

__device__ __forceinline__ unsigned ldg_u32(const uint8_t* p) {
    unsigned v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// P010: ten bits in the MSBs of a little-endian 16-bit word -> the 8-bit scale (exact: a multiple of 1/4)
__device__ __forceinline__ float p010_f(unsigned w) { return __fmul_rn(__uint2float_rn(w >> 6), 0.25f); }

__device__ __forceinline__ unsigned store8(float v) { return (unsigned)__float2int_rn(fminf(fmaxf(v, 0.f), 255.f)); }

// Tap counts are even numbers up to SVB_SCALE_MAX_TAPS; the passes are compiled per count (run-time bounds cost a compare per
// tap, as many issue slots as the multiply-adds themselves) and chosen by a switch that is uniform over the launch.
template <int NT>
__device__ __forceinline__ void hpass1(const float* __restrict__ wt, const float* __restrict__ src, int pitch, int rows, int rg, int col, float* __restrict__ out) {
    float w[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) w[k] = __ldg(wt + k);
    for (int i = rg; i < rows; i += 4) {
        const float* __restrict__ row = src + i * pitch;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < NT; ++k) acc = __fmaf_rn(w[k], row[k], acc);
        out[i * SVB_SCALE_TW + col] = acc;
    }
}
template <int NT>
__device__ __forceinline__ void hpass2(const float* __restrict__ wt, const float* __restrict__ srcU, const float* __restrict__ srcV, int pitch, int rows, int rg, int col,
                                       float* __restrict__ outU, float* __restrict__ outV) {
    float w[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) w[k] = __ldg(wt + k);
    for (int i = rg; i < rows; i += 4) {
        const float* __restrict__ ru = srcU + i * pitch;
        const float* __restrict__ rv = srcV + i * pitch;
        float au = 0.f, av = 0.f;
#pragma unroll
        for (int k = 0; k < NT; ++k) au = __fmaf_rn(w[k], ru[k], au), av = __fmaf_rn(w[k], rv[k], av);
        outU[i * SVB_SCALE_TW + col] = au;
        outV[i * SVB_SCALE_TW + col] = av;
    }
}
template <int NT>
__device__ __forceinline__ float vpass1(const float* __restrict__ w, const float* __restrict__ h) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < NT; ++k) acc = __fmaf_rn(__ldg(w + k), h[k * SVB_SCALE_TW], acc);
    return acc;
}
#define SVB_TAPS_SWITCH(n, CALL)                 \
    switch (n) {                                 \
    case 2: { constexpr int NT = 2; CALL; } break;   \
    case 4: { constexpr int NT = 4; CALL; } break;   \
    case 6: { constexpr int NT = 6; CALL; } break;   \
    case 8: { constexpr int NT = 8; CALL; } break;   \
    case 10: { constexpr int NT = 10; CALL; } break; \
    case 12: { constexpr int NT = 12; CALL; } break; \
    case 14: { constexpr int NT = 14; CALL; } break; \
    default: { constexpr int NT = 16; CALL; } break; \
    }

}  // namespace svb

extern "C" __global__ void __launch_bounds__(256) svb_scale_convert(const SvbScaleDesc d) {
    using namespace svb;
    extern __shared__ __align__(16) float sc_smem[];
    float* const hy = sc_smem;                                        // [spanYy][TW]  horizontally filtered luma rows
    float* const hu = hy + (size_t)d.spanYy * SVB_SCALE_TW;           // [spanCy][TW]
    float* const hv = hu + (size_t)d.spanCy * SVB_SCALE_TW;           // [spanCy][TW]
    float* const raw = hv + (size_t)d.spanCy * SVB_SCALE_TW;          // the tile's source window as floats: luma, then (U | V)
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, col = t & (SVB_SCALE_TW - 1), rg = t >> 6;
    const int x0 = blockIdx.x * SVB_SCALE_TW, y0 = blockIdx.y * SVB_SCALE_TH;
    const int x = min(x0 + col, d.dstW - 1);
    const int xlast = min(x0 + SVB_SCALE_TW, d.dstW) - 1, ylast = min(y0 + SVB_SCALE_TH, d.dstH) - 1;
    const int32_t* __restrict__ fYx = (const int32_t*)d.fYx;
    const int32_t* __restrict__ fCx = (const int32_t*)d.fCx;
    const int32_t* __restrict__ fYy = (const int32_t*)d.fYy;
    const int32_t* __restrict__ fCy = (const int32_t*)d.fCy;
    // source window of the tile (before clamping; firsts are monotone, so the tile's end columns / rows bound it)
    const int rx0 = __ldg(fYx + x0), nrx = __ldg(fYx + xlast) + d.nYx - rx0;
    const int ry0 = __ldg(fYy + y0), nry = __ldg(fYy + ylast) + d.nYy - ry0;
    const int cx0 = __ldg(fCx + x0), ncx = __ldg(fCx + xlast) + d.nCx - cx0;
    const int cy0 = __ldg(fCy + y0), ncr = __ldg(fCy + ylast) + d.nCy - cy0;
    const bool p010 = d.format == 1;
    const int cw = d.srcW >> 1, ch = d.srcH >> 1;
    const bool inX = rx0 >= 0 && rx0 + nrx <= d.srcW, inCX = cx0 >= 0 && cx0 + ncx <= cw;  // no column of the window needs clamping

    // ---- luma: stage the window as floats (every source sample is converted once, lanes read consecutive samples), then filter
    // horizontally: a thread owns one output COLUMN, keeps its tap weights in registers and walks down the rows
    // (eight independent loads per lane and row in flight: one load per iteration left the copy waiting on HBM latency)
    for (int r = warp; r < nry; r += 8) {
        const uint8_t* __restrict__ row = (const uint8_t*)d.srcY + (size_t)min(max(ry0 + r, 0), d.srcH - 1) * d.strideY;
        for (int c0 = lane; c0 < nrx; c0 += 256) {
            unsigned v[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int sx = inX ? rx0 + c0 + 32 * m : min(max(rx0 + c0 + 32 * m, 0), d.srcW - 1);
                v[m] = c0 + 32 * m < nrx ? (p010 ? ldg_u16(row + 2 * sx) : ldg_u8(row + sx)) : 0u;
            }
#pragma unroll
            for (int m = 0; m < 8; ++m)
                if (c0 + 32 * m < nrx) raw[r * d.spanYx + c0 + 32 * m] = p010 ? p010_f(v[m]) : __uint2float_rn(v[m]);
        }
    }
    __syncthreads();
    SVB_TAPS_SWITCH(d.nYx, hpass1<NT>((const float*)d.wYx + (size_t)x * d.nYx, raw + (__ldg(fYx + x) - rx0), d.spanYx, nry, rg, col, hy))
    __syncthreads();
    // ---- chroma: the same over the (U, V) pairs, de-interleaved while staging (the window reuses the luma window's memory)
    float* const rawU = raw;
    float* const rawV = raw + (size_t)d.spanCy * d.spanCx;
    for (int r = warp; r < ncr; r += 8) {
        const uint8_t* __restrict__ row = (const uint8_t*)d.srcC + (size_t)min(max(cy0 + r, 0), ch - 1) * d.strideC;
        for (int c0 = lane; c0 < ncx; c0 += 256) {
            unsigned v[8];
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int sx = inCX ? cx0 + c0 + 32 * m : min(max(cx0 + c0 + 32 * m, 0), cw - 1);
                v[m] = c0 + 32 * m < ncx ? (p010 ? ldg_u32(row + 4 * sx) : ldg_u16(row + 2 * sx)) : 0u;
            }
#pragma unroll
            for (int m = 0; m < 8; ++m)
                if (c0 + 32 * m < ncx) {
                    const int o = r * d.spanCx + c0 + 32 * m;
                    if (p010) rawU[o] = p010_f(v[m] & 0xffffu), rawV[o] = p010_f(v[m] >> 16);
                    else rawU[o] = __uint2float_rn(opaque(v[m] & 0xffu)), rawV[o] = __uint2float_rn(opaque(v[m] >> 8));
                }
        }
    }
    __syncthreads();
    {
        const int off = __ldg(fCx + x) - cx0;
        SVB_TAPS_SWITCH(d.nCx, hpass2<NT>((const float*)d.wCx + (size_t)x * d.nCx, rawU + off, rawV + off, d.spanCx, ncr, rg, col, hu, hv))
    }
    __syncthreads();

    // ---- vertical, colour, store ----------------------------------------------------------------------------------------
    if (x0 + col >= d.dstW) return;
    const float* __restrict__ wYy = (const float*)d.wYy;
    const float* __restrict__ wCy = (const float*)d.wCy;
    for (int j = rg; j < SVB_SCALE_TH; j += 4) {
        const int y = y0 + j;
        if (y >= d.dstH) break;
        const float* __restrict__ hyc = hy + (__ldg(fYy + y) - ry0) * SVB_SCALE_TW + col;
        const float* __restrict__ wy = wYy + (size_t)y * d.nYy;
        float Y, U, V;
        SVB_TAPS_SWITCH(d.nYy, Y = vpass1<NT>(wy, hyc))
        const int co = (__ldg(fCy + y) - cy0) * SVB_SCALE_TW + col;
        const float* __restrict__ wc = wCy + (size_t)y * d.nCy;
        SVB_TAPS_SWITCH(d.nCy, (U = vpass1<NT>(wc, hu + co), V = vpass1<NT>(wc, hv + co)))
        const float yy = __fmul_rn(1.164383f, __fsub_rn(Y, 16.f)), du = __fsub_rn(U, 128.f), dv = __fsub_rn(V, 128.f);
        const float R = __fmaf_rn(1.596027f, dv, yy);
        const float G = __fmaf_rn(-0.812968f, dv, __fmaf_rn(-0.391762f, du, yy));
        const float B = __fmaf_rn(2.017232f, du, yy);
        *(unsigned*)((uint8_t*)d.dst + (size_t)y * d.dstStride + 4 * (size_t)(x0 + col)) = store8(B) | (store8(G) << 8) | (store8(R) << 16) | 0xff000000u;
    }
}
