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
// One CTA = a 64 x tileH tile of the output (tileH = 16 unless the filter's vertical footprint needs a shorter tile to fit).
//   1. The tile's source window is staged in shared memory as floats, TRANSPOSED (column-major, rows contiguous): 16-byte loads
//      of 16 / 8 samples, four in flight per thread, each sample converted once.  Columns beyond the picture's edge are
//      replicated by a fix-up pass.  (Planes whose base or stride is not 16-byte aligned are gathered sample by sample.)
//   2. Horizontal pass: a thread owns one output column (tap weights in registers) and FOUR source rows at a time: one 16-byte
//      shared-memory load per tap feeds four fused multiply-adds (the row-major layout cost a load per multiply-add and was
//      bound by load issue, not arithmetic: profiles/r1_svb_scale_convert_ncu_full.json).  Luma rows, then U and V together.
//      A warp covers 16 columns x 2 row groups so that its 16-byte loads fall in eight different bank groups.
//   3. Vertical pass: a thread owns four adjacent output pixels of one row (row weights in registers, 16-byte loads of the
//      horizontally filtered rows), converts BT.601 limited-range YUV to RGB and stores 16 bytes of BGRA.
// The intermediate never touches HBM: algorithmic traffic = source planes once + BGRA once.  Tap counts up to 16 are compiled
// per count (weights in registers, loops unrolled); larger counts run the same passes with run-time loops.
#pragma once
#include "svb_device.cuh"

// tuning knobs (tools/ab_scale.sh)
#ifndef SVB_SCALE_MIN_CTAS
#define SVB_SCALE_MIN_CTAS 4  // registers: 64 at 4 CTAs per SM, 85 at 3
#endif
#define SVB_SCALE_GSTEP 4      // row groups a column's threads stride by (4 threads per column in both mappings)


namespace svb {

__device__ __forceinline__ unsigned ldg_u32(const uint8_t* p) {
    unsigned v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned store8(float v) { return (unsigned)__float2int_rn(fminf(fmaxf(v, 0.f), 255.f)); }

// ---- staging: one 16-byte chunk of a source row -> floats, written down `S` columns of the transposed window -------------
struct Chunk { unsigned w[4]; };
__device__ __forceinline__ Chunk ldg_chunk(const uint8_t* p) {
    Chunk c;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c.w[0]), "=r"(c.w[1]), "=r"(c.w[2]), "=r"(c.w[3]) : "l"(p));
    return c;
}
// the same 16 bytes gathered sample by sample with every index clamped into the row (planes whose base or stride is not
// 16-byte aligned, or whose rows are too short to read a whole chunk at the right edge)
template <int BPS>
__device__ __forceinline__ Chunk gather_chunk(const uint8_t* row, int a, int last) {
    Chunk c;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        unsigned v = 0;
#pragma unroll
        for (int b = 0; b < 4 / BPS; ++b) {
            const int sx = min(max(a + q * (4 / BPS) + b, 0), last);
            v |= (BPS == 1 ? ldg_u8(row + sx) : ldg_u16(row + 2 * sx)) << (8 * BPS * b);
        }
        c.w[q] = v;
    }
    return c;
}
// P010 samples are staged as the 10-bit integer (four times the 8-bit-scale value the definition filters); the horizontal
// weights are scaled by 1/4 instead, which is exact, so every fused multiply-add sees the same real product.
__device__ __forceinline__ float p010_i(unsigned half) { return __uint2float_rn(half >> 6); }

// Staging one plane's window.  PAIRS = 1: single-component plane (luma) into t0; PAIRS = 2: interleaved (U, V) into t0 / t1.
// BPS = bytes per component.  Columns are counted in samples (luma) or pairs (chroma); a chunk holds S = 16 / (BPS * PAIRS) of
// them.  A work item is one chunk of one window row; a thread has four of them in flight (four 16-byte loads), then converts and
// stores every sample into its transposed column.  Items are numbered rows fastest: consecutive lanes write consecutive floats of
// one transposed column.  (Measured against items of four rows with 16-byte transposed stores and the chroma window parked by
// cp.async: fewer instructions, slower -- profiles/r2_ab_scale.log.)
struct Window {  // one plane's staged window
    const uint8_t* plane;
    int stride, planeW, planeH, ax0, nch, ry0, nrows;
    bool vec;
};
__device__ __forceinline__ void item_of(int idx, int n, float inv, int& q, int& r) {
    q = __float2int_rz(__fmul_rn((float)idx + 0.5f, inv));  // idx / n, one step off at worst
    r = idx - q * n;
    if (r < 0) --q, r += n;
    else if (r >= n) ++q, r -= n;
}
template <int BPS, int PAIRS>
__device__ __forceinline__ Chunk load_chunk(const Window& wn, int ci, int r) {
    constexpr int S = 16 / (BPS * PAIRS);
    const uint8_t* __restrict__ row = wn.plane + (size_t)min(max(wn.ry0 + r, 0), wn.planeH - 1) * wn.stride;
    const int a = wn.ax0 + ci * S;
    if (wn.vec) return ldg_chunk(row + min(max(a, 0), ((wn.planeW - 1) / S) * S) * (BPS * PAIRS));  // chunks wholly outside the row are re-read from its ends and replaced by the fix-up
    if (PAIRS == 1) return gather_chunk<BPS>(row, a, wn.planeW - 1);
    Chunk c;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (BPS == 1) {
            const int s0 = min(max(a + 2 * q, 0), wn.planeW - 1), s1 = min(max(a + 2 * q + 1, 0), wn.planeW - 1);
            c.w[q] = ldg_u16(row + 2 * s0) | (ldg_u16(row + 2 * s1) << 16);
        } else {
            c.w[q] = ldg_u32(row + 4 * min(max(a + q, 0), wn.planeW - 1));
        }
    }
    return c;
}
template <int BPS, int PAIRS>
__device__ __forceinline__ void store_chunk(const Chunk& c, float* __restrict__ t0, float* __restrict__ t1, int PR, int ci, int r) {
    constexpr int S = 16 / (BPS * PAIRS), PER = 4 / BPS;
    float* __restrict__ o0 = t0 + ci * S * PR + r;
    float* __restrict__ o1 = t1 + ci * S * PR + r;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int b = 0; b < PER; ++b) {
            const int comp = q * PER + b;
            const unsigned raw = BPS == 1 ? (c.w[q] >> (8 * b)) & 0xffu : (c.w[q] >> (16 * b)) & 0xffffu;
            const float v = BPS == 1 ? __uint2float_rn(opaque(raw)) : p010_i(raw);
            if (PAIRS == 1) o0[comp * PR] = v;
            else ((comp & 1) ? o1 : o0)[(comp >> 1) * PR] = v;
        }
}
template <int BPS, int PAIRS>
__device__ __forceinline__ void stage_rows(const Window& wn, int t, float* __restrict__ t0, float* __restrict__ t1, int PR) {
    const int items = wn.nch * wn.nrows;
    const float inv = __frcp_rn((float)wn.nrows);
    for (int base = t; base < items; base += 1024) {
        Chunk c[4];
        int ci[4], r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (base + 256 * i < items) {
                item_of(base + 256 * i, wn.nrows, inv, ci[i], r[i]);
                c[i] = load_chunk<BPS, PAIRS>(wn, ci[i], r[i]);
            }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (base + 256 * i < items) store_chunk<BPS, PAIRS>(c[i], t0, t1, PR, ci[i], r[i]);
    }
}
// Columns of the staged window that lie left or right of the picture take the edge column's samples (clamp-to-edge).
__device__ __forceinline__ void fix_edges(float* __restrict__ tp, int PR, int nrows, int ax0, int ncols, int planeW, int warp, int lane) {
    const int left = min(max(-ax0, 0), ncols), right = min(max(planeW - ax0, 0), ncols);  // valid columns: [left, right)
    if (left == 0 && right == ncols) return;
    for (int c = warp; c < ncols; c += 8) {
        if (c >= left && c < right) continue;
        const float* __restrict__ from = tp + (size_t)(c < left ? left : right - 1) * PR;
        float* __restrict__ to = tp + (size_t)c * PR;
        for (int r = lane; r < nrows; r += 32) to[r] = from[r];
    }
}

// ---- horizontal pass: thread = (output column, four source rows); NT = 0: run-time tap count ----------------------------------
// w: the column's weights, fetched by the caller long before (16 registers; unused ones are dead code once NT is known).
template <int NT>
__device__ __forceinline__ void hpass1(const float (&w)[16], const float* __restrict__ wt, int nt, float ws, const float* __restrict__ colT, int PR, int groups, int g0, int col,
                                       float* __restrict__ out) {
    for (int g = g0; g < groups; g += SVB_SCALE_GSTEP) {
        const float* __restrict__ p = colT + 4 * g;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (NT) {
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                const float4 v = *(const float4*)(p + k * PR);
                acc.x = __fmaf_rn(w[k], v.x, acc.x), acc.y = __fmaf_rn(w[k], v.y, acc.y), acc.z = __fmaf_rn(w[k], v.z, acc.z), acc.w = __fmaf_rn(w[k], v.w, acc.w);
            }
        } else {
            for (int k = 0; k < nt; ++k) {
                const float wk = __fmul_rn(__ldg(wt + k), ws);
                const float4 v = *(const float4*)(p + k * PR);
                acc.x = __fmaf_rn(wk, v.x, acc.x), acc.y = __fmaf_rn(wk, v.y, acc.y), acc.z = __fmaf_rn(wk, v.z, acc.z), acc.w = __fmaf_rn(wk, v.w, acc.w);
            }
        }
        float* __restrict__ o = out + (4 * g) * SVB_SCALE_HP + col;
        o[0] = acc.x, o[SVB_SCALE_HP] = acc.y, o[2 * SVB_SCALE_HP] = acc.z, o[3 * SVB_SCALE_HP] = acc.w;
    }
}
template <int NT>
__device__ __forceinline__ void hpass2(const float (&w)[16], const float* __restrict__ wt, int nt, float ws, const float* __restrict__ colU, const float* __restrict__ colV, int PR,
                                       int groups, int g0, int col, float* __restrict__ outU, float* __restrict__ outV) {
    for (int g = g0; g < groups; g += SVB_SCALE_GSTEP) {
        const float* __restrict__ pu = colU + 4 * g;
        const float* __restrict__ pv = colV + 4 * g;
        float4 au = make_float4(0.f, 0.f, 0.f, 0.f), av = au;
        if (NT) {
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                const float4 u = *(const float4*)(pu + k * PR), v = *(const float4*)(pv + k * PR);
                au.x = __fmaf_rn(w[k], u.x, au.x), au.y = __fmaf_rn(w[k], u.y, au.y), au.z = __fmaf_rn(w[k], u.z, au.z), au.w = __fmaf_rn(w[k], u.w, au.w);
                av.x = __fmaf_rn(w[k], v.x, av.x), av.y = __fmaf_rn(w[k], v.y, av.y), av.z = __fmaf_rn(w[k], v.z, av.z), av.w = __fmaf_rn(w[k], v.w, av.w);
            }
        } else {
            for (int k = 0; k < nt; ++k) {
                const float wk = __fmul_rn(__ldg(wt + k), ws);
                const float4 u = *(const float4*)(pu + k * PR), v = *(const float4*)(pv + k * PR);
                au.x = __fmaf_rn(wk, u.x, au.x), au.y = __fmaf_rn(wk, u.y, au.y), au.z = __fmaf_rn(wk, u.z, au.z), au.w = __fmaf_rn(wk, u.w, au.w);
                av.x = __fmaf_rn(wk, v.x, av.x), av.y = __fmaf_rn(wk, v.y, av.y), av.z = __fmaf_rn(wk, v.z, av.z), av.w = __fmaf_rn(wk, v.w, av.w);
            }
        }
        float* __restrict__ ou = outU + (4 * g) * SVB_SCALE_HP + col;
        float* __restrict__ ov = outV + (4 * g) * SVB_SCALE_HP + col;
        ou[0] = au.x, ou[SVB_SCALE_HP] = au.y, ou[2 * SVB_SCALE_HP] = au.z, ou[3 * SVB_SCALE_HP] = au.w;
        ov[0] = av.x, ov[SVB_SCALE_HP] = av.y, ov[2 * SVB_SCALE_HP] = av.z, ov[3 * SVB_SCALE_HP] = av.w;
    }
}
// the column's weights into registers, scaled (see p010_i); slots beyond the tap count stay zero and unused
__device__ __forceinline__ void fetch_weights(const float* __restrict__ wt, int nt, float ws, float (&w)[16]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = k < nt && nt <= 16 ? __fmul_rn(__ldg(wt + k), ws) : 0.f;
}
// ---- vertical pass: four adjacent columns of one output row; the row's weights come from shared memory (staged at kernel start) ----
template <int NT>
__device__ __forceinline__ float4 vpass4(const float* __restrict__ w, const float* __restrict__ wg, int nt, const float* __restrict__ h) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (NT) {
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            const float wk = w[k];
            const float4 v = *(const float4*)(h + k * SVB_SCALE_HP);
            acc.x = __fmaf_rn(wk, v.x, acc.x), acc.y = __fmaf_rn(wk, v.y, acc.y), acc.z = __fmaf_rn(wk, v.z, acc.z), acc.w = __fmaf_rn(wk, v.w, acc.w);
        }
    } else {
        for (int k = 0; k < nt; ++k) {
            const float wk = __ldg(wg + k);
            const float4 v = *(const float4*)(h + k * SVB_SCALE_HP);
            acc.x = __fmaf_rn(wk, v.x, acc.x), acc.y = __fmaf_rn(wk, v.y, acc.y), acc.z = __fmaf_rn(wk, v.z, acc.z), acc.w = __fmaf_rn(wk, v.w, acc.w);
        }
    }
    return acc;
}
__device__ __forceinline__ unsigned bgra_of(float Y, float U, float V) {
    const float yy = __fmul_rn(1.164383f, __fsub_rn(Y, 16.f)), du = __fsub_rn(U, 128.f), dv = __fsub_rn(V, 128.f);
    const float R = __fmaf_rn(1.596027f, dv, yy);
    const float G = __fmaf_rn(-0.812968f, dv, __fmaf_rn(-0.391762f, du, yy));
    const float B = __fmaf_rn(2.017232f, du, yy);
    return store8(B) | (store8(G) << 8) | (store8(R) << 16) | 0xff000000u;
}
#define SVB_TAPS_SWITCH(n, CALL)                     \
    switch (n) {                                     \
    case 2: { constexpr int NT = 2; CALL; } break;   \
    case 4: { constexpr int NT = 4; CALL; } break;   \
    case 6: { constexpr int NT = 6; CALL; } break;   \
    case 8: { constexpr int NT = 8; CALL; } break;   \
    case 10: { constexpr int NT = 10; CALL; } break; \
    case 12: { constexpr int NT = 12; CALL; } break; \
    case 14: { constexpr int NT = 14; CALL; } break; \
    case 16: { constexpr int NT = 16; CALL; } break; \
    default: { constexpr int NT = 0; CALL; } break;  \
    }

// FIXED: the windows' row pitches are the compile-time SVB_SCALE_PITCH_Y / _C (every shared-memory address of the horizontal pass
// is then a register plus an immediate); otherwise they come from the descriptor (tiles whose vertical footprint is larger).
template <bool FIXED>
__device__ __forceinline__ void scale_body(const SvbScaleDesc& d) {
    extern __shared__ __align__(16) float sc_smem[];
    const int PRY = FIXED ? SVB_SCALE_PITCH_Y : d.pitchY, PRC = FIXED ? SVB_SCALE_PITCH_C : d.pitchC;
    const int rowsY4 = (d.spanYy + 3) & ~3, rowsC4 = (d.spanCy + 3) & ~3;
    float* const vw = sc_smem;                                    // [2][16 rows][16] vertical weights of the tile's rows (luma, chroma)
    float* const hy = vw + 2 * SVB_SCALE_TH * 16;                 // [rowsY4][TW]  horizontally filtered luma rows
    float* const hu = hy + rowsY4 * SVB_SCALE_HP;                 // [rowsC4][HP]
    float* const hv = hu + rowsC4 * SVB_SCALE_HP;                 // [rowsC4][HP]
    float* const raw = hv + rowsC4 * SVB_SCALE_HP;                // the tile's source window, transposed: [column][pitch rows]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // horizontal pass: a warp = 16 columns x 2 row groups, a quarter-warp = 4 adjacent columns x 2 row groups.  The 16-byte slot a lane
    // reads is (column * pitch/4 + group): with pitch/4 odd, four adjacent output columns at source stride 1 or 2 land in four different
    // slots mod 8 of one parity and the second row group fills the other parity -- no bank conflicts (32 columns of one group: two-way)
    // (-7 % kernel time at 2 : 1, profiles/r2_ab_scale.log)
    const int col = (warp & 3) * 16 + (lane & 3) + 4 * (lane >> 3), rg = 2 * (warp >> 2) + ((lane >> 2) & 1);
    const int x0 = blockIdx.x * SVB_SCALE_TW, y0 = blockIdx.y * d.tileH;
    const int x = min(x0 + col, d.dstW - 1);
    const int xlast = min(x0 + SVB_SCALE_TW, d.dstW) - 1, ylast = min(y0 + d.tileH, d.dstH) - 1;
    const int32_t* __restrict__ fYx = (const int32_t*)d.fYx;
    const int32_t* __restrict__ fCx = (const int32_t*)d.fCx;
    const int32_t* __restrict__ fYy = (const int32_t*)d.fYy;
    const int32_t* __restrict__ fCy = (const int32_t*)d.fCy;
    // source window of the tile (before clamping; firsts are monotone, so the tile's end columns / rows bound it)
    const int rx0 = __ldg(fYx + x0), rx1 = __ldg(fYx + xlast), ry0 = __ldg(fYy + y0), ry1 = __ldg(fYy + ylast);
    const int cx0 = __ldg(fCx + x0), cx1 = __ldg(fCx + xlast), cy0 = __ldg(fCy + y0), cy1 = __ldg(fCy + ylast);
    const int offY = __ldg(fYx + x), offC = __ldg(fCx + x);
    // the thread's row of the vertical pass
    const int j = t >> 4, c4 = (t & 15) * 4, y = min(y0 + j, d.dstH - 1);
    const int vy = __ldg(fYy + y), vc = __ldg(fCy + y);
    const bool p010 = d.format == 1;
    const float ws = p010 ? 0.25f : 1.0f;  // see p010_i
    const int nrx = rx1 + d.nYx - rx0, nry = ry1 + d.nYy - ry0, ncx = cx1 + d.nCx - cx0, ncr = cy1 + d.nCy - cy0;
    const int SY = p010 ? 8 : 16, SC = p010 ? 4 : 8;  // samples / pairs per 16-byte chunk
    Window wy, wc;
    wy.plane = (const uint8_t*)d.srcY, wy.stride = d.strideY, wy.planeW = d.srcW, wy.planeH = d.srcH;
    wy.ax0 = rx0 & ~(SY - 1), wy.nch = (rx0 + nrx - 1 - wy.ax0) / SY + 1, wy.ry0 = ry0, wy.nrows = nry, wy.vec = d.vecY != 0;
    wc.plane = (const uint8_t*)d.srcC, wc.stride = d.strideC, wc.planeW = d.srcW >> 1, wc.planeH = d.srcH >> 1;
    wc.ax0 = cx0 & ~(SC - 1), wc.nch = (cx0 + ncx - 1 - wc.ax0) / SC + 1, wc.ry0 = cy0, wc.nrows = ncr, wc.vec = d.vecC != 0;

    float w[16];
    {   // vertical weights of the tile's rows: thread t fetches element (row t / 16, tap t % 16) of both tables
        const int vr = min(y0 + (t >> 4), d.dstH - 1), vk = t & 15;
        vw[t] = vk < d.nYy && d.nYy <= 16 ? __ldg((const float*)d.wYy + (size_t)vr * d.nYy + vk) : 0.f;
        vw[256 + t] = vk < d.nCy && d.nCy <= 16 ? __ldg((const float*)d.wCy + (size_t)vr * d.nCy + vk) : 0.f;
    }
    // ---- luma ------------------------------------------------------------------------------------------------------------
    if (p010) stage_rows<2, 1>(wy, t, raw, raw, PRY);
    else stage_rows<1, 1>(wy, t, raw, raw, PRY);
    fetch_weights((const float*)d.wYx + (size_t)x * d.nYx, d.nYx, ws, w);
    __syncthreads();
    if (wy.ax0 < 0 || wy.ax0 + wy.nch * SY > d.srcW) {
        fix_edges(raw, PRY, nry, wy.ax0, wy.nch * SY, d.srcW, warp, lane);
        __syncthreads();
    }
    SVB_TAPS_SWITCH(d.nYx, hpass1<NT>(w, (const float*)d.wYx + (size_t)x * d.nYx, d.nYx, ws, raw + (offY - wy.ax0) * PRY, PRY, (nry + 3) >> 2, rg, col, hy))
    fetch_weights((const float*)d.wCx + (size_t)x * d.nCx, d.nCx, ws, w);
    __syncthreads();
    // ---- chroma: the same over the (U, V) pairs, de-interleaved while staging (the window reuses the luma window's memory) ------
    float* const rawU = raw;
    float* const rawV = raw + d.spanCx * PRC;
    if (p010) stage_rows<2, 2>(wc, t, rawU, rawV, PRC);
    else stage_rows<1, 2>(wc, t, rawU, rawV, PRC);
    __syncthreads();
    if (wc.ax0 < 0 || wc.ax0 + wc.nch * SC > wc.planeW) {
        fix_edges(rawU, PRC, ncr, wc.ax0, wc.nch * SC, wc.planeW, warp, lane);
        fix_edges(rawV, PRC, ncr, wc.ax0, wc.nch * SC, wc.planeW, warp, lane);
        __syncthreads();
    }
    {
        const int off = (offC - wc.ax0) * PRC;
        SVB_TAPS_SWITCH(d.nCx, hpass2<NT>(w, (const float*)d.wCx + (size_t)x * d.nCx, d.nCx, ws, rawU + off, rawV + off, PRC, (ncr + 3) >> 2, rg, col, hu, hv))
    }
    __syncthreads();
    // ---- vertical, colour, store ----------------------------------------------------------------------------------------
    if (j >= d.tileH || y0 + j >= d.dstH || x0 + c4 >= d.dstW) return;
    float4 Y, U, V;
    SVB_TAPS_SWITCH(d.nYy, Y = vpass4<NT>(vw + 16 * j, (const float*)d.wYy + (size_t)y * d.nYy, d.nYy, hy + (vy - ry0) * SVB_SCALE_HP + c4))
    {
        const int co = (vc - cy0) * SVB_SCALE_HP + c4;
        const float* __restrict__ wg = (const float*)d.wCy + (size_t)y * d.nCy;
        SVB_TAPS_SWITCH(d.nCy, (U = vpass4<NT>(vw + 256 + 16 * j, wg, d.nCy, hu + co), V = vpass4<NT>(vw + 256 + 16 * j, wg, d.nCy, hv + co)))
    }
    uint4 px;
    px.x = bgra_of(Y.x, U.x, V.x), px.y = bgra_of(Y.y, U.y, V.y), px.z = bgra_of(Y.z, U.z, V.z), px.w = bgra_of(Y.w, U.w, V.w);
    uint8_t* const o = (uint8_t*)d.dst + (size_t)y * d.dstStride + 4 * (size_t)(x0 + c4);
    if (d.vecDst && x0 + c4 + 3 < d.dstW) {
        *(uint4*)o = px;
    } else {
        const unsigned v[4] = {px.x, px.y, px.z, px.w};
#pragma unroll
        for (int m = 0; m < 4; ++m)
            if (x0 + c4 + m < d.dstW) ((unsigned*)o)[m] = v[m];
    }
}

}  // namespace svb

extern "C" __global__ void __launch_bounds__(256, SVB_SCALE_MIN_CTAS) svb_scale_convert(const SvbScaleDesc d) { svb::scale_body<true>(d); }
extern "C" __global__ void __launch_bounds__(256) svb_scale_convert_any(const SvbScaleDesc d) { svb::scale_body<false>(d); }
