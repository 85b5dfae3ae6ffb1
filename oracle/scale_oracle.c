/* scale_oracle.c -- CPU definition of the convert+scale operator (NV12 / P010 -> BGRA, bilinear or Lanczos-3).
 *
 * TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg).
 *
 * PARITY UNPINNED UPSTREAM.  The reference has no such operator: sample.pict.swift:19 is "// TODO: Higher bit-depth
 * formats", every Plane is built with bitDepth 8 (sample.pict.linux.swift:279-290), the only filter is the OpenCL linear
 * sampler (kernels.cl.swift:61), and the one name it reserves for an RGB target, img_bgra_bgra (compute.swift:54), has
 * only a half-written Metal body (kernels.metal:51-62).  BASELINE.json's config 5 (3840x2160 P010 -> 1920x1080 BGRA,
 * Lanczos) and the "NV12 -> BGRA -> scale" leg of config 2 name it nevertheless, with FFmpeg's libswscale as the CPU
 * comparator.  This file is therefore the DEFINITION the CUDA kernel must match bit for bit, and tests/test_scale.py
 * holds it against libswscale 9.1 (SWS_BILINEAR / SWS_LANCZOS, the copy bundled with the OpenCV wheel) within a
 * tolerance on smooth content -- swscale works in 14/15-bit fixed point and sites chroma differently, so bytes differ.
 *
 * Definition (swscale's initFilter scheme, in floating point):
 *   ratio r = srcN / dstN per axis and plane; filter scale s = max(1, r); support = 1*s (bilinear) or 3*s (Lanczos-3)
 *   taps  n = 2*ceil(support) source samples starting at first(x) = floor(c - support) + 1,
 *           c(x) = (x + 0.5) * r - 0.5 (pixel centres; chroma planes use their own size: centre-sited)
 *   weight  w_k = f((first + k - c) / s), f = 1 - |t| (bilinear) or sinc(t) sinc(t/3), |t| < 3 (Lanczos), computed in
 *           double, normalised to sum 1, rounded to float; source indices are clamped to the plane (edge replicate)
 *   samples in 8-bit scale as float: NV12 byte b -> b; P010 word w (10 bits in the MSBs, little endian) -> (w >> 6) / 4
 *   horizontal pass first: h[row][x] = fma-chain over k ascending, starting from 0; then the vertical pass over h the
 *   same way; every multiply-add is ONE fused operation (fmaf), nothing else is contracted (-ffp-contract=off)
 *   colour: BT.601 limited range in, full-range RGB out (swscale's default):
 *           yy = 1.164383f * (Y - 16);  R = fma(1.596027f, V - 128, yy)
 *           G = fma(-0.812968f, V - 128, fma(-0.391762f, U - 128, yy));  B = fma(2.017232f, U - 128, yy)
 *   store   BGRA bytes: rint (ties to even) of the value clamped to [0, 255]; A = 255
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { SC_NV12 = 0, SC_P010 = 1 };
enum { SC_BILINEAR = 0, SC_LANCZOS3 = 1 };

static double sc_weight(int filter, double t) {
    t = fabs(t);
    if (filter == SC_BILINEAR) return t < 1.0 ? 1.0 - t : 0.0;
    if (t >= 3.0) return 0.0;
    if (t < 1e-9) return 1.0;
    const double pt = M_PI * t;
    return (sin(pt) / pt) * (sin(pt / 3.0) / (pt / 3.0));
}

/* number of taps for srcN -> dstN */
int svo_scale_taps(int filter, int srcN, int dstN) {
    const double r = (double)srcN / (double)dstN, s = r > 1.0 ? r : 1.0;
    const double support = (filter == SC_BILINEAR ? 1.0 : 3.0) * s;
    return 2 * (int)ceil(support);
}

/* first[dstN], weights[dstN * taps] */
void svo_scale_table(int filter, int srcN, int dstN, int32_t* first, float* weights) {
    const double r = (double)srcN / (double)dstN, s = r > 1.0 ? r : 1.0;
    const double support = (filter == SC_BILINEAR ? 1.0 : 3.0) * s;
    const int n = 2 * (int)ceil(support);
    double* w = (double*)malloc(sizeof(double) * (size_t)n);
    for (int x = 0; x < dstN; ++x) {
        const double c = ((double)x + 0.5) * r - 0.5;
        const int f = (int)floor(c - support) + 1;
        double sum = 0.0;
        for (int k = 0; k < n; ++k) {
            w[k] = sc_weight(filter, ((double)(f + k) - c) / s);
            sum += w[k];
        }
        first[x] = f;
        for (int k = 0; k < n; ++k) weights[(size_t)x * n + k] = (float)(w[k] / sum);
    }
    free(w);
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static inline float sample_y(int fmt, const uint8_t* p, int stride, int x, int y) {
    if (fmt == SC_NV12) return (float)p[(size_t)y * stride + x];
    const uint8_t* q = p + (size_t)y * stride + 2 * (size_t)x;
    return (float)(((unsigned)q[0] | ((unsigned)q[1] << 8)) >> 6) * 0.25f;
}
static inline float sample_c(int fmt, const uint8_t* p, int stride, int x, int y, int comp) {
    if (fmt == SC_NV12) return (float)p[(size_t)y * stride + 2 * x + comp];
    const uint8_t* q = p + (size_t)y * stride + 4 * (size_t)x + 2 * comp;
    return (float)(((unsigned)q[0] | ((unsigned)q[1] << 8)) >> 6) * 0.25f;
}

static inline uint8_t store8(float v) {
    v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
    return (uint8_t)lrintf(v); /* default rounding mode: ties to even */
}

/* src: luma plane + interleaved chroma plane (W/2 x H/2 pairs); dst: BGRA, dstStride bytes per row.  Returns 0, or -1 on
 * a bad argument. */
int svo_scale_convert(int fmt, int filter, const uint8_t* srcY, int strideY, const uint8_t* srcC, int strideC, int srcW, int srcH, uint8_t* dst,
                      int dstStride, int dstW, int dstH) {
    if ((fmt != SC_NV12 && fmt != SC_P010) || (filter != SC_BILINEAR && filter != SC_LANCZOS3) || srcW < 2 || srcH < 2 || (srcW & 1) || (srcH & 1) ||
        dstW < 1 || dstH < 1)
        return -1;
    const int cw = srcW / 2, ch = srcH / 2;
    const int nyx = svo_scale_taps(filter, srcW, dstW), nyy = svo_scale_taps(filter, srcH, dstH);
    const int ncx = svo_scale_taps(filter, cw, dstW), ncy = svo_scale_taps(filter, ch, dstH);
    int32_t* fyx = malloc(sizeof(int32_t) * dstW); float* wyx = malloc(sizeof(float) * (size_t)dstW * nyx);
    int32_t* fyy = malloc(sizeof(int32_t) * dstH); float* wyy = malloc(sizeof(float) * (size_t)dstH * nyy);
    int32_t* fcx = malloc(sizeof(int32_t) * dstW); float* wcx = malloc(sizeof(float) * (size_t)dstW * ncx);
    int32_t* fcy = malloc(sizeof(int32_t) * dstH); float* wcy = malloc(sizeof(float) * (size_t)dstH * ncy);
    svo_scale_table(filter, srcW, dstW, fyx, wyx);
    svo_scale_table(filter, srcH, dstH, fyy, wyy);
    svo_scale_table(filter, cw, dstW, fcx, wcx);
    svo_scale_table(filter, ch, dstH, fcy, wcy);
    /* horizontal pass over every source row */
    float* hy = malloc(sizeof(float) * (size_t)srcH * dstW);
    float* hu = malloc(sizeof(float) * (size_t)ch * dstW);
    float* hv = malloc(sizeof(float) * (size_t)ch * dstW);
    for (int y = 0; y < srcH; ++y)
        for (int x = 0; x < dstW; ++x) {
            float acc = 0.f;
            for (int k = 0; k < nyx; ++k) acc = fmaf(wyx[(size_t)x * nyx + k], sample_y(fmt, srcY, strideY, clampi(fyx[x] + k, 0, srcW - 1), y), acc);
            hy[(size_t)y * dstW + x] = acc;
        }
    for (int y = 0; y < ch; ++y)
        for (int x = 0; x < dstW; ++x) {
            float au = 0.f, av = 0.f;
            for (int k = 0; k < ncx; ++k) {
                const int sx = clampi(fcx[x] + k, 0, cw - 1);
                au = fmaf(wcx[(size_t)x * ncx + k], sample_c(fmt, srcC, strideC, sx, y, 0), au);
                av = fmaf(wcx[(size_t)x * ncx + k], sample_c(fmt, srcC, strideC, sx, y, 1), av);
            }
            hu[(size_t)y * dstW + x] = au;
            hv[(size_t)y * dstW + x] = av;
        }
    /* vertical pass, colour conversion, store */
    for (int y = 0; y < dstH; ++y)
        for (int x = 0; x < dstW; ++x) {
            float Y = 0.f, U = 0.f, V = 0.f;
            for (int k = 0; k < nyy; ++k) Y = fmaf(wyy[(size_t)y * nyy + k], hy[(size_t)clampi(fyy[y] + k, 0, srcH - 1) * dstW + x], Y);
            for (int k = 0; k < ncy; ++k) {
                const size_t row = (size_t)clampi(fcy[y] + k, 0, ch - 1) * dstW + x;
                U = fmaf(wcy[(size_t)y * ncy + k], hu[row], U);
                V = fmaf(wcy[(size_t)y * ncy + k], hv[row], V);
            }
            const float yy = 1.164383f * (Y - 16.f), du = U - 128.f, dv = V - 128.f;
            const float R = fmaf(1.596027f, dv, yy);
            const float G = fmaf(-0.812968f, dv, fmaf(-0.391762f, du, yy));
            const float B = fmaf(2.017232f, du, yy);
            uint8_t* o = dst + (size_t)y * dstStride + 4 * (size_t)x;
            o[0] = store8(B), o[1] = store8(G), o[2] = store8(R), o[3] = 255;
        }
    free(fyx), free(wyx), free(fyy), free(wyy), free(fcx), free(wcx), free(fcy), free(wcy), free(hy), free(hu), free(hv);
    return 0;
}
