/*
 * mixer_oracle.c -- CPU restatement of SwiftVideo's VideoMixer compute path.
 * TEST INFRASTRUCTURE ONLY (see mixer_oracle.h for the rules and the parity pin).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fPIC -shared -pthread
 * (-ffp-contract=off is load-bearing: every mul and add must round separately).
 */
#include "mixer_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- OpenCL 1.2 builtins the kernels rely on ------------------------------------------- */

/* dot(float4,float4), summed left to right (kernels.cuda.swift:45-47 spells the order out). */
static inline float dot4(const float v[4], const float m[4]) {
    return v[0] * m[0] + v[1] * m[1] + v[2] * m[2] + v[3] * m[3];
}

/* vecmat4 (kernels.cl.swift:27): four dots against the four float4 rows. */
static inline void vecmat4(float out[4], const float v[4], const float m[16]) {
    out[0] = dot4(v, m + 0);
    out[1] = dot4(v, m + 4);
    out[2] = dot4(v, m + 8);
    out[3] = dot4(v, m + 12);
}

/* clamp(x, lo, hi) = fmin(fmax(x, lo), hi)  (OpenCL 1.2 section 6.12.4). */
static inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

/* write_imagef to CL_UNORM_INT8 = convert_uchar_sat_rte(f * 255.0f) (OpenCL 1.2 section 8.3.1.1). */
static inline uint8_t rte8(float f) {
    float v = f * 255.0f;
    if (!(v == v)) return 0; /* NaN saturates to 0 */
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)rintf(v); /* default rounding mode: nearest even */
}

/* read_imagef from CL_UNORM_INT8 (compute.cl.swift:548-558): c / 255.0f. */
static inline float rd(const svo_plane* p, int x, int y, int c) {
    return (float)p->data[(int64_t)y * p->stride + x * p->ncomp + c] / 255.0f;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* read_imagef with CLK_NORMALIZED_COORDS_TRUE | CLK_ADDRESS_CLAMP_TO_EDGE | CLK_FILTER_LINEAR
 * (kernels.cl.swift:61), OpenCL 1.2 section 8.2: all of the plane's components at once. */
static void lin(const svo_plane* p, float s, float t, float out[4]) {
    float u = s * (float)p->width;
    float v = t * (float)p->height;
    float fu = floorf(u - 0.5f), fv = floorf(v - 0.5f);
    float a = (u - 0.5f) - fu;
    float b = (v - 0.5f) - fv;
    int i0 = (int)fu, j0 = (int)fv;
    int i1 = clampi(i0 + 1, 0, p->width - 1), j1 = clampi(j0 + 1, 0, p->height - 1);
    i0 = clampi(i0, 0, p->width - 1);
    j0 = clampi(j0, 0, p->height - 1);
    for (int c = 0; c < 4; ++c) {
        if (c >= p->ncomp) {
            out[c] = (c == 3) ? 1.0f : 0.0f; /* missing channels read (0,0,0,1) */
            continue;
        }
        float t00 = rd(p, i0, j0, c), t10 = rd(p, i1, j0, c);
        float t01 = rd(p, i0, j1, c), t11 = rd(p, i1, j1, c);
        out[c] = (1.0f - a) * (1.0f - b) * t00 + a * (1.0f - b) * t10 + (1.0f - a) * b * t01 + a * b * t11;
    }
}

static const float RGB2YUV[16] = { /* kernels.cl.swift:96-99 -- 0.113 is as written upstream */
    0.299f, 0.587f, 0.113f, 0.f, -0.169f, -0.331f, 0.5f, 0.5f, 0.5f, -0.419f, -0.081f, 0.5f, 0.f, 0.f, 0.f, 1.f};

static inline int in01(float x, float y) { return x >= 0.f && y >= 0.f && x <= 1.f && y <= 1.f; }

/* ---- layouts (sample.pict.linux.swift:275-311) ----------------------------------------- */

int64_t svo_layout(svo_image* img, int32_t format, int32_t width, int32_t height, uint8_t* base) {
    if (width <= 0 || height <= 0) return SVO_ERR_BAD_INPUT;
    memset(img, 0, sizeof(*img));
    img->format = format;
    img->width = width;
    img->height = height;
    int64_t off = 0;
    switch (format) {
    case SVO_NV12:
        img->nplanes = 2;
        img->planes[0] = (svo_plane){base ? base : 0, width, height, width, 1};
        off = (int64_t)width * height;
        img->planes[1] = (svo_plane){base ? base + off : 0, width / 2, height / 2, width, 2};
        off += (int64_t)width * (height / 2);
        break;
    case SVO_Y420P:
        img->nplanes = 3;
        img->planes[0] = (svo_plane){base ? base : 0, width, height, width, 1};
        off = (int64_t)width * height;
        img->planes[1] = (svo_plane){base ? base + off : 0, width / 2, height / 2, width / 2, 1};
        off += (int64_t)(width / 2) * (height / 2);
        img->planes[2] = (svo_plane){base ? base + off : 0, width / 2, height / 2, width / 2, 1};
        off += (int64_t)(width / 2) * (height / 2);
        break;
    case SVO_BGRA:
    case SVO_RGBA:
        img->nplanes = 1;
        img->planes[0] = (svo_plane){base ? base : 0, width, height, width * 4, 4};
        off = (int64_t)width * 4 * height;
        break;
    /* ---- extensions: layouts by analogy with the cases above (componentsForPlane, sample.pict.swift:84-96) ---- */
    case SVO_NV21:
        img->nplanes = 2;
        img->planes[0] = (svo_plane){base ? base : 0, width, height, width, 1};
        off = (int64_t)width * height;
        img->planes[1] = (svo_plane){base ? base + off : 0, width / 2, height / 2, width, 2};
        off += (int64_t)width * (height / 2);
        break;
    case SVO_Y422P:
    case SVO_Y444P: {
        const int cw = format == SVO_Y444P ? width : width / 2;
        img->nplanes = 3;
        img->planes[0] = (svo_plane){base ? base : 0, width, height, width, 1};
        off = (int64_t)width * height;
        img->planes[1] = (svo_plane){base ? base + off : 0, cw, height, cw, 1};
        off += (int64_t)cw * height;
        img->planes[2] = (svo_plane){base ? base + off : 0, cw, height, cw, 1};
        off += (int64_t)cw * height;
        break;
    }
    case SVO_YUVS: /* packed 4:2:2, two bytes per pixel (sample.pict.linux.swift:283) */
        img->nplanes = 1;
        img->planes[0] = (svo_plane){base ? base : 0, width, height, width * 2, 2};
        off = (int64_t)width * 2 * height;
        break;
    default:
        return SVO_ERR_BAD_INPUT;
    }
    return off;
}

/* ---- clear (kernels.cl.swift:38-46,174-185,257-265) ------------------------------------ */

int svo_clear(svo_image* t) {
    switch (t->format) {
    case SVO_NV12:
    case SVO_Y420P:
        for (int y = 0; y < t->height; ++y) memset(t->planes[0].data + (int64_t)y * t->planes[0].stride, rte8(0.0f), t->width);
        for (int p = 1; p < t->nplanes; ++p)
            for (int y = 0; y < t->planes[p].height; ++y)
                memset(t->planes[p].data + (int64_t)y * t->planes[p].stride, rte8(0.5f),
                       (size_t)t->planes[p].width * t->planes[p].ncomp);
        return SVO_OK;
    case SVO_BGRA:
    case SVO_RGBA: /* img_clear_rgba maps to img_clear_bgra, compute.swift:101 */
        for (int y = 0; y < t->height; ++y) {
            uint8_t* row = t->planes[0].data + (int64_t)y * t->planes[0].stride;
            for (int x = 0; x < t->width; ++x) {
                row[4 * x + 0] = rte8(0.0f);
                row[4 * x + 1] = rte8(0.0f);
                row[4 * x + 2] = rte8(0.0f);
                row[4 * x + 3] = rte8(1.0f);
            }
        }
        return SVO_OK;
    case SVO_YUVS: /* EXTENSION, unpinned: img_clear_yuvs is named (compute.swift:58) but has no body anywhere; cleared like the other YUV
                      targets -- luma (even bytes: y cb y cr, sample.pict.swift:91-92) 0, chroma 0.5 */
        for (int y = 0; y < t->height; ++y) {
            uint8_t* row = t->planes[0].data + (int64_t)y * t->planes[0].stride;
            for (int x = 0; x < t->width; ++x) {
                row[2 * x + 0] = rte8(0.0f);
                row[2 * x + 1] = rte8(0.5f);
            }
        }
        return SVO_OK;
    default:
        return SVO_ERR_BAD_TARGET;
    }
}

/* ---- EXTENSION, pinned only by the text it restates: img_bgra_bgra exists upstream as a Metal body alone (kernels.metal:51-62, marked
 * "TODO: apply transformations"): nearest texel at trunc(gid * inputSize / outputSize), source-over with the source's own alpha, alpha 1
 * out.  Metal's BGRA8Unorm read/write swizzle cancels, so the arithmetic is per byte; UNORM8 conversions as everywhere else here. */
int svo_apply_bgra_bgra(svo_image* t, const svo_image* s, const svo_uniforms* un) {
    if (t->format != SVO_BGRA || s->format != SVO_BGRA) return SVO_ERR_KERNEL_NOT_FOUND;
    const float sx = un->inSize[0] / un->outSize[0], sy = un->inSize[1] / un->outSize[1];
    for (int y = 0; y < t->height; ++y)
        for (int x = 0; x < t->width; ++x) {
            int ix = (int)((float)x * sx), iy = (int)((float)y * sy);
            if (ix > s->width - 1) ix = s->width - 1;
            if (iy > s->height - 1) iy = s->height - 1;
            const uint8_t* in = s->planes[0].data + (int64_t)iy * s->planes[0].stride + 4 * ix;
            uint8_t* out = t->planes[0].data + (int64_t)y * t->planes[0].stride + 4 * x;
            const float a = (float)in[3] / 255.0f;
            for (int c = 0; c < 3; ++c) out[c] = rte8(((float)in[c] / 255.0f) * a + ((float)out[c] / 255.0f) * (1.0f - a));
            out[3] = 255;
        }
    return SVO_OK;
}

/* ---- one work-item of img_<src>_<dst> -------------------------------------------------- */

static inline void cur_chroma(const svo_image* t, int x, int y, float c[2]) {
    if (t->format == SVO_NV12) {
        c[0] = rd(&t->planes[1], x / 2, y / 2, 0);
        c[1] = rd(&t->planes[1], x / 2, y / 2, 1);
    } else {
        c[0] = rd(&t->planes[1], x / 2, y / 2, 0);
        c[1] = rd(&t->planes[2], x / 2, y / 2, 0);
    }
}

static inline void put_chroma(svo_image* t, int x, int y, float u, float v) {
    if (t->format == SVO_NV12) {
        uint8_t* p = t->planes[1].data + (int64_t)(y / 2) * t->planes[1].stride + (x / 2) * 2;
        p[0] = rte8(u);
        p[1] = rte8(v);
    } else {
        t->planes[1].data[(int64_t)(y / 2) * t->planes[1].stride + x / 2] = rte8(u);
        t->planes[2].data[(int64_t)(y / 2) * t->planes[2].stride + x / 2] = rte8(v);
    }
}

static inline void put_luma(svo_image* t, int x, int y, float v) {
    t->planes[0].data[(int64_t)y * t->planes[0].stride + x] = rte8(v);
}

static void work_item(svo_image* t, const svo_image* s, const svo_uniforms* un, int x, int y) {
    /* kernels.cl.swift:70-76 (identical prologue in all eight blend kernels) */
    float out_uv[2] = {(float)x / (float)t->width, (float)y / (float)t->height};
    float normpos[4] = {out_uv[0] * 2.f - 1.f, out_uv[1] * 2.f - 1.f, 0.f, 1.f};
    float tx[4], border[4], uv[4];
    vecmat4(tx, normpos, un->transform);
    vecmat4(border, normpos, un->borderMatrix);
    int chroma = (x % 2) == 0 && (y % 2) == 0;
    if (!in01(border[0], border[1])) return; /* :77 */
    vecmat4(uv, tx, un->textureTx);          /* :78 */
    float curY = rd(&t->planes[0], x, y, 0); /* :79 */
    float curC[2] = {0.f, 0.f};
    if (chroma) cur_chroma(t, x, y, curC); /* :80-83 */

    if (s->format == SVO_NV12 || s->format == SVO_Y420P || s->format == SVO_NV21 || s->format == SVO_Y422P || s->format == SVO_Y444P) {
        /* img_nv12_nv12 :84-105, img_y420p_nv12 :149-170, img_y420p_y420p :228-252 (and, as extensions, the same body over NV21 / 4:2:2 / 4:4:4 planes) */
        if (in01(tx[0], tx[1]) && in01(uv[0], uv[1])) {
            float luma[4], alpha = un->opacity;
            lin(&s->planes[0], uv[0], uv[1], luma);
            put_luma(t, x, y, curY * (1.f - alpha) + luma[0] * alpha);
            if (chroma) {
                float cb, cr, tmp[4];
                if (s->format == SVO_NV12) {
                    lin(&s->planes[1], uv[0], uv[1], tmp);
                    cb = tmp[0];
                    cr = tmp[1];
                } else if (s->format == SVO_NV21) { /* extension: (Cr, Cb) pairs */
                    lin(&s->planes[1], uv[0], uv[1], tmp);
                    cb = tmp[1];
                    cr = tmp[0];
                } else {
                    lin(&s->planes[1], uv[0], uv[1], tmp);
                    cb = tmp[0];
                    lin(&s->planes[2], uv[0], uv[1], tmp);
                    cr = tmp[0];
                }
                put_chroma(t, x, y, curC[0] * (1.f - alpha) + cb * alpha, curC[1] * (1.f - alpha) + cr * alpha);
            }
            return;
        }
        float fc[4] = {un->fillColor[0], un->fillColor[1], un->fillColor[2], 1.0f}, fill[4];
        vecmat4(fill, fc, RGB2YUV);
        float alpha = un->opacity * un->fillColor[3];
        put_luma(t, x, y, clampf(curY * (1.f - alpha) + fill[0] * alpha, 0.f, 1.f));
        if (chroma)
            put_chroma(t, x, y, clampf(curC[0] * (1.f - alpha) + fill[1] * alpha, -1.f, 1.f),
                       clampf(curC[1] * (1.f - alpha) + fill[2] * alpha, -1.f, 1.f));
        return;
    }

    /* img_{bgra,rgba}_{nv12,y420p}: kernels.cl.swift:509-529 / 445-464 / 308-332 / 376-400 */
    if (!in01(tx[0], tx[1])) return;
    float alpha = un->opacity * un->fillColor[3];
    float fc[4] = {un->fillColor[0] * alpha, un->fillColor[1] * alpha, un->fillColor[2] * alpha, 1.0f}, fill[4];
    vecmat4(fill, fc, RGB2YUV);
    float r0 = curY * (1.f - alpha) + fill[0] * alpha;
    float r1 = clampf(curC[0] * (1.f - alpha) + fill[1] * alpha, -1.f, 1.f);
    float r2 = clampf(curC[1] * (1.f - alpha) + fill[2] * alpha, -1.f, 1.f);
    if (in01(uv[0], uv[1])) {
        float px[4];
        lin(&s->planes[0], uv[0], uv[1], px);
        float rgba[4] = {px[0], px[1], px[2], px[3]};
        if (s->format == SVO_BGRA) { /* bytes are B,G,R,A read as x,y,z,w; swizzle (z,y,x,w) */
            rgba[0] = px[2];
            rgba[2] = px[0];
        }
        float a2 = rgba[3] * un->opacity;
        float pm[4] = {rgba[0] * a2, rgba[1] * a2, rgba[2] * a2, 1.0f}, yuv[4];
        vecmat4(yuv, pm, RGB2YUV);
        r0 = r0 * (1.f - a2) + yuv[0] * a2;
        r1 = r1 * (1.f - a2) + yuv[1] * a2;
        r2 = r2 * (1.f - a2) + yuv[2] * a2;
    }
    put_luma(t, x, y, r0);
    if (chroma) put_chroma(t, x, y, r1, r2);
}

static int check_pair(const svo_image* t, const svo_image* s) {
    if (t->format != SVO_NV12 && t->format != SVO_Y420P)
        return SVO_ERR_KERNEL_NOT_FOUND; /* no img_*_bgra kernel exists on Linux (compute.swift:54) */
    if (s->format == SVO_NV12 && t->format == SVO_Y420P)
        return SVO_ERR_KERNEL_NOT_FOUND; /* img_nv12_y420p is not in the enum (compute.swift:49-63) */
    if (s->format == SVO_NV21 && t->format == SVO_Y420P) return SVO_ERR_KERNEL_NOT_FOUND; /* extension: like img_nv12_y420p, not offered */
    if (s->format < SVO_NV12 || s->format > SVO_Y444P) return SVO_ERR_KERNEL_NOT_FOUND;
    if ((t->width & 1) || (t->height & 1)) return SVO_ERR_BAD_TARGET;
    return SVO_OK;
}

int svo_apply_rows(svo_image* t, const svo_image* s, const svo_uniforms* u, int y0, int y1) {
    int rc = check_pair(t, s);
    if (rc) return rc;
    for (int y = y0; y < y1; ++y)
        for (int x = 0; x < t->width; ++x) work_item(t, s, u, x, y);
    return SVO_OK;
}

int svo_apply(svo_image* t, const svo_image* s, const svo_uniforms* u) {
    return svo_apply_rows(t, s, u, 0, t->height);
}

int svo_mix(svo_image* t, const svo_image* layers, const svo_uniforms* us, int n) {
    int rc = svo_clear(t);
    for (int k = 0; k < n && rc == SVO_OK; ++k) rc = svo_apply(t, &layers[k], &us[k]);
    return rc;
}

/* ---- row-parallel variant: a work-item touches only its own luma byte and, on even/even,
 * its own chroma texel, so even-aligned row bands never interact across layers. ---------- */

typedef struct band_job {
    svo_image* t;
    const svo_image* layers;
    const svo_uniforms* us;
    int n, y0, y1, rc;
} band_job;

static void* band_main(void* p) {
    band_job* j = (band_job*)p;
    svo_image* t = j->t;
    /* clear this band */
    for (int y = j->y0; y < j->y1; ++y) memset(t->planes[0].data + (int64_t)y * t->planes[0].stride, 0, t->width);
    for (int pl = 1; pl < t->nplanes; ++pl)
        for (int y = j->y0 / 2; y < j->y1 / 2; ++y)
            memset(t->planes[pl].data + (int64_t)y * t->planes[pl].stride, 128,
                   (size_t)t->planes[pl].width * t->planes[pl].ncomp);
    for (int k = 0; k < j->n && j->rc == SVO_OK; ++k) j->rc = svo_apply_rows(t, &j->layers[k], &j->us[k], j->y0, j->y1);
    return 0;
}

int svo_mix_mt(svo_image* t, const svo_image* layers, const svo_uniforms* us, int n, int nthreads) {
    if (t->format != SVO_NV12 && t->format != SVO_Y420P) return svo_mix(t, layers, us, n);
    if ((t->width & 1) || (t->height & 1)) return SVO_ERR_BAD_TARGET;
    if (nthreads < 1) nthreads = 1;
    int pairs = t->height / 2;
    if (nthreads > pairs) nthreads = pairs;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    band_job* jobs = (band_job*)malloc(sizeof(band_job) * (size_t)nthreads);
    for (int i = 0; i < nthreads; ++i) {
        jobs[i] = (band_job){t, layers, us, n, 2 * (int)((int64_t)pairs * i / nthreads),
                             2 * (int)((int64_t)pairs * (i + 1) / nthreads), SVO_OK};
        pthread_create(&th[i], 0, band_main, &jobs[i]);
    }
    int rc = SVO_OK;
    for (int i = 0; i < nthreads; ++i) {
        pthread_join(th[i], 0);
        if (jobs[i].rc) rc = jobs[i].rc;
    }
    free(th);
    free(jobs);
    return rc;
}
