/*
 * mixer_oracle.h -- CPU restatement of SwiftVideo's VideoMixer compute path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker.
 *
 * What is restated (reference = unpause-live/SwiftVideo @113d3d9, paths relative
 * to /root/reference):
 *   - the ten OpenCL image kernels       Sources/SwiftVideo/kernels.cl.swift:38-532
 *   - OpenCL image binding / UNORM8       Sources/SwiftVideo/compute.cl.swift:264-344,532-581
 *   - clear-then-fold layer order        Sources/SwiftVideo/mix.video.swift:113-125
 *   - Linux plane layouts                Sources/SwiftVideo/sample.pict.linux.swift:275-311
 *   - ImageUniforms (236 B)              Sources/SwiftVideo/compute.swift:76-86
 *
 * Canonical semantics chosen where OpenCL 1.2 leaves latitude:
 *   fp32 everywhere, no FMA contraction (the reference's own CUDA path passes
 *   --fmad=false, compute.cuda.swift:177), dot() summed left to right
 *   (kernels.cuda.swift:45-47), linear sampler per OpenCL 1.2 section 8.2 with
 *   fp32 weights, UNORM8 read = c/255.0f, UNORM8 write = convert_uchar_sat_rte(f*255.0f).
 *
 * PARITY PIN: the reference holds no golden frames for this path and cannot run in
 * this container (no Swift, no OpenCL runtime).  The pin is oracle/_ref: the
 * reference's OpenCL kernel TEXT, extracted at build time from kernels.cl.swift and
 * compiled as C++ against an OpenCL-1.2 image/sampler shim (oracle/ref_cl/).  This
 * restatement is checked byte-for-byte against it (tests/test_oracle_vs_ref.py) and
 * the committed fixtures under tests/golden/ were generated from it.
 */
#ifndef SVB_MIXER_ORACLE_H
#define SVB_MIXER_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Byte-for-byte the struct uploaded by applyComputeImage (compute.swift:76-86,145-170):
 * three row-major float4[4] (already inverse.transpose'd on the host), then fill colour,
 * sizes, opacity, times.  236 bytes. */
typedef struct svo_uniforms {
    float transform[16];    /* off   0 */
    float textureTx[16];    /* off  64 */
    float borderMatrix[16]; /* off 128 */
    float fillColor[4];     /* off 192 */
    float inSize[2];        /* off 208 */
    float outSize[2];       /* off 216 */
    float opacity;          /* off 224 */
    float sampleTime;       /* off 228 */
    float targetTime;       /* off 232 */
} svo_uniforms;

/* PixelFormat cases that have kernels (sample.pict.swift:20-33). */
enum { SVO_NV12 = 0, SVO_Y420P = 1, SVO_BGRA = 2, SVO_RGBA = 3,
       /* EXTENSIONS -- PARITY UNPINNED: formats the reference's PixelFormat names (sample.pict.swift:22-27) but has no
        * kernel for.  As sources they run the img_nv12_nv12 / img_y420p_* body over their own planes (NV21 = (Cr, Cb) pairs; 4:2:2 and 4:4:4
        * chroma planes sampled at their own size); YUVS exists only as a clear target (img_clear_yuvs, compute.swift:58). */
       SVO_NV21 = 4, SVO_Y422P = 5, SVO_Y444P = 6, SVO_YUVS = 7 };

/* One Plane (sample.pict.swift:46-56) plus its bytes. width/height are the plane's own
 * size (chroma planes: W/2 x H/2), stride in bytes, ncomp = components per texel. */
typedef struct svo_plane {
    uint8_t* data;
    int32_t width, height, stride, ncomp;
} svo_plane;

typedef struct svo_image {
    int32_t format;
    int32_t width, height;
    int32_t nplanes;
    svo_plane planes[3];
} svo_image;

enum {
    SVO_OK = 0,
    SVO_ERR_KERNEL_NOT_FOUND = -1, /* ComputeError.computeKernelNotFound */
    SVO_ERR_BAD_TARGET = -2,       /* ComputeError.badTarget */
    SVO_ERR_BAD_INPUT = -3         /* ComputeError.badInputData */
};

/* Fill planes[] for a contiguous allocation laid out as planesForFormat/buffersForPlanes do
 * (sample.pict.linux.swift:275-311).  Returns total bytes, or <0. base may be NULL (sizing). */
/* EXTENSION: img_bgra_bgra after the Metal text (kernels.metal:51-62); both images BGRA */
int svo_apply_bgra_bgra(svo_image* t, const svo_image* s, const svo_uniforms* u);
int64_t svo_layout(svo_image* img, int32_t format, int32_t width, int32_t height, uint8_t* base);

/* img_clear_{nv12,y420p,bgra}  (kernels.cl.swift:38-46,174-185,257-265). */
int svo_clear(svo_image* target);

/* One applyComputeImage: kernel img_<src>_<dst> in place on target
 * (kernels.cl.swift:47-532 by format pair).  Rows [y0,y1) only; y0 must be even. */
int svo_apply_rows(svo_image* target, const svo_image* src, const svo_uniforms* u, int y0, int y1);
int svo_apply(svo_image* target, const svo_image* src, const svo_uniforms* u);

/* VideoMixer.mix fold (mix.video.swift:113-125): clear, then layers in the given
 * (already z-sorted) order, each re-quantised to 8 bits in the target. */
int svo_mix(svo_image* target, const svo_image* layers, const svo_uniforms* uniforms, int nlayers);

/* Same result, rows split over nthreads host threads (CPU baseline timing). */
int svo_mix_mt(svo_image* target, const svo_image* layers, const svo_uniforms* uniforms, int nlayers,
               int nthreads);

#ifdef __cplusplus
}
#endif
#endif
