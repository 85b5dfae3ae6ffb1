// ref_driver.cpp -- runs the reference's own OpenCL kernel text (compiled as C++ via cl_shim.hpp)
// the way compute.cl.swift runs it.  TEST INFRASTRUCTURE ONLY (see ../mixer_oracle.h).
//
// Argument binding follows runComputeKernel (OpenCL), /root/reference/Sources/SwiftVideo/compute.cl.swift:
//   :291-300  outputs are args 0..k-1
//   :301-314  when `blends`, the same outputs again as read-only "cur" args k..2k-1
//   :316-326  then every plane of every input image
//   :327      then the uniforms
//   :329-335  global size = target W x H
// Kernel choice follows VideoMixer.findKernel (mix.video.swift:142-146) + the name map (compute.swift:90-110).
#include <pthread.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../mixer_oracle.h"
#include "cl_shim.hpp"

thread_local int g_gid[2];
thread_local int g_gsize[2];

struct RefKernel {
    const char* name;
    void (*run)(ClImage** im, const void* u, int W, int H, int y0, int y1);
    int nimg;
    int has_uniforms;
};

#include "kernels_gen.inc"  // oracle/_ref/kernels_gen.inc (generated; -I oracle/_ref)

static_assert(sizeof(ClImage) == sizeof(svo_plane), "ClImage must mirror svo_plane");

static const RefKernel* find(const std::string& name) {
    for (const RefKernel& k : kRefKernels)
        if (name == k.name) return &k;
    return nullptr;
}

static const char* fmt_name(int f) {
    switch (f) {
    case SVO_NV12: return "nv12";
    case SVO_Y420P: return "y420p";
    case SVO_BGRA: return "bgra";
    case SVO_RGBA: return "rgba";
    default: return "invalid";
    }
}

extern "C" {

int svr_kernel_count(void) { return (int)(sizeof(kRefKernels) / sizeof(kRefKernels[0])); }
const char* svr_kernel_name(int i) { return kRefKernels[i].name; }

int svr_clear_rows(svo_image* t, int y0, int y1) {
    std::string name = std::string("img_clear_") + fmt_name(t->format);
    if (name == "img_clear_rgba") name = "img_clear_bgra";  // compute.swift:101
    const RefKernel* k = find(name);
    if (!k) return SVO_ERR_KERNEL_NOT_FOUND;
    ClImage* im[3];
    for (int i = 0; i < t->nplanes; ++i) im[i] = (ClImage*)&t->planes[i];
    if (k->nimg != t->nplanes) return SVO_ERR_BAD_TARGET;
    k->run(im, nullptr, t->width, t->height, y0, y1);
    return SVO_OK;
}
int svr_clear(svo_image* t) { return svr_clear_rows(t, 0, t->height); }

int svr_apply_rows(svo_image* t, const svo_image* s, const svo_uniforms* u, int y0, int y1) {
    const RefKernel* k = find(std::string("img_") + fmt_name(s->format) + "_" + fmt_name(t->format));
    if (!k) return SVO_ERR_KERNEL_NOT_FOUND;
    if ((t->width & 1) || (t->height & 1)) return SVO_ERR_BAD_TARGET;
    ClImage* im[9];
    int n = 0;
    for (int i = 0; i < t->nplanes; ++i) im[n++] = (ClImage*)&t->planes[i];  // outputs
    for (int i = 0; i < t->nplanes; ++i) im[n++] = (ClImage*)&t->planes[i];  // cur (blends: true)
    for (int i = 0; i < s->nplanes; ++i) im[n++] = (ClImage*)&s->planes[i];  // inputs
    if (n != k->nimg) return SVO_ERR_BAD_INPUT;
    alignas(16) unsigned char ubuf[240] = {0};  // device-side struct pads 236 -> 240
    std::memcpy(ubuf, u, sizeof(svo_uniforms));
    k->run(im, ubuf, t->width, t->height, y0, y1);
    return SVO_OK;
}
int svr_apply(svo_image* t, const svo_image* s, const svo_uniforms* u) { return svr_apply_rows(t, s, u, 0, t->height); }

// mix.video.swift:113-125 -- clear, then fold the (already z-sorted) layers
int svr_mix(svo_image* t, const svo_image* layers, const svo_uniforms* us, int n) {
    int rc = svr_clear(t);
    for (int k = 0; k < n && rc == SVO_OK; ++k) rc = svr_apply(t, &layers[k], &us[k]);
    return rc;
}

struct Band {
    svo_image* t;
    const svo_image* layers;
    const svo_uniforms* us;
    int n, y0, y1, rc;
};
static void* band_main(void* p) {
    Band* b = (Band*)p;
    b->rc = svr_clear_rows(b->t, b->y0, b->y1);
    for (int k = 0; k < b->n && b->rc == SVO_OK; ++k) b->rc = svr_apply_rows(b->t, &b->layers[k], &b->us[k], b->y0, b->y1);
    return nullptr;
}
// Same bytes as svr_mix: a work-item only touches its own luma byte and (even/even) its own chroma
// texel, so even-aligned row bands are independent across the whole fold.
int svr_mix_mt(svo_image* t, const svo_image* layers, const svo_uniforms* us, int n, int nthreads) {
    if ((t->width & 1) || (t->height & 1) || (t->format != SVO_NV12 && t->format != SVO_Y420P)) return svr_mix(t, layers, us, n);
    int pairs = t->height / 2;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > pairs) nthreads = pairs;
    std::vector<pthread_t> th(nthreads);
    std::vector<Band> bands(nthreads);
    for (int i = 0; i < nthreads; ++i) {
        bands[i] = Band{t, layers, us, n, 2 * (int)((int64_t)pairs * i / nthreads), 2 * (int)((int64_t)pairs * (i + 1) / nthreads), 0};
        pthread_create(&th[i], nullptr, band_main, &bands[i]);
    }
    int rc = SVO_OK;
    for (int i = 0; i < nthreads; ++i) {
        pthread_join(th[i], nullptr);
        if (bands[i].rc) rc = bands[i].rc;
    }
    return rc;
}

}  // extern "C"
