// kernels_fused.cuh -- one launch composites whole frames: clear + every layer in z-order, the running
// picture held in registers and re-quantised to 8 bits between layers exactly as the reference's
// clear-then-fold over an 8-bit target does (mix.video.swift:113-125; one RMW kernel per layer there).
//
// svb_mix_generic: any transform (rotation included), any source format.  One thread owns a 2x2 luma quad
// and its chroma texel, which is the reference's own ownership rule (chroma is produced by the even/even
// work-item, kernels.cl.swift:76,89-92).  It is the fallback of the tiled fast path in kernels_tiled.cuh.
#pragma once
#include "svb_device.cuh"

namespace svb {

__device__ __forceinline__ Src layer_src(const SvbLayerDesc* __restrict__ L) {
    Src s;
    s.format = L->format;
    s.p[0] = (const uint8_t*)L->plane[0];
    s.p[1] = (const uint8_t*)L->plane[1];
    s.p[2] = (const uint8_t*)L->plane[2];
    s.stride[0] = L->stride[0];
    s.stride[1] = L->stride[1];
    s.stride[2] = L->stride[2];
    s.w = L->width;
    s.h = L->height;
    s.cw = SVB_FORMAT_CHROMA_W(L->format, L->width);
    s.ch = SVB_FORMAT_CHROMA_H(L->format, L->height);
    return s;
}

}  // namespace svb

#define SVB_GEN_BX 32
#define SVB_GEN_BY 8

extern "C" __global__ void __launch_bounds__(SVB_GEN_BX* SVB_GEN_BY)
    svb_mix_generic(const SvbFrameDesc* __restrict__ frames) {
    using namespace svb;
    const SvbFrameDesc* __restrict__ F = frames + blockIdx.z;
    const int W = F->width, H = F->height;
    const int x = 2 * (blockIdx.x * SVB_GEN_BX + threadIdx.x), y = 2 * (blockIdx.y * SVB_GEN_BY + threadIdx.y);
    if (x >= W || y >= H) return;
    const bool nv12 = F->format == SVB_NV12;
    uint8_t* const oY = (uint8_t*)F->out_plane[0] + (size_t)y * F->out_stride[0] + x;
    uint8_t* const oU = (uint8_t*)F->out_plane[1] + (size_t)(y >> 1) * F->out_stride[1] + (nv12 ? x : (x >> 1));
    uint8_t* const oV = nv12 ? oU + 1 : (uint8_t*)F->out_plane[2] + (size_t)(y >> 1) * F->out_stride[2] + (x >> 1);

    unsigned Y[4] = {0, 0, 0, 0}, Cu = 128, Cv = 128;  // img_clear_*: Y=0, chroma=0.5 -> 128
    if (F->flags & SVB_FRAME_LOAD_CUR) {
        const uchar2 r0 = *(const uchar2*)oY, r1 = *(const uchar2*)(oY + F->out_stride[0]);
        Y[0] = r0.x, Y[1] = r0.y, Y[2] = r1.x, Y[3] = r1.y;
        Cu = *oU, Cv = *oV;
    }
    const float fW = (float)W, fH = (float)H;
    for (int l = 0; l < F->nlayers; ++l) {
        const SvbLayerDesc* __restrict__ L = &F->layers[l];
        const Src s = layer_src(L);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float oy, ou, ov;
            const bool chroma = q == 0;
            if (eval_pixel(&L->u, s, x + (q & 1), y + (q >> 1), fW, fH, chroma, unorm(opaque(Y[q])), unorm(opaque(Cu)), unorm(opaque(Cv)), oy, ou, ov)) {
                Y[q] = rte8(oy);
                if (chroma) {
                    Cu = rte8(ou);
                    Cv = rte8(ov);
                }
            }
        }
    }
    *(uchar2*)oY = make_uchar2(Y[0], Y[1]);
    *(uchar2*)(oY + F->out_stride[0]) = make_uchar2(Y[2], Y[3]);
    if (nv12) {
        *(uchar2*)oU = make_uchar2(Cu, Cv);
    } else {
        *oU = (uint8_t)Cu;
        *oV = (uint8_t)Cv;
    }
}
