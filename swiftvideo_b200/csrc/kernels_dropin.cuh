// kernels_dropin.cuh -- per-layer kernels with the reference's CUDA kernel ABI, so the module this file is
// compiled into can be handed to the reference's own runComputeKernel unchanged:
//   /root/reference/Sources/SwiftVideo/compute.cuda.swift:294-303
//     params  = [out planes..., in planes..., ImageUniforms*, int32 inStride[]]   (each a CUdeviceptr)
//     block   = (gcd(W,16), gcd(H,16), 1), grid = (W/bx, H/by, 1), 0 B smem, NULL stream
//   entry-point names = ComputeKernel case names (compute.swift:49-63; lookup compute.cuda.swift:204-209,226)
// Output strides are not passed: like the reference kernels (kernels.cuda.swift:151,205) they are inferred
// from the launch, W = gridDim.x*blockDim.x; chroma stride is W for NV12 (sample.pict.linux.swift:280)
// and W/2 for Y420P (:287-288).  Input plane sizes come from uniforms->inSize (kernels.cuda.swift:152).
//
// Unlike the reference's three CUDA kernels these reproduce the OpenCL results (SURVEY.md section 2.3 lists the
// CUDA bugs we do not copy) and all ten kernels exist.
#pragma once
#include "svb_device.cuh"

namespace svb {

template <int DST>
__device__ __forceinline__ void dropin_blend(uint8_t* __restrict__ o0, uint8_t* __restrict__ o1, uint8_t* __restrict__ o2,
                                             const Src& s, const ImageUniforms* __restrict__ U) {
    const int W = gridDim.x * blockDim.x, H = gridDim.y * blockDim.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const bool chroma = ((x | y) & 1) == 0;
    uint8_t* py = o0 + (size_t)y * W + x;
    uint8_t *pu = nullptr, *pv = nullptr;
    float cu = 0.f, cv = 0.f;
    if (chroma) {
        if (DST == SVB_NV12) {
            pu = o1 + (size_t)(y >> 1) * W + (x >> 1) * 2;
            pv = pu + 1;
        } else {
            pu = o1 + (size_t)(y >> 1) * (W >> 1) + (x >> 1);
            pv = o2 + (size_t)(y >> 1) * (W >> 1) + (x >> 1);
        }
        cu = unorm(opaque(*pu));
        cv = unorm(opaque(*pv));
    }
    float oy, ou, ov;
    if (!eval_pixel(U, s, x, y, (float)W, (float)H, chroma, unorm(opaque(*py)), cu, cv, oy, ou, ov)) return;
    *py = (uint8_t)rte8(oy);
    if (chroma) {
        *pu = (uint8_t)rte8(ou);
        *pv = (uint8_t)rte8(ov);
    }
}

__device__ __forceinline__ Src make_src(int format, const uint8_t* p0, const uint8_t* p1, const uint8_t* p2,
                                        const ImageUniforms* U, const int* inStride) {
    Src s;
    s.format = format;
    s.p[0] = p0;
    s.p[1] = p1;
    s.p[2] = p2;
    s.w = (int)__ldg(&U->inSize[0]);
    s.h = (int)__ldg(&U->inSize[1]);
    s.cw = SVB_FORMAT_CHROMA_W(format, s.w);
    s.ch = SVB_FORMAT_CHROMA_H(format, s.h);
    s.stride[0] = __ldg(inStride);
    s.stride[1] = SVB_FORMAT_IS_YUV(format) ? __ldg(inStride + 1) : 0;
    s.stride[2] = SVB_FORMAT_IS_YUV(format) && !SVB_FORMAT_IS_SEMIPLANAR(format) ? __ldg(inStride + 2) : 0;
    return s;
}

}  // namespace svb

extern "C" {

__global__ void img_clear_nv12(uint8_t* __restrict__ oY, uint8_t* __restrict__ oC) {
    const int W = gridDim.x * blockDim.x;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    oY[(size_t)y * W + x] = 0;
    if (((x | y) & 1) == 0) *(uchar2*)(oC + (size_t)(y >> 1) * W + x) = make_uchar2(128, 128);
}

__global__ void img_clear_y420p(uint8_t* __restrict__ oY, uint8_t* __restrict__ oU, uint8_t* __restrict__ oV) {
    const int W = gridDim.x * blockDim.x;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    oY[(size_t)y * W + x] = 0;
    if (((x | y) & 1) == 0) {
        oU[(size_t)(y >> 1) * (W >> 1) + (x >> 1)] = 128;
        oV[(size_t)(y >> 1) * (W >> 1) + (x >> 1)] = 128;
    }
}

__global__ void img_clear_bgra(uint8_t* __restrict__ o) {
    const int W = gridDim.x * blockDim.x;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    ((uchar4*)o)[(size_t)y * W + x] = make_uchar4(0, 0, 0, 255);
}

__global__ void img_nv12_nv12(uint8_t* oY, uint8_t* oC, const uint8_t* iY, const uint8_t* iC, const ImageUniforms* U,
                              const int* inStride) {
    svb::dropin_blend<SVB_NV12>(oY, oC, nullptr, svb::make_src(SVB_NV12, iY, iC, nullptr, U, inStride), U);
}
__global__ void img_y420p_nv12(uint8_t* oY, uint8_t* oC, const uint8_t* iY, const uint8_t* iU, const uint8_t* iV,
                               const ImageUniforms* U, const int* inStride) {
    svb::dropin_blend<SVB_NV12>(oY, oC, nullptr, svb::make_src(SVB_Y420P, iY, iU, iV, U, inStride), U);
}
__global__ void img_bgra_nv12(uint8_t* oY, uint8_t* oC, const uint8_t* iP, const ImageUniforms* U, const int* inStride) {
    svb::dropin_blend<SVB_NV12>(oY, oC, nullptr, svb::make_src(SVB_BGRA, iP, nullptr, nullptr, U, inStride), U);
}
__global__ void img_rgba_nv12(uint8_t* oY, uint8_t* oC, const uint8_t* iP, const ImageUniforms* U, const int* inStride) {
    svb::dropin_blend<SVB_NV12>(oY, oC, nullptr, svb::make_src(SVB_RGBA, iP, nullptr, nullptr, U, inStride), U);
}
__global__ void img_y420p_y420p(uint8_t* oY, uint8_t* oU, uint8_t* oV, const uint8_t* iY, const uint8_t* iU,
                                const uint8_t* iV, const ImageUniforms* U, const int* inStride) {
    svb::dropin_blend<SVB_Y420P>(oY, oU, oV, svb::make_src(SVB_Y420P, iY, iU, iV, U, inStride), U);
}
__global__ void img_bgra_y420p(uint8_t* oY, uint8_t* oU, uint8_t* oV, const uint8_t* iP, const ImageUniforms* U,
                               const int* inStride) {
    svb::dropin_blend<SVB_Y420P>(oY, oU, oV, svb::make_src(SVB_BGRA, iP, nullptr, nullptr, U, inStride), U);
}
__global__ void img_rgba_y420p(uint8_t* oY, uint8_t* oU, uint8_t* oV, const uint8_t* iP, const ImageUniforms* U,
                               const int* inStride) {
    svb::dropin_blend<SVB_Y420P>(oY, oU, oV, svb::make_src(SVB_RGBA, iP, nullptr, nullptr, U, inStride), U);
}

// ---- operators the reference names but has no OpenCL / CUDA kernel for ------------------------------------------
// Sources: the same body as img_nv12_nv12 / img_y420p_* with the source's own plane plumbing (the sampler works on each plane's own size,
// so 4:2:2 and 4:4:4 chroma need nothing else).  Names follow findKernel's img_<source>_<target> rule (mix.video.swift:142-146).
__global__ void img_nv21_nv12(uint8_t* oY, uint8_t* oC, const uint8_t* iY, const uint8_t* iC, const ImageUniforms* U, const int* inStride) {
    svb::dropin_blend<SVB_NV12>(oY, oC, nullptr, svb::make_src(SVB_NV21, iY, iC, nullptr, U, inStride), U);
}
__global__ void img_y422p_nv12(uint8_t* oY, uint8_t* oC, const uint8_t* iY, const uint8_t* iU, const uint8_t* iV, const ImageUniforms* U, const int* inStride) {
    svb::dropin_blend<SVB_NV12>(oY, oC, nullptr, svb::make_src(SVB_Y422P, iY, iU, iV, U, inStride), U);
}
__global__ void img_y444p_nv12(uint8_t* oY, uint8_t* oC, const uint8_t* iY, const uint8_t* iU, const uint8_t* iV, const ImageUniforms* U, const int* inStride) {
    svb::dropin_blend<SVB_NV12>(oY, oC, nullptr, svb::make_src(SVB_Y444P, iY, iU, iV, U, inStride), U);
}
__global__ void img_y422p_y420p(uint8_t* oY, uint8_t* oU, uint8_t* oV, const uint8_t* iY, const uint8_t* iU, const uint8_t* iV, const ImageUniforms* U,
                                const int* inStride) {
    svb::dropin_blend<SVB_Y420P>(oY, oU, oV, svb::make_src(SVB_Y422P, iY, iU, iV, U, inStride), U);
}
__global__ void img_y444p_y420p(uint8_t* oY, uint8_t* oU, uint8_t* oV, const uint8_t* iY, const uint8_t* iU, const uint8_t* iV, const ImageUniforms* U,
                                const int* inStride) {
    svb::dropin_blend<SVB_Y420P>(oY, oU, oV, svb::make_src(SVB_Y444P, iY, iU, iV, U, inStride), U);
}

// img_bgra_bgra: the only text upstream has is the Metal body (kernels.metal:51-62, "TODO: apply transformations"): the source texel at
// trunc(gid * inputSize / outputSize) -- no filtering, no transform -- laid source-over onto the target with the source's own alpha, result
// alpha 1.  Metal reads a BGRA8Unorm texture as (r, g, b, a) = bytes (2, 1, 0, 3) / 255 and writes with round-to-nearest-even; the channel
// order cancels out (the same swizzle on the way in and out), so the arithmetic is per byte.
__global__ void img_bgra_bgra(uint8_t* o, const uint8_t* i, const ImageUniforms* U, const int* inStride) {
    using namespace svb;
    const int W = gridDim.x * blockDim.x;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const float sx = __fdiv_rn(__ldg(&U->inSize[0]), __ldg(&U->outSize[0])), sy = __fdiv_rn(__ldg(&U->inSize[1]), __ldg(&U->outSize[1]));
    const int ix = min((int)mul((float)x, sx), (int)__ldg(&U->inSize[0]) - 1), iy = min((int)mul((float)y, sy), (int)__ldg(&U->inSize[1]) - 1);
    const uchar4 in = *(const uchar4*)(i + (size_t)iy * __ldg(inStride) + 4 * ix);
    uchar4* const out = (uchar4*)o + (size_t)y * W + x;
    const uchar4 cur = *out;
    const float a = unorm(opaque(in.w)), na = sub(1.f, a);
    uchar4 r;
    r.x = (uint8_t)rte8(add(mul(unorm(opaque(in.x)), a), mul(unorm(opaque(cur.x)), na)));
    r.y = (uint8_t)rte8(add(mul(unorm(opaque(in.y)), a), mul(unorm(opaque(cur.y)), na)));
    r.z = (uint8_t)rte8(add(mul(unorm(opaque(in.z)), a), mul(unorm(opaque(cur.z)), na)));
    r.w = 255;
    *out = r;
}

// img_clear_yuvs: named in the enum (compute.swift:58), no body anywhere upstream.  'yuvs' is packed 4:2:2, two bytes per pixel, luma in the
// even bytes (componentsForPlane, sample.pict.swift:91-92: y cb y cr); cleared like the other YUV targets: Y = 0, chroma = 0.5 -> 128.
__global__ void img_clear_yuvs(uint8_t* __restrict__ o) {
    const int W = gridDim.x * blockDim.x;
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    ((uchar2*)o)[(size_t)y * W + x] = make_uchar2(0, 128);
}

}  // extern "C"
