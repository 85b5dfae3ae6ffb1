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
    s.cw = s.w / 2;
    s.ch = s.h / 2;
    s.stride[0] = __ldg(inStride);
    s.stride[1] = format == SVB_NV12 || format == SVB_Y420P ? __ldg(inStride + 1) : 0;
    s.stride[2] = format == SVB_Y420P ? __ldg(inStride + 2) : 0;
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

}  // extern "C"
