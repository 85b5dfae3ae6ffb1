// kernels.cu -- every device entry point of the compositor, compiled to one sm_100a cubin that the host
// library embeds and loads with cuModuleLoadData (the call the reference itself makes, compute.cuda.swift:193).
//   kernels_dropin.cuh  the ten per-layer kernels under the reference's names and parameter convention
//   kernels_fused.cuh   svb_mix_generic: clear + N layers in one launch, any transform / format
//   kernels_tiled.cuh   svb_mix_tiled:   the TMA-staged tile path for separable YUV layers
//   kernels_strip.cuh   the unit-level layer bodies (taps carried from row to row, half-texel body) and svb_strip_tables, the pre-pass
//   kernels_ring.cuh    svb_mix_ring:    128x32 tiles planned and staged once per CTA, composited by eight free-running warps through a three-stage mbarrier ring
//   kernels_gather.cuh  svb_mix_gather: the same compositor with the texture unit fetching the bilinear footprints (no staging, no barriers)
//   kernels_scale.cuh   svb_scale_convert: NV12 / P010 -> BGRA with a bilinear / Lanczos-3 resize (an extension; no upstream counterpart)
#include "kernels_dropin.cuh"
#include "kernels_fused.cuh"
#include "kernels_tiled.cuh"
#include "kernels_gather.cuh"
#include "kernels_strip.cuh"
#include "kernels_ring.cuh"
#include "kernels_scale.cuh"

// 256 bytes -> 256 floats with the division-free UNORM8 read, and with a true division: the parity tests
// compare the two on the device (tests/test_gpu_parity.py::test_unorm_identity).
extern "C" __global__ void svb_selftest_unorm(float* fast, float* divided) {
    const unsigned c = threadIdx.x;
    fast[c] = svb::unorm(c);
    divided[c] = svb::unorm_div(c);
}
