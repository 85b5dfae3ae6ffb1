"""-m gpu: the drop-in boundary proved WITHOUT this repo's host side.  The only artefact used is svb200_kernels.cubin; everything else is
the CUDA driver API through ctypes, called exactly as the reference's runComputeKernel calls it (compute.cuda.swift:260-306):
cuModuleLoadData -> cuModuleGetFunction(<ComputeKernel case name>) -> per call: cuMemAlloc + cuMemcpyHtoD of the 236-byte ImageUniforms and of
the int32 stride array, a parameter list of pointers to CUdeviceptr in the order [out planes..., in planes..., uniforms, inStride], a launch
of (W / gcd(W,16)) x (H / gcd(H,16)) blocks of gcd(W,16) x gcd(H,16) threads with no shared memory on the NULL stream, cuCtxSynchronize.
The bytes must equal the oracle's clear-then-fold."""
import ctypes as C
import math
from pathlib import Path

import numpy as np
import pytest

import scenes
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


class Driver:
    def __init__(self):
        self.cu = C.CDLL("libcuda.so.1")
        self.ck(self.cu.cuInit(0), "cuInit")                                       # compute.cuda.swift:97
        dev = C.c_int()
        self.ck(self.cu.cuDeviceGet(C.byref(dev), 0), "cuDeviceGet")               # :138
        self.ctx = C.c_void_p()
        self.ck(self.cu.cuDevicePrimaryCtxRetain(C.byref(self.ctx), dev), "cuDevicePrimaryCtxRetain")
        self.ck(self.cu.cuCtxPushCurrent_v2(self.ctx), "cuCtxPushCurrent")         # :309
        self.dev = dev
        self.module = C.c_void_p()
        image = (ROOT / "swiftvideo_b200" / "svb200_kernels.cubin").read_bytes()
        self.ck(self.cu.cuModuleLoadData(C.byref(self.module), image), "cuModuleLoadData")  # :193

    def ck(self, rc, what):
        assert rc == 0, f"{what}: CUresult {rc}"

    def function(self, name):
        f = C.c_void_p()
        self.ck(self.cu.cuModuleGetFunction(C.byref(f), self.module, name.encode()), f"cuModuleGetFunction({name})")  # :194
        return f

    def upload(self, data):
        """cuMemAlloc + synchronous cuMemcpyHtoD, as uploadComputeBuffer does (:330-342, :404-410)"""
        buf = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        p = C.c_uint64()
        self.ck(self.cu.cuMemAlloc_v2(C.byref(p), C.c_size_t(max(buf.size, 1))), "cuMemAlloc")
        self.ck(self.cu.cuMemcpyHtoD_v2(p, buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size)), "cuMemcpyHtoD")
        return p

    def download(self, p, n):
        out = np.zeros(n, dtype=np.uint8)
        self.ck(self.cu.cuMemcpyDtoH_v2(out.ctypes.data_as(C.c_void_p), p, C.c_size_t(n)), "cuMemcpyDtoH")  # :353
        return out

    def launch(self, fn, W, H, buffers):
        """runComputeKernel's launch (:290-303): every parameter is a pointer to the CUdeviceptr"""
        bx, by = math.gcd(W, 16), math.gcd(H, 16)
        params = (C.c_void_p * len(buffers))(*[C.cast(C.pointer(b), C.c_void_p) for b in buffers])
        self.ck(self.cu.cuLaunchKernel(fn, W // bx, H // by, 1, bx, by, 1, 0, None, params, None), "cuLaunchKernel")
        self.ck(self.cu.cuCtxSynchronize(), "cuCtxSynchronize")                    # :316

    def free(self, *ptrs):
        for p in ptrs:
            self.cu.cuMemFree_v2(p)

    def close(self):
        self.cu.cuModuleUnload(self.module)
        popped = C.c_void_p()
        self.cu.cuCtxPopCurrent_v2(C.byref(popped))
        self.cu.cuDevicePrimaryCtxRelease_v2(self.dev)


def _planes(img):
    return [np.ascontiguousarray(img.data[off:off + stride * h]) for off, w, h, stride, nc in img.layout]


@pytest.mark.parametrize("src_fmt,dst_fmt", [(O.NV12, O.NV12), (O.Y420P, O.Y420P), (O.Y420P, O.NV12), (O.BGRA, O.NV12), (O.RGBA, O.Y420P)],
                         ids=["nv12_nv12", "y420p_y420p", "y420p_nv12", "bgra_nv12", "rgba_y420p"])
def test_reference_driver_sequence_against_the_module(src_fmt, dst_fmt):
    drv = Driver()
    try:
        W, H = 320, 180  # 16 x 4 threads per block
        canvas = (W, H)
        layers = [scenes.random_image(src_fmt, 320, 180, 501), scenes.random_image(src_fmt, 200, 120, 502)]
        us = [scenes.layer_uniforms(canvas, (320, 180), (0, 0), canvas, z=1, opacity=1.0),
              scenes.layer_uniforms(canvas, (200, 120), (37, 21), (240, 130), z=2, opacity=0.6, fill=(0.2, 0.6, 0.4, 0.7), border=(3, 3, 3, 3))]
        want = O.Image(dst_fmt, W, H)
        assert O.best()[0].mix(want, layers, us) == 0
        names = {O.NV12: "nv12", O.Y420P: "y420p", O.BGRA: "bgra", O.RGBA: "rgba"}
        target = O.Image(dst_fmt, W, H)
        target.data[:] = 0xA5
        out = [drv.upload(p) for p in _planes(target)]
        drv.launch(drv.function(f"img_clear_{names[dst_fmt]}"), W, H, out)      # mix.video.swift:118: clear kernels take the outputs only
        fn = drv.function(f"img_{names[src_fmt]}_{names[dst_fmt]}")
        for img, u in zip(layers, us):                                           # :119-124: one applyComputeImage per layer, in z order
            ins = [drv.upload(p) for p in _planes(img)]
            ub = drv.upload(np.frombuffer(bytes(u), dtype=np.uint8)[:236])       # MemoryLayout<ImageUniforms>.size == 236
            sb = drv.upload(np.array([l[3] for l in img.layout], dtype=np.int32))
            drv.launch(fn, W, H, out + ins + [ub, sb])
            drv.free(*ins, ub, sb)
        got = np.concatenate([drv.download(p, l[3] * l[2]) for p, l in zip(out, target.layout)])
        drv.free(*out)
        d = np.nonzero(got != want.data)[0]
        assert d.size == 0, f"{d.size} bytes differ, first at {d[0]}: got {got[d[0]]} want {want.data[d[0]]}"
    finally:
        drv.close()
