"""ctypes binding of the CPU checker (oracle/liboracle.so and oracle/_ref/libsvref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent

NV12, Y420P, BGRA, RGBA = 0, 1, 2, 3
# extensions (parity unpinned, mixer_oracle.h): sources the reference names without a kernel, and the packed 4:2:2 clear target
NV21, Y422P, Y444P, YUVS = 4, 5, 6, 7
FORMAT_NAMES = {NV12: "nv12", Y420P: "y420p", BGRA: "bgra", RGBA: "rgba", NV21: "nv21", Y422P: "y422p", Y444P: "y444p", YUVS: "yuvs"}
OK, ERR_KERNEL_NOT_FOUND, ERR_BAD_TARGET, ERR_BAD_INPUT = 0, -1, -2, -3


class Uniforms(C.Structure):
    """ImageUniforms, 236 bytes (reference compute.swift:76-86)."""

    _fields_ = [
        ("transform", C.c_float * 16),
        ("textureTx", C.c_float * 16),
        ("borderMatrix", C.c_float * 16),
        ("fillColor", C.c_float * 4),
        ("inSize", C.c_float * 2),
        ("outSize", C.c_float * 2),
        ("opacity", C.c_float),
        ("sampleTime", C.c_float),
        ("targetTime", C.c_float),
    ]


assert C.sizeof(Uniforms) == 236


class _Plane(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("stride", C.c_int32), ("ncomp", C.c_int32)]


class _Image(C.Structure):
    _fields_ = [("format", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("nplanes", C.c_int32), ("planes", _Plane * 3)]


def plane_layout(fmt, w, h):
    """[(offset, width, height, stride, ncomp)] as planesForFormat lays a contiguous sample out
    (reference sample.pict.linux.swift:275-311)."""
    if fmt == NV12:
        return [(0, w, h, w, 1), (w * h, w // 2, h // 2, w, 2)], w * h + w * (h // 2)
    if fmt == Y420P:
        c = (w // 2) * (h // 2)
        return [(0, w, h, w, 1), (w * h, w // 2, h // 2, w // 2, 1), (w * h + c, w // 2, h // 2, w // 2, 1)], w * h + 2 * c
    if fmt in (BGRA, RGBA):
        return [(0, w, h, 4 * w, 4)], 4 * w * h
    if fmt == NV21:
        return [(0, w, h, w, 1), (w * h, w // 2, h // 2, w, 2)], w * h + w * (h // 2)
    if fmt in (Y422P, Y444P):
        cw = w if fmt == Y444P else w // 2
        return [(0, w, h, w, 1), (w * h, cw, h, cw, 1), (w * h + cw * h, cw, h, cw, 1)], w * h + 2 * cw * h
    if fmt == YUVS:
        return [(0, w, h, 2 * w, 2)], 2 * w * h
    raise ValueError(fmt)


class Image:
    """A contiguous 8-bit picture in the reference's Linux plane layout."""

    def __init__(self, fmt, width, height, data=None, strides=None):
        """strides: optional per-plane row strides in bytes (decoder-style padded rows); default = the reference's layout."""
        self.format, self.width, self.height = fmt, width, height
        self.layout, self.nbytes = plane_layout(fmt, width, height)
        if strides is not None:
            off, lay = 0, []
            for (_, w, h, st, nc), s2 in zip(self.layout, strides):
                assert s2 >= w * nc
                lay.append((off, w, h, int(s2), nc))
                off += int(s2) * h
            self.layout, self.nbytes = lay, off
        if data is None:
            data = np.zeros(self.nbytes, dtype=np.uint8)
        data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        assert data.size == self.nbytes, (data.size, self.nbytes)
        self.data = data

    def copy(self):
        return Image(self.format, self.width, self.height, self.data.copy(), [l[3] for l in self.layout])

    def padded_planes(self):
        """Each plane as a (rows x stride) array, padding included."""
        return [self.data[off : off + stride * h].reshape(h, stride) for off, w, h, stride, nc in self.layout]

    def plane(self, i):
        off, w, h, stride, nc = self.layout[i]
        return self.data[off : off + stride * h].reshape(h, stride)[:, : w * nc]

    def _c(self):
        im = _Image()
        im.format, im.width, im.height, im.nplanes = self.format, self.width, self.height, len(self.layout)
        base = self.data.ctypes.data
        for i, (off, w, h, stride, nc) in enumerate(self.layout):
            im.planes[i] = _Plane(base + off, w, h, stride, nc)
        return im


def build(ref_root="/root/reference"):
    """Compile liboracle.so (always) and _ref/libsvref.so (when the reference tree is present)."""
    subprocess.run(["make", "-C", str(HERE), f"REF={ref_root}"], check=True, capture_output=True)


class _Lib:
    def __init__(self, path, prefix):
        self.path = str(path)
        self.lib = C.CDLL(self.path)
        self.prefix = prefix
        for name in ("clear", "apply", "mix", "mix_mt"):
            getattr(self.lib, f"{prefix}_{name}").restype = C.c_int

    def clear(self, target):
        t = target._c()
        return getattr(self.lib, f"{self.prefix}_clear")(C.byref(t))

    def apply(self, target, src, uniforms):
        t, s = target._c(), src._c()
        return getattr(self.lib, f"{self.prefix}_apply")(C.byref(t), C.byref(s), C.byref(uniforms))

    def apply_bgra_bgra(self, target, src, uniforms):
        """EXTENSION: img_bgra_bgra after upstream's Metal text (the restatement only)."""
        fn = getattr(self.lib, f"{self.prefix}_apply_bgra_bgra")
        fn.restype = C.c_int
        t, s = target._c(), src._c()
        return fn(C.byref(t), C.byref(s), C.byref(uniforms))

    def mix(self, target, layers, uniforms, threads=0):
        """clear + fold layers (already z-sorted) into target, in place. Returns the status code."""
        n = len(layers)
        t = target._c()
        ls = (_Image * max(n, 1))(*[l._c() for l in layers])
        us = (Uniforms * max(n, 1))(*uniforms)
        if threads and threads > 1:
            return getattr(self.lib, f"{self.prefix}_mix_mt")(C.byref(t), ls, us, n, int(threads))
        return getattr(self.lib, f"{self.prefix}_mix")(C.byref(t), ls, us, n)


_cache = {}


def port():
    """The hand-written restatement (mixer_oracle.c)."""
    if "port" not in _cache:
        p = HERE / "liboracle.so"
        if not p.exists():
            build()
        _cache["port"] = _Lib(p, "svo")
    return _cache["port"]


def ref_available():
    return (HERE / "_ref" / "libsvref.so").exists()


def ref():
    """The reference's own OpenCL kernel text compiled as C++ (oracle/_ref, built from /root/reference)."""
    if "ref" not in _cache:
        p = HERE / "_ref" / "libsvref.so"
        if not p.exists():
            raise FileNotFoundError(f"{p} missing: run `make -C oracle` where /root/reference is mounted")
        _cache["ref"] = _Lib(p, "svr")
    return _cache["ref"]


def best():
    """(lib, kind): the reference-text build when present, else the port."""
    return (ref(), "reference") if ref_available() else (port(), "port")


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---- convert + scale operator (scale_oracle.c): NV12 / P010 -> BGRA, bilinear or Lanczos-3 ----------------------------
SC_NV12, SC_P010 = 0, 1
SC_BILINEAR, SC_LANCZOS3 = 0, 1


def _scale_lib():
    lib = port().lib
    if not getattr(lib, "_scale_ready", False):
        lib.svo_scale_taps.restype = C.c_int
        lib.svo_scale_taps.argtypes = [C.c_int, C.c_int, C.c_int]
        lib.svo_scale_table.restype = None
        lib.svo_scale_table.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.svo_scale_convert.restype = C.c_int
        lib.svo_scale_convert.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib._scale_ready = True
    return lib


def scale_table(filt, src_n, dst_n):
    """(first[dst_n] int32, weights[dst_n, taps] float32) of one axis."""
    lib = _scale_lib()
    n = lib.svo_scale_taps(filt, src_n, dst_n)
    first = np.zeros(dst_n, dtype=np.int32)
    w = np.zeros((dst_n, n), dtype=np.float32)
    lib.svo_scale_table(filt, src_n, dst_n, first.ctypes.data, w.ctypes.data)
    return first, w


def scale_src_layout(fmt, w, h):
    """(bytes per sample, luma stride, chroma stride, total bytes) of a contiguous NV12 / P010 picture."""
    if fmt not in (SC_NV12, SC_P010):
        raise ValueError(fmt)
    bps = 1 if fmt == SC_NV12 else 2
    return bps, w * bps, w * bps, w * h * bps + w * (h // 2) * bps


def scale_convert(fmt, filt, src, src_w, src_h, dst_w, dst_h):
    """src: contiguous uint8 buffer (luma plane, then the interleaved chroma plane).  Returns dst_h x dst_w x 4 BGRA bytes."""
    lib = _scale_lib()
    bps, sy, sc, total = scale_src_layout(fmt, src_w, src_h)
    src = np.ascontiguousarray(src, dtype=np.uint8).reshape(-1)
    assert src.size == total, (src.size, total)
    dst = np.zeros((dst_h, dst_w, 4), dtype=np.uint8)
    base = src.ctypes.data
    rc = lib.svo_scale_convert(fmt, filt, base, sy, base + sy * src_h, sc, src_w, src_h, dst.ctypes.data, dst_w * 4, dst_w, dst_h)
    if rc != 0:
        raise ValueError("svo_scale_convert: bad argument")
    return dst
