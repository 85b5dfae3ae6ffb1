"""libswscale through ctypes (the copy bundled with the OpenCV wheel): the CPU comparator BASELINE.json names for the
convert/scale operator.  Test and bench infrastructure only."""
import ctypes as C
import glob
import os

import numpy as np

SWS_BILINEAR, SWS_LANCZOS, SWS_ACCURATE_RND, SWS_FULL_CHR_H_INT = 2, 0x200, 0x40000, 0x2000
_state = {}


def load():
    if "sws" not in _state:
        import cv2
        libdir = os.path.join(os.path.dirname(cv2.__file__), "..", "opencv_python_headless.libs")
        avutil = C.CDLL(glob.glob(libdir + "/libavutil*")[0], mode=C.RTLD_GLOBAL)
        sws = C.CDLL(glob.glob(libdir + "/libswscale*")[0])
        avutil.av_get_pix_fmt.argtypes = [C.c_char_p]
        sws.sws_getContext.restype = C.c_void_p
        sws.sws_getContext.argtypes = [C.c_int] * 7 + [C.c_void_p] * 3
        sws.sws_scale.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        sws.sws_freeContext.argtypes = [C.c_void_p]
        _state["sws"], _state["avutil"] = sws, avutil
    return _state["sws"], _state["avutil"]


def available():
    try:
        load()
        return True
    except Exception:
        return False


class Scaler:
    """One SwsContext: `src_fmt` (nv12 / p010le) src_w x src_h -> bgra dst_w x dst_h."""

    def __init__(self, src_fmt, src_w, src_h, dst_w, dst_h, flags):
        sws, avutil = load()
        self.sws = sws
        self.bps = 2 if src_fmt == "p010le" else 1
        self.src_w, self.src_h, self.dst_w, self.dst_h = src_w, src_h, dst_w, dst_h
        self.ctx = sws.sws_getContext(src_w, src_h, avutil.av_get_pix_fmt(src_fmt.encode()), dst_w, dst_h, avutil.av_get_pix_fmt(b"bgra"), flags, None, None, None)
        if not self.ctx:
            raise RuntimeError("sws_getContext failed")
        self.dst = np.zeros((dst_h, dst_w, 4), np.uint8)

    def run(self, src):
        """src: contiguous uint8 buffer, luma plane then interleaved chroma plane."""
        stride = self.src_w * self.bps
        sp = (C.c_void_p * 4)(src.ctypes.data, src.ctypes.data + stride * self.src_h, None, None)
        dp = (C.c_void_p * 4)(self.dst.ctypes.data, None, None, None)
        self.sws.sws_scale(self.ctx, sp, (C.c_int * 4)(stride, stride, 0, 0), 0, self.src_h, dp, (C.c_int * 4)(self.dst_w * 4, 0, 0, 0))
        return self.dst

    def close(self):
        if self.ctx:
            self.sws.sws_freeContext(self.ctx)
            self.ctx = None


def smooth_picture(fmt_bits, w, h, seed=0):
    """A smooth synthetic picture (ramps + low-frequency waves, video range) as a contiguous NV12 (8) or P010 (10) buffer."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    ph = rng.uniform(0, 6.28, 4)
    Y = 16 + 219 * (0.5 + 0.25 * np.sin(xx / w * 6.3 + ph[0]) + 0.2 * np.cos(yy / h * 5.1 + ph[1]))
    cy, cx = np.mgrid[0:h // 2, 0:w // 2].astype(np.float64)
    U = 128 + 90 * np.sin(cx / (w / 2) * 4.0 + ph[2]) * np.cos(cy / (h / 2) * 3.0)
    V = 128 + 90 * np.cos(cx / (w / 2) * 3.1 + ph[3]) * np.sin(cy / (h / 2) * 4.4 + 1.0)
    Y, U, V = np.clip(Y, 16, 235), np.clip(U, 16, 240), np.clip(V, 16, 240)
    if fmt_bits == 8:
        uv = np.stack([U, V], axis=-1).reshape(h // 2, w)
        return np.concatenate([np.rint(Y).astype(np.uint8).reshape(-1), np.rint(uv).astype(np.uint8).reshape(-1)])
    y10 = (np.rint(Y * 4).astype(np.uint16) << 6)
    uv10 = (np.rint(np.stack([U, V], axis=-1) * 4).astype(np.uint16) << 6).reshape(h // 2, w)
    return np.concatenate([y10.reshape(-1), uv10.reshape(-1)]).view(np.uint8)
