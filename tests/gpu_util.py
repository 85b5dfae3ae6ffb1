"""Helpers shared by the -m gpu tests: move oracle images onto the device and back through the C ABI."""
import numpy as np

import swiftvideo_b200 as sv
from oracle import oracle as O

FMT = {O.NV12: sv.NV12, O.Y420P: sv.Y420P, O.BGRA: sv.BGRA, O.RGBA: sv.RGBA, O.NV21: sv.NV21, O.Y422P: sv.Y422P, O.Y444P: sv.Y444P, O.YUVS: sv.YUVS}

_ctx = None


def context():
    global _ctx
    if _ctx is None:
        _ctx = sv.make_compute_context(0)
    return _ctx


def to_gpu(ctx, img, asset="layer", pinned=False):
    """O.Image -> uploaded PictureSample (same plane layout and strides on both sides)."""
    default, _ = O.plane_layout(img.format, img.width, img.height)
    if [l[3] for l in img.layout] != [l[3] for l in default]:  # decoder-style padded rows
        from swiftvideo_b200 import api
        return api.picture_sample_from_planes(img.width, img.height, FMT[img.format], img.padded_planes(), asset, "test").upload(ctx)
    p = sv.create_picture_sample(img.width, img.height, FMT[img.format], asset, "test", pinned_from=ctx if pinned else None)
    p.set_host_bytes(img.data)
    return p.upload(ctx)


def gpu_target(ctx, fmt, w, h, asset="target"):
    p = sv.create_picture_sample(w, h, FMT[fmt], asset, "test")
    p.set_host_bytes(np.full(O.Image(fmt, w, h).nbytes, 0xA5, dtype=np.uint8))  # stale bytes the clear must overwrite
    return p.upload(ctx)


def fetch(ctx, pict):
    return pict.download(ctx, retain_gpu_buffer=True).host_bytes().copy()


def gpu_case(ctx, case, mode, target_strides=None):
    """Run a scenes.Case through svb_compose with the oracle's own uniforms; returns the target bytes (padding included
    when the target has padded rows)."""
    layers = [to_gpu(ctx, l, f"l{i}") for i, l in enumerate(case.layers)]
    if target_strides is None:
        target = gpu_target(ctx, case.target_fmt, case.canvas[0], case.canvas[1])
    else:
        t = O.Image(case.target_fmt, case.canvas[0], case.canvas[1], strides=target_strides)
        t.data[:] = 0xA5
        target = to_gpu(ctx, t, "target")
    sv.compose(ctx, target, layers, case.uniforms, mode)
    return fetch(ctx, target)


def first_diff(a, b, case=None):
    d = np.nonzero(a != b)[0]
    if d.size == 0:
        return "identical"
    return f"{d.size} bytes differ, first at {d[0]}: got {a[d[0]]} want {b[d[0]]}"
