"""-m gpu: SURVEY.md 8(f-3) -- the operators the reference names but never implemented, through the C ABI, against their definitions in
oracle/mixer_oracle.c (PARITY UNPINNED: no reference kernel exists for any of them; img_bgra_bgra follows upstream's Metal text):
NV21 / Y422P / Y444P sources into NV12 / Y420P targets (fused -> svb_mix_generic, and the per-layer drop-in kernels), img_bgra_bgra with a
BGRA mixer, img_clear_yuvs."""
import ctypes as C

import numpy as np
import pytest

import scenes
import swiftvideo_b200 as sv
from gpu_util import FMT, context, fetch, first_diff, gpu_target, to_gpu
from oracle import oracle as O
from swiftvideo_b200 import api

pytestmark = pytest.mark.gpu
MODES = [(sv.MixMode.FUSED, "fused"), (sv.MixMode.GENERIC, "generic"), (sv.MixMode.PER_LAYER, "per_layer")]


def _stack(target_fmt, canvas, fmts, seed):
    layers, us = [], []
    geo = [((0, 0), canvas, 1.0, 0.0), ((8, 6), (70, 44), 0.6, 0.0), ((30, 20), (60, 40), 0.8, 0.35), ((-6, 30), (50, 30), 1.0, 0.0)]
    for k, f in enumerate(fmts):
        size = (64, 48) if k else (96, 64)
        layers.append(scenes.random_image(f, size[0], size[1], seed + k))
        pos, dsz, op, rot = geo[k % len(geo)]
        us.append(scenes.layer_uniforms(canvas, size, pos, dsz, z=k + 1, opacity=op, rotation=rot, border=(2, 1, 2, 1) if k == 1 else (0, 0, 0, 0),
                                        fill=(0.2, 0.5, 0.7, 0.8)))
    return layers, us


@pytest.mark.parametrize("mode,mname", MODES)
@pytest.mark.parametrize("target_fmt,fmts", [(O.NV12, [O.NV21, O.Y422P, O.Y444P, O.NV12]), (O.NV12, [O.Y444P, O.NV21, O.BGRA, O.Y422P]),
                                             (O.Y420P, [O.Y422P, O.Y444P, O.Y420P, O.RGBA])])
def test_new_sources(mode, mname, target_fmt, fmts):
    ctx = context()
    canvas = (96, 64)
    layers, us = _stack(target_fmt, canvas, fmts, 8100 + fmts[0])
    want = O.Image(target_fmt, *canvas)
    assert O.port().mix(want, layers, us) == 0
    target = gpu_target(ctx, target_fmt, *canvas)
    sv.compose(ctx, target, [to_gpu(ctx, l, f"l{i}") for i, l in enumerate(layers)], us, mode)
    got = fetch(ctx, target)
    assert (got == want.data).all(), f"{mname}: {first_diff(got, want.data)}"


def test_new_sources_through_the_mixer_at_size():
    """A 1080p mixer with a 4:4:4 graphic and an NV21 camera among NV12 layers: findKernel resolves img_y444p_nv12 / img_nv21_nv12 and the
    frame goes through svb_mix_generic."""
    from test_gpu_mixer import _oracle_mix, _place
    ctx = context()
    canvas = (1920, 1080)
    fmts = [O.NV12, O.NV21, O.Y444P, O.Y422P]
    imgs = [scenes.random_image(f, 640, 360, 8300 + i) for i, f in enumerate(fmts)]
    gpu = [to_gpu(ctx, im, f"a{i}") for i, im in enumerate(imgs)]
    placed = [_place(gpu[0], canvas, (640, 360), (0, 0), canvas, z=0), _place(gpu[1], canvas, (640, 360), (100, 80), (800, 450), z=1, opacity=0.75),
              _place(gpu[2], canvas, (640, 360), (900, 500), (960, 540), z=2, opacity=0.9), _place(gpu[3], canvas, (640, 360), (1200, 40), (640, 360), z=3)]
    mixer = sv.VideoMixer(ctx, canvas[0], canvas[1], sv.NV12, asset_id="m", workspace_id="w")
    for p in placed:
        assert mixer.push(p)
    got = fetch(ctx, mixer.mix(1000))
    want = _oracle_mix(O.NV12, canvas, placed, imgs, lib=O.port())   # the extension formats exist in the restatement only
    assert (got == want).all(), first_diff(got, want)
    assert api.default_compute_kernel_from_string("img_y444p_nv12") > api.KERNEL_CUSTOM
    with pytest.raises(sv.ComputeError):                              # like img_nv12_y420p: not offered
        api.default_compute_kernel_from_string("img_nv21_y420p")
    mixer.close()


def test_bgra_mixer_uses_img_bgra_bgra():
    """VideoMixer(outputFormat: .BGRA) with BGRA layers: findKernel gives img_clear_bgra + img_bgra_bgra (mix.video.swift:142-146) -- on
    upstream's Linux build that kernel does not exist; here it is the Metal text's (nearest texel, source-over, no transform)."""
    ctx = context()
    canvas = (128, 64)
    rng = np.random.default_rng(12)
    imgs = []
    for size in ((128, 64), (64, 32), (256, 128)):
        im = O.Image(O.BGRA, *size)
        im.data[:] = rng.integers(0, 256, im.nbytes, dtype=np.uint8)
        imgs.append(im)
    imgs[0].data.reshape(-1, 4)[:, 3] = 255
    mixer = sv.VideoMixer(ctx, canvas[0], canvas[1], sv.BGRA, asset_id="m", workspace_id="w")
    placed = []
    for k, im in enumerate(imgs):
        p = to_gpu(ctx, im, f"g{k}").animate(canvas, (0, 0, float(k)), canvas)
        placed.append(p)
        assert mixer.push(p)
    got = fetch(ctx, mixer.mix(1000))
    want = O.Image(O.BGRA, *canvas)
    assert O.port().clear(want) == 0
    for im in imgs:
        u = O.Uniforms()
        u.inSize[:] = [im.width, im.height]
        u.outSize[:] = list(canvas)
        assert O.port().apply_bgra_bgra(want, im, u) == 0
    assert (got == want.data).all(), first_diff(got, want.data)
    mixer.close()


def test_clear_yuvs():
    ctx = context()
    t = sv.create_picture_sample(64, 32, sv.YUVS, "t", "w")
    t.set_host_bytes(np.full(64 * 32 * 2, 0x5A, dtype=np.uint8))
    g = t.upload(ctx)
    api.run_compute_kernel(ctx, [], g, api.default_compute_kernel_from_string("img_clear_yuvs"))
    got = fetch(ctx, g)
    want = O.Image(O.YUVS, 64, 32)
    assert O.port().clear(want) == 0
    assert (got == want.data).all() and (got.reshape(-1, 2) == [0, 128]).all()
