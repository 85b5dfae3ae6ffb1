"""CPU: the oracle itself.  (1) the hand-written restatement against the reference's own OpenCL kernel text
compiled for the host (oracle/_ref, when built); (2) both against the committed golden fixtures, which were
generated from oracle/_ref; (3) known answers that follow from the kernel text by hand."""
import numpy as np
import pytest

import scenes
from oracle import oracle as O

CASES = scenes.parity_cases()
GOLDEN = scenes.golden_cases()
needs_ref = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_port_equals_reference_text(case):
    rc1, a = scenes.run_case(O.port(), case)
    rc2, b = scenes.run_case(O.ref(), case)
    assert rc1 == rc2 == 0
    assert (a.data == b.data).all()


@pytest.mark.parametrize("case,want", GOLDEN, ids=[c.name for c, _ in GOLDEN])
def test_port_against_golden(case, want):
    rc, got = scenes.run_case(O.port(), case)
    assert rc == 0 and (got.data == want).all()
    rc, got = scenes.run_case(O.port(), case, threads=3)  # row-band variant used for CPU timing
    assert rc == 0 and (got.data == want).all()


@needs_ref
def test_reference_text_against_golden():
    for case, want in GOLDEN:
        rc, got = scenes.run_case(O.ref(), case, threads=2)
        assert rc == 0 and (got.data == want).all(), case.name


def test_golden_matches_scene_code():
    """The fixtures were made from scenes.parity_cases(); if the scene code drifts, say so."""
    assert [c.name for c, _ in GOLDEN] == [c.name for c in CASES]


def test_clear_known_answer():
    """img_clear_*: Y = 0, chroma = 0.5 -> 128 after UNORM8 round-to-even, BGRA = (0,0,0,255)
    (kernels.cl.swift:43-44,181-183,262)."""
    for fmt, planes in ((O.NV12, [0, 128]), (O.Y420P, [0, 128, 128])):
        t = O.Image(fmt, 16, 8)
        t.data[:] = 7
        assert O.port().clear(t) == 0
        for i, v in enumerate(planes):
            assert (t.plane(i) == v).all()
    t = O.Image(O.BGRA, 4, 2)
    O.port().clear(t)
    assert (t.data.reshape(-1, 4) == [0, 0, 0, 255]).all()


def test_half_pixel_shift_known_answer():
    """out_uv = gid/size has no +0.5 (kernels.cl.swift:72), so a same-size full-canvas layer samples at texel
    corners: every output is the mean of a 2x2 neighbourhood (clamped at the top/left edge)."""
    w, h = 32, 16
    src = O.Image(O.NV12, w, h)
    src.plane(0)[:, :] = (np.arange(w)[None, :] * 4 + np.arange(h)[:, None] * 8) % 256
    src.plane(1)[:, :] = 128
    t = O.Image(O.NV12, w, h)
    u = scenes.layer_uniforms((w, h), (w, h), (0, 0), (w, h), z=1)
    assert O.port().mix(t, [src], [u]) == 0
    y = src.plane(0).astype(np.float64)
    yy, xx = np.mgrid[0:h, 0:w]
    pred = (y[np.maximum(yy - 1, 0), np.maximum(xx - 1, 0)] + y[np.maximum(yy - 1, 0), xx] + y[yy, np.maximum(xx - 1, 0)] + y[yy, xx]) / 4
    assert np.abs(t.plane(0).astype(np.float64) - pred).max() <= 0.5 + 1e-9
    assert (t.plane(1) == 128).all()


def test_opacity_zero_and_fill_known_answers():
    """opacity 0: cur*(1-0) + src*0 leaves the clear; a fill-only region takes RGB2YUV(fill) (kernels.cl.swift:96-105)."""
    canvas = (32, 16)
    src = scenes.random_image(O.NV12, 16, 8, 5)
    t = O.Image(O.NV12, *canvas)
    assert O.port().mix(t, [src], [scenes.layer_uniforms(canvas, (16, 8), (0, 0), canvas, z=1, opacity=0.0)]) == 0
    assert (t.plane(0) == 0).all() and (t.plane(1) == 128).all()
    # letterboxed 16x8 picture in a 32x4... the bars left and right of an aspect-fit picture are pure fill
    u = scenes.layer_uniforms(canvas, (8, 16), (0, 0), canvas, z=1, opacity=1.0, fill=(1.0, 1.0, 1.0, 1.0), aspect="fit")
    tall = scenes.random_image(O.NV12, 8, 16, 6)
    assert O.port().mix(t, [tall], [u]) == 0
    f32 = np.float32
    luma = f32(f32(f32(1) * f32(0.299) + f32(1) * f32(0.587)) + f32(1) * f32(0.113))  # 0.999: the 0.113 typo is upstream's
    assert t.plane(0)[4, 1] == int(np.rint(min(float(luma), 1.0) * 255))
    assert t.plane(0)[4, 30] == t.plane(0)[4, 1]


def test_unsupported_pairs_and_sizes():
    t = O.Image(O.Y420P, 16, 8)
    nv = scenes.random_image(O.NV12, 16, 8, 1)
    u = scenes.layer_uniforms((16, 8), (16, 8), (0, 0), (16, 8))
    assert O.port().apply(t, nv, u) == O.ERR_KERNEL_NOT_FOUND          # img_nv12_y420p is not in the enum
    b = O.Image(O.BGRA, 16, 8)
    assert O.port().apply(b, scenes.random_image(O.BGRA, 16, 8, 2), u) == O.ERR_KERNEL_NOT_FOUND  # no img_bgra_bgra on Linux
    odd = O.Image(O.NV12, 18, 8)
    odd.width = 17
    assert O.port().apply(odd, nv, u) == O.ERR_BAD_TARGET
