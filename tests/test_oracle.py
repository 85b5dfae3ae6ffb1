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


def test_cfg1_cpu_plumbing():
    """BASELINE.json configs[0] ("cfg 1", the reference's own CPU-runnable case): ONE 640x360 NV12 sample in CPU buffers through
    img_clear_nv12 + img_nv12_nv12 placed over the whole 640x360 canvas -- 345 600 source bytes in, 345 600 target bytes out
     -- and the same picture through the NV12 -> BGRA leg against libswscale.  No GPU anywhere in this test."""
    w, h = 640, 360
    src = scenes.random_image(O.NV12, w, h, scenes.cfg_seed(1, 0, 0))
    assert src.nbytes == 345600
    u = scenes.layer_uniforms((w, h), (w, h), (0, 0), (w, h), z=1.0)
    outs = []
    for lib in [O.port()] + ([O.ref()] if O.ref_available() else []):
        for threads in (0, 4):
            t = O.Image(O.NV12, w, h)
            t.data[:] = 0xA5
            assert lib.mix(t, [src], [u], threads) == 0
            outs.append(t.data.copy())
    assert all((o == outs[0]).all() for o in outs[1:])                  # restatement == reference text, single- and multi-threaded
    got = O.Image(O.NV12, w, h)
    got.data[:] = outs[0]
    # the kernel has no half-texel offset (kernels.cl.swift:72): a same-size picture lands as the rounded mean of each 2x2
    # neighbourhood (edge-clamped), in luma and in each chroma component -- predicted here in float64, so allow the half code of rounding
    for plane, ncomp in ((0, 1), (1, 2)):
        a = src.plane(plane).astype(np.float64).reshape(src.plane(plane).shape[0], -1, ncomp)
        yy, xx = np.mgrid[0:a.shape[0], 0:a.shape[1]]
        ym, xm = np.maximum(yy - 1, 0), np.maximum(xx - 1, 0)
        pred = (a[ym, xm] + a[ym, xx] + a[yy, xm] + a[yy, xx]) / 4
        out = got.plane(plane).astype(np.float64).reshape(a.shape)
        assert np.abs(out - pred).max() <= 0.5 + 1e-6
    # the BGRA leg of the same configuration: the convert+scale definition at 1:1 against libswscale's NV12 -> BGRA
    import swscale_util as S
    if S.available():
        smooth = S.smooth_picture(8, w, h, seed=1)
        ours = O.scale_convert(O.SC_NV12, O.SC_BILINEAR, smooth, w, h, w, h).astype(np.int32)
        sc = S.Scaler("nv12", w, h, w, h, S.SWS_BILINEAR | S.SWS_ACCURATE_RND | S.SWS_FULL_CHR_H_INT)
        ref = sc.run(smooth).astype(np.int32)
        sc.close()
        assert np.abs(ours - ref).max() <= 1                              # TOLERANCE: one 8-bit code value


# ---- extensions: operators the reference names but never implemented.  PARITY UNPINNED: these tests hold the
# ---- definitions in oracle/mixer_oracle.c against properties that follow from the reference kernels they extend. -------------------

def _ext_scene(src_fmt, tgt_fmt, canvas=(96, 64), size=(64, 48), seed=11):
    src = scenes.random_image(src_fmt, size[0], size[1], seed)
    u = scenes.layer_uniforms(canvas, size, (10, 6), (70, 50), z=1, opacity=0.7, border=(2, 2, 2, 2), fill=(0.3, 0.6, 0.1, 0.9))
    return src, u


def test_extension_nv21_is_nv12_with_swapped_pairs():
    src, u = _ext_scene(O.NV12, O.NV12)
    swapped = O.Image(O.NV21, src.width, src.height)
    swapped.plane(0)[:] = src.plane(0)
    swapped.plane(1)[:, 0::2] = src.plane(1)[:, 1::2]
    swapped.plane(1)[:, 1::2] = src.plane(1)[:, 0::2]
    a, b = O.Image(O.NV12, 96, 64), O.Image(O.NV12, 96, 64)
    assert O.port().mix(a, [src], [u]) == 0 and O.port().mix(b, [swapped], [u]) == 0
    assert (a.data == b.data).all()
    assert O.port().mix(O.Image(O.Y420P, 96, 64), [swapped], [u]) == O.ERR_KERNEL_NOT_FOUND   # like img_nv12_y420p: not offered


@pytest.mark.parametrize("fmt", [O.Y422P, O.Y444P])
@pytest.mark.parametrize("tgt", [O.NV12, O.Y420P])
def test_extension_planar_sources_share_the_y420p_body(fmt, tgt):
    """Luma is filtered exactly as for a Y420P source; chroma planes are sampled at their own size, so constant chroma planes give what
    a Y420P source with the same constants gives."""
    base, u = _ext_scene(O.Y420P, tgt)
    base.plane(1)[:] = 90
    base.plane(2)[:] = 200
    other = O.Image(fmt, base.width, base.height)
    other.plane(0)[:] = base.plane(0)
    other.plane(1)[:] = 90
    other.plane(2)[:] = 200
    a, b = O.Image(tgt, 96, 64), O.Image(tgt, 96, 64)
    assert O.port().mix(a, [base], [u]) == 0 and O.port().mix(b, [other], [u]) == 0
    assert (a.data == b.data).all()
    # and a chroma plane with structure is really read at full height (4:2:2) / full size (4:4:4): a vertical ramp survives
    other.plane(1)[:] = (np.arange(other.plane(1).shape[0])[:, None] * 3) % 256
    c = O.Image(tgt, 96, 64)
    assert O.port().mix(c, [other], [u]) == 0
    assert (c.data != b.data).any()


def test_extension_bgra_bgra_follows_the_metal_text():
    """kernels.metal:51-62: nearest texel at trunc(gid * in/out), source-over with the source alpha, alpha 1 out."""
    rng = np.random.default_rng(4)
    src = O.Image(O.BGRA, 32, 16)
    src.data[:] = rng.integers(0, 256, src.nbytes, dtype=np.uint8)
    tgt = O.Image(O.BGRA, 16, 8)
    tgt.data[:] = rng.integers(0, 256, tgt.nbytes, dtype=np.uint8)
    before = tgt.data.copy().reshape(8, 16, 4)
    u = O.Uniforms()
    u.inSize[:] = [32, 16]
    u.outSize[:] = [16, 8]
    s = src.data.reshape(16, 32, 4).copy()
    s[:, :, 3] = np.where(np.arange(32)[None, :] % 4 == 0, 255, s[:, :, 3])   # opaque columns
    s[:, :, 3] = np.where(np.arange(32)[None, :] % 4 == 2, 0, s[:, :, 3])     # transparent columns
    src.data[:] = s.reshape(-1)
    assert O.port().apply_bgra_bgra(tgt, src, u) == 0
    out = tgt.data.reshape(8, 16, 4)
    assert (out[:, :, 3] == 255).all()
    pick = s[::2, ::2]                                                        # texel (2x, 2y)
    assert (out[:, 0::2, :3] == pick[:, 0::2, :3]).all()                      # source columns 0, 4, 8...: opaque -> copied
    assert (out[:, 1::2, :3] == before[:, 1::2, :3]).all()                    # source columns 2, 6, ...: transparent -> untouched
    # a half-transparent texel: rint((s * a + d * (1 - a)) * 255) within a code of the float64 value
    s[:, :, 3] = 128
    src.data[:] = s.reshape(-1)
    tgt.data[:] = before.reshape(-1)
    assert O.port().apply_bgra_bgra(tgt, src, u) == 0
    a = 128 / 255
    pred = s[::2, ::2, :3].astype(np.float64) * a + before[:, :, :3].astype(np.float64) * (1 - a)
    assert np.abs(tgt.data.reshape(8, 16, 4)[:, :, :3] - pred).max() <= 0.5 + 1e-4
    assert O.port().apply_bgra_bgra(O.Image(O.NV12, 16, 8), src, u) == O.ERR_KERNEL_NOT_FOUND


def test_extension_clear_yuvs():
    t = O.Image(O.YUVS, 16, 4)
    t.data[:] = 9
    assert O.port().clear(t) == 0
    assert (t.data.reshape(-1, 2) == [0, 128]).all()
