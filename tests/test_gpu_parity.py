"""-m gpu: the CUDA path through the C ABI against the oracle -- bit-exact bytes.

Every case runs in four compose modes: svb_mix_strip (FUSED, the product path: a warp stages and composites its own 64x8
unit), svb_mix_tiled (FUSED_TILED, the CTA-per-tile TMA kernel of round 1), the generic fused kernel (GENERIC) and the
reference's own per-layer launch sequence over the drop-in kernels (PER_LAYER).  The checker is O.best(): the reference's own
OpenCL kernel text compiled as C++ (oracle/_ref) wherever that build is present, else the C restatement."""
import numpy as np
import pytest

import scenes
import swiftvideo_b200 as sv
from gpu_util import context, fetch, first_diff, gpu_case, gpu_target, to_gpu
from oracle import oracle as O

pytestmark = pytest.mark.gpu

MODES = [(sv.MixMode.FUSED, "fused"), (sv.MixMode.FUSED_RING, "fused_ring"), (sv.MixMode.FUSED_TILED, "fused_tiled"),
         (sv.MixMode.GENERIC, "generic"), (sv.MixMode.PER_LAYER, "per_layer")]
CHECKER = O.best()[0]
# the fused compositor with its taps through the texture unit (svb_mix_gather): same plans, tables and bytes
MODES_GATHER = MODES + [(sv.MixMode.FUSED_GATHER, "fused_gather")]
CASES = scenes.parity_cases()


def test_unorm_identity():
    """The division-free UNORM8 read equals c/255.0f for every byte, on the device and against numpy."""
    fast, divided = context().selftest_unorm()
    want = np.arange(256, dtype=np.float32) / np.float32(255.0)
    assert (fast.view(np.uint32) == divided.view(np.uint32)).all()
    assert (fast.view(np.uint32) == want.view(np.uint32)).all()


@pytest.mark.parametrize("mode,mname", MODES, ids=[m[1] for m in MODES])
@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_small_scenes(case, mode, mname):
    rc, want = scenes.run_case(CHECKER, case)
    assert rc == 0
    got = gpu_case(context(), case, mode)
    assert (got == want.data).all(), f"{case.name}/{mname}: {first_diff(got, want.data)}"


def _tiled_cases():
    """Bigger canvases so that tiles lie fully inside layers (the TMA-staged table path) and on their edges."""
    cases = []
    canvas = (640, 352)
    rng_seed = 7000
    for sf, tf in ((O.NV12, O.NV12), (O.Y420P, O.NV12), (O.Y420P, O.Y420P)):
        layers = [scenes.random_image(sf, 640, 352, rng_seed), scenes.random_image(sf, 960, 540, rng_seed + 1),
                  scenes.random_image(sf, 320, 180, rng_seed + 2), scenes.random_image(sf, 400, 300, rng_seed + 3),
                  scenes.random_image(sf, 512, 288, rng_seed + 4)]
        us = [scenes.layer_uniforms(canvas, (640, 352), (0, 0), canvas, z=1, opacity=1.0),                       # 1:1
              scenes.layer_uniforms(canvas, (960, 540), (40, 20), (480, 270), z=2, opacity=0.6),                 # 2x downscale
              scenes.layer_uniforms(canvas, (320, 180), (100, 60), (500, 280), z=3, opacity=0.45),               # upscale
              scenes.layer_uniforms(canvas, (400, 300), (-30, 100), (400, 250), z=4, opacity=0.8,                # off-canvas, fit + fill
                                    fill=(0.3, 0.5, 0.7, 0.6), aspect="fit", border=(5, 5, 5, 5)),
              scenes.layer_uniforms(canvas, (512, 288), (600, 330), (-560, -300), z=5, opacity=0.7)]             # mirrored in x and y
        cases.append(scenes.Case(f"tiled_{O.FORMAT_NAMES[sf]}_{O.FORMAT_NAMES[tf]}", tf, canvas, layers, us))
        rng_seed += 10
    # heavy downscale (footprint does not fit the staged box -> global taps) and a 17-layer stack (two passes)
    src = scenes.random_image(O.NV12, 1920, 1080, 7100)
    cases.append(scenes.Case("tiled_downscale_4x", O.NV12, (480, 256), [src],
                             [scenes.layer_uniforms((480, 256), (1920, 1080), (0, 0), (480, 256), z=1)]))
    many_l, many_u = [], []
    for k in range(17):
        many_l.append(scenes.random_image(O.NV12, 160 + 16 * k, 96 + 8 * k, 7200 + k))
        many_u.append(scenes.layer_uniforms((512, 288), (160 + 16 * k, 96 + 8 * k), (10 * k, 6 * k), (300, 170), z=k + 1, opacity=0.3 + 0.04 * k))
    cases.append(scenes.Case("tiled_17_layers", O.NV12, (512, 288), many_l, many_u))
    # opaque pictures above other layers: tiles they cover completely skip everything underneath (planner occlusion)
    canvas = (640, 352)
    occ_l = [scenes.random_image(O.NV12, 640, 352, 7400), scenes.random_image(O.BGRA, 300, 200, 7401), scenes.random_image(O.NV12, 480, 270, 7402),
             scenes.random_image(O.NV12, 320, 180, 7403), scenes.random_image(O.NV12, 256, 144, 7404)]
    occ_u = [scenes.layer_uniforms(canvas, (640, 352), (0, 0), canvas, z=1, opacity=1.0),
             scenes.layer_uniforms(canvas, (300, 200), (50, 40), (300, 200), z=2, opacity=0.7),
             scenes.layer_uniforms(canvas, (480, 270), (64, 32), (384, 224), z=3, opacity=1.0),   # covers whole tiles: hides layers 0-1 there
             scenes.layer_uniforms(canvas, (320, 180), (200, 100), (320, 180), z=4, opacity=0.5),
             scenes.layer_uniforms(canvas, (256, 144), (300, 150), (256, 160), z=5, opacity=1.0)]  # on top, hides 0-3 in its interior tiles
    cases.append(scenes.Case("tiled_occluders", O.NV12, canvas, occ_l, occ_u))
    # odd plane strides / widths that break the tiled kernel's alignment preconditions fall back to generic
    src = scenes.random_image(O.NV12, 250, 130, 7300)
    cases.append(scenes.Case("unaligned_250x130_src", O.NV12, (384, 224), [src],
                             [scenes.layer_uniforms((384, 224), (250, 130), (0, 0), (384, 224), z=1, opacity=0.9)]))
    return cases


TILED = _tiled_cases()


@pytest.mark.parametrize("mode,mname", MODES_GATHER, ids=[m[1] for m in MODES_GATHER])
@pytest.mark.parametrize("case", TILED, ids=[c.name for c in TILED])
def test_tiled_scenes(case, mode, mname):
    rc, want = scenes.run_case(CHECKER, case, threads=O.host_threads())
    assert rc == 0
    got = gpu_case(context(), case, mode)
    assert (got == want.data).all(), f"{case.name}/{mname}: {first_diff(got, want.data)}"


def test_cfg2_full_size():
    """BASELINE config 2 at full size: 1920x1080 NV12 -> 1280x720 NV12, bytes against the oracle."""
    canvas, tf, layers, us = scenes.cfg2_scene()
    case = scenes.Case("cfg2", tf, canvas, layers, us)
    rc, want = scenes.run_case(CHECKER, case, threads=O.host_threads())
    assert rc == 0
    for mode, mname in MODES:
        got = gpu_case(context(), case, mode)
        assert (got == want.data).all(), f"cfg2/{mname}: {first_diff(got, want.data)}"


@pytest.mark.parametrize("nlayers", [4, 8])
def test_cfg34_full_size(nlayers):
    """BASELINE configs 3 and 4 (one stream) at 3840x2160: fused and generic against the oracle, and against each
    other and the per-layer sequence (a size-independent cross-check: three independent code paths, same bytes)."""
    canvas, tf, layers, us = scenes.cfg34_scene(nlayers)
    case = scenes.Case(f"cfg{3 if nlayers == 4 else 4}", tf, canvas, layers, us)
    rc, want = scenes.run_case(CHECKER, case, threads=O.host_threads())
    assert rc == 0
    ctx = context()
    outs = {}
    for mode, mname in MODES_GATHER:
        outs[mname] = gpu_case(ctx, case, mode)
        assert (outs[mname] == want.data).all(), f"{case.name}/{mname}: {first_diff(outs[mname], want.data)}"


def test_cfg4_variants_full_size():
    """cfg 4's geometry at 3840x2160 in the two variants bench.py offers beside the headline: (1) YUV420P layers and target -- the
    format SwiftVideo composes in on Linux (composer.swift:52-56); (2) the two topmost pictures-in-picture as RGBA / BGRA overlays
    with per-pixel alpha, which take the table-driven RGBA body of the tiled kernel.  Fused path against the oracle, bytes."""
    ctx = context()
    canvas = (3840, 2160)
    geo = scenes.cfg34_geometry(8)
    # (1) planar
    layers = [scenes.random_image(O.Y420P, ssz[0], ssz[1], scenes.cfg_seed(4, 1, k)) for k, (ssz, _, _, _) in enumerate(geo)]
    us = [scenes.layer_uniforms(canvas, ssz, pos, dsz, z=float(k + 1), opacity=op) for k, (ssz, pos, dsz, op) in enumerate(geo)]
    case = scenes.Case("cfg4_y420p", O.Y420P, canvas, layers, us)
    rc, want = scenes.run_case(CHECKER, case, threads=O.host_threads())
    assert rc == 0
    for mode, mname in ((sv.MixMode.FUSED, "fused"), (sv.MixMode.FUSED_RING, "fused_ring"), (sv.MixMode.FUSED_TILED, "fused_tiled"),
                        (sv.MixMode.FUSED_GATHER, "fused_gather")):
        got = gpu_case(ctx, case, mode)
        assert (got == want.data).all(), f"cfg4_y420p/{mname}: {first_diff(got, want.data)}"
    # (2) RGBA / BGRA overlays on an NV12 stack (4 layers keep the oracle's time down)
    geo = scenes.cfg34_geometry(8)[:2] + scenes.cfg34_geometry(8)[5:7]
    fmts = [O.NV12, O.NV12, O.RGBA, O.BGRA]
    layers = [scenes.random_image(f, ssz[0], ssz[1], scenes.cfg_seed(4, 2, k)) for k, (f, (ssz, _, _, _)) in enumerate(zip(fmts, geo))]
    us = [scenes.layer_uniforms(canvas, ssz, pos, dsz, z=float(k + 1), opacity=op) for k, (ssz, pos, dsz, op) in enumerate(geo)]
    case = scenes.Case("cfg4_rgba_overlays", O.NV12, canvas, layers, us)
    rc, want = scenes.run_case(CHECKER, case, threads=O.host_threads())
    assert rc == 0
    for mode, mname in MODES_GATHER:
        got = gpu_case(ctx, case, mode)
        assert (got == want.data).all(), f"cfg4_rgba_overlays/{mname}: {first_diff(got, want.data)}"


def test_properties_full_size():
    """Size-independent properties at 4K: (1) an opaque full-canvas top layer hides everything below it;
    (2) composing twice gives the same bytes (no state leaks between launches); (3) opacity 0 leaves the clear."""
    ctx = context()
    canvas, tf, layers, us = scenes.cfg34_scene(4)
    gl = [to_gpu(ctx, l, f"p{i}") for i, l in enumerate(layers)]
    top_only = gpu_target(ctx, tf, *canvas)
    sv.compose(ctx, top_only, [gl[0]], [us[0]], sv.MixMode.FUSED)
    stacked = gpu_target(ctx, tf, *canvas)
    sv.compose(ctx, stacked, gl[1:] + [gl[0]], us[1:] + [us[0]], sv.MixMode.FUSED)  # layer 0 (opacity 1, full canvas) on top
    a, b = fetch(ctx, top_only), fetch(ctx, stacked)
    assert (a == b).all(), first_diff(a, b)
    again = gpu_target(ctx, tf, *canvas)
    sv.compose(ctx, again, gl[1:] + [gl[0]], us[1:] + [us[0]], sv.MixMode.FUSED)
    assert (fetch(ctx, again) == b).all()
    # opacity 0 everywhere: cur*(1-0) + v*0 == cur, so the picture stays at the clear values
    zero = [scenes.layer_uniforms(canvas, (l.width, l.height), (0, 0), canvas, z=i + 1, opacity=0.0) for i, l in enumerate(layers)]
    cleared = gpu_target(ctx, tf, *canvas)
    sv.compose(ctx, cleared, gl, zero, sv.MixMode.FUSED)
    want = O.Image(tf, *canvas)
    CHECKER.clear(want)
    assert (fetch(ctx, cleared) == want.data).all()


def _padded(img, pads):
    """The same picture with decoder-style padded rows (stride = row bytes + pad); padding bytes are junk."""
    strides = [l[1] * l[4] + p for l, p in zip(img.layout, pads)]
    out = O.Image(img.format, img.width, img.height, strides=strides)
    out.data[:] = 0x5C
    for i in range(len(img.layout)):
        w, nc = img.layout[i][1], img.layout[i][4]
        out.plane(i)[:, :] = img.plane(i)[:, : w * nc]
    return out


@pytest.mark.parametrize("mode,mname", MODES, ids=[m[1] for m in MODES])
@pytest.mark.parametrize("pads", [(64, 64, 32), (16, 16, 16), (6, 6, 6), (6, 10, 2)],
                         ids=["pad64_uv_differ", "pad16", "pad6_no_tma", "pad_unaligned"])
def test_padded_strides(pads, mode, mname):
    """FFmpeg hands the mixer planes whose stride is its linesize, not the width (dec.video.ffmpeg.swift:183): padded
    sources (16-byte multiples keep the TMA path, others fall back) must give the same bytes as tight ones."""
    base = TILED[1]  # y420p sources -> nv12 target: three planes per source
    layers = [_padded(l, pads) for l in base.layers]
    case = scenes.Case("padded", base.target_fmt, base.canvas, layers, base.uniforms)
    rc, want = scenes.run_case(CHECKER, base, threads=O.host_threads())
    assert rc == 0
    got = gpu_case(context(), case, mode)
    assert (got == want.data).all(), f"padded/{mname}: {first_diff(got, want.data)}"


def test_padded_target():
    """A target with padded rows (fused and generic take explicit output strides; the per-layer kernels infer the stride
    from the launch like the reference's, kernels.cuda.swift:151,205, so they only accept tight targets).  A pad of 64 keeps the
    rows 16-byte aligned (svb_mix_ring's whole units leave as 16-byte stores), a pad of 6 does not (two bytes per lane and row)."""
    base = TILED[0]
    rc, want = scenes.run_case(CHECKER, base, threads=O.host_threads())
    assert rc == 0
    W, H = base.canvas
    for padb in (64, 6):
        S = W + padb
        for mode, mname in MODES[:4]:
            got = gpu_case(context(), base, mode, target_strides=[S, S])
            tight = np.concatenate([got[: S * H].reshape(H, S)[:, :W].reshape(-1),
                                    got[S * H :].reshape(H // 2, S)[:, :W].reshape(-1)])
            assert (tight == want.data).all(), f"padded target {padb}/{mname}: {first_diff(tight, want.data)}"
            pad = np.concatenate([got[: S * H].reshape(H, S)[:, W:].reshape(-1), got[S * H :].reshape(H // 2, S)[:, W:].reshape(-1)])
            assert (pad == 0xA5).all(), "padding bytes of the target were written"
    with pytest.raises(sv.ComputeError) as e:
        gpu_case(context(), base, sv.MixMode.PER_LAYER, target_strides=[W + 64, W + 64])
    assert "badTarget" in str(e.value)


GOLDEN = scenes.golden_cases()


@pytest.mark.parametrize("mode,mname", MODES_GATHER, ids=[m[1] for m in MODES_GATHER])
def test_golden_vectors_direct(mode, mname):
    """tests/golden/cases.npz (inputs, uniforms and the bytes oracle/_ref -- the reference's kernel text -- produced) fed straight to
    the CUDA path: no oracle runs in this test."""
    for case, want in GOLDEN:
        got = gpu_case(context(), case, mode)
        assert (got == want).all(), f"golden {case.name}/{mname}: {first_diff(got, want)}"


@pytest.mark.parametrize("mode,mname", MODES, ids=[m[1] for m in MODES])
def test_equal_z_is_deterministic(mode, mname):
    """Equal zIndex has no defined order upstream (an unstable sort over a dictionary, mix.video.swift:115); here ties keep the
    order of the layer list, so the same call gives the same bytes every time, and those bytes are the fold in list order."""
    canvas = (640, 352)
    layers = [scenes.random_image(O.NV12, 640, 352, 9100), scenes.random_image(O.NV12, 400, 240, 9101), scenes.random_image(O.NV12, 400, 240, 9102)]
    us = [scenes.layer_uniforms(canvas, (640, 352), (0, 0), canvas, z=1, opacity=1.0),
          scenes.layer_uniforms(canvas, (400, 240), (60, 40), (400, 240), z=2, opacity=0.5),
          scenes.layer_uniforms(canvas, (400, 240), (120, 70), (400, 240), z=2, opacity=0.5)]  # same z as the layer before, overlapping it
    case = scenes.Case("equal_z", O.NV12, canvas, layers, us)
    rc, want = scenes.run_case(CHECKER, case)
    assert rc == 0
    a = gpu_case(context(), case, mode)
    b = gpu_case(context(), case, mode)
    assert (a == b).all() and (a == want.data).all(), f"equal_z/{mname}: {first_diff(a, want.data)}"


def test_table_cache_follows_the_geometry():
    """svb_mix_ring keeps a batch's coordinate tables while the batch's geometry is what it was when the batch's buffer (one of eight,
    used in turn) was last filled, and runs its table pre-pass otherwise (svb_table_cache, include/svb200.h).  Launch patterns with a
    period of eight make every launch of rounds 1 and 3 find its buffer's tables in place, and every launch of rounds 0 and 2 find
    another scene's; with the cache off the pre-pass runs every time.  Every output is compared with the reference's."""
    ctx = context()
    cases = [TILED[0], TILED[1 % len(TILED)], TILED[2 % len(TILED)]]
    wants = []
    for c in cases:
        rc, want = scenes.run_case(CHECKER, c, threads=O.host_threads())
        assert rc == 0
        wants.append(want.data)
    try:
        for enabled in (True, False, True):
            ctx.table_cache(enabled)
            for rnd in range(4):
                for i in range(8):
                    k = (i + rnd // 2) % len(cases)
                    got = gpu_case(ctx, cases[k], sv.MixMode.FUSED_RING)
                    assert (got == wants[k]).all(), f"cache {enabled}, round {rnd}, launch {i}, {cases[k].name}: {first_diff(got, wants[k])}"
    finally:
        ctx.table_cache(True)
