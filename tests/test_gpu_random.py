"""-m gpu: seeded random scenes through the fused path against the oracle -- geometry the hand-written cases do not
think of (sub-pixel positions, extreme scales, layers hanging off every edge, mirrored, thin borders, mixed formats)."""
import numpy as np
import pytest

import scenes
import swiftvideo_b200 as sv
from gpu_util import context, first_diff, gpu_case
from oracle import oracle as O

pytestmark = pytest.mark.gpu

SRC_FORMATS = [O.NV12, O.Y420P, O.BGRA, O.RGBA]


def random_case(seed):
    rng = np.random.default_rng(seed)
    tf = [O.NV12, O.Y420P][int(rng.integers(0, 2))]
    cw = int(rng.choice([132, 256, 384, 640, 644]))   # 132 and 644: W % 4 == 0 but partial tiles; others tile-aligned
    ch = int(rng.choice([34, 64, 128, 180, 360]))
    if rng.random() < 0.15:
        cw += 2                                        # W % 4 == 2: the tiled kernel's precondition fails -> generic kernel
    n = int(rng.integers(1, 7))
    layers, us = [], []
    for k in range(n):
        fmts = [f for f in SRC_FORMATS if not (f == O.NV12 and tf == O.Y420P)]  # img_nv12_y420p does not exist
        sf = fmts[int(rng.integers(0, len(fmts)))] if rng.random() < 0.35 else (O.Y420P if tf == O.Y420P else O.NV12)
        sw, sh = int(rng.integers(1, 200)) * 2, int(rng.integers(1, 120)) * 2
        scale = float(np.exp(rng.uniform(np.log(0.2), np.log(5.0))))
        dw, dh = max(2.0, sw * scale * rng.uniform(0.8, 1.25)), max(2.0, sh * scale * rng.uniform(0.8, 1.25))
        if rng.random() < 0.3:                         # whole-pixel placement at native size (the unit-step case)
            dw, dh = float(sw), float(sh)
            pos = (float(rng.integers(-sw // 2, cw)), float(rng.integers(-sh // 2, ch)))
        else:
            pos = (float(rng.uniform(-dw * 0.6, cw)), float(rng.uniform(-dh * 0.6, ch)))
        if rng.random() < 0.1:
            dw = -dw                                   # mirrored
        if rng.random() < 0.1:
            dh = -dh
        rot = float(rng.uniform(-0.6, 0.6)) if rng.random() < 0.15 else 0.0
        border = tuple(float(v) for v in rng.integers(0, 6, 4)) if rng.random() < 0.3 else (0.0, 0.0, 0.0, 0.0)
        aspect = ["none", "fit", "fill"][int(rng.integers(0, 3))]
        opacity = float(rng.choice([1.0, 1.0, rng.uniform(0, 1), rng.uniform(0, 1), 0.0, 1.25]))
        fill = tuple(float(v) for v in rng.uniform(0, 1, 4)) if rng.random() < 0.5 else (0.0, 0.0, 0.0, 0.0)
        layers.append(scenes.random_image(sf, sw, sh, seed * 100 + k, "uniform" if rng.random() < 0.7 else "ramp"))
        us.append(scenes.layer_uniforms((cw, ch), (sw, sh), pos, (dw, dh), rotation=rot, z=k + 1, opacity=opacity, fill=fill, border=border,
                                        aspect=aspect))
    return scenes.Case(f"random{seed}", tf, (cw, ch), layers, us)


@pytest.mark.parametrize("seed", range(48))
def test_random_scene(seed):
    case = random_case(seed)
    rc, want = scenes.run_case(O.best()[0], case)  # the reference's kernel text where that build is present
    assert rc == 0
    for mode, name in ((sv.MixMode.FUSED, "fused"), (sv.MixMode.FUSED_RING, "fused_ring"), (sv.MixMode.FUSED_TILED, "fused_tiled"),
                       (sv.MixMode.GENERIC, "generic")):
        got = gpu_case(context(), case, mode)
        assert (got == want.data).all(), f"seed {seed}/{name}: {first_diff(got, want.data)}"
