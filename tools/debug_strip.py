"""Manual debugging aid (not a test): compose scenes through one mode and report where bytes differ from the oracle, by 64x8 unit.
    python tools/debug_strip.py [mode=0]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import scenes
import swiftvideo_b200 as sv
from gpu_util import context, gpu_case
from oracle import oracle as O
import test_gpu_parity as T

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cases = []
for base in T.TILED[:3]:
    cases += [scenes.Case(f"{base.name}_only{i}", base.target_fmt, base.canvas, [base.layers[i]], [base.uniforms[i]]) for i in range(len(base.layers))]
    cases += [scenes.Case(f"{base.name}_01", base.target_fmt, base.canvas, base.layers[:2], base.uniforms[:2]), base]
cases += T.TILED[3:]
bad = 0
for case in cases:
    rc, want = scenes.run_case(O.port(), case, threads=O.host_threads())
    try:
        got = gpu_case(context(), case, mode)
    except Exception as e:
        print(case.name, "EXCEPTION", e); bad += 1; continue
    W, H = case.canvas
    d = np.nonzero(got != want.data)[0]
    if d.size == 0:
        print(case.name, "identical"); continue
    bad += 1
    luma = d[d < W * H]
    chroma = d[d >= W * H] - W * H
    ys, xs = luma // W, luma % W
    print(case.name, f"{d.size} differ; luma {luma.size} chroma {chroma.size}")
    if luma.size:
        print("   luma x[%d,%d] y[%d,%d] first" % (xs.min(), xs.max(), ys.min(), ys.max()), [(int(x), int(y), int(got[i]), int(want.data[i])) for x, y, i in list(zip(xs, ys, luma))[:6]])
        units = sorted(set((int(x) // 64, int(y) // 8) for x, y in zip(xs, ys)))
        print("   units with luma diffs (%d):" % len(units), units[:30])
        print("   rows in unit:", sorted(set(int(y) % 8 for y in ys)), " cols in unit (mod 64):", sorted(set(int(x) % 64 for x in xs))[:20])
    if chroma.size:
        if case.target_fmt == O.NV12:
            cy, cx = chroma // W, chroma % W
            print("   chroma(nv12) x[%d,%d] y[%d,%d] first" % (cx.min(), cx.max(), cy.min(), cy.max()), [(int(x), int(y), int(got[W * H + i]), int(want.data[W * H + i])) for x, y, i in list(zip(cx, cy, chroma))[:6]])
        else:
            print("   chroma(planar) first offsets", chroma[:8])
print("cases with differences:", bad)
