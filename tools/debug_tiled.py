"""Manual debugging aid (not a test): compose scenes through one mode and report where bytes differ from the oracle."""
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
base = T.TILED[0]
cases = [scenes.Case(f"only_layer{i}", base.target_fmt, base.canvas, [base.layers[i]], [base.uniforms[i]]) for i in range(len(base.layers))]
cases += [scenes.Case("layers01", base.target_fmt, base.canvas, base.layers[:2], base.uniforms[:2]), base]
for case in cases:
    rc, want = scenes.run_case(O.port(), case)
    got = gpu_case(context(), case, mode)
    W, H = case.canvas
    d = np.nonzero(got != want.data)[0]
    if d.size == 0:
        print(case.name, "identical"); continue
    luma = d[d < W * H]
    ys, xs = luma // W, luma % W
    print(case.name, f"{d.size} differ; luma {luma.size}: x[{xs.min() if luma.size else -1},{xs.max() if luma.size else -1}] y[{ys.min() if luma.size else -1},{ys.max() if luma.size else -1}]",
          "first", [(int(x), int(y), int(got[i]), int(want.data[i])) for x, y, i in list(zip(xs, ys, luma))[:6]])
    if luma.size:
        tiles = sorted(set((int(x) // 128, int(y) // 32) for x, y in zip(xs, ys)))
        print("   tiles with luma diffs:", tiles[:40])
