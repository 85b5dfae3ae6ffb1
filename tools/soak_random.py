#!/usr/bin/env python3
"""One-off soak: tests/test_gpu_random.py's scene generator over a range of seeds far beyond the 48 of the suite, default compositor
(svb_mix_ring) against the reference's kernel text.   python tools/soak_random.py [first [count]]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import scenes  # noqa: E402
import swiftvideo_b200 as sv  # noqa: E402
from gpu_util import context, first_diff, gpu_case  # noqa: E402
from oracle import oracle as O  # noqa: E402
from test_gpu_random import random_case  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 48
count = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
bad, skipped, t0 = [], 0, time.time()
for seed in range(first, first + count):
    try:
        case = random_case(seed)
    except Exception:  # (a degenerate draw: the generator's matrix inverse fails)
        skipped += 1
        continue
    rc, want = scenes.run_case(O.best()[0], case)
    assert rc == 0
    got = gpu_case(context(), case, sv.MixMode.FUSED)
    if not (got == want.data).all():
        bad.append((seed, first_diff(got, want.data)))
print(f"{count} random scenes (seeds {first}..{first + count - 1}), checker: {O.best()[1]}: {len(bad)} differ, {skipped} draws skipped (degenerate), {time.time() - t0:.0f} s")
for b in bad[:20]:
    print("  seed", b[0], b[1])
sys.exit(1 if bad else 0)
