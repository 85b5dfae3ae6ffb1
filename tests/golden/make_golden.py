#!/usr/bin/env python3
"""Generate tests/golden/cases.npz from oracle/_ref -- the reference's own OpenCL kernel text
(/root/reference/Sources/SwiftVideo/kernels.cl.swift) compiled for the host by `make -C oracle`.

Run where /root/reference is mounted:   python tests/golden/make_golden.py
Every case stores its inputs (layer bytes, the 236-byte ImageUniforms of each layer) and the bytes the reference
kernels produced, so the fixtures do not depend on numpy's RNG or on this script's scene code staying unchanged.
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    ref = O.ref()
    out = {}
    names = []
    for case in scenes.parity_cases():
        rc, target = scenes.run_case(ref, case)
        assert rc == 0, case.name
        names.append(case.name)
        k = case.name
        out[f"{k}/meta"] = np.array([case.target_fmt, case.canvas[0], case.canvas[1], len(case.layers)], dtype=np.int32)
        out[f"{k}/out"] = target.data.copy()
        for i, (l, u) in enumerate(zip(case.layers, case.uniforms)):
            out[f"{k}/l{i}/meta"] = np.array([l.format, l.width, l.height], dtype=np.int32)
            out[f"{k}/l{i}/data"] = l.data.copy()
            out[f"{k}/l{i}/uniforms"] = np.frombuffer(C.string_at(C.addressof(u), 236), dtype=np.uint8).copy()
    out["names"] = np.array(names)
    path = Path(__file__).resolve().parent / "cases.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size} bytes, {len(names)} cases)")


if __name__ == "__main__":
    main()
