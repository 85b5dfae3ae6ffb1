#!/usr/bin/env python3
"""Generate tests/golden/scale_cases.npz: fixtures of the convert+scale operator (NV12 / P010 -> BGRA, bilinear / Lanczos-3).

The reference has no such operator, so the fixtures hold BOTH sides of its pin: the bytes of libswscale 9.1 -- the comparator
BASELINE.json names, in its accurate mode (SWS_ACCURATE_RND | SWS_FULL_CHR_H_INT) -- and the bytes of the definition
(oracle/scale_oracle.c) on the same smooth inputs.  tests/test_scale.py checks, without libswscale at hand, that the
definition still produces its bytes and stays within one code value of the stored libswscale bytes; on the GPU the kernel
must reproduce the definition's bytes.

Run where the OpenCV wheel (libswscale) is installed:   python tests/golden/make_scale_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import swscale_util as S  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [  # (name, swscale pixel format, bits, oracle format, oracle filter, swscale flag, source size, destination size)
    ("nv12_bilinear_down", "nv12", 8, O.SC_NV12, O.SC_BILINEAR, S.SWS_BILINEAR, (192, 108), (128, 72)),
    ("nv12_bilinear_same", "nv12", 8, O.SC_NV12, O.SC_BILINEAR, S.SWS_BILINEAR, (96, 54), (96, 54)),
    ("nv12_lanczos_up", "nv12", 8, O.SC_NV12, O.SC_LANCZOS3, S.SWS_LANCZOS, (64, 36), (160, 90)),
    ("p010_lanczos_half", "p010le", 10, O.SC_P010, O.SC_LANCZOS3, S.SWS_LANCZOS, (256, 144), (128, 72)),
    ("p010_bilinear_up", "p010le", 10, O.SC_P010, O.SC_BILINEAR, S.SWS_BILINEAR, (80, 46), (142, 80)),
    ("nv12_lanczos_ragged", "nv12", 8, O.SC_NV12, O.SC_LANCZOS3, S.SWS_LANCZOS, (130, 70), (97, 53)),
]


def main():
    out, names = {}, []
    for name, sfmt, bits, fmt, filt, flag, src, dst in CASES:
        pic = S.smooth_picture(bits, src[0], src[1], seed=11)
        ours = O.scale_convert(fmt, filt, pic, src[0], src[1], dst[0], dst[1])
        sc = S.Scaler(sfmt, src[0], src[1], dst[0], dst[1], flag | S.SWS_ACCURATE_RND | S.SWS_FULL_CHR_H_INT)
        ref = sc.run(pic).copy()
        sc.close()
        d = np.abs(ours.astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= 1, (name, d.max())
        names.append(name)
        out[f"{name}/meta"] = np.array([fmt, filt, src[0], src[1], dst[0], dst[1]], dtype=np.int32)
        out[f"{name}/src"] = pic
        out[f"{name}/definition"] = ours
        out[f"{name}/swscale"] = ref
    out["names"] = np.array(names)
    path = Path(__file__).resolve().parent / "scale_cases.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size} bytes, {len(names)} cases)")


if __name__ == "__main__":
    main()
