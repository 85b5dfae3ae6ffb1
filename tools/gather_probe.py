"""2-GPU probe for svb_gather_picture: frames composited on GPU 1 gathered onto GPU 0 with peer copies.
   gpurun --gpus 2 -- 'python tools/gather_probe.py'   ->  one JSON line (frames/s, GB/s), not a bench.py value."""
import json
import sys
import time

from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

import swiftvideo_b200 as sv

W, H, N = 3840, 2160, 64
if sv.available_compute_devices() < 2:
    print(json.dumps({"unavailable": "needs two GPUs"}))
    sys.exit(0)
ctx0, ctx1 = sv.make_compute_context(0), sv.make_compute_context(1)
rng = np.random.default_rng(1)
frames = []
for i in range(8):
    p = sv.create_picture_sample(W, H, sv.NV12, f"f{i}", "w", pinned_from=ctx1)
    p.set_host_bytes(rng.integers(0, 256, W * H * 3 // 2, dtype=np.uint8))
    frames.append(p.upload(ctx1))
ref = frames[3].download(ctx1, retain_gpu_buffer=True).host_bytes().copy()
got = frames[3].gather(ctx0).download(ctx0).host_bytes()
assert (got == ref).all()
for _ in range(8):  # warm: pool blocks on GPU 0, peer access enabled
    [f.gather(ctx0, wait=False) for f in frames][-1].wait()
best = None
for rep in range(5):
    t0 = time.perf_counter()
    outs = [frames[i % 8].gather(ctx0, wait=False) for i in range(N)]
    for o in outs:
        o.wait()
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
    del outs
nbytes = W * H * 3 // 2
print(json.dumps({"what": "svb_gather_picture GPU1 -> GPU0, 4K NV12 frames, 64 queued then joined, best of 5 (host clock)",
                  "frames_per_s": round(N / best, 1), "gbs": round(N * nbytes / best / 1e9, 1), "bytes_per_frame": nbytes}))
