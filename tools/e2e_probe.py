"""Where the end-to-end tick spends its time (manual probe, not a test): per-object calls vs svb_video_mixer_tick_many, host time and device time."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import swiftvideo_b200 as sv
import bench
ctx = sv.make_compute_context(0)
geo = bench.geometry(); S = 8; NL = bench.NLAYERS; CANVAS = bench.CANVAS
host = [[None] * NL for _ in range(S)]
for s in range(S):
    for k, (ssz, pos, dsz, op) in enumerate(geo):
        h = sv.create_picture_sample(ssz[0], ssz[1], sv.NV12, f"s{s}l{k}", "b", pinned_from=ctx)
        h.set_host_bytes(np.random.default_rng(s * 16 + k).integers(0, 256, size=ssz[0] * ssz[1] * 3 // 2, dtype=np.uint8))
        host[s][k] = h.animate(CANVAS, (pos[0], pos[1], float(k)), dsz, transparency=1.0 - op)
mixers = [sv.VideoMixer(ctx, CANVAS[0], CANVAS[1], sv.NV12, asset_id=f"m{s}", workspace_id="b") for s in range(S)]
def per_object(i):
    for s in range(S):
        mixers[s].push_many([host[s][k].upload(ctx, retain_cpu_buffer=False, wait=False) for k in range(NL)])
    outs = sv.VideoMixer.mix_many(mixers, i, wait=False)
    return [o.download(ctx, retain_gpu_buffer=True, wait=False) for o in outs]
def one_call(i):
    return sv.VideoMixer.tick_many(mixers, host, i, wait=False)
def upload_only(i):
    return [[host[s][k].upload(ctx, retain_cpu_buffer=False, wait=False) for k in range(NL)] for s in range(S)]
def upload_mix(i):
    for s in range(S):
        mixers[s].push_many([host[s][k].upload(ctx, retain_cpu_buffer=False, wait=False) for k in range(NL)])
    return sv.VideoMixer.mix_many(mixers, i, wait=False)
for name, fn in (("upload_only", upload_only), ("upload_mix", upload_mix), ("per_object", per_object), ("one_call", one_call), ("per_object", per_object), ("one_call", one_call)):
    for i in range(4): last = fn(i)
    ctx.synchronize()
    t0 = time.perf_counter(); hq = 0.0
    N = 12
    for i in range(N):
        h0 = time.perf_counter(); last = fn(4 + i); hq += time.perf_counter() - h0
    ctx.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name:12s} {dt / N * 1e3:7.2f} ms/step  host {hq / N * 1e3:6.2f} ms/step  -> {S * N / dt:7.1f} frames/s", flush=True)
    del last
