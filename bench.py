#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric: 4K 8-layer composite frames/s (+ achieved HBM GB/s) per B200.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 our CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W]  the reference's kernels on the host cores

Workload (BASELINE.md cfg 4): every stream is one VideoMixer with a 3840x2160 NV12 target and 8
NV12 layers -- one 3840x2160 full-canvas (opacity 1) and seven 1920x1080 scaled to 1600x900 (opacity 0.55..0.85).
8 streams per GPU (64 over 8 GPUs), frames of different streams are independent: streams are sharded over
ranks with no collective on the data path ("scaling": "weak").  One STEP = one tick of every stream of this
rank = 8 composited 4K frames.

  value   frames/s with layers resident in HBM: per step the layers are pushed to their mixers and all 8
          mixers are folded into one fused launch (descriptor upload + kernel), no host sync inside the region.
  e2e     the same metric through the public call sequence a SwiftVideo pipeline makes, with HOST buffers:
          uploadComputePicture of every layer (pinned host -> device), VideoMixer.mix, downloadComputePicture of
          the composited frame (device -> pinned host), all inside the timed region.
  roofline  algorithmic bytes of one launch (46 656 000 B/frame x 8 frames) / that launch's device time
          (CUDA events around the kernel on its stream) against the measured HBM copy bandwidth.
  cpu_baseline / --impl reference
          the reference's own OpenCL kernel text compiled for the host (oracle/_ref) -- or the C port when that
          build is absent -- folding the same 4K 8-layer frame on all host threads.

L2 hygiene: the 8 streams of a step read 8 x 34.2 MB of distinct sources and write 8 x 12.4 MB of distinct
targets (373 MB per step, the backing ring makes it 10 distinct targets per stream), three times the 126 MB L2;
sources alternate between two device copies on successive steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

CANVAS = (3840, 2160)
NLAYERS = 8
STREAMS_PER_GPU = 8
ALG_BYTES_PER_FRAME = 12441600 + 7 * 3110400 + 12441600  # sources read once + target written once
DEFAULT_COMPOSITOR = "svb_mix_ring"  # what MixMode.FUSED runs unless SVB_COMPOSITOR says otherwise (mix_video.cpp: compositorChoice)
METRIC = "4K 8-layer composite frames/sec"
WORKLOAD = ("cfg4: 3840x2160 NV12 target, 8 NV12 layers (1x 3840x2160 full canvas + 7x 1920x1080 -> 1600x900), "
            f"{STREAMS_PER_GPU} streams per GPU")


def shard_streams(rank, world, per_gpu=STREAMS_PER_GPU):
    """Global stream ids owned by `rank`: streams are independent, so sharding is a plain partition (no collective)."""
    return list(range(rank * per_gpu, (rank + 1) * per_gpu))


def max_over_ranks(ms, world, dist=None, device=None):
    """The step time of the job is the slowest rank's device time."""
    if world <= 1:
        return float(ms)
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def geometry():
    import scenes
    if CANVAS == (1280, 720):  # cfg 2: one 1080p picture over the whole 720p canvas
        return [((1920, 1080), (0, 0), (1280, 720), 1.0)]
    return scenes.cfg34_geometry(NLAYERS)


def configure_workload(name):
    """The headline is cfg 4; cfg 3 and cfg 2 (BASELINE.md's other compositor rows) run through the same code as side workloads."""
    global CANVAS, NLAYERS, ALG_BYTES_PER_FRAME, WORKLOAD
    if name == "cfg3":
        CANVAS, NLAYERS = (3840, 2160), 4
        ALG_BYTES_PER_FRAME = 12441600 + 3 * 3110400 + 12441600
        WORKLOAD = (f"cfg3 (side workload, not the headline): 3840x2160 NV12 target, 4 NV12 layers (1x 3840x2160 full canvas + 3x 1920x1080 -> 1600x900), {STREAMS_PER_GPU} streams per GPU")
    elif name == "cfg2mix":
        CANVAS, NLAYERS = (1280, 720), 1
        ALG_BYTES_PER_FRAME = 3110400 + 1382400
        WORKLOAD = (f"cfg2 (side workload, not the headline): 1280x720 NV12 target, one 1920x1080 NV12 picture over the whole canvas (clear + img_nv12_nv12), {STREAMS_PER_GPU} streams per GPU "
                    "-- 36 MB per step: FITS in the 126 MB L2")


def workload_config(args, world):
    """The `config` object of the JSON line: what is measured, identical for both arms (`--impl ours` / `--impl reference`)."""
    S = STREAMS_PER_GPU
    return {"workload": (WORKLOAD if args.pip_opacity is None else WORKLOAD + f" -- NOT the headline: picture-in-picture opacity forced to {args.pip_opacity}") +
                        (f" -- NOT the headline: the top {args.rgba_pips} pictures-in-picture are RGBA overlays" if args.rgba_pips else "") +
                        (" -- NOT the headline: YUV420P layers and target instead of NV12" if args.format == "y420p" else ""),
            "mode": args.mode, "streams_total": S * world, "frames_per_step": S * world, "parallelism": f"streams sharded, {S}/GPU, no collective",
            "l2": (f"{ALG_BYTES_PER_FRAME * S / 1e6:.0f} MB of distinct sources+targets per step (> 126 MB L2); sources alternate between two device copies"
                   if ALG_BYTES_PER_FRAME * S > 126e6 else
                   f"{ALG_BYTES_PER_FRAME * S / 1e6:.0f} MB of distinct sources+targets per step: FITS in the 126 MB L2; sources alternate between two device copies "
                   f"({2 * ALG_BYTES_PER_FRAME * S / 1e6:.0f} MB over two steps)"),
            "bit_exact_vs_oracle": "tests/test_gpu_parity.py::test_cfg2_full_size" if CANVAS == (1280, 720) else "tests/test_gpu_parity.py::test_cfg34_full_size"}


def e2e_limiter(e2e_fps, link, host_ms, step_ms, resident_step_ms, world):
    """Name what bounds the end-to-end number, from what the same process measured: the host link with every rank copying at once
    (PCIe and the host-memory path behind it, shared by the ranks of a box), the calling thread, or the GPU itself."""
    ceiling = link["duplex_all_ranks"]["frames_per_s"]
    if e2e_fps >= 0.85 * ceiling:
        return (f"host link: {e2e_fps / ceiling:.2f} of what the box moves with all {world} rank(s) copying one step's bytes both ways at once "
                f"({link['duplex_all_ranks']['h2d_gbs_per_gpu']} GB/s in + {link['duplex_all_ranks']['d2h_gbs_per_gpu']} GB/s out per GPU); "
                f"the GPU needs {resident_step_ms:.2f} ms of the {step_ms:.2f} ms step, the calling thread {host_ms:.2f} ms")
    if host_ms >= 0.8 * step_ms:
        return f"calling thread (vCPU): busy {host_ms:.2f} ms of the {step_ms:.2f} ms step"
    if resident_step_ms >= 0.8 * step_ms:
        return f"GPU: the resident step alone takes {resident_step_ms:.2f} ms of {step_ms:.2f} ms"
    return f"unattributed: {e2e_fps / ceiling:.2f} of the link ceiling, calling thread {host_ms:.2f} ms, GPU {resident_step_ms:.2f} ms of the {step_ms:.2f} ms step"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa(index):
    """Best effort: run this rank (and first-touch its pinned buffers) on the CPU cores next to its GPU, so that host<->device
    copies do not cross the socket interconnect.  Returns a short description for the bench line."""
    try:
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)], capture_output=True, text=True,
                             timeout=20).stdout.strip().lower()
        if bdf.startswith("00000000:"):
            bdf = bdf[4:]
        base = Path("/sys/bus/pci/devices") / bdf
        node = int((base / "numa_node").read_text().strip())
        cpulist = (base / "local_cpulist").read_text().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= set(os.sched_getaffinity(0))
        if node >= 0 and cpus:
            os.sched_setaffinity(0, cpus)
            return f"numa node {node} ({len(cpus)} cpus)"
        # no NUMA information (these VMs report numa_node -1): give every rank of the node its own even share of the allowed cores, so
        # that eight ranks do not pile their copy-issuing threads onto the same few
        world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", 1)))
        allowed = sorted(os.sched_getaffinity(0))
        if world > 1 and len(allowed) >= world:
            per = len(allowed) // world
            mine = set(allowed[index * per:(index + 1) * per])
            os.sched_setaffinity(0, mine)
            return f"numa node {node}: even split, cores {min(mine)}-{max(mine)} of {len(allowed)}"
        return f"numa node {node} (not bound: single rank)"
    except Exception as e:
        return f"unbound ({type(e).__name__})"


# ---- our arm -------------------------------------------------------------------------------------------------

def measure_link(torch, device, nbytes=256 << 20, reps=5):
    """Pinned-memory copy rate of this box's host link, each direction alone (best of `reps`, CUDA events): the bound of the
    e2e leg, whose every step moves its inputs in and its results out."""
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=device)
    out = {}
    for name, (dst, src) in {"h2d_gbs": (d, h), "d2h_gbs": (h, d)}.items():
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dst.copy_(src, non_blocking=True)
            b.record()
            b.synchronize()
            best = min(best, a.elapsed_time(b))
        out[name] = round(nbytes / (best / 1e3) / 1e9, 1)
    return out


def measure_link_duplex(torch, device, in_bytes, out_bytes, barrier, reps=6):
    """What the e2e leg can reach at most on this box: every rank copies one step's input bytes host -> device and one step's output
    bytes device -> host AT THE SAME TIME (two streams, pinned memory), all ranks together (`barrier` lines them up).  Returns this
    rank's seconds per step-equivalent (best of reps); the caller takes the max over ranks."""
    hi = torch.empty(in_bytes, dtype=torch.uint8).pin_memory()
    di = torch.empty(in_bytes, dtype=torch.uint8, device=device)
    ho = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
    do = torch.empty(out_bytes, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)
    best = 1e30
    for _ in range(reps):
        barrier()
        a, b, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_event(a), s2.wait_event(a)
        with torch.cuda.stream(s1):
            di.copy_(hi, non_blocking=True)
            b.record()
        with torch.cuda.stream(s2):
            ho.copy_(do, non_blocking=True)
            c.record()
        b.synchronize(), c.synchronize()
        best = min(best, max(a.elapsed_time(b), a.elapsed_time(c)) / 1e3)
    return best


def run_ours(args):
    import torch
    import torch.distributed as dist

    import scenes
    import swiftvideo_b200 as sv
    from oracle import oracle as O

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local) if not args.no_numa_bind else "unbound (--no-numa-bind)"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sv.make_compute_context(local)
    geo = geometry()
    S = STREAMS_PER_GPU
    rng_base = 1000 * 4
    yuv_fmt = sv.Y420P if args.format == "y420p" else sv.NV12  # same bytes per picture; y420p is what SwiftVideo composes on Linux (composer.swift:52-56)

    # ---- synthetic layers: pinned host pictures (e2e source) and two device copies (device-resident source)
    host = [[None] * NLAYERS for _ in range(S)]
    dev = [[[None] * NLAYERS for _ in range(S)] for _ in range(2)]
    for s in range(S):
        gstream = shard_streams(rank, world, S)[s]
        for k, (ssz, pos, dsz, op) in enumerate(geo):
            if args.pip_opacity is not None and k > 0:
                op = args.pip_opacity
            rng = np.random.default_rng(rng_base + 16 * gstream + k)
            rgba = k >= NLAYERS - args.rgba_pips  # side experiment: the topmost pictures-in-picture as RGBA overlays (text / logo layers)
            h = sv.create_picture_sample(ssz[0], ssz[1], sv.RGBA if rgba else yuv_fmt, f"s{gstream}l{k}", "bench", pinned_from=ctx)
            h.set_host_bytes(rng.integers(0, 256, size=ssz[0] * ssz[1] * (4 if rgba else 1) * (2 if rgba else 3) // 2, dtype=np.uint8))
            # PictureAnimator.impl in native code: matrix = ortho(canvas) * T(pos) * S(size), opacity = 1 - transparency
            host[s][k] = h.animate(CANVAS, (pos[0], pos[1], float(k)), dsz, transparency=1.0 - op)
            for c in range(2):
                dev[c][s][k] = host[s][k].upload(ctx)
    mixers = [sv.VideoMixer(ctx, CANVAS[0], CANVAS[1], yuv_fmt, asset_id=f"mixer{rank * S + s}", workspace_id="bench") for s in range(S)]
    mode = {"fused": sv.MixMode.FUSED, "generic": sv.MixMode.GENERIC, "per_layer": sv.MixMode.PER_LAYER, "fused_gather": sv.MixMode.FUSED_GATHER, "fused_tiled": sv.MixMode.FUSED_TILED, "fused_ring": sv.MixMode.FUSED_RING}[args.mode]
    for m in mixers:
        m.set_mode(mode)
    for i in range(10):  # setup, untimed: fill every mixer's backing ring (10 targets, allocated on first use upstream too)
        sv.VideoMixer.mix_many(mixers, -1 - i, wait=False)
    ctx.synchronize()

    def step_resident(i):
        for s in range(S):
            mixers[s].push_many(dev[i & 1][s])
        return sv.VideoMixer.mix_many(mixers, i, wait=False)

    def step_e2e(i):  # host buffers in, host buffers out: ONE call across the boundary (upload, push, mix, download of all S mixers)
        return sv.VideoMixer.tick_many(mixers, host, i, wait=False)

    def step_per_mixer(i):  # the reference's own call pattern: one VideoMixer.mix(at:) per mixer (mix.video.swift:95-99) = one frame per launch
        outs = []
        for s in range(S):
            mixers[s].push_many(dev[i & 1][s])
            outs.append(mixers[s].mix(i, wait=False))
        return outs

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, timing=True):
        last = None
        for i in range(warmup):
            last = fn(i)  # (kept until the next step returns, exactly like the timed steps: the buffer pools then reach their steady state here)
        ctx.synchronize()
        barrier()
        if timing:
            ctx.launch_timing(True)  # (re-)start the per-launch device timing and the in-library host timing: the warm-up's first-use costs stay out
        timer = sv.Timer(ctx)
        launches0 = sv.kernel_launch_count()
        timer.start()
        last = None
        h0 = time.perf_counter()
        for i in range(steps):
            last = fn(warmup + i)
        host_ms = (time.perf_counter() - h0) * 1e3  # time the host needs to queue the steps (device runs behind it)
        timer.stop()
        ms = timer.elapsed_ms()
        for o in last:
            o.wait()
        barrier()
        launches = sv.kernel_launch_count() - launches0
        timer.close()
        ms = max_over_ranks(ms, world, dist, "cuda")
        return ms, launches, host_ms

    # clocks and throttle reasons are sampled (nvidia-smi, every 50 ms) from here to the end of the e2e leg: the GPU is
    # busy throughout, and the resident leg alone can be shorter than one sampling period
    # set-up, untimed: one turn of the library's eight table buffers with the workload's geometry, so that the warm-up and the timed steps
    # see what a mixer that has been running for more than eight ticks sees (coordinate tables in place, `table_cache` below)
    last = None
    for i in range(8):
        last = step_resident(-100 - i)
    for o in last:
        o.wait()
    last = None
    ctx.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    # ---- value: layers resident in HBM
    ctx.launch_timing(True)
    ms, launches, host_ms = timed(step_resident, args.steps, args.warmup)
    kern_ms, kern_n = ctx.launch_timing_read()
    host_c_ms, host_c_calls, host_c_wait = ctx.host_timing_read(with_wait=True)
    ctx.launch_timing(False)
    frames = S * world * args.steps
    value = frames / (ms / 1e3)
    # (the same leg with the library's table cache off -- every launch runs its coordinate-table pre-pass, as the library did before the
    # cache existed -- so that the line shows what the cache is worth and nothing hides behind it: it keeps tables derived from the layers'
    # UNIFORMS while those do not change, never pixels; every frame is composited in full in either leg)
    ctx.table_cache(False)
    nc_ms, _, _ = timed(step_resident, args.steps, args.warmup, timing=False)
    ctx.table_cache(True)

    # ---- the same, one frame per launch (S launches per step): the reference's call pattern
    ctx.launch_timing(True)
    pm_ms, _, pm_host_ms = timed(step_per_mixer, args.steps, args.warmup)
    pm_kern_ms, pm_kern_n = ctx.launch_timing_read()
    pm_c_ms, pm_c_calls, pm_c_wait = ctx.host_timing_read(with_wait=True)
    ctx.launch_timing(False)

    # ---- e2e: host buffers in, host buffers out
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    if os.environ.get("SVB_DEBUG_POOL"):
        print("[bench] e2e leg starts", file=sys.stderr, flush=True)
    e_ms, _, e_host_ms = timed(step_e2e, e2e_steps, max(4, min(args.warmup, 6)), timing=False)
    if os.environ.get("SVB_DEBUG_POOL"):
        print("[bench] e2e leg ends", file=sys.stderr, flush=True)
    clocks = sampler.stop()
    e2e_value = S * world * e2e_steps / (e_ms / 1e3)
    h2d = S * sum(ssz[0] * ssz[1] * 3 // 2 for ssz, _, _, _ in geo)
    d2h = S * CANVAS[0] * CANVAS[1] * 3 // 2

    link = measure_link(torch, torch.device("cuda", local))
    # the ceiling of the e2e leg, measured: every rank moves one step's bytes in and out at once, all ranks together
    duplex_s = measure_link_duplex(torch, torch.device("cuda", local), h2d, d2h, barrier)
    duplex_s = max_over_ranks(duplex_s * 1e3, world, dist, "cuda") / 1e3
    link["duplex_all_ranks"] = {"frames_per_s": round(S * world / duplex_s, 1), "h2d_gbs_per_gpu": round(h2d / duplex_s / 1e9, 1), "d2h_gbs_per_gpu": round(d2h / duplex_s / 1e9, 1),
                                "what": "one step's input bytes host->device and output bytes device->host at the same time on two streams, every rank at once (best of 6)"}
    # both directions run at once (separate copy engines); indicative only -- the H2D rate of one big copy varies run to run (39-53 GB/s seen)
    link["frames_per_s_at_this_copy_rate"] = round(S * world / max(h2d / (link["h2d_gbs"] * 1e9), d2h / (link["d2h_gbs"] * 1e9)), 1)

    peak, peak_src = peaks()
    roof = None
    if kern_n and args.mode != "per_layer":  # the per-layer sequence has no single dominant launch to put on a roofline
        per_launch_ms = kern_ms / kern_n
        # warm-up launches are inside the timing window too; they run the same work, so the average stands
        achieved = ALG_BYTES_PER_FRAME * S / (per_launch_ms / 1e3) / 1e9
        roof = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": None, "peak_source": peak_src, "kernel": {"fused": {"gather": "svb_mix_gather", "tma": "svb_mix_tiled", "tiled": "svb_mix_tiled", "ring": "svb_mix_ring"}.get(os.environ.get("SVB_COMPOSITOR", ""), DEFAULT_COMPOSITOR), "fused_gather": "svb_mix_gather", "fused_tiled": "svb_mix_tiled", "fused_ring": "svb_mix_ring"}.get(args.mode, "svb_mix_generic"), "kernel_ms_per_launch": round(per_launch_ms, 4),
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_FRAME * S, "launches_timed": int(kern_n)}
        tr = ROOT / "profiles" / "traffic.json"
        if tr.exists() and args.mode.startswith("fused"):
            try:
                roof["traffic"] = json.loads(tr.read_text()).get(roof["kernel"] + "_dram_bytes_per_launch")
            except Exception:
                pass

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms / args.steps, 4),
        # wall time the calling thread spends per step; it INCLUDES back-pressure: the host may run eight launches ahead, then it waits for the GPU
        "host_queue_ms_per_step": round(host_ms / args.steps, 4),
        # the C++ side alone (planner + driver calls) per compose call, back-pressure excluded and shown beside it
        "host_queue_in_library_ms_per_step": round(host_c_ms / max(1, host_c_calls), 4),
        "host_backpressure_ms_per_step": round(host_c_wait / max(1, host_c_calls), 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 (fp32 arithmetic, no FMA contraction)", "data": "synthetic (uniform u8 planes, seeded)",
        "config": workload_config(args, world), "host_affinity": numa,
        "table_cache": {"enabled": True, "what": "coordinate tables of a batch are kept while its layers' uniforms, sizes and formats do not change (svb_table_cache, include/svb200.h); pixels are never cached",
                        "primed_in_setup_steps": 8,
                        "value_without": round(frames / (nc_ms / 1e3), 2), "ms_per_step_without": round(nc_ms / args.steps, 4)},
        "e2e": {"value": round(e2e_value, 2), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": round(e_ms / e2e_steps, 4), "host_queue_ms_per_step": round(e_host_ms / e2e_steps, 4), "calls_per_step": 1,
                "frac_of_link_ceiling": round(e2e_value / link["duplex_all_ranks"]["frames_per_s"], 3), "host_link": link,
                "limiter": e2e_limiter(e2e_value, link, e_host_ms / e2e_steps, e_ms / e2e_steps, ms / args.steps, world)},
        # the reference's own call pattern, for comparison with the headline: one VideoMixer.mix(at:) per mixer = one 4K frame per launch
        "one_frame_per_launch": {"value": round(frames / (pm_ms / 1e3), 2), "unit": "frames/s", "ms_per_step": round(pm_ms / args.steps, 4),
                                 "launches_per_step": S, "kernel_ms_per_launch": round(pm_kern_ms / max(1, pm_kern_n), 4),
                                 "host_queue_ms_per_step": round(pm_host_ms / args.steps, 4),
                                 "host_queue_in_library_ms_per_launch": round(pm_c_ms / max(1, pm_c_calls), 4),
                                 "host_backpressure_ms_per_launch": round(pm_c_wait / max(1, pm_c_calls), 4)},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
        "hbm_gbs_per_gpu_algorithmic": round(ALG_BYTES_PER_FRAME * value / world / 1e9, 1),
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(budget_s=args.cpu_seconds)
        try:
            line["cpu_baseline"]["swscale_scale_stage_only"] = swscale_scale_stage()
        except Exception as e:  # context figure only
            line["cpu_baseline"]["swscale_scale_stage_only"] = {"unavailable": str(e)[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    for m in mixers:
        m.close()
    if world > 1:
        dist.destroy_process_group()



# ---- side workloads: the convert+scale operator (BASELINE cfg 5 and cfg 2's convert leg) -- NOT the headline --------------

SCALE_WORKLOADS = {
    # name: (source format, filter, source size, destination size, swscale names)
    "cfg5": ("p010", "lanczos3", (3840, 2160), (1920, 1080)),
    "cfg2": ("nv12", "bilinear", (1920, 1080), (1280, 720)),
    "cfg2chain": ("nv12", "bilinear", (1920, 1080), (1280, 720)),
}


def run_scale(args):
    """One step = 8 frames of `svb_scale_convert` (one launch each) over 8 distinct device-resident sources; value, e2e,
    roofline and cpu_baseline (libswscale on the host cores, accurate mode off = its default speed path) as for the headline."""
    import torch

    import swiftvideo_b200 as sv
    import swscale_util as SW
    from oracle import oracle as O
    fmt_name, filt_name, (sw, sh), (dw, dh) = SCALE_WORKLOADS[args.workload]
    fmt = sv.P010 if fmt_name == "p010" else sv.NV12
    filt = sv.FILTER_LANCZOS3 if filt_name == "lanczos3" else sv.FILTER_BILINEAR
    bps = 2 if fmt_name == "p010" else 1
    src_bytes, dst_bytes = sw * sh * 3 // 2 * bps, dw * dh * 4
    torch.cuda.set_device(0)
    ctx = sv.make_compute_context(0)
    N = 8
    rng = np.random.default_rng(5000)
    hosts, devs = [], []
    for i in range(N):
        h = sv.create_picture_sample(sw, sh, fmt, f"src{i}", "bench", pinned_from=ctx)
        data = rng.integers(0, 256, size=src_bytes, dtype=np.uint8) if bps == 1 else (rng.integers(0, 1024, size=src_bytes // 2, dtype=np.uint16) << 6).view(np.uint8)
        h.set_host_bytes(data)
        hosts.append(h)
        devs.append(h.upload(ctx))
    ctx.synchronize()

    chain = args.workload == "cfg2chain"
    mixers = [sv.VideoMixer(ctx, dw, dh, sv.NV12, asset_id=f"mixer{k}", workspace_id="bench") for k in range(N)] if chain else []

    def compose_back(bgra, i):
        """cfg 2's chain: the scaled BGRA pictures go back to NV12 through the reference's img_bgra_nv12 (the mixer, one layer)"""
        for k in range(N):
            mixers[k].push(bgra[k].animate((dw, dh), (0, 0, 0.0), (dw, dh), revision="src"))
        return sv.VideoMixer.mix_many(mixers, 1000 * (i + 1), wait=False)

    def step_resident(i):
        outs = [devs[k].scale_convert(ctx, dw, dh, sv.BGRA, filt, wait=False) for k in range(N)]
        return compose_back(outs, i) if chain else outs

    def step_e2e(i):
        outs = [hosts[k].upload(ctx, retain_cpu_buffer=False, wait=False).scale_convert(ctx, dw, dh, sv.BGRA, filt, wait=False) for k in range(N)]
        if chain:
            outs = compose_back(outs, i)
        return [o.download(ctx, retain_gpu_buffer=True, wait=False) for o in outs]

    def timed(fn, steps, warmup):
        last = None
        for i in range(warmup):  # same shape as the timed loop (the previous step's outputs die while this step is queued), so that
            last = fn(i)         # the device and pinned-host pools hold every block the loop needs before the clock starts
        ctx.synchronize()
        torch.cuda.synchronize()
        timer = sv.Timer(ctx)
        l0 = sv.kernel_launch_count()
        timer.start()
        for i in range(steps):
            last = fn(warmup + i)
        timer.stop()
        ms = timer.elapsed_ms()
        for o in last:
            o.wait()
        ctx.synchronize()
        timer.close()
        return ms, sv.kernel_launch_count() - l0

    sampler = ClockSampler(0)
    sampler.start()
    # (the chain's mixers fill their backing rings of ten targets during warm-up: device allocations synchronise)
    ms, launches = timed(step_resident, args.steps, max(args.warmup, 12) if chain else args.warmup)
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    e_ms, _ = timed(step_e2e, e2e_steps, 3)
    clocks = sampler.stop()
    value = N * args.steps / (ms / 1e3)
    per_launch_ms = ms / (N * args.steps)
    peak, peak_src = peaks()
    alg = src_bytes + dst_bytes
    if chain:  # + the BGRA picture read back and the NV12 frame written by the compositor
        alg += dst_bytes + dw * dh * 3 // 2
        dst_bytes = dw * dh * 3 // 2
    achieved = alg / (per_launch_ms / 1e3) / 1e9
    line = {
        "metric": f"{sw}x{sh} {fmt_name} -> {dw}x{dh} bgra {filt_name} frames/sec (side workload, not the headline)", "value": round(value, 2), "unit": "frames/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/u16 in, fp32 fused multiply-add arithmetic, u8 out", "data": "synthetic (uniform random codes, seeded)",
        "config": {"workload": f"{args.workload}: {sw}x{sh} {fmt_name} -> {dw}x{dh} BGRA, {filt_name}, {N} frames per step over {N} distinct sources" +
                               (" -> NV12 through the compositor's img_bgra_nv12 (BASELINE cfg 2's convert -> scale -> convert chain: 2 kernels + the table pre-pass per step)" if chain else ""),
                   "l2": f"{N * (src_bytes + dst_bytes) / 1e6:.0f} MB of distinct sources+targets per step (> 126 MB L2)" if N * (src_bytes + dst_bytes) > 126e6 else
                         f"{N * (src_bytes + dst_bytes) / 1e6:.0f} MB per step: FITS in the 126 MB L2 (the sources are re-read from L2, not HBM)",
                   "bit_exact_vs_oracle": "tests/test_scale.py::test_cfg2_chain_full_size" if chain else "tests/test_scale.py::test_gpu_scale_full_size"},
        "e2e": {"value": round(N * e2e_steps / (e_ms / 1e3), 2), "unit": "frames/s", "h2d_bytes_per_step": N * src_bytes, "d2h_bytes_per_step": N * dst_bytes,
                "steps": e2e_steps, "ms_per_step": round(e_ms / e2e_steps, 4)},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                     "kernel": "svb_scale_convert" + (" + svb_mix_ring" if chain else ""), "kernel_ms_per_launch": round(per_launch_ms, 4), "algorithmic_bytes_per_launch": alg,
                     "note": ("per FRAME of the chain (scale launch + its share of the compose launch), = step time / frames" if chain else
                              "launch time = step time / launches (back to back on one stream, gaps included)")},
    }
    if not args.no_cpu_baseline and SW.available():
        # libswscale, the comparator BASELINE.json names, one SwsContext per host thread, default (fast) mode
        from concurrent.futures import ThreadPoolExecutor
        threads = min(O.host_threads(), 64)
        pic = hosts[0].host_bytes().copy()
        flag = SW.SWS_LANCZOS if filt_name == "lanczos3" else SW.SWS_BILINEAR

        def worker(_):
            sc = SW.Scaler("p010le" if bps == 2 else "nv12", sw, sh, dw, dh, flag)
            n, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < min(args.cpu_seconds, 6.0):
                sc.run(pic)
                n += 1
            dt = time.perf_counter() - t0
            sc.close()
            return n / dt

        with ThreadPoolExecutor(threads) as ex:
            fps = sum(ex.map(worker, range(threads)))
        t0 = time.perf_counter()
        O.scale_convert(O.SC_P010 if bps == 2 else O.SC_NV12, O.SC_LANCZOS3 if filt_name == "lanczos3" else O.SC_BILINEAR, pic, sw, sh, dw, dh)
        line["cpu_baseline"] = {"value": round(fps, 2), "unit": "frames/s", "cores": threads, "kind": "libswscale 9.1 (external comparator; the reference has no such operator)",
                                "sample": f"{min(args.cpu_seconds, 6.0):.0f} s of whole frames per thread, one SwsContext per thread",
                                "definition_single_thread_frames_per_s": round(1.0 / (time.perf_counter() - t0), 3)}
    print(json.dumps(line), flush=True)


# ---- the reference's kernels on the host cores ---------------------------------------------------------------

def cpu_scene():
    import scenes
    from oracle import oracle as O
    canvas, tf, layers, us = scenes.cfg2_scene() if CANVAS == (1280, 720) else scenes.cfg34_scene(NLAYERS)
    return O.Image(tf, canvas[0], canvas[1]), layers, us


def cpu_reference(budget_s=15.0, steps=None, warmup=0):
    """Frames/s of the reference kernels (oracle/_ref: the reference's OpenCL kernel text compiled for the host;
    falls back to the C port) on all host threads, one 4K 8-layer frame per step."""
    from oracle import oracle as O
    lib, kind = O.best()
    threads = O.host_threads()
    target, layers, us = cpu_scene()
    for _ in range(warmup):
        lib.mix(target, layers, us, threads=threads)
    t0 = time.perf_counter()
    assert lib.mix(target, layers, us, threads=threads) == 0
    first = time.perf_counter() - t0
    if steps is None:
        steps = int(max(1, min(40, budget_s / max(first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.mix(target, layers, us, threads=threads)
    dt = time.perf_counter() - t0
    return {"value": round(steps / dt, 3), "unit": "frames/s", "cores": threads, "kind": kind,
            "sample": f"{steps} frames of one stream of the workload (3840x2160, 8 layers), rows split over {threads} threads",
            "ms_per_frame": round(dt / steps * 1e3, 2)}


def swscale_scale_stage(seconds=3.0):
    """Context only (BASELINE.json names swscale; the reference itself never calls it): FFmpeg libswscale
    (the copy bundled in the OpenCV wheel, via ctypes) doing just the SCALE stage of one frame of the workload -- one
    3840x2160 NV12 1:1 pass and seven 1920x1080 -> 1600x900 NV12 bilinear scales, no blending (swscale has none) -- on all
    host threads, one frame per thread at a time."""
    import ctypes as C
    import glob
    from concurrent.futures import ThreadPoolExecutor

    import cv2
    from oracle import oracle as O
    libdir = os.path.join(os.path.dirname(cv2.__file__), "..", "opencv_python_headless.libs")
    avutil = C.CDLL(glob.glob(libdir + "/libavutil*")[0], mode=C.RTLD_GLOBAL)
    sws = C.CDLL(glob.glob(libdir + "/libswscale*")[0])
    avutil.av_get_pix_fmt.argtypes = [C.c_char_p]
    nv12 = avutil.av_get_pix_fmt(b"nv12")
    sws.sws_getContext.restype = C.c_void_p
    sws.sws_getContext.argtypes = [C.c_int] * 7 + [C.c_void_p] * 3
    sws.sws_scale.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    threads = min(O.host_threads(), 64)
    jobs = [((3840, 2160), (3840, 2160), 1), ((1920, 1080), (1600, 900), 7)]

    def worker(_):
        ctxs = []
        for (sw, sh), (dw, dh), n in jobs:
            ctx = sws.sws_getContext(sw, sh, nv12, dw, dh, nv12, 2, None, None, None)  # SWS_BILINEAR
            src = np.zeros(sw * sh * 3 // 2, np.uint8)
            dst = np.zeros(dw * dh * 3 // 2, np.uint8)
            sp = (C.c_void_p * 4)(src.ctypes.data, src.ctypes.data + sw * sh, None, None)
            dp = (C.c_void_p * 4)(dst.ctypes.data, dst.ctypes.data + dw * dh, None, None)
            ctxs.append((ctx, sp, (C.c_int * 4)(sw, sw, 0, 0), sh, dp, (C.c_int * 4)(dw, dw, 0, 0), n, src, dst))
        frames, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for ctx, sp, ss, sh, dp, ds, n, _, _ in ctxs:
                for _ in range(n):
                    sws.sws_scale(ctx, sp, ss, 0, sh, dp, ds)
            frames += 1
        return frames, time.perf_counter() - t0

    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(worker, range(threads)))
    fps = sum(f / dt for f, dt in res)
    return {"value": round(fps, 2), "unit": "frames/s", "cores": threads,
            "what": "libswscale 9.1 SWS_BILINEAR, scale stage only (8 NV12 scales per frame, no blend)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", 1))
    t0 = time.perf_counter()
    base = cpu_reference(steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": base["ms_per_frame"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 (fp32 arithmetic, no FMA contraction)", "data": "synthetic (uniform u8 planes, seeded)",
            "config": workload_config(args, world), "note": "host CPU only: one step = one 4K 8-layer frame of one stream of this workload (rows split over all host threads)",
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 2)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to the CPU cores local to its GPU")
    ap.add_argument("--mode", default="fused", choices=["fused", "fused_ring", "fused_tiled", "fused_gather", "generic", "per_layer"],
                    help="compose strategy: fused (default, svb_mix_ring), fused_tiled (svb_mix_tiled, the CTA-per-tile TMA compositor of round 1), fused_gather (svb_mix_gather: taps through the texture unit), "
                         "generic (svb_mix_generic), per_layer (the reference's own launch sequence over the drop-in kernels: clear + one "
                         "launch per layer)")
    ap.add_argument("--pip-opacity", type=float, default=None,
                    help="side experiment, not the headline: opacity of layers 1..7 (1.0 = opaque pictures, which let the planner skip "
                         "whatever they cover)")
    ap.add_argument("--format", default="nv12", choices=["nv12", "y420p"], help="side experiment, not the headline: y420p layers and target (the Linux Composer's format)")
    ap.add_argument("--rgba-pips", type=int, default=0, help="side experiment, not the headline: the topmost N pictures-in-picture are RGBA overlays")
    ap.add_argument("--workload", default="cfg4", choices=["cfg4", "cfg3", "cfg2mix", "cfg5", "cfg2", "cfg2chain"],
                    help="cfg4 (default) = the headline; cfg3 / cfg2mix = BASELINE.md's other compositor rows through the same code; "
                         "cfg5 / cfg2 = the convert+scale operator's side workloads (1 GPU, our arm only); cfg2chain = cfg 2's whole chain: convert+scale to BGRA, then "
                         "img_bgra_nv12 through the compositor")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload in ("cfg3", "cfg2mix"):
        configure_workload(args.workload)
    if args.workload in ("cfg5", "cfg2", "cfg2chain") and args.impl == "ours":
        run_scale(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
