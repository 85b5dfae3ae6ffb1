"""CPU, world_size 2 over gloo: the multi-rank plumbing of bench.py -- stream sharding is a partition, the job's step
time is the max over ranks, and under `--impl reference` only rank 0 works."""
import json
import os
import subprocess
import sys
import textwrap
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _torchrun(script, nproc=2, timeout=240):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", script]
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_shard_and_max_over_ranks(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import json, os, sys
        sys.path.insert(0, {str(ROOT)!r})
        import torch.distributed as dist
        import bench
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        dist.init_process_group("gloo")
        mine = bench.shard_streams(rank, world)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        ms = bench.max_over_ranks(10.0 + 5.0 * rank, world, dist)
        if rank == 0:
            print("RESULT " + json.dumps({{"streams": gathered, "ms": ms}}))
        dist.destroy_process_group()
    """))
    r = _torchrun(str(script))
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0]
    res = json.loads(line[7:])
    flat = [s for part in res["streams"] for s in part]
    assert flat == list(range(16)) and len(res["streams"][0]) == 8    # a partition, 8 streams per rank
    assert res["ms"] == 15.0                                           # the slowest rank decides


def test_reference_arm_runs_on_rank0_only(tmp_path):
    """`bench.py --impl reference` under torchrun: rank 0 prints the line, the other ranks exit 0 without work.
    (steps kept tiny: one 4K 8-layer frame costs about a second on this box's cores.)"""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29732", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["unit"] == "frames/s"


def test_both_arms_print_the_same_config():
    """`bench.py --impl ours` and `--impl reference` describe the workload with one function: the driver compares the two lines' `config`."""
    import argparse
    sys.path.insert(0, str(ROOT))
    import bench
    args = argparse.Namespace(pip_opacity=None, rgba_pips=0, format="nv12", mode="fused")
    for world in (1, 2, 8):
        a, b = bench.workload_config(args, world), bench.workload_config(args, world)
        assert a == b and a["streams_total"] == bench.STREAMS_PER_GPU * world
        assert a["workload"].startswith("cfg4: 3840x2160 NV12 target, 8 NV12 layers")
    src = (ROOT / "bench.py").read_text()
    assert src.count('"config": workload_config(args, world)') == 2  # the line of either arm
