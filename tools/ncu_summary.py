#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full capture of one kernel) into the JSON summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rNN_<kernel>.json [frames_per_launch alg_bytes_per_frame]
"""
import csv
import json
import subprocess
import sys

KEEP = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
smsp__inst_executed.sum smsp__issue_active.avg.pct_of_peak_sustained_active sm__warps_active.avg.pct_of_peak_sustained_active
launch__registers_per_thread launch__grid_size launch__block_size launch__shared_mem_per_block_dynamic launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem smsp__cycles_active.avg sm__cycles_elapsed.max sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum lts__t_sector_hit_rate.pct
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio""".split()


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    launches = []
    for vals in rows[2:]:
        d = {h: vals[i] for i, h in enumerate(hdr)}
        m = {k: {"value": float(d[k]) if d[k].replace(".", "", 1).replace("-", "", 1).isdigit() else d[k], "unit": units[hdr.index(k)]} for k in KEEP if k in d}
        launches.append({"kernel": d.get("Kernel Name"), "metrics": m})
    summary = {"report": rep, "launches": launches}
    if len(sys.argv) > 4 and launches:
        frames, alg = int(sys.argv[3]), int(sys.argv[4])
        m = launches[0]["metrics"]
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
        traffic = sum(m[k]["value"] * scale[m[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        tscale = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
        secs = m["gpu__time_duration.sum"]["value"] * tscale[m["gpu__time_duration.sum"]["unit"]]
        summary["derived"] = {"frames_per_launch": frames, "algorithmic_bytes_per_launch": frames * alg, "dram_bytes_per_launch": traffic,
                              "dram_over_algorithmic": traffic / (frames * alg), "seconds_under_profiler": secs,
                              "thread_instructions_per_output_pixel": m["smsp__inst_executed.sum"]["value"] * 32 / (frames * 3840 * 2160)}
    json.dump(summary, open(out, "w"), indent=1)
    print(json.dumps(summary.get("derived", {}), indent=1))


if __name__ == "__main__":
    main()
