#!/usr/bin/env python3
"""profiles/r1_results_table.md from the bench lines under profiles/ (BASELINE.md section 2.3's template)."""
import json
from pathlib import Path

P = Path(__file__).resolve().parent.parent / "profiles"


def L(n):
    return json.loads((P / n).read_text())


def main():
    c4, c42, c3, c2 = L("r1_bench_1gpu.json"), L("r1_bench_2gpu.json"), L("r1_bench_cfg3.json"), L("r1_bench_cfg2_mix.json")
    c5, c2s, ref = L("r1_bench_cfg5_scale.json"), L("r1_bench_cfg2_scale.json"), L("r1_bench_reference_arm.json")
    tr = L("traffic.json")["svb_mix_tiled_dram_bytes_per_launch"]
    rows = []

    def row(cfg, g, streams, d, dram, exact, cpu):
        r = d["roofline"]
        rows.append(f"| {cfg} | {g} | {streams} | {d['value']:,.0f} | {r['achieved']:,.0f} | {r['achieved'] / 8000 * 100:.1f} % | {r['frac'] * 100:.1f} % | {dram} | {exact} | {cpu} |")

    row("2 (mixer: clear + `img_nv12_nv12`, 1080p -> 720p)", 1, 8, c2, "— (36 MB per step: L2-resident, said so in the line)", "bit-exact (`test_cfg2_full_size`)",
        f"{c2['cpu_baseline']['value']:.0f} (reference kernels, {c2['cpu_baseline']['cores']} threads)")
    row("2 (convert+scale: NV12 -> BGRA bilinear 1080p -> 720p)", 1, 8, c2s, "— (54 MB per step: L2-resident)", "bit-exact vs definition; <= 1 code vs swscale",
        f"{c2s['cpu_baseline']['value']:,.0f} (libswscale, {c2s['cpu_baseline']['cores']} threads)")
    row("3 (4K, 4 layers)", 1, 8, c3, "—", "bit-exact (`test_cfg34_full_size[4]`)", f"{c3['cpu_baseline']['value']:.1f} (reference kernels, {c3['cpu_baseline']['cores']} threads)")
    row("4 (4K, 8 layers) — the headline", 1, 8, c4, f"{tr / 8 / 1e6:.1f} MB (algorithmic 46.7 MB)", "bit-exact (`test_cfg34_full_size[8]`)",
        f"{c4['cpu_baseline']['value']:.1f} (reference kernels, {c4['cpu_baseline']['cores']} threads); swscale scale stage alone {c4['cpu_baseline']['swscale_scale_stage_only']['value']:,.0f}")
    row("4", 2, 16, c42, "as above", "as above", "—")
    row("5 (4K P010 -> 1080p BGRA, Lanczos-3)", 1, 8, c5, "25.2 MB read; the 8.3 MB of BGRA were still in L2 when the launch ended (algorithmic 33.2 MB)",
        "bit-exact vs definition; <= 1 code vs swscale (tolerance)", f"{c5['cpu_baseline']['value']:,.0f} (libswscale, {c5['cpu_baseline']['cores']} threads)")
    txt = """# Round-1 results in the shape of BASELINE.md section 2.3 (1 x B200 unless stated; `python bench.py [--workload ...]`)

`frames/s` = whole job, layers resident in HBM.  `alg. GB/s` = algorithmic bytes per launch / device time of the dominant kernel,
per GPU.  N = 4 and N = 8 are left to the driver's scaling run (`SCALE_r01.json`); streams are independent, nothing is exchanged.
Made by `tools/results_table.py` from the bench lines beside it.

| cfg | GPUs | streams | frames/s | alg. GB/s per GPU | % of 8.0 TB/s | % of measured copy (6 541 GB/s) | DRAM bytes / frame (ncu) | parity | CPU frames/s on this box |
|---|---|---|---|---|---|---|---|---|---|
""" + "\n".join(rows) + f"""

End to end through host buffers (H2D and D2H inside the timed region): cfg 4 {c4['e2e']['value']:,.0f} frames/s on one GPU, {c42['e2e']['value']:,.0f} on two;
cfg 3 {c3['e2e']['value']:,.0f}; cfg 2 (mixer) {c2['e2e']['value']:,.0f}; cfg 5 {c5['e2e']['value']:,.0f} -- all bound by the host link.
Reference arm (`bench.py --impl reference`, the reference's kernel text on {ref['cpu_baseline']['cores']} host threads): {ref['value']:.2f} frames/s on cfg 4.
"""
    (P / "r1_results_table.md").write_text(txt)
    print(txt)


if __name__ == "__main__":
    main()
