# Evidence run for profiles/r2_*: the GPU suite, launch list, full ncu captures of the two dominant kernels, the bench lines.
set -x
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2_launches_run.log 2>&1
# -s 13: the 10 ring-fill launches of the set-up (no layers) and three of its eight priming steps are skipped: the launch captured is a full
# 8-frame step of the workload (every step of the set-up's priming, of the warm-up and of the timed region is the same launch)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svb_mix_ring -s 13 -c 1 -f -o gpurun_out/r2_prof_ring python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2_ncu_ring.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:svb_scale -s 30 -c 1 -f -o gpurun_out/r2_prof_scale python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -3 gpurun_out/r2_bench.err
cat gpurun_out/r2_bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2_bench_reference.json
cat gpurun_out/r2_bench_reference.json
for w in cfg5 cfg2 cfg2chain cfg3 cfg2mix; do timeout 300 python bench.py --workload $w 2>/dev/null | tail -1 > gpurun_out/r2_bench_$w.json; done
python -c "import __graft_entry__ as g; g.smoke()"
# afterwards, here: tools/ncu_summary.py, tools/ncu_exec_mix.py, tools/sass_mix.py -> profiles/
