set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
for w in cfg5 cfg2 cfg2chain; do timeout 300 python bench.py --workload $w 2>/dev/null | tail -1 > gpurun_out/r2_bench_$w.json; cut -c1-400 gpurun_out/r2_bench_$w.json; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:svb_scale -s 30 -c 1 -f -o gpurun_out/r2_prof_scale python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
