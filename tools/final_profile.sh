# Evidence run for profiles/: the GPU suite, launch list, full ncu capture of the dominant kernel, the bench lines.
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/launches_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svb_mix_tiled -s 14 -c 1 -f -o gpurun_out/prof_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:svb_scale -s 30 -c 1 -f -o gpurun_out/prof_scale_final python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -3 gpurun_out/bench_final.err
cat gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json
cat gpurun_out/bench_reference.json
timeout 300 python bench.py --workload cfg5 2>/dev/null | tail -1 > gpurun_out/bench_cfg5.json
timeout 300 python bench.py --workload cfg2 2>/dev/null | tail -1 > gpurun_out/bench_cfg2.json
cat gpurun_out/bench_cfg5.json gpurun_out/bench_cfg2.json
python -c "import __graft_entry__ as g; g.smoke()"
