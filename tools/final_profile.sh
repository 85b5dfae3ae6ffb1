# Evidence run for profiles/: launch list, full ncu capture of the dominant kernel, and the bench line itself.
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/launches_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svb_mix_tiled -s 14 -c 1 -o gpurun_out/prof_final python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_final.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -3 gpurun_out/bench_final.err
cat gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json
cat gpurun_out/bench_reference.json
