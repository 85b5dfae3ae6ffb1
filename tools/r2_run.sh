set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 300 python tools/e2e_probe.py 2>&1 | tee gpurun_out/r2_e2e_probe.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
cut -c1-2400 gpurun_out/r2_bench_b.json
