set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 300 python tools/e2e_probe.py 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_b.json'))
print(d['value'], d['ms_per_step'], d['e2e'])"
