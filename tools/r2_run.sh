set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
SVB_HOST_PROFILE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
grep "svb host" gpurun_out/r2_bench_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_b.json'))
print({k:d[k] for k in ('value','ms_per_step','host_queue_ms_per_step','host_queue_in_library_ms_per_step','host_backpressure_ms_per_step')}, d['e2e']['value'], d['one_frame_per_launch'])
PY
