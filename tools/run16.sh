for m in generic per_layer; do timeout 600 python bench.py --mode $m --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 2> gpurun_out/bench_modes.err | tee gpurun_out/bench_mode_$m.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d.get('roofline'))"; done
tail -3 gpurun_out/bench_modes.err
