timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench6.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'])"
tail -3 gpurun_out/bench6.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svb_mix_tiled -s 14 -c 1 -o gpurun_out/prof_r1e python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu5.log 2>&1; tail -2 gpurun_out/ncu5.log
