timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
B() { timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench5.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'])"; }
B ctas2
SVB_NVCC_DEFS="-DSVB_TILED_MIN_CTAS=3" python -m swiftvideo_b200.build --force > /dev/null 2>&1
B ctas3
python -m swiftvideo_b200.build --force > /dev/null 2>&1
tail -3 gpurun_out/bench5.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svb_mix_tiled -s 14 -c 1 -o gpurun_out/prof_r1d python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu4.log 2>&1; tail -2 gpurun_out/ncu4.log
