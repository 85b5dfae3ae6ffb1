set -x
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
B() { timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 3 "$@" 2>> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['e2e']['value'], d['host_queue_ms_per_step'], d['clocks'].get('sm_mhz'))"; }
B --mode fused
B --mode fused
B --mode fused
B --mode fused_tiled
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svb_mix_ring -s 13 -c 1 -f -o gpurun_out/prof_ring python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_ring.log 2>&1
