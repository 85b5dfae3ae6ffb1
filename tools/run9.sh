B() { timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench9.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'])"; }
B h32
SVB_NVCC_DEFS="-DSVB_TILE_H=16 -DSVB_TILED_MIN_CTAS=4" python -m swiftvideo_b200.build --force > /dev/null 2>&1
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
B h16x4
SVB_NVCC_DEFS="-DSVB_TILE_H=8 -DSVB_TILED_MIN_CTAS=8" python -m swiftvideo_b200.build --force > /dev/null 2>&1
B h8x8
python -m swiftvideo_b200.build --force > /dev/null 2>&1
tail -3 gpurun_out/bench9.err
