#!/bin/bash
# Run prebuilt library variants (swiftvideo_b200/variants/libsvb200_<name>.so, built here with SVB_NVCC_DEFS) on the GPU box without
# paying for compilation there:  tools/ab_prebuilt.sh <name> ...     (SVB_COMPOSITOR is taken from the environment)
for name in "$@"; do
    cp swiftvideo_b200/variants/libsvb200_$name.so swiftvideo_b200/libsvb200.so || continue
    timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cfg34 or tiled or variants" 2>&1 | tail -1
    timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 2>> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'])"
done
cp swiftvideo_b200/variants/libsvb200_default.so swiftvideo_b200/libsvb200.so
