#!/bin/bash
# A/B a tuning knob of svb_mix_gather on the GPU box:  tools/ab_gather.sh "<label>" "<SVB_NVCC_DEFS>" ...
export SVB_COMPOSITOR=gather
while [ $# -ge 2 ]; do
    SVB_NVCC_DEFS="$2" python -m swiftvideo_b200.build --force > /dev/null 2>&1 || { echo "build failed for $1"; shift 2; continue; }
    timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cfg34 or tiled or variants" 2>&1 | tail -1
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 2>> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'])"
    shift 2
done
python -m swiftvideo_b200.build --force > /dev/null 2>&1
