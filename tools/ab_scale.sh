#!/bin/bash
# A/B a tuning knob of svb_scale_convert on the GPU box:  tools/ab_scale.sh "<label>" "<SVB_NVCC_DEFS>" ...
while [ $# -ge 2 ]; do
    SVB_NVCC_DEFS="$2" python -m swiftvideo_b200.build --force > /dev/null 2>&1 || { echo "build failed for $1"; shift 2; continue; }
    timeout 600 python -m pytest tests/test_scale.py -m gpu -q -x 2>&1 | tail -1
    for w in cfg5 cfg2; do timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2>>gpurun_out/scale.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['config']['workload'][:30], d['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['e2e']['value'])"; done
    shift 2
done
python -m swiftvideo_b200.build --force > /dev/null 2>&1
