#!/bin/bash
# A/B a kernel tuning knob on the GPU box:  tools/ab.sh "<label>" "<SVB_NVCC_DEFS>" ...   (pairs; "" = default build)
# e.g.  gpurun -- 'bash tools/ab.sh base "" ctas3 "-DSVB_TILED_MIN_CTAS=3"'
B() { timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['e2e']['value'], d['clocks'].get('sm_mhz'))"; }
while [ $# -ge 2 ]; do
    SVB_NVCC_DEFS="$2" python -m swiftvideo_b200.build --force > /dev/null 2>&1 || { echo "build failed for $1"; shift 2; continue; }
    timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cfg34 or tiled or small" 2>&1 | tail -1
    B "$1"; B "$1"
    shift 2
done
python -m swiftvideo_b200.build --force > /dev/null 2>&1
