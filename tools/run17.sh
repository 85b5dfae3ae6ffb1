nvidia-smi topo -m 2>/dev/null | head -8
for f in "" "--no-numa-bind"; do timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline $f 2> gpurun_out/bench17.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$f', d['value'], d['e2e'], d['config']['host_affinity'], d['clocks'])"; done
tail -3 gpurun_out/bench17.err
