timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
B() { timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench15.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'])"; }
B precomp; B precomp
tail -3 gpurun_out/bench15.err
