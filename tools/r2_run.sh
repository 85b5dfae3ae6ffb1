set -x
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['e2e']['value'], d['host_queue_ms_per_step'])"
python -c "import __graft_entry__ as g; g.smoke()"
