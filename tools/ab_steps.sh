# step time against kernel time of prebuilt library variants at two timed-region lengths:  tools/ab_steps.sh <name> ...
for name in "$@"; do
    cp swiftvideo_b200/variants/libsvb200_$name.so swiftvideo_b200/libsvb200.so || continue
    for steps in 20 200; do
        timeout 200 python bench.py --steps $steps --warmup 5 --no-cpu-baseline --e2e-steps 2 2>> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name steps $steps: step', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms_per_launch'], 'launches', d['gpu_launches'])"
    done
done
cp swiftvideo_b200/variants/libsvb200_default.so swiftvideo_b200/libsvb200.so
