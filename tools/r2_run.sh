SVB_DEBUG_POOL=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err
awk '/e2e leg starts/{on=1} on{print} /e2e leg ends/{on=0}' gpurun_out/bench_dbg.err | sort | uniq -c | sort -rn | head
grep -c cuMemHostAlloc gpurun_out/bench_dbg.err; grep -c "cuMemAlloc " gpurun_out/bench_dbg.err
