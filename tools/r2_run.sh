# run 19: verify everything added since run 18, bench numbers, HEAD profile of the ring compositor
set -x
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -3 gpurun_out/r2_bench.err
cat gpurun_out/r2_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2_launches_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svb_mix_ring -s 13 -c 1 -f -o gpurun_out/r2_prof_ring python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2_ncu_ring.log 2>&1
tail -2 gpurun_out/r2_ncu_ring.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
