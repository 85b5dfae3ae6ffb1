set -x
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/r2_sanitizer_$tool.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_handoff.py tests/test_gpu_formats.py -m gpu -q -x -k "fused_ring or fused-0 or handoff or consumer or new_sources or bgra_mixer or padded or tiled_scenes" 2>&1 | tail -2
  tail -3 gpurun_out/r2_sanitizer_$tool.log
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_sanitizer_scale.log python -m pytest tests/test_scale.py -m gpu -q -x -k "bit_exact or strides" 2>&1 | tail -2
tail -3 gpurun_out/r2_sanitizer_scale.log
