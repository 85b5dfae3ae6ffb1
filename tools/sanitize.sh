# compute-sanitizer over the compositor's parity tests (profiles/r2_sanitizer.txt): memcheck, synccheck, racecheck
set -x
SEL="fused_ring or fused-0 or handoff or consumer or new_sources or bgra_mixer or padded or tiled_scenes"
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/r2_sanitizer_$tool.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_handoff.py tests/test_gpu_formats.py -m gpu -q -x -k "$SEL" 2>&1 | tail -2
  tail -3 gpurun_out/r2_sanitizer_$tool.log
done
grep -E "Race reported|and (Read|Write) access|Write access|Read access" gpurun_out/r2_sanitizer_racecheck.log | sed -E 's/0x[0-9a-f]+//g; s/\[[0-9]+ hazards\]//' | sort | uniq -c | sort -rn | head -30 > gpurun_out/r2_sanitizer_racecheck_sites.txt
