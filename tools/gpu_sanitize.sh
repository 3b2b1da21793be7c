mkdir -p gpurun_out
timeout 1000 compute-sanitizer --tool racecheck --racecheck-report analysis --launch-timeout 0 python -m pytest tests/test_gpu_parity.py -x -q -k "plane_stages or extraction_matches" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck.log | head; grep -E "Race reported|hazard" gpurun_out/sanitize_racecheck.log | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -20
