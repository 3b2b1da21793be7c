# final validation of a round: GPU parity tests, driver smoke, default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err
echo "bench stdout lines: $(grep -c . gpurun_out/validate_bench.json)"; cut -c1-200 gpurun_out/validate_bench.json
