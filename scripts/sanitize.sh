# compute-sanitizer passes over the small GPU cases (memcheck: out-of-bounds / misaligned / leaks of the kernels and
# of the block cache; racecheck: shared-memory hazards of the tile kernel). Output: gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_dexelize.py -m gpu -q -x -k "bit_exact or errors" > gpurun_out/sanitizer_memcheck_dexelize.log 2>&1; echo "memcheck dexelize rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
timeout 900 compute-sanitizer --tool memcheck --leak-check full --error-exitcode 9 python scripts/sanitize_ops.py > gpurun_out/sanitizer_memcheck_ops.log 2>&1; echo "memcheck ops rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_ops.py > gpurun_out/sanitizer_racecheck_ops.log 2>&1; echo "racecheck ops rc=$?"
timeout 1200 compute-sanitizer --tool synccheck --error-exitcode 9 python scripts/sanitize_ops.py > gpurun_out/sanitizer_synccheck_ops.log 2>&1; echo "synccheck ops rc=$?"
for f in gpurun_out/sanitizer_*.log; do echo $f; tail -n 3 $f; done
