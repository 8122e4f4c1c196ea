# Round 2: compute-sanitizer over the code that changed this round - smoke and the small operations (tile kernel, list
# insertion by bisection, pass-2 variants), and the banded host-buffer call with aligned copy boundaries, launches left
# out and the SM-driven download (two pipelined tests at full pipeline size).
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck smoke rc=$?"
timeout 900 compute-sanitizer --tool memcheck --leak-check full --error-exitcode 9 python scripts/sanitize_ops.py > gpurun_out/r2_sanitizer_memcheck_ops.log 2>&1; echo "memcheck ops rc=$?"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "copy_boundaries or launches_left_out" > gpurun_out/r2_sanitizer_memcheck_pipeline.log 2>&1; echo "memcheck pipeline rc=$?"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_ops.py > gpurun_out/r2_sanitizer_racecheck_ops.log 2>&1; echo "racecheck ops rc=$?"
for f in gpurun_out/r2_sanitizer_*.log; do echo $f; tail -n 4 $f; done
