# Round 2: block-aligned survivor entries - parity of the tile kernel against the unpruned kernel, C5 / C4 step times for
# two list capacities, the GPU test suite.
set -x
python scripts/gpu_pass1.py > gpurun_out/r2v_gpu_pass1.log 2>&1; tail -3 gpurun_out/r2v_gpu_pass1.log
python scripts/quick_c5.py > gpurun_out/r2v_quick_lcap32.log 2>&1; cat gpurun_out/r2v_quick_lcap32.log
VO_LIB=build/lib_lcap24.so python scripts/quick_c5.py > gpurun_out/r2v_quick_lcap24.log 2>&1; cat gpurun_out/r2v_quick_lcap24.log
python scripts/run_vol.py lattice 512 10 5 dilation 4 2>&1 | tail -2
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2v_pytest.log 2>&1; tail -5 gpurun_out/r2v_pytest.log
