set -x
python scripts/gpu_pass1.py > gpurun_out/r2am_gpu_pass1.log 2>&1; grep -c "True" gpurun_out/r2am_gpu_pass1.log; tail -1 gpurun_out/r2am_gpu_pass1.log; grep lattice gpurun_out/r2am_gpu_pass1.log
python scripts/run_vol.py lattice 512 10 5 dilation 5 2>&1 | tail -2
python scripts/run_vol.py torus_z 2048 34 32 closing 4 2>&1 | tail -1
python scripts/quick_c5.py
(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2am_pytest.log 2>&1; tail -5 gpurun_out/r2am_pytest.log
