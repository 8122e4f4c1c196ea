set -x
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile" -s 6 -c 3 -o gpurun_out/r2al_c3 python scripts/run_vol.py lattice 512 10 5 dilation 3 > gpurun_out/r2al_c3.log 2>&1
tail -2 gpurun_out/r2al_c3.log
python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -m gpu -x -q -k "pipelined or rows or window" 2>&1 | tail -2
