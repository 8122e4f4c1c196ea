python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -m gpu -x -q -k "pipelined or rows or window or config5" 2>&1 | tail -2
python scripts/e2e_ab.py - pipe_predict=off pipe_predict=on,bands=10 pipe_predict=on,bands=12 pipe_predict=on,bands=16 pipe_predict=on,band_split=1 > gpurun_out/r2ba_e2e_ab.txt 2>&1; cat gpurun_out/r2ba_e2e_ab.txt
