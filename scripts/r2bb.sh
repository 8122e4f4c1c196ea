python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -m gpu -x -q -k "pipelined or rows or window or config5" 2>&1 | tail -2
python scripts/e2e_ab.py - copy_batch=off pipe_ahead=off pipe_ahead=off,bands=12 bands=12 pipe_ahead=off,bands=16 pipe_ahead=off,copy_align=4096 copy_align=4096 > gpurun_out/r2bb_e2e_ab.txt 2>&1; cat gpurun_out/r2bb_e2e_ab.txt
