set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or rows or config5 or offsets" 2>&1 | tail -3
python scripts/e2e_ab.py - pipe_ahead=off pipe_lean=off,pipe_order_one=off,pipe_ahead=off band_split=1 band_split=1,pipe_ctas=4 bands=6 bands=10 > gpurun_out/r2af_e2e_ab.txt 2>&1; cat gpurun_out/r2af_e2e_ab.txt
