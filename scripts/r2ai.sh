set -x
python -m pytest tests/test_gpu_parity.py tests/test_multigpu.py -m gpu -x -q -k "pipelined or rows or config5 or offsets or window" 2>&1 | tail -3
python scripts/e2e_ab.py - copy_align=0 copy_align=128 copy_align=4096 copy_align=256,band_split=1 copy_align=256,bands=6 copy_align=256,bands=12 copy_align=256,copy_out=32 > gpurun_out/r2ai_e2e_ab.txt 2>&1; cat gpurun_out/r2ai_e2e_ab.txt
