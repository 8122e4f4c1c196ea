set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or rows or config5 or offsets" 2>&1 | tail -3
python scripts/e2e_opts.py copy_out=0 copy_out=32 copy_out=16 copy_out=64 copy_out=148 copy_out=32,band_split=1 copy_out=32,pipe_ctas=4 copy_out=32,bands=6 copy_out=32,bands=12 copy_out=32,band_weights=1:2:3:3:3:3:2:1 > gpurun_out/r2ac_e2e_opts.txt 2>&1; cat gpurun_out/r2ac_e2e_opts.txt
