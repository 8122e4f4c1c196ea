set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or rows or config5" 2>&1 | tail -3
python scripts/e2e_opts.py pipe_lean=off,pipe_order_one=off band_split=1 band_split=1,pipe_ctas=4 band_split=1,pipe_ctas=2 pipe_ctas=4 band_split=1,pipe_ctas=4,band_weights=2:3:4:4:4:4:3:2:1 band_split=1,pipe_ctas=4,bands=6 band_split=1,pipe_ctas=4,bands=12 > gpurun_out/r2x_e2e_opts.txt 2>&1; cat gpurun_out/r2x_e2e_opts.txt
export VO_LIB=build/lib_ktrace.so
python scripts/ktrace_e2e.py gpurun_out/kt2_default.csv > gpurun_out/kt2_default.log 2>&1; cat gpurun_out/kt2_default.log
python scripts/ktrace_e2e.py gpurun_out/kt2_split1_ctas4.csv band_split=1 pipe_ctas=4 > gpurun_out/kt2_split1_ctas4.log 2>&1; cat gpurun_out/kt2_split1_ctas4.log
