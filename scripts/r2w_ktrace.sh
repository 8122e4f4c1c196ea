set -x
export VO_LIB=build/lib_ktrace.so
python scripts/ktrace_e2e.py gpurun_out/kt_default.csv > gpurun_out/kt_default.log 2>&1; cat gpurun_out/kt_default.log
python scripts/ktrace_e2e.py gpurun_out/kt_split1.csv band_split=1 > gpurun_out/kt_split1.log 2>&1; cat gpurun_out/kt_split1.log
python scripts/ktrace_e2e.py gpurun_out/kt_w12.csv pipe_warps=12 > gpurun_out/kt_w12.log 2>&1; cat gpurun_out/kt_w12.log
