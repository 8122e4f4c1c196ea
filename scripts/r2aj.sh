export VO_LIB=build/lib_ktrace.so
python scripts/e2e_dry.py
python scripts/e2e_dry.py copy_align=0
python scripts/ktrace_e2e.py gpurun_out/kt4_aligned.csv > gpurun_out/kt4_aligned.log 2>&1; cat gpurun_out/kt4_aligned.log
VO_TRACE=1 python scripts/e2e_once.py 8 6 2> gpurun_out/trace_aligned.txt; python scripts/trace_fmt.py gpurun_out/trace_aligned.txt
