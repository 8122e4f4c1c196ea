set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile|k_pass2_rows|k_thresh" -s 10 -c 5 -o gpurun_out/r1c_kernels python scripts/run_c5.py 2048 32 3 > gpurun_out/p.log 2>&1
tail -2 gpurun_out/p.log
cat gpurun_out/bench_r1c.json | cut -c1-400
