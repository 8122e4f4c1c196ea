set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile|k_pass2_rows|k_thresh" -s 10 -c 5 -o gpurun_out/r1d_kernels python scripts/run_c5.py 2048 32 3 > gpurun_out/p.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile" -s 6 -c 3 -o gpurun_out/r1d_ero python scripts/run_vol.py torus_z 2048 34 32 erosion 3 > gpurun_out/p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile|k_thresh" -s 8 -c 4 -o gpurun_out/r1d_c3 python scripts/run_vol.py lattice 512 10 5 dilation 3 > gpurun_out/p3.log 2>&1
tail -1 gpurun_out/p.log gpurun_out/p2.log gpurun_out/p3.log
cat gpurun_out/bench_r1d.json | cut -c1-300
