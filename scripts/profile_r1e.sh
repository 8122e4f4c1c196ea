set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_r1e.log 2>&1
tail -3 gpurun_out/pytest_r1e.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1e.log 2>&1; tail -2 gpurun_out/smoke_r1e.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1e_ref.json 2> gpurun_out/bench_r1e_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile|k_pass2_rows|k_thresh" -s 10 -c 5 -o gpurun_out/r1e_kernels python scripts/run_c5.py 2048 32 3 > gpurun_out/p.log 2>&1
tail -1 gpurun_out/p.log
cat gpurun_out/bench_r1e.json | cut -c1-600
cat gpurun_out/bench_r1e_ref.json | cut -c1-300
