set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_r1g.log 2>&1
tail -3 gpurun_out/pytest_r1g.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1g.log 2>&1; tail -2 gpurun_out/smoke_r1g.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1g_ref.json 2> gpurun_out/bench_r1g_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pass1_tile|k_pass2_rows|k_thresh" -s 10 -c 5 -o gpurun_out/r1g_kernels python scripts/run_c5.py 2048 32 3 > gpurun_out/p.log 2>&1
tail -1 gpurun_out/p.log
timeout 900 python scripts/bench_configs.py > gpurun_out/configs.log 2>&1; tail -3 gpurun_out/configs.log
timeout 300 python scripts/dexelize_time.py gpurun_out/dexelize_times.jsonl 2> gpurun_out/dex.err
cat gpurun_out/bench_r1g.json | cut -c1-300
python scripts/pcie_probe.py > gpurun_out/pcie_probe.json 2>/dev/null; cat gpurun_out/pcie_probe.json
