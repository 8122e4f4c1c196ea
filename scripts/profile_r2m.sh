# Round 2, state at re-entry: tests, smoke, both bench arms, launch list of the bench, every config. One gpurun call, one GPU.
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2m_pytest.log 2>&1
tail -3 gpurun_out/r2m_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; tail -2 gpurun_out/r2m_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/r2m_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/b.log 2>&1
timeout 900 python scripts/bench_configs.py > gpurun_out/r2m_configs.log 2>&1; tail -3 gpurun_out/r2m_configs.log
cut -c1-400 gpurun_out/r2m_bench.json
cut -c1-400 gpurun_out/r2m_bench_ref.json
nproc; free -g | head -2
