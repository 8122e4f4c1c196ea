mkdir -p gpurun_out
python scripts/op_times.py 2048 32 34 > gpurun_out/op_times_cache_on.log 2>&1
VO_BLOCK_CACHE=off python scripts/op_times.py 2048 32 34 > gpurun_out/op_times_cache_off.log 2>&1
tail -11 gpurun_out/op_times_cache_on.log; tail -5 gpurun_out/op_times_cache_off.log | cut -c1-150
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_r1f.log 2>&1; tail -3 gpurun_out/pytest_r1f.log | head -1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; cut -c1-200 gpurun_out/bench_r1f.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_r1f.json
timeout 300 python -m pytest tests/test_dexelize.py -m gpu -x -q 2>&1 | tail -1
timeout 300 python scripts/dexelize_time.py gpurun_out/dexelize_times.jsonl 2> gpurun_out/dex.err; cut -c1-230 gpurun_out/dexelize_times.jsonl
