python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "copy_boundaries or launches_left_out or pipelined" 2>&1 | tail -3
python scripts/bench_configs.py > gpurun_out/r2as_configs.log 2>&1; grep roofline gpurun_out/r2as_configs.log | cut -c1-700
