set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "dual or tile_kernel_forced or z_bounds or seeded or config4 or config5" 2>&1 | tail -15
timeout 300 python scripts/dual_ab.py > gpurun_out/dual_ab.jsonl 2> gpurun_out/dual_ab.err; cat gpurun_out/dual_ab.jsonl; tail -5 gpurun_out/dual_ab.err
