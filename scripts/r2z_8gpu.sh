set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2z_bench_${N}gpu.json 2> gpurun_out/r2z_bench_${N}gpu.err; echo "bench rc=$?"
tail -1 gpurun_out/r2z_bench_${N}gpu.json | cut -c1-300
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r2z_pytest_multigpu_${N}gpu.log 2>&1; tail -2 gpurun_out/r2z_pytest_multigpu_${N}gpu.log
