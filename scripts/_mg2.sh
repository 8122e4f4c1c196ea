set -x
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2t_bench_${N}gpu.json 2> gpurun_out/r2t_bench_${N}gpu.err
tail -3 gpurun_out/r2t_bench_${N}gpu.err
cut -c1-600 gpurun_out/r2t_bench_${N}gpu.json
timeout 600 python -m pytest tests/test_multigpu.py -x -q 2>&1 | tail -3
