set -x
mkdir -p gpurun_out
N=${1:-4}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/slab_check.py > gpurun_out/slab_check_$N.log 2>&1; echo "slab_check rc=$?"
grep -c "True" gpurun_out/slab_check_$N.log; grep "False" gpurun_out/slab_check_$N.log | head
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_${N}gpu.json | cut -c1-400
tail -3 gpurun_out/bench_${N}gpu.err
