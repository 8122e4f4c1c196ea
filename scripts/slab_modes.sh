for mode in serial overlap; do
VO_SLAB=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{"metric"' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$mode', 'value', round(d['value']/1e9,3), 'ms', round(d['ms_per_step'],3), d['pass_ms'])"
done
