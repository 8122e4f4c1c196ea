# halo exchange A/B on N GPUs (N = $1): NCCL p2p channel limits and SMs reserved for the NCCL kernels, 8192^2 dilation
N=$1
run() { echo "== $*"; env "$@" VO_OPS=dilation timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/mg_steps.py 8192 32 5 2>&1 | grep -E '"rank": (0|1|3),' | grep '"step": [34]' | cut -c1-260; }
run A=1
run NCCL_MAX_P2P_NCHANNELS=8
run NCCL_MAX_P2P_NCHANNELS=4
run NCCL_MAX_P2P_NCHANNELS=8 VO_SLAB_RESERVE=16
run VO_SLAB_RESERVE=32
