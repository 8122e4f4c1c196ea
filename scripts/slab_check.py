"""Multi-GPU parity: torchrun --nproc-per-node N scripts/slab_check.py
Every rank dilates its y-slab (halo over NCCL) and compares with its rows of the single-GPU result."""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist
from voroffset_b200 import _lib, morpho, slab, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = _lib.Context(local)
op = morpho.make_operator("ours", ctx)
ok = True
for name, vol, R in [("blobs n160 R9.5", synth.blobs(160, count=40, padding=10, seed=3), 9.5),
                     ("torus_z n512 R16", synth.torus_z(512), 16.0),
                     ("lattice n128 R5", synth.lattice(128, padding=6), 5.0)]:
    full, _, _ = op.dilation(vol, R)
    mine = slab.shard_rows(vol, rank, world)
    d = morpho.DeviceVolume.upload(ctx, mine)
    sd = slab.SlabDilation(slab.CudaSlabBackend(ctx), rank, world)
    out = sd.dilate(d, R).download()
    want = slab.shard_rows(full, rank, world)
    same = out.bit_equal(want)
    ok &= same
    print(f"rank {rank}/{world} {name}: slab == single-GPU rows: {same} (halo {sd.last_halo_bytes} B, pass1 {sd.last_ms[0]:.3f} ms, pass2 {sd.last_ms[1]:.3f} ms)", flush=True)
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
