"""Per-step timings and halo statistics of the multi-GPU operators (one process per GPU under torchrun):
    python -m torch.distributed.run --nproc-per-node N scripts/mg_steps.py [n=2048] [R=32] [steps=6]"""
import json, os, sys, time
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from voroffset_b200 import morpho, multigpu, slab, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
R = float(sys.argv[2]) if len(sys.argv) > 2 else 32.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mg = multigpu.MultiGpu.from_torch_distributed(local)
ctx = mg.contexts[0]
if os.environ.get("VO_SLAB_RESERVE"):
    ctx.set_option("slab_reserve", os.environ["VO_SLAB_RESERVE"])
OPS = os.environ.get("VO_OPS", "dilation,erosion,opening,closing").split(",")
full = synth.torus_z(n, padding=int(R) + 2) if n <= 4096 else None
if full is None:      # large grids: only this rank's rows are generated
    import types
    b = slab.slab_bounds(n, world)[rank]
    own_rows = synth.torus_z_rows(n, b[0], b[1])
    full = types.SimpleNamespace(nx=own_rows.nx, ny=n, zmin=own_rows.zmin, zmax=own_rows.zmax)
if n <= 4096:
    y0, y1 = slab.slab_bounds(full.ny, world)[rank]
    c0, c1 = y0 * full.nx, y1 * full.nx
    off = full.off[c0:c1 + 1].astype("int64")
    own = full.like(full.nx, y1 - y0, (off - off[0]).astype("uint32"), full.spans[off[0]:off[-1]])
else:
    own = own_rows
d = morpho.DeviceVolume.upload(ctx, own)
for opn in OPS:
    for i in range(steps):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t = time.perf_counter()
        ctx.mark(0)
        outs, t1, t2 = mg.morph_dev(opn, [d], R, full.zmin, full.zmax)
        ctx.mark(1)
        wall = (time.perf_counter() - t) * 1e3
        st = mg.stats(0)
        print(json.dumps({"rank": rank, "op": opn, "step": i, "ms": round(ctx.elapsed_ms(0, 1), 3), "wall_ms": round(wall, 3),
                          "t1": round(t1, 3), "t2": round(t2, 3), **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}}), flush=True)
        outs[0].free()
mg.close()
dist.destroy_process_group()
