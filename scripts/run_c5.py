"""Runs a few resident C5-style operations (for ncu captures). Usage: run_c5.py [n] [R] [iters] [op] [padding] [shape]"""
import sys
sys.path.insert(0, ".")
from voroffset_b200 import synth, morpho, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
R = float(sys.argv[2]) if len(sys.argv) > 2 else 32.0
it = int(sys.argv[3]) if len(sys.argv) > 3 else 2
opn = sys.argv[4] if len(sys.argv) > 4 else "dilation"
pad = int(sys.argv[5]) if len(sys.argv) > 5 else 0
shape = sys.argv[6] if len(sys.argv) > 6 else "torus_z"
ctx = _lib.Context(0)
op = morpho.make_operator("ours", ctx)
vol = getattr(synth, shape)(n, padding=pad)
d = morpho.DeviceVolume.upload(ctx, vol)
for i in range(it):
    ctx.mark(0)
    out, t1, t2 = op.morph_dev(opn, d, R)
    ctx.mark(1)
    print(i, opn, "total_ms", ctx.elapsed_ms(0, 1), "last primitive", t1, t2, ctx.last_profile(), "k_out", out.info()[2] / (vol.nx * vol.ny), flush=True)
    out.free()
