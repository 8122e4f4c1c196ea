"""Runs a few resident C5 dilations (for ncu captures). Usage: run_c5.py [n] [R] [iters]"""
import sys
sys.path.insert(0, ".")
from voroffset_b200 import synth, morpho, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
R = float(sys.argv[2]) if len(sys.argv) > 2 else 32.0
it = int(sys.argv[3]) if len(sys.argv) > 3 else 2
ctx = _lib.Context(0)
op = morpho.make_operator("ours", ctx)
d = morpho.DeviceVolume.upload(ctx, synth.torus_z(n))
for i in range(it):
    out, t1, t2 = op.morph_dev("dilation", d, R)
    print(i, t1, t2, ctx.last_profile(), flush=True)
    out.free()
