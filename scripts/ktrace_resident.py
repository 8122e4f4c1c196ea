"""Kernel trace of one resident dilation (library built with -DVO_KTRACE, passed as VO_LIB).
Usage: ktrace_resident.py out.csv [shape n padding R]"""
import sys
sys.path.insert(0, ".")
from voroffset_b200 import synth, morpho, _lib
shape, n, pad, R = (sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])) if len(sys.argv) > 5 else ("torus_z", 2048, 0, 32.0)
ctx = _lib.Context(0); op = morpho.make_operator("ours", ctx)
d = morpho.DeviceVolume.upload(ctx, getattr(synth, shape)(n, padding=pad))
for _ in range(4):
    r, _, _ = op.morph_dev("dilation", d, R); r.free()
ctx.set_option("ktrace", "400000")
r, t1, t2 = op.morph_dev("dilation", d, R); r.free()
ctx.set_option("ktrace_dump", sys.argv[1])
print("traced dilation passes ms", t1, t2)
