"""Resident operations on a synthetic volume (for ncu launch lists). Usage: run_vol.py shape n padding R op iters [key=value ...]"""
import sys
sys.path.insert(0, ".")
from voroffset_b200 import synth, morpho, _lib
shape, n, pad, R, opn, it = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), sys.argv[5], int(sys.argv[6])
ctx = _lib.Context(0); op = morpho.make_operator("ours", ctx)
for kv in sys.argv[7:]:
    ctx.set_option(*kv.split("="))
vol = getattr(synth, shape)(n, padding=pad)
d = morpho.DeviceVolume.upload(ctx, vol)
for i in range(it):
    ctx.mark(0); out, t1, t2 = op.morph_dev(opn, d, R); ctx.mark(1)
    print(i, shape, opn, "total_ms", round(ctx.elapsed_ms(0, 1), 4), "passes", round(t1, 4), round(t2, 4), flush=True)
    out.free()
