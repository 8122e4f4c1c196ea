"""Resident operation on a synthetic volume, timed over many iterations: min / median / max of the device time.
Usage: time_vol.py shape n padding R op iters [key=value ...]   (VO_LIB picks the build)"""
import sys, statistics
sys.path.insert(0, ".")
from voroffset_b200 import synth, morpho, _lib
shape, n, pad, R, opn, it = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), sys.argv[5], int(sys.argv[6])
ctx = _lib.Context(0); op = morpho.make_operator("ours", ctx)
for kv in sys.argv[7:]:
    ctx.set_option(*kv.split("="))
vol = getattr(synth, shape)(n, padding=pad)
d = morpho.DeviceVolume.upload(ctx, vol)
ts, p1, p2 = [], [], []
for i in range(it + 3):
    ctx.mark(0); out, t1, t2 = op.morph_dev(opn, d, R); ctx.mark(1)
    if i >= 3:
        ts.append(ctx.elapsed_ms(0, 1)); p1.append(t1); p2.append(t2)
    out.free()
print(shape, n, pad, R, opn, " ".join(sys.argv[7:]), "ms min/med/max %.4f %.4f %.4f" % (min(ts), statistics.median(ts), max(ts)),
      "pass1 med %.4f pass2 med %.4f" % (statistics.median(p1), statistics.median(p2)), flush=True)
