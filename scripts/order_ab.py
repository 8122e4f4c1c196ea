"""A/B of the pass-1 tile order (expensive tiles first vs row-major) on a resident volume: pass times from vo_last_profile."""
import sys
sys.path.insert(0, ".")
import numpy as np
from voroffset_b200 import synth, morpho, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
R = float(sys.argv[2]) if len(sys.argv) > 2 else 32.0
pad = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ops = sys.argv[4].split(",") if len(sys.argv) > 4 else ["dilation"]
ctx = _lib.Context(0); op = morpho.make_operator("ours", ctx)
d = morpho.DeviceVolume.upload(ctx, synth.torus_z(n, padding=pad))
for opn in ops:
    for mode in ("off", "on", "off", "on"):
        ctx.set_option("tile_order", mode)
        ts, p1, k1 = [], [], []
        for i in range(12):
            ctx.mark(0); out, t1, t2 = op.morph_dev(opn, d, R); ctx.mark(1); out.free()
            ts.append(ctx.elapsed_ms(0, 1)); p1.append(t1); k1.append(ctx.last_profile()[0])
        print(opn, "tile_order", mode, "total ms", round(float(np.median(ts[3:])), 4), "pass1", round(float(np.median(p1[3:])), 4),
              "k_pass1_tile", round(float(np.median(k1[3:])), 4), flush=True)
