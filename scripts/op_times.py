"""Steady-state times of the four operations on a resident volume (n, R, padding from argv), ten calls each."""
import sys
sys.path.insert(0, ".")
import numpy as np
from voroffset_b200 import synth, morpho, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
R = float(sys.argv[2]) if len(sys.argv) > 2 else 32.0
pad = int(sys.argv[3]) if len(sys.argv) > 3 else 34
import os
ctx = _lib.Context(0); op = morpho.make_operator("ours", ctx)
if os.environ.get("VO_BLOCK_CACHE"): ctx.set_option("block_cache", os.environ["VO_BLOCK_CACHE"]); print("block_cache", os.environ["VO_BLOCK_CACHE"])
d = morpho.DeviceVolume.upload(ctx, synth.torus_z(n, padding=pad))
for opn in ("dilation", "erosion", "opening", "closing"):
    ts = []
    for i in range(10):
        ctx.mark(0); out, t1, t2 = op.morph_dev(opn, d, R); ctx.mark(1); out.free()
        ts.append(ctx.elapsed_ms(0, 1))
    print(opn, "ms:", " ".join(f"{t:.2f}" for t in ts), "| median", round(float(np.median(ts[3:])), 3), f"| last primitive: pass1 {t1:.3f} pass2 {t2:.3f}", flush=True)
# closing step by step: dilation, then the erosion of its result
dil, _, _ = op.morph_dev("dilation", d, R)
for i in range(6):
    ctx.mark(0); out, t1, t2 = op.morph_dev("erosion", dil, R); ctx.mark(1); out.free()
    print("erosion of the dilated volume:", round(ctx.elapsed_ms(0, 1), 3), f"pass1 {t1:.3f} pass2 {t2:.3f}", flush=True)
