"""Development helper: tile vs simple pass-1 kernels, parity + timing (run under gpurun)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from voroffset_b200 import synth, morpho, _lib

ctx = _lib.Context(0)
op = morpho.make_operator("ours", ctx)
bad = 0
cases = [
    ("torus_x n128 p8", synth.torus_x(128, padding=8), 5.5),
    ("torus_z n160 p12", synth.torus_z(160, padding=12), 10.0),
    ("blobs n96", synth.blobs(96, padding=9), 8.0),
    ("blobs many", synth.blobs(80, count=120, padding=5, rmin=0.02, rmax=0.07, seed=9), 3.3),
    ("random k8", synth.random_volume(40, 32, kmax=8, padding=6), 4.3),
    ("lattice n96", synth.lattice(96, padding=8), 5.0),
    ("lattice n256 R12", synth.lattice(256, padding=14), 12.0),
    ("one row", synth.random_volume(50, 1, kmax=4, padding=0, seed=5), 6.0),
    ("R40 blobs", synth.blobs(128, padding=45), 40.5),
    ("R0.75", synth.blobs(40, padding=2, seed=3), 0.75),
    ("torus_z 512 R16", synth.torus_z(512), 16.0),
    ("torus_z 1024 R16", synth.torus_z(1024), 16.0),
    ("torus_z 2048 R32", synth.torus_z(2048), 32.0),
]
for name, v, R in cases:
    res = {}
    for mode in ("simple", "tile"):
        ctx.set_option("pass1", mode)
        d = morpho.DeviceVolume.upload(ctx, v)
        ts = []
        for i in range(3):
            out, t1, t2 = op.morph_dev("dilation", d, R)
            k1, k2 = ctx.last_profile()
            ts.append((t1, t2, k1, k2))
            if i < 2: out.free()
        res[mode] = out.download()
        t1, t2, k1, k2 = ts[-1]
        print(f"{name:22s} {mode:6s} pass1 {t1:8.3f} ms (kernel {k1:8.3f})  pass2 {t2:7.3f} ms (kernel {k2:7.3f})  segs {res[mode].numSegments()}", flush=True)
        out.free(); d.free()
    same = res["simple"].bit_equal(res["tile"])
    print(f"{name:22s} tile == simple: {same}", flush=True)
    if not same: bad += 1
print("BAD", bad)
sys.exit(1 if bad else 0)
