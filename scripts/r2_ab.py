"""Round-2 A/B timings on one GPU: operations of the BASELINE configs under library options
(usage: python scripts/r2_ab.py [case ...]; output: one JSON line per (case, options))."""
import json, sys
sys.path.insert(0, ".")
import numpy as np
from voroffset_b200 import synth, morpho, _lib

CASES = {
    "c5_dil": (lambda: synth.torus_z(2048), 32.0, "dilation"),
    "c5_ero": (lambda: synth.torus_z(2048, padding=34), 32.0, "erosion"),
    "c5_clo": (lambda: synth.torus_z(2048, padding=34), 32.0, "closing"),
    "c4_dil": (lambda: synth.torus_z(1024, padding=18), 16.0, "dilation"),
    "c4_ero": (lambda: synth.torus_z(1024, padding=18), 16.0, "erosion"),
    "c3_dil": (lambda: synth.lattice(512, padding=10), 5.0, "dilation"),
    "c1_dil": (lambda: synth.torus_x(256), 8.0, "dilation"),
}
OPTS = [{}, {"scan": "classic"}, {"tile_dbuf": "on"}, {"tile_dbuf": "off"}, {"tile_lean": "off"}, {"tile_lean": "on"}]
DEFAULTS = {"scan": "fused", "tile_dbuf": "auto", "tile_lean": "auto"}

ctx = _lib.Context(0)
op = morpho.make_operator("ours", ctx)
for name in (sys.argv[1:] or list(CASES)):
    gen, R, opn = CASES[name]
    vol = gen()
    d = morpho.DeviceVolume.upload(ctx, vol)
    for opts in OPTS:
        for k, v in {**DEFAULTS, **opts}.items():
            ctx.set_option(k, v)
        ms, p1, p2, k1, k2 = [], [], [], [], []
        for i in range(8):
            ctx.mark(0)
            out, t1, t2 = op.morph_dev(opn, d, R)
            ctx.mark(1)
            if i >= 3:
                ms.append(ctx.elapsed_ms(0, 1)); p1.append(t1); p2.append(t2)
                a, b = ctx.last_profile(); k1.append(a); k2.append(b)
            out.free()
        print(json.dumps({"case": name, "opts": opts, "ms": round(float(np.median(ms)), 4), "pass1": round(float(np.median(p1)), 4),
                          "pass2": round(float(np.median(p2)), 4), "k_pass1": round(float(np.median(k1)), 4), "k_pass2": round(float(np.median(k2)), 4)}), flush=True)
    d.free()
