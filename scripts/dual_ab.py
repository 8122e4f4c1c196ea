"""Erosion / opening / closing at C5 and C4 sizes with the dual form on and off (resident, CUDA events): ms per call."""
import json, sys
sys.path.insert(0, ".")
import numpy as np
from voroffset_b200 import synth, morpho, _lib

ctx = _lib.Context(0)
op = morpho.make_operator("ours", ctx)
for name, vol, R in (("c5", synth.torus_z(2048, padding=34), 32.0), ("c4", synth.torus_z(1024, padding=18), 16.0)):
    d = morpho.DeviceVolume.upload(ctx, vol)
    res = {}
    for mode in ("general", "auto"):
        ctx.set_option("erosion", mode)
        for opn in ("erosion", "opening", "closing"):
            ms = []
            for i in range(7):
                l0 = ctx.launches
                ctx.mark(0)
                r, t1, t2 = op.morph_dev(opn, d, R)
                ctx.mark(1)
                if i >= 2: ms.append(ctx.elapsed_ms(0, 1))
                k1, k2 = ctx.last_profile()
                res.setdefault((mode, opn), r.download())
                r.free()
            print(json.dumps({"case": name, "erosion_mode": mode, "operation": opn, "ms": round(float(np.median(ms)), 4),
                              "time_1": round(t1, 4), "time_2": round(t2, 4), "k_pass1": round(k1, 4), "k_pass2": round(k2, 4),
                              "launches": ctx.launches - l0}), flush=True)
    for opn in ("erosion", "opening", "closing"):
        a, b = res[("general", opn)], res[("auto", opn)]
        same = a.bit_equal(b)
        dmax = float(np.abs(a.spans - b.spans).max()) if a.same_topology(b) and a.spans.size else -1.0
        print(json.dumps({"case": name, "operation": opn, "dual_equals_general_bitwise": bool(same), "same_topology": bool(a.same_topology(b)), "max_abs_diff": dmax}), flush=True)
    d.free()
