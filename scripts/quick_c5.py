"""C5 / C4 resident dilation: step, pass and kernel times (medians of 8 calls). Usage: quick_c5.py [key=value ...]"""
import sys
sys.path.insert(0, ".")
import numpy as np
from voroffset_b200 import synth, morpho, _lib
ctx = _lib.Context(0); op = morpho.make_operator("ours", ctx)
for kv in sys.argv[1:]:
    ctx.set_option(*kv.split("="))
for name, vol, R in (("c5", synth.torus_z(2048), 32.0), ("c4", synth.torus_z(1024, padding=18), 16.0)):
    d = morpho.DeviceVolume.upload(ctx, vol)
    rec = []
    for i in range(11):
        ctx.mark(0); r, t1, t2 = op.morph_dev("dilation", d, R); ctx.mark(1)
        k1, k2 = ctx.last_profile(); r.free()
        if i >= 3: rec.append((ctx.elapsed_ms(0, 1), t1, t2, k1, k2))
    m = np.median(np.array(rec), axis=0)
    print(name, "step %.4f pass1 %.4f pass2 %.4f k_tile %.4f k_pass2 %.4f  pass1-k_tile %.4f" % (m[0], m[1], m[2], m[3], m[4], m[1] - m[3]), flush=True)
    d.free()
