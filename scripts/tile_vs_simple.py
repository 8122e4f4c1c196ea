import sys; sys.path.insert(0, ".")
from voroffset_b200 import synth, morpho, _lib
ctx = _lib.Context(0); op = morpho.make_operator("ours", ctx)
for name, vol, R in (("C3 lattice", synth.lattice(512, padding=10), 5.0), ("C1 torus_x", synth.torus_x(256), 8.0), ("blobs256 R8", synth.blobs(256, padding=10), 8.0)):
    d = morpho.DeviceVolume.upload(ctx, vol)
    for mode in ("simple", "tile"):
        ctx.set_option("pass1", mode)
        for i in range(4):
            ctx.mark(0); out, t1, t2 = op.morph_dev("dilation", d, R); ctx.mark(1); out.free()
        print(name, vol.nx, vol.ny, mode, "total", round(ctx.elapsed_ms(0, 1), 3), "p1", round(t1, 3), "p2", round(t2, 3), flush=True)
