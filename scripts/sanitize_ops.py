"""Small instances of every operation (tile kernel forced, multi-interval columns, erosion, composites, 2D, xor,
host-buffer pipeline) for a compute-sanitizer run; results are checked against the oracle."""
import sys
sys.path.insert(0, ".")
from oracle.cpu import Oracle
from voroffset_b200 import _lib, image2d, morpho, synth

ctx = _lib.Context(0)
orc = Oracle(threads=4)
ok = True
for name, vol, R in [("torus_z 96", synth.torus_z(96, padding=7), 6.5), ("lattice 64", synth.lattice(64, padding=4), 3.0),
                     ("blobs 80", synth.blobs(80, count=20, padding=5, seed=2), 4.2)]:
    for method in ("ours", "brute_force"):
        op = morpho.make_operator(method, ctx)
        for opn in ("dilation", "erosion", "opening", "closing"):
            got, _, _ = morpho.apply_operation(op, opn, vol, R)
            want = orc.morph3d(vol, opn, R, method)
            same = got.same_topology(want)
            ok &= same
            print(name, method, opn, "topology equal:", same, flush=True)
ctx.set_option("pass1", "tile")
vol = synth.torus_z(200, padding=0)
got, _, _ = morpho.make_operator("ours", ctx).dilation(vol, 9.0)
ok &= got.bit_equal(orc.morph3d(vol, "dilation", 9.0, "ours"))
img = synth.star_image(96, 96, 9)
d = image2d.DoubleCompressedImage.from_image(img, ctx)
d.dilate(3.0 / 96)
ok &= d.bit_equal(orc.morph2d(img, "dilate", 3.0 / 96))
# mesh -> dexels -> offset on the device, everything released explicitly so that --leak-check sees a clean exit
from voroffset_b200.dexelize import dexelize_dev, grid_for
V, F = synth.boxes_mesh(40)
grid = grid_for(V, None, 2, 72)
dv, _ = dexelize_dev(ctx, V, F, grid)
ok &= dv.download(grid).bit_equal(orc.dexelize(V, F, grid))
out, _, _ = morpho.make_operator("ours", ctx).morph_dev("closing", dv, 3.0)
out.free(); dv.free(); d.free() if hasattr(d, "free") else None
del d
ctx.close()
print("all equal:", ok)
sys.exit(0 if ok else 1)
