"""Quick GPU parity sweep against the CPU checkers (development helper; run under gpurun)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from voroffset_b200 import synth, morpho, image2d, _lib
from oracle.cpu import Oracle, Reference, reference_available

o = Oracle(threads=16)
ref = Reference() if reference_available() else None
ctx = _lib.Context(0)
bad = 0

def cmp(a, b, name):
    global bad
    topo = a.same_topology(b)
    md = float(np.abs(a.spans - b.spans).max()) if topo and a.spans.size else (0.0 if topo else -1)
    bit = a.bit_equal(b)
    print(f"{name:60s} topo={topo} bit={bit} maxdiff={md:.3g} segs={a.numSegments()}/{b.numSegments()}", flush=True)
    if not topo: bad += 1
    return bit

cases = [
    ("torus_x n128 p8 R5.5", synth.torus_x(128, padding=8), 5.5),
    ("torus_z n96 p6 R4.3", synth.torus_z(96, padding=6), 4.3),
    ("blobs n80 p9 R8", synth.blobs(80, padding=9), 8.0),
    ("random 40x32 k8 p6 R4.3", synth.random_volume(40, 32, kmax=8, padding=6), 4.3),
    ("lattice n96 p8 R5", synth.lattice(96, padding=8), 5.0),
    ("tiny 3x2 R2.5", synth.random_volume(3, 2, kmax=3, seed=1, padding=0), 2.5),
    ("empty 5x4 R3", synth.random_volume(5, 4, kmax=0, seed=1, padding=1), 3.0),
    ("R0.5", synth.random_volume(9, 7, kmax=3, seed=2, padding=1), 0.5),
]
for name, v, R in cases:
    for method in ("ours", "brute_force"):
        op = morpho.make_operator(method, ctx)
        for opn in ("dilation", "erosion", "opening", "closing"):
            t = time.time()
            got, t1, t2 = morpho.apply_operation(op, opn, v, R)
            dt = time.time() - t
            want = o.morph3d(v, opn, R, method)
            ok = cmp(got, want, f"{name} {method} {opn} [{dt*1e3:.1f} ms, p1 {t1:.2f} p2 {t2:.2f}]")
            if opn in ("dilation", "erosion") and not ok: bad += 1
            if ref is not None and opn in ("dilation",) and v.nx * v.ny < 30000:
                cmp(got, ref.morph3d(v, opn, R, method), "     vs reference")

for name, img in [("rand2d", synth.random_image(200, 300, kmax=5)), ("stars", synth.star_image(256, 256, 16))]:
    for opn, r in [("dilate", 8.0 / img.rows), ("dilate", 5.5 / img.rows), ("erode", 4.0), ("erode", 2.5),
                   ("open", 3.0 / img.rows), ("close", 3.0 / img.rows), ("negate", 0.0)]:
        d = image2d.DoubleCompressedImage.from_image(img, ctx)
        if opn == "negate": d.negate()
        else: getattr(d, opn)(r)
        w = o.morph2d(img, opn, r)
        print(f"2D {name} {opn} {r:.4g}: bit={d.bit_equal(w)} segs={d.numSegments()}/{w.numSegments()}", flush=True)
        if not d.bit_equal(w): bad += 1

# xor
a = synth.blobs(64, padding=4, seed=1); b = synth.blobs(64, padding=4, seed=2)
op = morpho.make_operator("ours", ctx)
vol, x = op.calculateXor(a, b)
wvol, wx = o.xor3d(a, b)
print("xor", x.bit_equal(wx), vol, wvol, flush=True)
if ref is not None:
    rvol, rx = ref.xor3d(a, b)
    print("xor vs ref", x.bit_equal(rx), vol, rvol)
if not x.bit_equal(wx): bad += 1

# a bigger timing case
v = synth.torus_z(512)
op = morpho.make_operator("ours", ctx)
for i in range(3):
    t = time.time(); got, t1, t2 = op.dilation(v, 16.0); dt = time.time() - t
    print(f"torus_z 512 R16 ours: e2e {dt*1e3:.1f} ms pass1 {t1:.2f} pass2 {t2:.2f} segs {got.numSegments()}")
want = o.morph3d(v, "dilation", 16.0, "ours")
cmp(got, want, "torus_z 512 R16 ours vs oracle")
print("launches", ctx.launches)
print("BAD", bad)
sys.exit(1 if bad else 0)
