"""Times every BASELINE.json config on one GPU (resident data, CUDA events on the library stream) next to the
reference's own code on a bounded CPU sample. Output: one JSON line per config (also written to
gpurun_out/configs.jsonl). Development / reporting helper; bench.py is the contract benchmark."""
import json, os, sys, time
sys.path.insert(0, ".")
import numpy as np
from voroffset_b200 import synth, morpho, image2d, _lib

ctx = _lib.Context(0)
out = []
# pieces per column of the REFERENCE's mid volume for these exact volumes (its own first pass, oracle/_ref
# ref3d_mid_count = VoronoiVorPower.cpp:50-65, run once on the CPU): SURVEY.md 8(d)'s k_mid
K_MID = {"C1 torus_x -n 256 -r 8": 24.139, "C3 lattice -n 512 -p 10 -r 5": 23.590, "C4 torus_z -n 1024 -p 18 -r 16": 18.461,
         "C5 torus_z -n 2048 -r 32": 38.825}
try:
    PEAK = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    PEAK = 6548.2

def timed3d(name, vol, R, opn, method, reps=5):
    op = morpho.make_operator(method, ctx)
    d = morpho.DeviceVolume.upload(ctx, vol)
    ms = []
    for i in range(reps + 2):
        ctx.mark(0)
        r, t1, t2 = op.morph_dev(opn, d, R)
        ctx.mark(1)
        if i >= 2: ms.append(ctx.elapsed_ms(0, 1))
        k1, k2 = ctx.last_profile()
        nseg = r.info()[2]
        r.free()
    ncols = vol.nx * vol.ny
    prim = 2 if opn in ("opening", "closing") else 1
    rec = {"config": name, "grid": [vol.nx, vol.ny], "radius": R, "operation": opn, "method": method,
           "k_in": round(vol.numSegments() / ncols, 3), "k_out": round(nseg / ncols, 3),
           "ms": round(float(np.median(ms)), 4), "Mcolumns_per_s": round(ncols * prim / np.median(ms) / 1e3, 1)}
    if method == "ours" and opn == "dilation" and name in K_MID and k1 > 0:
        # the contract's roofline (bench.py: roofline) for the dominant kernel of pass 1 and for the whole dilation
        k_in, k_out, k_mid = vol.numSegments() / ncols, nseg / ncols, K_MID[name]
        b1, b2 = (4 + 16 * k_in) + (4 + 24 * k_mid), (4 + 24 * k_mid) + (4 + 16 * k_out)
        rec["roofline"] = {"bound": "hbm", "k_mid": k_mid, "algorithmic_bytes_per_column": round(b1, 1), "k_pass1_ms": round(k1, 4),
                           "achieved": round(b1 * ncols / (k1 * 1e-3) / 1e9, 1), "peak": PEAK, "unit": "GB/s",
                           "frac": round(b1 * ncols / (k1 * 1e-3) / 1e9 / PEAK, 3),
                           "whole_dilation_frac": round((b1 + b2) * ncols / (float(np.median(ms)) * 1e-3) / 1e9 / PEAK, 3)}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    d.free()

timed3d("C1 torus_x -n 256 -r 8", synth.torus_x(256), 8.0, "dilation", "ours")
timed3d("C3 lattice -n 512 -p 10 -r 5", synth.lattice(512, padding=10), 5.0, "dilation", "ours")
timed3d("C3 lattice -n 512 -p 10 -r 5", synth.lattice(512, padding=10), 5.0, "dilation", "brute_force")
v4 = synth.torus_z(1024, padding=18)
for opn in ("dilation", "erosion", "opening", "closing"):
    timed3d("C4 torus_z -n 1024 -p 18 -r 16", v4, 16.0, opn, "ours")
timed3d("C4 torus_z -n 1024 -p 18 -r 16", v4, 16.0, "dilation", "brute_force", reps=2)
v5 = synth.torus_z(2048)
timed3d("C5 torus_z -n 2048 -r 32", v5, 32.0, "dilation", "ours")
v5p = synth.torus_z(2048, padding=34)
for opn in ("erosion", "opening", "closing"):
    timed3d("C5 torus_z -n 2048 -p 34 -r 32", v5p, 32.0, opn, "ours", reps=3)
timed3d("C5 torus_z -n 2048 -r 32", v5, 32.0, "dilation", "brute_force", reps=1)

# C2: 2D, 2048 rows
for npoly in (64, 1024):
    img = synth.star_image(2048, 2048, npoly)
    for opn, r in (("dilate", 16.0 / 2048), ("erode", 4.0)):
        ms = []
        for i in range(6):
            d = image2d.DoubleCompressedImage.from_image(img, ctx)
            t = time.perf_counter()
            getattr(d, opn)(r)
            wall = (time.perf_counter() - t) * 1e3
            if i >= 2: ms.append((d.last_ms, wall))
        rec = {"config": f"C2 2D 2048 rows, {npoly} polygons", "intervals_in": img.numSegments(), "intervals_out": d.numSegments(),
               "operation": opn, "r": r, "device_ms": round(float(np.median([m[0] for m in ms])), 4),
               "call_ms_incl_copies": round(float(np.median([m[1] for m in ms])), 4)}
        print(json.dumps(rec), flush=True)
        out.append(rec)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/configs.jsonl", "w") as f:
    for r in out: f.write(json.dumps(r) + "\n")
