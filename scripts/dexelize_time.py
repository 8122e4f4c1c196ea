"""Times the device dexeliser (vo_dexelize_dev) on BASELINE-sized grids and, beside it, the host loop of the re-hosted
offset3d (-x noop: single thread with column buckets, the counterpart of the reference's serial compute_sign loop,
Dexelize.cpp:182-219) on the same mesh. Usage: python scripts/dexelize_time.py [out.jsonl]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from voroffset_b200 import _lib, synth  # noqa: E402
from voroffset_b200.dexelize import dexelize_dev, grid_for, save_obj  # noqa: E402

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
ctx = _lib.Context(0)
cases = [
    ("torus 1024x256 quads", synth.torus_mesh(1024, 256), 2048, 0),
    ("torus 2048x512 quads", synth.torus_mesh(2048, 512), 2048, 0),
    ("torus 256x128 quads", synth.torus_mesh(256, 128), 1024, 0),
    ("box, 12 facets", synth.box_mesh(), 2048, 0),
    ("3000 boxes", synth.boxes_mesh(3000), 1024, 4),
]
for name, (V, F), n, pad in cases:
    grid = grid_for(V, None, pad, n)
    times = []
    for _ in range(6):
        dv, ms = dexelize_dev(ctx, V, F, grid)
        times.append(ms)
        nsp = dv.info()[2]
        dv.free()
    t0 = time.perf_counter()
    dv, _ = dexelize_dev(ctx, V, F, grid)
    wall = (time.perf_counter() - t0) * 1e3
    dv.free()
    rec = {"mesh": name, "facets": int(F.shape[0]), "grid": [grid.nx, grid.ny], "intervals": int(nsp),
           "device_ms": float(np.median(times[1:])), "call_wall_ms_incl_mesh_upload": wall,
           "columns_per_s": grid.nx * grid.ny / (float(np.median(times[1:])) * 1e-3)}
    if "--host" in sys.argv or F.shape[0] <= 600000:
        with tempfile.TemporaryDirectory() as d:
            save_obj(os.path.join(d, "m.obj"), V, F)
            t0 = time.perf_counter()
            r = subprocess.run([os.path.join(ROOT, "voroffset_b200", "cpp", "bin", "offset3d"), os.path.join(d, "m.obj"),
                                "-n", str(n), "-p", str(pad), "-x", "noop"], capture_output=True, text=True)
            rec["host_loop_wall_ms_incl_obj_parse"] = (time.perf_counter() - t0) * 1e3 if r.returncode == 0 else None
    out.write(json.dumps(rec) + "\n")
    out.flush()
