"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref = the reference's own sources
compiled in place, see oracle/Makefile). Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

Each fixture stores a seeded input volume (CSR + metadata) and the reference's outputs for every
operation and both methods. The GPU box has no /root/reference: tests read these files instead.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.cpu import Reference  # noqa: E402
from voroffset_b200 import synth  # noqa: E402

OPS = ("dilation", "erosion", "opening", "closing")
METHODS = ("ours", "brute_force")

CASES_3D = {
    # name: (generator, radius, ops)
    "c1_torus_x_n256_r8": (lambda: synth.torus_x(256, padding=0), 8.0, ("dilation",)),       # BASELINE config 1
    "torus_x_n64_p6_r4.3": (lambda: synth.torus_x(64, padding=6), 4.3, OPS),
    "torus_z_n64_p7_r5.5": (lambda: synth.torus_z(64, padding=7), 5.5, OPS),
    "blobs_n48_p7_r6": (lambda: synth.blobs(48, padding=7), 6.0, OPS),
    "random_20x16_k5_p5_r3.7": (lambda: synth.random_volume(20, 16, kmax=5, padding=5, seed=21), 3.7, OPS),
    "lattice_n48_p6_r3": (lambda: synth.lattice(48, padding=6), 3.0, OPS),                    # config 3 in small
}

CASES_2D = {
    "stars_128": (lambda: synth.star_image(128, 128, 9, seed=4), [("dilate", 6.0 / 128), ("erode", 3.0), ("open", 2.0 / 128), ("close", 2.0 / 128), ("negate", 0.0)]),
    "rand2d_96x128": (lambda: synth.random_image(96, 128, kmax=5, seed=12), [("dilate", 4.5 / 96), ("erode", 2.5), ("open", 1.5 / 96), ("close", 1.5 / 96), ("negate", 0.0)]),
}


def main():
    ref = Reference()
    for name, (gen, radius, ops) in CASES_3D.items():
        v = gen()
        out = dict(in_off=v.off, in_spans=v.spans, nx=v.nx, ny=v.ny, origin=np.array(v.origin), extent=np.array(v.extent),
                   spacing=v.spacing, padding=v.padding, radius=radius, ops=np.array(ops))
        for op in ops:
            for m in METHODS:
                r = ref.morph3d(v, op, radius, m)
                assert (r.nx, r.ny) == (v.nx, v.ny)
                out[f"{op}__{m}__off"] = r.off
                out[f"{op}__{m}__spans"] = r.spans
        if "dilation" in ops:
            out["k_mid_pieces"] = ref.mid_count(v, radius)
        path = os.path.join(HERE, f"vol3d_{name}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB")
    for name, (gen, oplist) in CASES_2D.items():
        img = gen()
        out = dict(in_off=img.off, in_spans=img.spans, rows=img.rows, width=img.width,
                   ops=np.array([o for o, _ in oplist]), rs=np.array([r for _, r in oplist]))
        for i, (op, r) in enumerate(oplist):
            res = ref.morph2d(img, op, r)
            out[f"{i}__off"] = res.off
            out[f"{i}__spans"] = res.spans
        path = os.path.join(HERE, f"img2d_{name}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB")
    ingest(ref)


def ingest_curves():
    """Seeded closed polygons for DoubleCompressedImage::fromImage: stars in a grid, a rectangle with vertices ON scan
    lines and an edge along one, and a curve whose crossings go in front of the ray's content."""
    import math
    rng = np.random.default_rng(77)
    w, h, curves = 120, 90, []
    for k in range(12):
        cx, cy = (k % 4 + 0.5) * h / 4 + rng.uniform(-1, 1), (k // 4 + 0.5) * w / 3 + rng.uniform(-1, 1)
        n = int(rng.integers(3, 9))
        a = np.linspace(0, 2 * math.pi, 2 * n, endpoint=False) + rng.uniform(0, 1)
        rr = np.where(np.arange(2 * n) % 2 == 0, 9.0, 4.0)
        curves.append(np.stack([cx + rr * np.cos(a), cy + rr * np.sin(a)], 1))
    curves.append(np.array([[5, 5], [12, 5], [12, 9], [5, 9.5]], float))
    curves.append(np.array([[30, 1.0], [30, 2.5], [33, 2.5], [36, 4], [36, 1.0]], float))
    return w, h, curves


def ingest(ref):
    """2D ingestion fixture: fromImage (DoubleCompressedImage.cpp:25-111) and transposeInPlace (:478-584) outputs of
    the reference."""
    w, h, curves = ingest_curves()
    img = ref.from_image(w, h, curves)
    src = synth.random_image(37, 53, kmax=4, seed=4)
    tr = ref.transposed(src)
    path = os.path.join(HERE, "ingest2d.npz")
    np.savez_compressed(path, w=w, h=h, curve_off=np.cumsum([0] + [len(c) for c in curves]), curve_pts=np.concatenate(curves),
                        img_off=img.off, img_spans=img.spans, t_in_off=src.off, t_in_spans=src.spans, t_rows=src.rows,
                        t_width=src.width, t_off=tr.off, t_spans=tr.spans)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
