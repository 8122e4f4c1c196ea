"""Known-answer fixtures for the dexeliser (mesh -> dexel volume, src/vor3d/Dexelize.cpp:56-225), derived with EXACT
rational arithmetic, independently of oracle/oracle.c and of the CUDA kernels:

    python tests/golden/make_dexel_known.py        ->  tests/golden/dexel_known.npz

The reference's dexeliser needs geogram and holds no test of its own, so these pin the part that is
easy to get wrong: the simulation-of-simplicity tie rules of point_in_triangle_2d (Dexelize.cpp:56-92). Every mesh is a
closed convex polyhedron (or two nested ones) whose vertices, edges and face diagonals pass EXACTLY through column
centres of the grid ((x + 0.5) spacing + origin with dyadic numbers). For a column whose centre lies strictly inside the
footprint - even when it lies on an edge or a vertex shared by several facets - the vertical line crosses the surface
exactly twice per shell: the expected crossings are the distinct exact heights of the facets whose closed projection
contains the centre. A centre exactly on the silhouette may count as inside or outside, but its crossing count must be
even and its crossings, if any, the exact heights there. Columns outside have none.
"""
import os
from fractions import Fraction as Fr

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def box(x0, x1, y0, y1, z0, z1, flip_top=False):
    V = [(x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0), (x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)]
    top = [(4, 5, 6), (4, 6, 7)] if not flip_top else [(4, 5, 7), (5, 6, 7)]
    F = [(0, 2, 1), (0, 3, 2)] + top + [(0, 1, 5), (0, 5, 4), (1, 2, 6), (1, 6, 5), (2, 3, 7), (2, 7, 6), (3, 0, 4), (3, 4, 7)]
    return V, F


def tetra():
    V = [(0.375, 0.375, 0.25), (2.375, 0.625, 0.25), (0.875, 2.125, 0.25), (0.875, 0.625, 1.875)]
    F = [(0, 2, 1), (0, 1, 3), (1, 2, 3), (2, 0, 3)]
    return V, F


def octa():
    c = (1.375, 1.375)
    V = [(c[0] - 1.0, c[1], 1.0), (c[0] + 1.0, c[1], 1.0), (c[0], c[1] - 1.0, 1.0), (c[0], c[1] + 1.0, 1.0), (c[0], c[1], 0.125), (c[0], c[1], 1.75)]
    F = [(0, 2, 5), (2, 1, 5), (1, 3, 5), (3, 0, 5), (2, 0, 4), (1, 2, 4), (3, 1, 4), (0, 3, 4)]
    return V, F


def merge(parts):
    """Several closed shells in one mesh: (V, F, facet count of each shell)."""
    V, F = [], []
    for v, f in parts:
        F += [tuple(i + len(V) for i in t) for t in f]
        V += v
    return V, F, [len(f) for _, f in parts]


CASES = {
    # name: (mesh, origin (x, y), spacing, nx, ny)
    "box_diagonals": (box(0.125, 2.125, 0.125, 1.125, 0.3, 1.7), (0.0, 0.0), 0.25, 10, 6),
    "box_flipped_top": (box(0.375, 1.875, 0.375, 1.875, -0.4, 0.9, flip_top=True), (0.0, 0.0), 0.25, 9, 9),
    "tetrahedron": (tetra(), (0.0, 0.0), 0.25, 11, 10),
    "octahedron": (octa(), (0.0, 0.0), 0.25, 12, 12),
    "nested_boxes": (merge([box(0.125, 2.625, 0.125, 2.625, 0.1, 2.9), box(0.625, 1.625, 0.875, 2.125, 0.8, 1.6, flip_top=True)]), (0.0, 0.0), 0.25, 12, 12),
    "fine_grid_tetra": (tetra(), (0.0625, 0.0625), 0.125, 20, 18),
}


def cross(ax, ay, bx, by):
    return ax * by - ay * bx


def classify(tri, p):
    """(closed containment of p in the xy-projection of tri, exact height there)."""
    (x1, y1, z1), (x2, y2, z2), (x3, y3, z3) = tri
    px, py = p
    d = cross(x2 - x1, y2 - y1, x3 - x1, y3 - y1)
    if d == 0:
        return False, None
    w1 = cross(x2 - px, y2 - py, x3 - px, y3 - py) / d
    w2 = cross(x3 - px, y3 - py, x1 - px, y1 - py) / d
    w3 = 1 - w1 - w2
    if w1 < 0 or w2 < 0 or w3 < 0:
        return False, None
    return True, w1 * z1 + w2 * z2 + w3 * z3


def footprint_state(Vf, Ff, p):
    """+1 strictly inside the union of the projected facets' interiors or on an edge interior to the footprint, 0 on the
    silhouette, -1 outside: a point is strictly inside iff a small circle around it is covered - checked exactly by
    testing the four diagonal neighbours at a distance far below every feature (the meshes live on a 1/16 lattice)."""
    eps = Fr(1, 1 << 20)
    inside_any = any(classify([Vf[i] for i in t], p)[0] for t in Ff)
    if not inside_any:
        return -1
    for dx, dy in ((eps, eps * 3 / 7), (-eps, eps * 5 / 11), (eps * 2 / 3, -eps), (-eps * 4 / 9, -eps)):
        q = (p[0] + dx, p[1] + dy)
        if not any(classify([Vf[i] for i in t], q)[0] for t in Ff):
            return 0
    return 1


def main():
    out = {"names": np.array(list(CASES))}
    for name, (mesh, origin, spacing, nx, ny) in CASES.items():
        V, F = mesh[0], mesh[1]
        shells, k = [], 0
        for n in (mesh[2] if len(mesh) > 2 else [len(F)]):
            shells.append(F[k:k + n])
            k += n
        Vf = [tuple(Fr(c) for c in v) for v in V]
        s, ox, oy = Fr(spacing), Fr(origin[0]), Fr(origin[1])
        state = np.zeros(nx * ny, dtype=np.int8)
        zoff, zs = [0], []
        for y in range(ny):
            for x in range(nx):
                p = ((x + Fr(1, 2)) * s + ox, (y + Fr(1, 2)) * s + oy)
                # per shell: strictly inside (its two exact heights are expected), outside, or on its silhouette (then the
                # column as a whole is only required to hold an even number of crossings out of the exact heights)
                sts = [footprint_state(Vf, sh, p) for sh in shells]
                st = 0 if 0 in sts else (1 if 1 in sts else -1)
                state[x + nx * y] = st
                if st >= 0:
                    use = [sh for sh, q in zip(shells, sts) if q >= 0]
                    hs = sorted({h / s for sh in use for ok, h in (classify([Vf[i] for i in t], p) for t in sh) if ok})
                    if st == 1:
                        assert len(hs) == 2 * sum(q == 1 for q in sts), (name, x, y, hs)
                    zs += [float(h) for h in hs]
                zoff.append(len(zs))
        n_in, n_edge = int((state == 1).sum()), int((state == 0).sum())
        assert n_in > 0 and n_edge > 0, name                    # every case must exercise the ties
        out[f"{name}__V"] = np.array(V, dtype=np.float64)
        out[f"{name}__F"] = np.array(F, dtype=np.int32)
        out[f"{name}__grid"] = np.array([origin[0], origin[1], spacing, nx, ny], dtype=np.float64)
        out[f"{name}__state"] = state
        out[f"{name}__zoff"] = np.array(zoff, dtype=np.int64)
        out[f"{name}__z"] = np.array(zs, dtype=np.float64)
        print(f"{name}: {nx}x{ny} columns, {n_in} inside, {n_edge} on the silhouette, {len(F)} facets")
    path = os.path.join(HERE, "dexel_known.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
