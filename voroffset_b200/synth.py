"""Seeded synthetic dexel volumes / images in generic position.

These stand in for `vor3d::create_dexels` output (src/vor3d/Dexelize.cpp:231-274): the grid follows
the `CompressedVolume` constructor (CompressedVolume.cpp:11-23), column centres are
`(i + 0.5) * spacing + origin` (Dexelize.cpp:166-225) and z is stored as `world_z / spacing`.
All shapes are shifted by small irrational offsets so that no two interval endpoints coincide
exactly (the reference's brute_force method is tie-unstable on degenerate inputs).
"""
from __future__ import annotations

import math

import numpy as np

from .volume import CompressedVolume, DexelImage, csr_from_lists

_JX, _JY, _JZ = 0.0137, 0.0071, 0.0093  # generic-position jitter, in dexels


def _grid(extent, n, padding, min_corner):
    spacing = max(extent) / n
    vol = CompressedVolume.from_box(min_corner, extent, spacing, padding)
    xs = (np.arange(vol.nx) + 0.5) * spacing + vol.origin[0]
    ys = (np.arange(vol.ny) + 0.5) * spacing + vol.origin[1]
    return vol, xs, ys, spacing


def _single_interval_volume(vol, lo, hi, mask):
    """One interval [lo, hi] where mask, none elsewhere. lo/hi/mask are (ny, nx)."""
    cnt = mask.reshape(-1).astype(np.int64)
    off = np.zeros(cnt.size + 1, dtype=np.int64)
    np.cumsum(cnt, out=off[1:])
    spans = np.stack([lo.reshape(-1)[mask.reshape(-1)], hi.reshape(-1)[mask.reshape(-1)]], axis=1)
    return vol.like(vol.nx, vol.ny, off, spans)


def torus_z(n: int, padding: int = 0, major: float = 1.0, minor: float = 0.35) -> CompressedVolume:
    """Torus with axis z (square footprint, at most one interval per column): configs C4 / C5."""
    ext = (2 * (major + minor), 2 * (major + minor), 2 * minor)
    corner = (-(major + minor), -(major + minor), -minor)
    vol, xs, ys, sp = _grid(ext, n, padding, corner)
    X, Y = np.meshgrid(xs + _JX * sp, ys + _JY * sp)
    rho = np.sqrt(X * X + Y * Y)
    d2 = minor * minor - (rho - major) ** 2
    mask = d2 > 0
    half = np.sqrt(np.where(mask, d2, 0.0))
    zc = _JZ * sp
    return _single_interval_volume(vol, (zc - half) / sp, (zc + half) / sp, mask)


def torus_z_rows(n: int, y0: int, y1: int, padding: int = 0, major: float = 1.0, minor: float = 0.35) -> CompressedVolume:
    """Rows [y0, y1) of `torus_z(n, padding)` (a y-slab of a grid sharded over GPUs) without building the whole grid:
    the same elementwise arithmetic, hence the same bits."""
    ext = (2 * (major + minor), 2 * (major + minor), 2 * minor)
    corner = (-(major + minor), -(major + minor), -minor)
    vol, xs, ys, sp = _grid(ext, n, padding, corner)
    y0, y1 = max(0, y0), min(vol.ny, y1)
    X, Y = np.meshgrid(xs + _JX * sp, ys[y0:y1] + _JY * sp)
    rho = np.sqrt(X * X + Y * Y)
    d2 = minor * minor - (rho - major) ** 2
    mask = d2 > 0
    half = np.sqrt(np.where(mask, d2, 0.0))
    zc = _JZ * sp
    cnt = mask.reshape(-1).astype(np.int64)
    off = np.zeros(cnt.size + 1, dtype=np.int64)
    np.cumsum(cnt, out=off[1:])
    lo, hi = (zc - half) / sp, (zc + half) / sp
    spans = np.stack([lo.reshape(-1)[mask.reshape(-1)], hi.reshape(-1)[mask.reshape(-1)]], axis=1)
    return vol.like(vol.nx, y1 - y0, off, spans)


def torus_x(n: int, padding: int = 0, major: float = 1.0, minor: float = 0.35) -> CompressedVolume:
    """Torus with axis x (up to two intervals per column): config C1 (grid 67 x 256 at n=256)."""
    ext = (2 * minor, 2 * (major + minor), 2 * (major + minor))
    corner = (-minor, -(major + minor), -(major + minor))
    vol, xs, ys, sp = _grid(ext, n, padding, corner)
    X, Y = np.meshgrid(xs + _JX * sp, ys + _JY * sp)
    zc = _JZ * sp
    # point (x, y, z) inside  <=>  (sqrt(y^2+z^2) - major)^2 + x^2 < minor^2
    w2 = minor * minor - X * X
    lists = []
    flat_w2 = w2.reshape(-1)
    flat_y = Y.reshape(-1)
    for w2c, y in zip(flat_w2, flat_y):
        if w2c <= 0:
            lists.append(())
            continue
        w = math.sqrt(w2c)
        ro, ri = major + w, major - w          # outer / inner radius of the annulus in the (y,z) plane
        if abs(y) >= ro:
            lists.append(())
            continue
        zo = math.sqrt(ro * ro - y * y)
        if abs(y) >= ri:
            lists.append(((zc - zo) / sp, (zc + zo) / sp))
        else:
            zi = math.sqrt(ri * ri - y * y)
            lists.append(((zc - zo) / sp, (zc - zi) / sp, (zc + zi) / sp, (zc + zo) / sp))
    off, spans = csr_from_lists(lists)
    return vol.like(vol.nx, vol.ny, off, spans)


def lattice(n: int, padding: int = 10, seed: int = 3) -> CompressedVolume:
    """Filigree-like jittered lattice of axis-aligned bars (config C3: -n 512 -p 10).

    Bars of width 5..8 dexels with period 16..24 along x, y and z; every bar edge is at
    `k*period + eps*(1 + 0.37 k)` so no two endpoints coincide.
    """
    rng = np.random.RandomState(seed)
    eps = 0.0137
    ext = (1.0, 1.0, 1.0)
    vol, xs, ys, sp = _grid(ext, n, padding, (0.0, 0.0, 0.0))

    def bars(length):
        out, k, pos = [], 0, 2.0 + rng.uniform(0, 4)
        while pos < length - 10:
            w = rng.uniform(5.0, 8.0)
            a = pos + eps * (1 + 0.37 * k)
            out.append((a, a + w))
            pos += rng.uniform(16.0, 24.0)
            k += 1
        return out

    bx, by, bz = bars(n), bars(n), bars(n)
    p = padding

    def inside(bs, coord):
        m = np.zeros(coord.shape, dtype=bool)
        for a, b in bs:
            m |= (coord > a) & (coord < b)
        return m

    cx = np.arange(vol.nx) - p + 0.5 + _JX
    cy = np.arange(vol.ny) - p + 0.5 + _JY
    in_x = inside(bx, cx)
    in_y = inside(by, cy)
    grid_ok_x = (cx > 0) & (cx < n)
    grid_ok_y = (cy > 0) & (cy < n)
    z0, z1 = 1.0 + _JZ, n - 1.0 - _JZ * 0.7

    def jit(i, j, k):                                    # per-column pseudo-random jitter in [0, 0.02)
        t = math.sin(i * 12.9898 + j * 78.233 + k * 37.719) * 43758.5453
        return 0.02 * (t - math.floor(t))

    lists = []
    for j in range(vol.ny):
        for i in range(vol.nx):
            if not (grid_ok_x[i] and grid_ok_y[j]):
                lists.append(())
            elif in_x[i] and in_y[j]:
                lists.append((z0 + jit(i, j, 0), z1 - jit(i, j, 1)))                      # z-bar
            elif in_x[i] or in_y[j]:
                ev = []
                for k, (a, b) in enumerate(bz):          # x- or y-bar seen end-on: one interval per z level
                    ev += [a + jit(i, j, 2 * k + 2), b + jit(i, j, 2 * k + 3)]
                lists.append(ev)
            else:
                lists.append(())
    off, spans = csr_from_lists(lists)
    return vol.like(vol.nx, vol.ny, off, spans)


def random_volume(nx: int, ny: int, kmax: int = 8, seed: int = 7, zrange: float = 64.0,
                  fill: float = 1.0, padding: int = 0) -> CompressedVolume:
    """Multi-interval stress volume: k ~ U{0..kmax} intervals per column, lengths U[0.3, 6].

    `padding` behaves like offset3d's -p: `padding` empty columns on every side and `padding`
    dexels of head-room below and above the data in z (CompressedVolume.cpp:17-20,
    VoronoiVorPower.cpp:28-29), which closing / opening need to stay inside [zmin, zmax].
    """
    rng = np.random.RandomState(seed)
    inner = {}
    zhi = 1.0
    for j in range(ny):
        for i in range(nx):
            k = rng.randint(0, kmax + 1) if rng.uniform() < fill else 0
            ev, z = [], rng.uniform(0, 5)
            for _ in range(k):
                z += rng.uniform(0.05, zrange / max(kmax, 1))
                a = z
                z += rng.uniform(0.3, 6.0)
                ev += [a, z]
            inner[(i, j)] = ev
            if ev:
                zhi = max(zhi, ev[-1])
    zhi = math.ceil(zhi) + 1.0
    p = padding
    gx, gy = nx + 2 * p, ny + 2 * p
    lists = [inner.get((i - p, j - p), ()) for j in range(gy) for i in range(gx)]
    off, spans = csr_from_lists(lists)
    return CompressedVolume(gx, gy, off, spans, origin=(-float(p), -float(p), -float(p)),
                            extent=(float(nx), float(ny), zhi), spacing=1.0, padding=p)


def blobs(n: int, count: int = 24, seed: int = 5, padding: int = 0, rmin: float = 0.06, rmax: float = 0.2) -> CompressedVolume:
    """Union of random balls in the unit cube: smooth multi-interval columns (thick enough to survive
    erosion / opening), used by the parity tests next to the tori."""
    rng = np.random.RandomState(seed)
    vol, xs, ys, sp = _grid((1.0, 1.0, 1.0), n, padding, (0.0, 0.0, 0.0))
    cen = rng.uniform(0.2, 0.8, size=(count, 3))
    rad = rng.uniform(rmin, rmax, size=count)
    X, Y = np.meshgrid(xs + _JX * sp, ys + _JY * sp)
    per = [[] for _ in range(vol.nx * vol.ny)]
    for (cx, cy, cz), r in zip(cen, rad):
        d2 = r * r - (X - cx) ** 2 - (Y - cy) ** 2
        idx = np.nonzero(d2.reshape(-1) > 0)[0]
        half = np.sqrt(d2.reshape(-1)[idx])
        for c, h in zip(idx, half):
            per[c].append(((cz - h) / sp, (cz + h) / sp))
    lists = []
    for iv in per:
        iv.sort()
        merged = []
        for a, b in iv:
            if merged and a <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], b)
            else:
                merged.append([a, b])
        lists.append([v for ab in merged for v in ab])
    off, spans = csr_from_lists(lists)
    return vol.like(vol.nx, vol.ny, off, spans)


def star_image(rows: int = 2048, width: int = 2048, n_polys: int = 64, seed: int = 1234) -> DexelImage:
    """Config C2: star polygons (5..12 spikes) scan-converted at row centres, rows = sweep axis.

    This plays the role of `DoubleCompressedImage::fromImage` (src/vor2d/DoubleCompressedImage.cpp:
    25-111), which is upstream of the hot path; parity is pinned at the dexel-image boundary.
    """
    rng = np.random.RandomState(seed)
    g = int(math.ceil(math.sqrt(n_polys)))
    cell_r, cell_c = rows / g, width / g
    per_row = [[] for _ in range(rows)]
    for p in range(n_polys):
        cy = (p // g + 0.5) * cell_r + rng.uniform(-0.1, 0.1) * cell_r
        cx = (p % g + 0.5) * cell_c + rng.uniform(-0.1, 0.1) * cell_c
        nv = rng.randint(5, 13)
        ang0 = rng.uniform(0, 2 * math.pi)
        pts = []
        for v in range(2 * nv):
            rad = min(cell_r, cell_c) * (rng.uniform(0.25, 0.35) if v % 2 == 0 else rng.uniform(0.09, 0.18))
            a = ang0 + math.pi * v / nv
            pts.append((cy + rad * math.sin(a), cx + rad * math.cos(a)))
        pts = np.array(pts)
        r0 = max(0, int(math.floor(pts[:, 0].min())) - 1)
        r1 = min(rows - 1, int(math.ceil(pts[:, 0].max())) + 1)
        q = np.roll(pts, -1, axis=0)
        for i in range(r0, r1 + 1):
            yy = i + 0.5 + 0.00173
            crossing = (pts[:, 0] <= yy) != (q[:, 0] <= yy)
            if not crossing.any():
                continue
            t = (yy - pts[crossing, 0]) / (q[crossing, 0] - pts[crossing, 0])
            xs = np.sort(pts[crossing, 1] + t * (q[crossing, 1] - pts[crossing, 1]))
            xs = np.clip(xs, 0.0, float(width))
            for k in range(0, len(xs) - 1, 2):
                if xs[k + 1] > xs[k]:
                    per_row[i].append((xs[k], xs[k + 1]))
    lists = []
    for i in range(rows):
        iv = sorted(per_row[i])
        merged = []
        for a, b in iv:
            if merged and a <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], b)
            else:
                merged.append([a, b])
        lists.append([v for ab in merged for v in ab])
    return DexelImage.from_lists(width, lists)


def random_image(rows: int, width: int, kmax: int = 6, seed: int = 11) -> DexelImage:
    rng = np.random.RandomState(seed)
    lists = []
    for _ in range(rows):
        k = rng.randint(0, kmax + 1)
        pts = np.sort(rng.uniform(0.01, width - 0.01, size=2 * k))
        lists.append(pts.tolist())
    return DexelImage.from_lists(width, lists)


# ---- triangle meshes (input of the dexeliser, Dexelize.cpp:231-274) -----------------------------------------------
def torus_mesh(nu: int = 256, nv: int = 128, major: float = 1.0, minor: float = 0.35, axis: str = "z"):
    """Closed triangulated torus, `nu` x `nv` quads cut in two; vertices slightly rotated / shifted so that no edge is
    parallel to the grid and no vertex sits on a column centre (generic position)."""
    u = (np.arange(nu) + 0.137) * (2 * math.pi / nu)
    v = (np.arange(nv) + 0.291) * (2 * math.pi / nv)
    U, W = np.meshgrid(u, v, indexing="ij")
    rho = major + minor * np.cos(W)
    P = np.stack([rho * np.cos(U), rho * np.sin(U), minor * np.sin(W)], axis=-1).reshape(-1, 3)
    if axis == "x":
        P = P[:, [2, 0, 1]]
    P = P + np.array([1.3e-4, -2.1e-4, 0.7e-4])
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = (i * nv + j).reshape(-1)
    b = (((i + 1) % nu) * nv + j).reshape(-1)
    c = (((i + 1) % nu) * nv + (j + 1) % nv).reshape(-1)
    d = (i * nv + (j + 1) % nv).reshape(-1)
    F = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.int32)
    return np.ascontiguousarray(P), np.ascontiguousarray(F)


def box_mesh(lo=(-1.0, -0.8, -0.6), hi=(1.0, 0.8, 0.6), tilt: float = 0.0113):
    """Twelve large triangles (every facet covers a big part of the grid); `tilt` shears the box so that the side walls
    are not vertical (a vertical facet has a flat projection and never yields a crossing, Dexelize.cpp:150-157)."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    P = np.array([[x, y, z] for z in (lo[2], hi[2]) for y in (lo[1], hi[1]) for x in (lo[0], hi[0])], dtype=np.float64)
    P[:, 0] += tilt * P[:, 2] + 0.37 * tilt * P[:, 1]
    P[:, 1] += 0.71 * tilt * P[:, 2]
    quads = [(0, 2, 3, 1), (4, 5, 7, 6), (0, 1, 5, 4), (2, 6, 7, 3), (0, 4, 6, 2), (1, 3, 7, 5)]
    F = np.array([t for q in quads for t in ((q[0], q[1], q[2]), (q[0], q[2], q[3]))], dtype=np.int32)
    return P, F


def boxes_mesh(count: int = 200, seed: int = 5, span: float = 1.0):
    """`count` small sheared boxes at random places (overlapping: several crossings per column, like the lattice)."""
    rng = np.random.default_rng(seed)
    Vs, Fs = [], []
    for k in range(count):
        c = rng.uniform(-span, span, 3)
        h = rng.uniform(0.03, 0.12, 3) * span
        P, F = box_mesh(c - h, c + h, tilt=0.0113 + 0.001 * (k % 7))
        Fs.append(F + 8 * k)
        Vs.append(P)
    return np.concatenate(Vs), np.concatenate(Fs).astype(np.int32)


def open_patch_mesh(n: int = 24):
    """An open height-field patch above a closed box: columns under the patch only see an odd number of crossings."""
    P, F = box_mesh((-1.0, -1.0, -0.5), (1.0, 1.0, 0.1))
    g = np.linspace(-0.613, 0.577, n)
    X, Y = np.meshgrid(g, g, indexing="ij")
    Z = 0.6 + 0.1 * np.sin(3 * X) * np.cos(2 * Y)
    Q = np.stack([X, Y, Z], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(n - 1), np.arange(n - 1), indexing="ij")
    a = (i * n + j).reshape(-1) + 8
    G = np.concatenate([np.stack([a, a + n, a + n + 1], 1), np.stack([a, a + n + 1, a + 1], 1)]).astype(np.int32)
    return np.concatenate([P, Q]), np.concatenate([F, G]).astype(np.int32)
