import glob
import os

import numpy as np

from voroffset_b200.volume import CompressedVolume, DexelImage

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Endpoint tolerance for COMPOSITES (opening / closing) of method 'ours' only. The second primitive of a
# composite runs on an input whose surfaces are exact ball offsets, so many (dx,dy) candidates tie
# mathematically and differ by an ulp in fp64; which one the reference's Voronoi pruning keeps is not
# reproducible by any other evaluation order (the reference's own 'ours' and 'brute_force' differ the same
# way). Topology must still match exactly. Primitives are compared bit for bit.
COMPOSITE_TOL = 1e-11   # dexels; north_star allows 1e-5


def golden_3d():
    return sorted(glob.glob(os.path.join(GOLDEN, "vol3d_*.npz")))


def golden_2d():
    return sorted(glob.glob(os.path.join(GOLDEN, "img2d_*.npz")))


def load_3d(path):
    z = np.load(path)
    vol = CompressedVolume(int(z["nx"]), int(z["ny"]), z["in_off"], z["in_spans"], tuple(z["origin"]),
                           tuple(z["extent"]), float(z["spacing"]), int(z["padding"]))
    return z, vol, float(z["radius"]), [str(o) for o in z["ops"]]


def expected_3d(z, vol, op, method):
    return vol.like(vol.nx, vol.ny, z[f"{op}__{method}__off"], z[f"{op}__{method}__spans"])


def load_2d(path):
    z = np.load(path)
    img = DexelImage(int(z["rows"]), int(z["width"]), z["in_off"], z["in_spans"])
    ops = [(str(o), float(r)) for o, r in zip(z["ops"], z["rs"])]
    return z, img, ops


def assert_same(got, want, op, method, what=""):
    assert got.same_topology(want), f"{what}: interval counts differ per column ({op}, {method})"
    if op in ("dilation", "erosion") or method == "brute_force":
        assert got.bit_equal(want), f"{what}: endpoints are not bit-identical ({op}, {method})"
    else:
        if got.spans.size:
            assert np.abs(got.spans - want.spans).max() <= COMPOSITE_TOL, f"{what}: endpoint beyond tolerance ({op}, {method})"


def checksum(vol) -> int:
    """Order-sensitive 64-bit checksum of a CSR volume (offsets and endpoint bit patterns)."""
    a = np.ascontiguousarray(vol.off).view(np.uint32).astype(np.uint64)
    b = np.ascontiguousarray(vol.spans).view(np.uint64).reshape(-1)
    h = np.uint64(1469598103934665603)
    with np.errstate(over="ignore"):
        for arr in (a, b):
            if arr.size:
                w = (np.arange(arr.size, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ arr
                h = h ^ np.bitwise_xor.reduce(w * np.uint64(0x100000001B3))
    return int(h)


def row_checksums(vol) -> np.ndarray:
    """Order-sensitive 64-bit checksum of the endpoint bit patterns of every row (uint64[ny]); with the per-column
    counts it pins a whole volume. Rows without intervals give 0."""
    nx, ny = vol.nx, vol.ny
    bits = np.ascontiguousarray(vol.spans, dtype=np.float64).view(np.uint64).reshape(-1)        # 2 per interval
    row_off = 2 * np.ascontiguousarray(vol.off).astype(np.int64)[::nx][:ny + 1]                  # first endpoint of each row
    out = np.zeros(ny, dtype=np.uint64)
    if bits.size == 0:
        return out
    row_of = np.repeat(np.arange(ny), np.diff(row_off))
    pos = (np.arange(bits.size, dtype=np.int64) - row_off[row_of]).astype(np.uint64)              # position inside its row
    with np.errstate(over="ignore"):
        w = ((pos * np.uint64(0x9E3779B97F4A7C15)) ^ bits) * np.uint64(0x100000001B3)
        w ^= w >> np.uint64(29)
    nz = np.nonzero(np.diff(row_off) > 0)[0]
    out[nz] = np.bitwise_xor.reduceat(w, row_off[nz])
    return out


def row_sums(vol):
    """(sum of z1, sum of z2) per row, float64[ny] each (np.add.reduceat: a fixed left-to-right order)."""
    nx, ny = vol.nx, vol.ny
    row_off = np.ascontiguousarray(vol.off).astype(np.int64)[::nx][:ny + 1]
    z1, z2 = np.zeros(ny), np.zeros(ny)
    nz = np.nonzero(np.diff(row_off) > 0)[0]
    if nz.size:
        sp = np.ascontiguousarray(vol.spans, dtype=np.float64).reshape(-1, 2)
        z1[nz] = np.add.reduceat(sp[:, 0], row_off[nz])
        z2[nz] = np.add.reduceat(sp[:, 1], row_off[nz])
    return z1, z2


# ---- full-size digests of the reference's output (tests/golden/make_golden_full.py) ----------------------------
def golden_full(name):
    return np.load(os.path.join(GOLDEN, f"full_{name}.npz"))


def assert_digest(vol, z, key, exact=True, what=""):
    """`vol` against the digest `key` ("in", "dilation", ...) of a full-size golden: per-column interval counts always
    exact; endpoint bits through the per-row checksums (exact=True) or per-row endpoint sums within COMPOSITE_TOL per
    interval (composites of 'ours')."""
    cnt = vol.counts()
    want = z[f"{key}__counts"]
    assert cnt.shape == want.shape and np.array_equal(cnt, want), f"{what} {key}: interval counts differ in {int((cnt != want).sum())} columns"
    assert int(z[f"{key}__nseg"]) == vol.numSegments()
    if exact:
        got = row_checksums(vol)
        bad = np.nonzero(got != z[f"{key}__rowsum"])[0]
        assert bad.size == 0, f"{what} {key}: endpoint bits differ in {bad.size} rows (first: {bad[:5].tolist()})"
    else:
        z1, z2 = row_sums(vol)
        per_row = cnt.reshape(vol.ny, vol.nx).sum(axis=1)
        tol = per_row * COMPOSITE_TOL + 1e-7
        assert np.all(np.abs(z1 - z[f"{key}__z1sum"]) <= tol) and np.all(np.abs(z2 - z[f"{key}__z2sum"]) <= tol), f"{what} {key}: endpoint sums beyond tolerance"
