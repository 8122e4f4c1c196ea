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
