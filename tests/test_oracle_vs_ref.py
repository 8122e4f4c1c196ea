"""Pins the restatement against the compiled reference (oracle/_ref) on fresh seeded inputs.
CPU only; skipped where oracle/_ref could not be built (no /root/reference)."""
import numpy as np
import pytest

import util
from voroffset_b200 import synth

CASES = [
    ("torus_x", lambda: synth.torus_x(48, padding=5), 3.3),
    ("torus_z", lambda: synth.torus_z(40, padding=6), 5.0),
    ("blobs", lambda: synth.blobs(36, padding=5, seed=8), 4.0),
    ("random_thin", lambda: synth.random_volume(14, 12, kmax=6, padding=4, seed=3), 2.2),
    ("random_sparse", lambda: synth.random_volume(16, 10, kmax=3, padding=8, seed=5, fill=0.4), 7.7),
    ("single_column", lambda: synth.random_volume(1, 1, kmax=2, padding=4, seed=6), 3.5),
    ("integer_radius", lambda: synth.blobs(30, padding=6, seed=2), 5.0),
    ("radius_below_one", lambda: synth.blobs(24, padding=2, seed=3), 0.75),
]


@pytest.mark.parametrize("name,gen,radius", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("method", ["ours", "brute_force"])
def test_oracle_equals_reference_3d(oracle, reference, name, gen, radius, method):
    vol = gen()
    for op in ("dilation", "erosion", "opening", "closing"):
        want = reference.morph3d(vol, op, radius, method)
        got = oracle.morph3d(vol, op, radius, method)
        util.assert_same(got, want, op, method, f"oracle vs reference [{name}]")


def test_reference_threads_do_not_change_results(reference):
    vol = synth.blobs(40, padding=5, seed=11)
    a = reference.morph3d(vol, "dilation", 4.5, "ours", threads=1)
    b = reference.morph3d(vol, "dilation", 4.5, "ours", threads=4)
    assert a.bit_equal(b)


def test_ours_and_brute_force_agree_in_topology(reference):
    vol = synth.blobs(40, padding=6, seed=12)
    a = reference.morph3d(vol, "dilation", 5.5, "ours")
    b = reference.morph3d(vol, "dilation", 5.5, "brute_force")
    assert a.same_topology(b)
    assert np.abs(a.spans - b.spans).max() < 1e-12


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_equals_reference_2d(oracle, reference, seed):
    img = synth.random_image(60, 90, kmax=5, seed=seed)
    for op, r in [("dilate", 3.5 / 60), ("dilate", 6.0 / 60), ("erode", 2.0), ("erode", 3.7), ("open", 1.5 / 60),
                  ("close", 2.5 / 60), ("negate", 0.0)]:
        assert oracle.morph2d(img, op, r).bit_equal(reference.morph2d(img, op, r)), (op, r)


def test_xor_matches_reference(oracle, reference):
    a = synth.blobs(32, padding=3, seed=1)
    b = synth.blobs(32, padding=3, seed=2)
    va, xa = oracle.xor3d(a, b)
    vb, xb = reference.xor3d(a, b)
    assert xa.bit_equal(xb)
    assert abs(va - vb) <= 1e-12 * max(1.0, abs(vb))
    # xor of a volume with itself is empty (calculateXor is how the reference compares methods)
    v0, x0 = reference.xor3d(a, a)
    assert v0 == 0 and x0.numSegments() == 0
