"""The plain-C restatement (oracle/oracle.c) against the committed outputs of the reference itself
(tests/golden, produced by tests/golden/make_golden.py from oracle/_ref). CPU only."""
import pytest

import util


@pytest.mark.parametrize("path", util.golden_3d(), ids=lambda p: p.split("vol3d_")[-1][:-4])
def test_oracle_matches_reference_outputs_3d(oracle, path):
    z, vol, radius, ops = util.load_3d(path)
    for op in ops:
        for method in ("ours", "brute_force"):
            got = oracle.morph3d(vol, op, radius, method)
            util.assert_same(got, util.expected_3d(z, vol, op, method), op, method, "oracle vs golden")


@pytest.mark.parametrize("path", util.golden_2d(), ids=lambda p: p.split("img2d_")[-1][:-4])
def test_oracle_matches_reference_outputs_2d(oracle, path):
    z, img, ops = util.load_2d(path)
    for i, (op, r) in enumerate(ops):
        got = oracle.morph2d(img, op, r)
        assert got.off.tolist() == z[f"{i}__off"].tolist(), (op, r)
        assert (got.spans.view("u8") == z[f"{i}__spans"].view("u8")).all(), (op, r)


def test_golden_set_is_present():
    assert len(util.golden_3d()) >= 5 and len(util.golden_2d()) >= 2


def test_oracle_from_image_matches_the_reference_fixture(oracle):
    """fromImage of the reference (tests/golden/ingest2d.npz, written by make_golden.py from oracle/_ref)."""
    import os
    import numpy as np
    z = np.load(os.path.join(util.GOLDEN, "ingest2d.npz"))
    co = z["curve_off"]
    curves = [z["curve_pts"][co[k]:co[k + 1]] for k in range(len(co) - 1)]
    got = oracle.from_image(int(z["w"]), int(z["h"]), curves)
    assert got.off.tolist() == z["img_off"].tolist() and int(got.off[-1]) > 100
    assert (got.spans.view("u8") == z["img_spans"].view("u8")).all()


def test_oracle_matches_the_reference_at_config4_size(oracle):
    """Full-size pin: the oracle's dilation of BASELINE config 4 (1060 x 1060 columns, R = 16) against the digest of the
    reference's own result (tests/golden/full_c4_*.npz, written by tests/golden/make_golden_full.py from oracle/_ref);
    also catches a drifting input generator. ~10 s with 8 threads."""
    from voroffset_b200 import synth
    z = util.golden_full("c4_torus_z_n1024_p18_r16")
    vol = synth.torus_z(1024, padding=18)
    util.assert_digest(vol, z, "in", what="synthetic input")
    util.assert_digest(oracle.morph3d(vol, "dilation", float(z["radius"]), "ours"), z, "dilation", what="oracle vs reference digest")


def test_full_size_inputs_are_reproducible():
    from voroffset_b200 import synth
    z = util.golden_full("c5_torus_z_n2048_r32")
    util.assert_digest(synth.torus_z(2048), z, "in", what="synthetic input")
    assert int(z["dilation__nseg"]) == 2790447
