"""Host-side logic: CSR containers, the reference's text format, synthetic generators. CPU only."""
import io

import numpy as np
import pytest

from voroffset_b200 import synth
from voroffset_b200.volume import CompressedVolume, DexelImage, csr_from_lists


def test_grid_follows_reference_constructor():
    # CompressedVolume.cpp:11-23: grid = ceil(extent/spacing) + 2*padding, origin -= padding*spacing
    v = CompressedVolume.from_box((0.0, 0.0, 0.0), (1.0, 2.0, 0.5), 0.1, 3)
    assert (v.nx, v.ny) == (10 + 6, 20 + 6)
    assert v.origin == pytest.approx((-0.3, -0.3, -0.3))
    assert v.zmin == pytest.approx(-3.0)
    assert v.zmax == pytest.approx(-3.0 + 6 + 5.0)


def test_column_indexing_is_x_fastest():
    v = CompressedVolume.from_lists(3, 2, [[], [1, 2], [], [3, 4, 5, 6], [], []])
    assert v.at(1, 0).tolist() == [1, 2]
    assert v.at(0, 1).tolist() == [3, 4, 5, 6]
    assert v.numSegments() == 3


def test_text_round_trip_matches_reference_format():
    v = synth.random_volume(4, 3, kmax=3, seed=1, padding=1)
    s = v.dumps()
    lines = s.strip().split("\n")
    assert len(lines) == 5 + v.nx * v.ny                    # CompressedVolume.cpp:133-152
    w = CompressedVolume.load(io.StringIO(s))
    assert w.bit_equal(v) and w.padding == v.padding and w.spacing == v.spacing


def test_odd_event_count_rejected():
    with pytest.raises(ValueError):
        csr_from_lists([[1.0, 2.0, 3.0]])


def test_generators_are_seeded_and_generic():
    a, b = synth.torus_z(64), synth.torus_z(64)
    assert a.bit_equal(b)
    assert 0.5 < a.numSegments() / (a.nx * a.ny) < 0.7      # k_in ~ 0.60
    c1 = synth.torus_x(256)
    assert (c1.nx, c1.ny) == (67, 256)                       # config C1 grid
    lat = synth.lattice(64, padding=4)
    ends = lat.spans.reshape(-1)
    assert len(np.unique(ends)) > 0.2 * ends.size


def test_dexel_image_validity():
    img = synth.random_image(20, 30, kmax=4, seed=2)
    assert img.isValid()
    bad = DexelImage.from_lists(30, [[5.0, 40.0]])
    assert not bad.isValid()                                  # last event > width
