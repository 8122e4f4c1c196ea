"""Parity of the CUDA path against the oracle, through the C ABI (ctypes -> libvoroffset_b200.so).

Bit-exact for dilation / erosion (both methods) and for every brute_force composite; opening / closing
of 'ours' are compared with identical topology and |dz| <= util.COMPOSITE_TOL (see tests/util.py).
"""
import math

import numpy as np
import pytest

import util
from voroffset_b200 import _lib, image2d, morpho, synth
from voroffset_b200.volume import CompressedVolume, DexelImage

pytestmark = pytest.mark.gpu

OPS = ("dilation", "erosion", "opening", "closing")


# ---- golden vectors produced by the reference itself -----------------------------------------------
@pytest.mark.parametrize("path", util.golden_3d(), ids=lambda p: p.split("vol3d_")[-1][:-4])
@pytest.mark.parametrize("method", ["ours", "brute_force"])
def test_golden_3d(ctx, path, method):
    z, vol, radius, ops = util.load_3d(path)
    op = morpho.make_operator(method, ctx)
    for name in ops:
        got, t1, t2 = morpho.apply_operation(op, name, vol, radius)
        util.assert_same(got, util.expected_3d(z, vol, name, method), name, method, "cuda vs golden")
        assert t1 >= 0 and t2 >= 0


@pytest.mark.parametrize("path", util.golden_2d(), ids=lambda p: p.split("img2d_")[-1][:-4])
def test_golden_2d(ctx, path):
    z, img, ops = util.load_2d(path)
    for i, (name, r) in enumerate(ops):
        d = image2d.DoubleCompressedImage.from_image(img, ctx)
        d.negate() if name == "negate" else getattr(d, name)(r)
        want = DexelImage(img.rows, img.width, z[f"{i}__off"], z[f"{i}__spans"])
        assert d.bit_equal(want), (name, r)


# ---- seeded inputs against the oracle ---------------------------------------------------------------
SEEDED = [
    ("torus_x_n128_p8", lambda: synth.torus_x(128, padding=8), 5.5),
    ("torus_z_n160_p12", lambda: synth.torus_z(160, padding=12), 10.0),
    ("blobs_n96_p9", lambda: synth.blobs(96, padding=9), 8.0),
    ("blobs_many_n80", lambda: synth.blobs(80, count=120, padding=5, rmin=0.02, rmax=0.07, seed=9), 3.3),
    ("random_k8", lambda: synth.random_volume(40, 32, kmax=8, padding=6), 4.3),
    ("random_sparse", lambda: synth.random_volume(33, 47, kmax=3, padding=9, seed=9, fill=0.3), 7.7),
    ("lattice_n96", lambda: synth.lattice(96, padding=8), 5.0),
    ("ragged_31x3", lambda: synth.random_volume(31, 3, kmax=5, padding=0, seed=4), 2.5),
    ("one_row", lambda: synth.random_volume(50, 1, kmax=4, padding=0, seed=5), 6.0),
    ("one_column", lambda: synth.random_volume(1, 1, kmax=3, padding=0, seed=6), 4.0),
    ("radius_lt_1", lambda: synth.blobs(40, padding=2, seed=3), 0.75),
    ("radius_integer", lambda: synth.blobs(48, padding=7, seed=2), 6.0),
    ("all_empty", lambda: synth.random_volume(9, 7, kmax=0, padding=2), 3.0),
]


@pytest.mark.parametrize("name,gen,radius", SEEDED, ids=[c[0] for c in SEEDED])
@pytest.mark.parametrize("method", ["ours", "brute_force"])
def test_seeded_3d(ctx, oracle, name, gen, radius, method):
    vol = gen()
    op = morpho.make_operator(method, ctx)
    for opn in OPS:
        got, _, _ = morpho.apply_operation(op, opn, vol, radius)
        util.assert_same(got, oracle.morph3d(vol, opn, radius, method), opn, method, f"cuda vs oracle [{name}]")


TILE_CASES = [c for c in SEEDED if c[0] not in ("all_empty",)] + [
    ("lattice_n128_R9", lambda: synth.lattice(128, padding=10), 9.0),
    ("blobs_n128_R20", lambda: synth.blobs(128, padding=22, seed=11), 20.5),
    ("random_k8_wide", lambda: synth.random_volume(150, 24, kmax=8, padding=3, seed=12), 6.4),
]


@pytest.mark.parametrize("name,gen,radius", TILE_CASES, ids=[c[0] for c in TILE_CASES])
def test_tile_kernel_forced_3d(ctx, oracle, name, gen, radius):
    """Small grids normally take the one-thread-per-slot pass 1; force the pruned tile kernel (pass1_tile.cuh)
    so that its dominance logic is checked bit for bit against the oracle on every seeded input."""
    vol = gen()
    op = morpho.make_operator("ours", ctx)
    ctx.set_option("pass1", "tile")
    try:
        for opn in OPS:
            got, _, _ = morpho.apply_operation(op, opn, vol, radius)
            util.assert_same(got, oracle.morph3d(vol, opn, radius, "ours"), opn, "ours", f"tile kernel vs oracle [{name}]")
    finally:
        ctx.set_option("pass1", "auto")


def test_complex_classes_in_single_interval_tiles(ctx, oracle):
    """One interval per column at unrelated heights: most classes of the first tile launch are 'complex' (their survivors do
    not union to one interval). The 20-warp variant hands them to the redo launch, the 16-warp variant folds them inline, and
    'auto' moves from the first to the second after the first call: same bits every way."""
    vol = synth.random_volume(150, 40, kmax=1, padding=3, seed=21)
    want = oracle.morph3d(vol, "dilation", 6.4, "ours")
    op = morpho.make_operator("ours", ctx)
    ctx.set_option("pass1", "tile")
    try:
        for mode in ("redo", "inline", "auto", "auto", "redo"):
            ctx.set_option("tile_general", mode)
            got, _, _ = op.dilation(vol, 6.4)
            assert got.bit_equal(want), mode
            d = morpho.DeviceVolume.upload(ctx, vol)
            out, _, _ = op.morph_dev("dilation", d, 6.4)
            assert out.download().bit_equal(want), mode
            out.free(); d.free()
        # the same input leaves one output column in sixty with three intervals and more: pass 2's register-only union (capacity 2) hands
        # them to the redo launch, the list-capable kernel folds them itself, 'auto' goes from one to the other
        assert want.counts().max() >= 3
        for mode in ("registers", "lists", "auto", "auto", "registers"):
            ctx.set_option("pass2_union", mode)
            got, _, _ = op.dilation(vol, 6.4)
            assert got.bit_equal(want), mode
            d = morpho.DeviceVolume.upload(ctx, vol)
            out, _, _ = op.morph_dev("dilation", d, 6.4)
            assert out.download().bit_equal(want), mode
            out.free(); d.free()
    finally:
        ctx.set_option("pass1", "auto")
        ctx.set_option("tile_general", "auto")
        ctx.set_option("pass2_union", "auto")


def _height_field(nx, ny, seed, padding, holes=0.03, thick=(0.4, 30.0), rough=3.0):
    """One interval per column (a few columns empty): rough lower and upper surfaces, so that the eroded intervals
    [a + h, b - h] of neighbouring columns cut each other in every way (U <= L, clamping, pinched-off columns)."""
    rng = np.random.RandomState(seed)
    p = padding
    gx, gy = nx + 2 * p, ny + 2 * p
    yy, xx = np.mgrid[0:ny, 0:nx]
    base = 6.0 + 4.0 * np.sin(xx / 7.0) * np.cos(yy / 5.0) + rng.uniform(0, rough, (ny, nx))
    top = base + rng.uniform(thick[0], thick[1], (ny, nx))
    keep = rng.uniform(size=(ny, nx)) >= holes
    lists = []
    for j in range(gy):
        for i in range(gx):
            inner = p <= i < gx - p and p <= j < gy - p and keep[j - p, i - p]
            lists.append((base[j - p, i - p], top[j - p, i - p]) if inner else ())
    zhi = math.ceil(float(top.max())) + 1.0
    return CompressedVolume.from_lists(gx, gy, lists, origin=(-float(p), -float(p), -float(p)), extent=(float(nx), float(ny), zhi),
                                       spacing=1.0, padding=p)


DUAL_CASES = [
    ("torus_z_n160_p12", lambda: synth.torus_z(160, padding=12), 10.0),
    ("torus_z_n96_R0.75", lambda: synth.torus_z(96, padding=3), 0.75),
    ("torus_z_n128_R6", lambda: synth.torus_z(128, padding=8), 6.0),
    ("height_field_rough", lambda: _height_field(90, 60, 1, 6), 3.7),
    ("height_field_thin", lambda: _height_field(75, 44, 2, 9, holes=0.0, thick=(0.1, 9.0)), 8.0),
    ("height_field_no_padding", lambda: _height_field(64, 64, 3, 0, holes=0.002), 2.2),
    ("height_field_wide", lambda: _height_field(300, 40, 4, 5, holes=0.001, thick=(20.0, 40.0), rough=1.0), 12.5),
]


@pytest.mark.parametrize("name,gen,radius", DUAL_CASES, ids=[c[0] for c in DUAL_CASES])
def test_erosion_dual_form(ctx, oracle, name, gen, radius):
    """Volumes with at most one interval per column take the erosion in DUAL form (vo_lib.cu: erode_dual - the hull of
    the mirrored intervals through the tile kernel, no complement volumes). "erosion" = "dual" makes the library fail
    instead of falling back, so this really is that path: bit for bit against the oracle and against the general
    complement - dilate - complement path; the composites (which chain it) against the oracle."""
    vol = gen()
    op = morpho.make_operator("ours", ctx)
    want = oracle.morph3d(vol, "erosion", radius, "ours")
    ctx.set_option("pass1", "tile")
    try:
        ctx.set_option("erosion", "dual")
        got, t1, t2 = morpho.apply_operation(op, "erosion", vol, radius)
        util.assert_same(got, want, "erosion", "ours", f"dual erosion vs oracle [{name}]")
        assert t1 >= 0 and t2 >= 0
        ctx.set_option("erosion", "general")
        gen_, _, _ = morpho.apply_operation(op, "erosion", vol, radius)
        assert got.bit_equal(gen_), f"dual and general erosion differ [{name}]"
        ctx.set_option("erosion", "auto")
        for opn in ("opening", "closing"):
            r, _, _ = morpho.apply_operation(op, opn, vol, radius)
            util.assert_same(r, oracle.morph3d(vol, opn, radius, "ours"), opn, "ours", f"{opn} with dual erosion [{name}]")
    finally:
        ctx.set_option("erosion", "auto")
        ctx.set_option("pass1", "auto")


def test_erosion_dual_form_declines_what_does_not_qualify(ctx, oracle):
    """A second interval in ONE column, an interval starting exactly at zmin - 1 ... the dual form must notice while it
    runs (k_thresh) and the call must fall back to the general path: same result as the oracle; with "erosion" = "dual"
    the call fails instead."""
    base = _height_field(80, 50, 7, 6, holes=0.2)
    # (a) one column with two intervals (the total still fits "at most one per column on average")
    lists = [tuple(base.at(x, y)) for y in range(base.ny) for x in range(base.nx)]
    c = 20 + base.nx * 25
    a0, b0 = lists[c] if lists[c] else (5.0, 9.0)
    lists[c] = (a0, a0 + 0.25 * (b0 - a0), a0 + 0.5 * (b0 - a0), b0)
    meta = dict(origin=base.origin, extent=base.extent, spacing=base.spacing, padding=base.padding)
    two = CompressedVolume.from_lists(base.nx, base.ny, lists, **meta)
    # (b) an interval that starts exactly at the lower bound of the complement (zmin - 1): negate_ray erases that event
    sp = base.spans.copy()
    sp[3, 0] = base.zmin - 1.0
    low = CompressedVolume(base.nx, base.ny, base.off, sp, **meta)
    op = morpho.make_operator("ours", ctx)
    ctx.set_option("pass1", "tile")
    try:
        for what, vol in (("two intervals in a column", two), ("interval at the bound", low)):
            ctx.set_option("erosion", "auto")
            got, _, _ = morpho.apply_operation(op, "erosion", vol, 4.4)
            util.assert_same(got, oracle.morph3d(vol, "erosion", 4.4, "ours"), "erosion", "ours", what)
            ctx.set_option("erosion", "dual")
            with pytest.raises(_lib.VoroffsetError):
                morpho.apply_operation(op, "erosion", vol, 4.4)
    finally:
        ctx.set_option("erosion", "auto")
        ctx.set_option("pass1", "auto")


def test_row_window_call_matches_the_rows_of_the_whole_grid(ctx):
    """vo_morph3d_rows: a y-slab handed over with its ghost rows (a window of the host CSR, offsets not starting at 0)
    must give, bit for bit, the rows the call on the whole grid gives - through the banded pipeline (config-5 size) and
    through the plain path (small grids, the other operations)."""
    op = morpho.make_operator("ours", ctx)
    big, R = synth.torus_z(2048), 32.0
    J = int(R)
    whole, _, _ = op.dilation(big, R)
    for y0, y1 in ((0, 700), (700, 1500), (1500, 2048), (1017, 1091)):
        e0, e1 = max(0, y0 - J), min(big.ny, y1 + J)
        got, t1, t2 = op.morph_rows("dilation", big, R, e0, e1, y0 - e0, y1 - e0)
        want = _rows_of(whole, y0, y1)
        assert got.bit_equal(want), f"window [{y0}, {y1}) of the config-5 dilation differs from the whole-grid rows"
    small, r = synth.torus_z(160, padding=12), 5.5
    j = int(r)
    for opn, ghost in (("dilation", j), ("erosion", j), ("opening", 2 * j), ("closing", 2 * j)):
        whole, _, _ = morpho.apply_operation(op, opn, small, r)
        for y0, y1 in ((0, 60), (60, 130), (130, small.ny)):
            e0, e1 = max(0, y0 - ghost), min(small.ny, y1 + ghost)
            got, _, _ = op.morph_rows(opn, small, r, e0, e1, y0 - e0, y1 - e0)
            assert got.bit_equal(_rows_of(whole, y0, y1)), f"{opn}: window [{y0}, {y1}) differs"


def _rows_of(vol, y0, y1):
    c0, c1 = y0 * vol.nx, y1 * vol.nx
    off = vol.off[c0:c1 + 1].astype(np.int64)
    return vol.like(vol.nx, y1 - y0, (off - off[0]).astype(np.uint32), vol.spans[off[0]:off[-1]])


def test_erosion_with_data_touching_the_z_bounds(ctx, oracle):
    """Erosion of 'ours' prunes with the clip range of negateInv (pass1_tile.cuh: saturated endpoints). A tight
    bounding box makes the lowest interval start EXACTLY at zmin and the highest end exactly at zmax
    (negate_ray's == tests, MorphologyOperators.cpp:241,250): the complement then has columns without a lower /
    upper part next to columns with one. Forced tile kernel, all four operations, against the oracle."""
    for seed, kmax, R in ((21, 3, 4.5), (22, 6, 7.2), (23, 1, 9.0)):
        v = synth.random_volume(70, 40, kmax=kmax, padding=0, seed=seed, fill=0.8)
        lo, hi = float(v.spans[:, 0].min()), float(v.spans[:, 1].max())
        # snap a fifth of the columns to the bounds so that many of them touch
        sp = v.spans.copy()
        first = v.off[:-1][np.diff(v.off.astype(np.int64)) > 0].astype(np.int64)
        last = v.off[1:][np.diff(v.off.astype(np.int64)) > 0].astype(np.int64) - 1
        meta = dict(origin=(0.0, 0.0, lo), extent=(float(v.nx), float(v.ny), hi - lo), spacing=1.0, padding=0)
        zmax = CompressedVolume(v.nx, v.ny, v.off, sp, **meta).zmax      # lo + (hi - lo), as the reference computes it
        sp[:, 1] = np.minimum(sp[:, 1], zmax)
        sp[first[::5], 0] = lo
        sp[last[::7], 1] = zmax
        vol = CompressedVolume(v.nx, v.ny, v.off, sp, **meta)
        assert vol.zmin == lo and vol.zmax == zmax and sp[:, 0].min() == lo and sp[:, 1].max() == zmax
        op = morpho.make_operator("ours", ctx)
        for mode in ("auto", "tile"):
            ctx.set_option("pass1", mode)
            try:
                for opn in OPS:
                    got, _, _ = morpho.apply_operation(op, opn, vol, R)
                    util.assert_same(got, oracle.morph3d(vol, opn, R, "ours"), opn, "ours", f"z-bounds seed {seed} [{mode}]")
            finally:
                ctx.set_option("pass1", "auto")


def test_config5_all_operations_ours_vs_brute_force_2048(ctx):
    """The north-star size (n = 2048, R = 32) for all four operations: the pruned two-pass method against the
    brute-force sphere union (an independent kernel without any pruning), both on the GPU. Identical topology;
    endpoints within 1e-11 dexel (the two methods round their caps differently)."""
    vol = synth.torus_z(2048, padding=34)
    ours, brute = morpho.make_operator("ours", ctx), morpho.make_operator("brute_force", ctx)
    d = morpho.DeviceVolume.upload(ctx, vol)
    for opn in OPS:
        a, _, _ = ours.morph_dev(opn, d, 32.0)
        b, _, _ = brute.morph_dev(opn, d, 32.0)
        ha, hb = a.download(), b.download()
        a.free(); b.free()
        assert ha.same_topology(hb), opn
        assert np.abs(ha.spans - hb.spans).max() <= 1e-11, opn
    # opening is anti-extensive, closing extensive (single-interval columns)
    ci = vol.counts()
    o, _, _ = ours.morph_dev("opening", d, 32.0)
    c, _, _ = ours.morph_dev("closing", d, 32.0)
    co, cc = o.download().counts(), c.download().counts()
    o.free(); c.free()
    assert np.all(ci[co > 0] >= 1) and np.all(cc[ci > 0] >= 1)


def test_many_layers_use_the_redo_path(ctx, oracle):
    # 40 thin layers per column: the running union outgrows the fast capacity (16) and is redone
    rng = np.random.RandomState(0)
    lists = []
    for _ in range(12 * 10):
        z = np.cumsum(rng.uniform(0.4, 0.9, size=80)) + rng.uniform(0, 0.3)
        lists.append(z.tolist())
    vol = CompressedVolume.from_lists(12, 10, lists, origin=(0, 0, -5.0), extent=(12.0, 10.0, 70.0), spacing=1.0, padding=0)
    for method in ("ours", "brute_force"):
        op = morpho.make_operator(method, ctx)
        got, _, _ = op.dilation(vol, 0.3)
        assert got.bit_equal(oracle.morph3d(vol, "dilation", 0.3, method))
        assert got.counts().max() > 16
    # with neighbours in reach (R = 1.3) the transient lists are long although the result is short
    for method in ("ours", "brute_force"):
        got, _, _ = morpho.make_operator(method, ctx).dilation(vol, 1.3)
        assert got.bit_equal(oracle.morph3d(vol, "dilation", 1.3, method))


def test_lists_beyond_the_redo_capacity_take_the_last_resort_launch(ctx, oracle):
    # columns of 700 and 1500 thin layers: the running union outgrows CAP_BIG = 512 as well; the primitive is repeated with the
    # redo launches in their CAP_HUGE form (lists in global memory) instead of failing with VO_ERR_OVERFLOW (round 1 did)
    rng = np.random.RandomState(5)
    lists = []
    for c in range(6 * 5):
        n = 1500 if c == 7 else 700 if c == 20 else 40
        z = np.cumsum(rng.uniform(0.7, 0.9, size=2 * n)) + rng.uniform(0, 0.3)
        lists.append(z.tolist())
    dense = CompressedVolume.from_lists(6, 5, lists, origin=(0, 0, -5.0), extent=(6.0, 5.0, 2700.0), spacing=1.0, padding=0)
    # ... and a mostly empty grid with one such column, which the tile kernel accepts (its sorted-list union hands the slot to
    # the redo launch of the one-thread-per-slot kernel)
    lists = []
    for c in range(64 * 4):
        n = 700 if c == 64 * 2 + 10 else 3 if c % 7 == 0 else 0
        lists.append((np.cumsum(rng.uniform(0.7, 0.9, size=2 * n)) + rng.uniform(0, 0.3)).tolist())
    sparse = CompressedVolume.from_lists(64, 4, lists, origin=(0, 0, -5.0), extent=(64.0, 4.0, 1300.0), spacing=1.0, padding=0)
    for vol, mode in ((dense, "auto"), (sparse, "auto"), (sparse, "tile")):
        ctx.set_option("pass1", mode)
        try:
            for method in ("ours", "brute_force"):
                for radius in (0.3, 1.3):
                    got, _, _ = morpho.make_operator(method, ctx).dilation(vol, radius)
                    assert got.bit_equal(oracle.morph3d(vol, "dilation", radius, method)), (mode, method, radius)
                    if radius < 1:
                        assert got.counts().max() > 512
        finally:
            ctx.set_option("pass1", "auto")
    # the calls that follow are back on the normal launches
    small = synth.blobs(32, padding=2, seed=4)
    got, _, _ = morpho.make_operator("ours", ctx).dilation(small, 3.0)
    assert got.bit_equal(oracle.morph3d(small, "dilation", 3.0, "ours"))
    # 2D rows
    rows = []
    for i in range(8):
        n = 700 if i == 3 else 20
        rows.append((np.cumsum(rng.uniform(2.0, 3.0, size=2 * n)) + 1.0).tolist())
    img = DexelImage.from_lists(5000, rows)
    d = image2d.DoubleCompressedImage.from_image(img, ctx)
    d.dilate(0.3 / 8)
    want = oracle.morph2d(img, "dilate", 0.3 / 8)
    assert d.bit_equal(want) and int(np.diff(want.off.astype(np.int64)).max()) > 512


def test_zero_radius_is_identity(ctx):
    vol = synth.blobs(32, padding=2, seed=4)
    got, _, _ = morpho.make_operator("ours", ctx).dilation(vol, 0.0)
    assert got.bit_equal(vol)


def test_single_point_column_dilates_to_a_ball(ctx):
    # analytic check (SURVEY.md 8(c) pin 4): half-length at offset d is sqrt(R^2 - d^2)
    R, n = 6.5, 21
    lists = [[] for _ in range(n * n)]
    lists[10 + n * 10] = [3.0, 3.0]
    vol = CompressedVolume.from_lists(n, n, lists, origin=(0, 0, -10.0), extent=(float(n), float(n), 30.0), spacing=1.0)
    got, _, _ = morpho.make_operator("ours", ctx).dilation(vol, R)
    for y in range(n):
        for x in range(n):
            d2 = (x - 10) ** 2 + (y - 10) ** 2
            col = got.at(x, y)
            if d2 <= R * R and abs(x - 10) <= np.floor(np.sqrt(R * R - (y - 10) ** 2)):
                assert col.size == 2
                assert col[1] - 3.0 == pytest.approx(np.sqrt(R * R - d2), abs=1e-12)
                assert 3.0 - col[0] == pytest.approx(np.sqrt(R * R - d2), abs=1e-12)
            else:
                assert col.size == 0


# ---- error behaviour ----------------------------------------------------------------------------------
def test_errors_are_reported_not_swallowed(ctx):
    vol = synth.blobs(16, padding=1, seed=1)
    op = morpho.make_operator("ours", ctx)
    with pytest.raises(_lib.VoroffsetError):
        op.dilation(vol, -1.0)
    with pytest.raises(_lib.VoroffsetError):
        op.dilation(vol, float("nan"))
    with pytest.raises(ValueError):
        morpho.make_operator("fancy", ctx)
    with pytest.raises(ValueError):
        morpho.apply_operation(op, "smoothing", vol, 1.0)
    import ctypes as C
    poff, pspans, n = _lib._u32p(), _lib._f64p(), C.c_uint64()
    rc = ctx.lib.vo_morph3d(ctx.handle, 9, 0, vol.nx, vol.ny, 0.0, 1.0, _lib.ptr(vol.off), _lib.ptr(vol.spans), 2.0,
                            C.byref(poff), C.byref(pspans), C.byref(n), None, None)
    assert rc == 1 and ctx.lib.vo_last_error(ctx.handle) == b"Operation"
    rc = ctx.lib.vo_morph3d(ctx.handle, 0, 7, vol.nx, vol.ny, 0.0, 1.0, _lib.ptr(vol.off), _lib.ptr(vol.spans), 2.0,
                            C.byref(poff), C.byref(pspans), C.byref(n), None, None)
    assert rc == 1 and ctx.lib.vo_last_error(ctx.handle) == b"Invalid method"


# ---- resident API ---------------------------------------------------------------------------------------
def test_resident_round_trip_rows_and_concat(ctx):
    vol = synth.blobs(48, padding=3, seed=6)
    d = morpho.DeviceVolume.upload(ctx, vol)
    assert d.download().bit_equal(vol)
    a, b, c = d.rows(0, 10), d.rows(10, 31), d.rows(31, vol.ny)
    assert morpho.concat_rows(ctx, [a, b, c]).download().bit_equal(vol)
    part = b.download()
    assert part.ny == 21 and part.at(5, 0).tolist() == vol.at(5, 10).tolist()


def test_split_passes_equal_fused_call(ctx):
    import ctypes as C
    vol = synth.blobs(64, padding=8, seed=7)
    R = 7.3
    op = morpho.make_operator("ours", ctx)
    fused, _, _ = op.dilation(vol, R)
    d = morpho.DeviceVolume.upload(ctx, vol)
    mid = C.c_void_p()
    ctx.check(ctx.lib.vo_pass1_dev(ctx.handle, d.handle, R, C.byref(mid), None))
    pieces = []
    for y0, y1 in [(0, 17), (17, 18), (18, vol.ny)]:
        h = C.c_void_p()
        ctx.check(ctx.lib.vo_pass2_dev(ctx.handle, mid, y0, y1, C.byref(h), None))
        pieces.append(morpho.DeviceVolume(ctx, h, vol))
    ctx.lib.vo_dmid_free(ctx.handle, mid)
    assert morpho.concat_rows(ctx, pieces).download().bit_equal(fused)


def test_resident_operator_matches_host_call(ctx):
    vol = synth.torus_x(96, padding=6)
    op = morpho.make_operator("ours", ctx)
    d = morpho.DeviceVolume.upload(ctx, vol)
    for opn in OPS:
        r, _, _ = op.morph_dev(opn, d, 4.4)
        h, _, _ = morpho.apply_operation(op, opn, vol, 4.4)
        assert r.download().bit_equal(h)


def test_overlapped_slab_api_matches_single_call(ctx):
    """vo_slab_begin / vo_slab_finish (the multi-GPU step, voroffset_b200/slab.py) on ONE device: the volume is cut
    into three y-slabs, every slab is dilated with the halos its neighbours would send, the rows must be the
    single-call result bit for bit."""
    import ctypes as C
    import torch
    vol = synth.torus_z(640, padding=0)
    R, J = 20.5, 20
    op = morpho.make_operator("ours", ctx)
    whole, _, _ = op.dilation(vol, R)
    d = morpho.DeviceVolume.upload(ctx, vol)
    dev = torch.device("cuda", ctx.device)
    bounds = [(0, 200), (200, 430), (430, vol.ny)]
    pieces = []
    for y0, y1 in bounds:
        own = d.rows(y0, y1)
        halos = []
        for (h0, h1), present in (((y0 - J, y0), y0 > 0), ((y1, y1 + J), y1 < vol.ny)):
            if not present:
                halos.append(None)
                continue
            off = torch.empty(J * vol.nx + 1, dtype=torch.int32, device=dev)
            sp = torch.empty(2 * 200000, dtype=torch.float64, device=dev)
            n = C.c_uint64(0)
            ctx.check(ctx.lib.vo_dvol_rows_to(ctx.handle, d.handle, h0, h1, off.data_ptr(), sp.data_ptr(), 200000, C.byref(n)))
            halos.append((off, sp, int(n.value)))
        slab = C.c_void_p()
        rc = ctx.lib.vo_slab_begin(ctx.handle, own.handle, R, int(halos[0] is not None), int(halos[1] is not None),
                                   halos[0][2] + 100 if halos[0] else 0, halos[1][2] + 7 if halos[1] else 0,
                                   None, None, 0, None, None, 0, None, C.byref(slab))
        assert rc == 0, ctx.lib.vo_last_error(ctx.handle)
        args = []
        for h in halos:
            args += [None, None, 0] if h is None else [h[0].data_ptr(), h[1].data_ptr(), h[2]]
        out = C.c_void_p()
        ctx.check(ctx.lib.vo_slab_finish(ctx.handle, slab, *args, C.byref(out), None, None))
        pieces.append(morpho.DeviceVolume(ctx, out, vol))
        own.free()
    assert morpho.concat_rows(ctx, pieces).download().bit_equal(whole)
    # a small grid is declined (the caller then takes the plain path)
    small = morpho.DeviceVolume.upload(ctx, synth.blobs(24, padding=2, seed=3))
    slab = C.c_void_p()
    assert ctx.lib.vo_slab_begin(ctx.handle, small.handle, 3.0, 0, 1, 0, 1000, None, None, 0, None, None, 0, None, C.byref(slab)) == 1
    # the device-side halo packing (what a rank sends): offsets from 0, [count, overflow], spans when they fit
    J2, nx2 = 20, vol.nx
    off = torch.zeros(J2 * nx2 + 3, dtype=torch.int32, device=dev)
    sp = torch.zeros(2 * 50000, dtype=torch.float64, device=dev)
    rc = ctx.lib.vo_slab_begin(ctx.handle, d.handle, R, 0, 1, 0, 1000, None, None, 0, off.data_ptr(), sp.data_ptr(), 50000,
                               torch.cuda.current_stream(dev).cuda_stream, C.byref(slab))
    assert rc == 0
    ctx.lib.vo_slab_abort(ctx.handle, slab)
    want = vol.off[(vol.ny - J2) * nx2:].astype(np.int64)
    got = off.cpu().numpy()
    assert np.array_equal(got[:J2 * nx2 + 1], want - want[0]) and got[J2 * nx2 + 1] == want[-1] - want[0] and got[J2 * nx2 + 2] == 0
    assert np.array_equal(sp.cpu().numpy()[:2 * (want[-1] - want[0])].reshape(-1, 2), vol.spans[want[0]:want[-1]])


# ---- xor --------------------------------------------------------------------------------------------------
def test_xor(ctx, oracle):
    a, b = synth.blobs(64, padding=4, seed=1), synth.blobs(64, padding=4, seed=2)
    op = morpho.make_operator("ours", ctx)
    vol, x = op.calculateXor(a, b)
    wvol, wx = oracle.xor3d(a, b)
    assert x.bit_equal(wx)
    assert vol == pytest.approx(wvol, rel=1e-12)
    v0, x0 = op.calculateXor(a, a)
    assert v0 == 0 and x0.numSegments() == 0
    # the reference's own use: ours vs brute_force have zero xor volume
    d1, _, _ = op.dilation(a, 5.5)
    d2, _, _ = morpho.make_operator("brute_force", ctx).dilation(a, 5.5)
    v, _ = op.calculateXor(d1, d2)
    assert v == 0


# ---- 2D ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [1, 2])
def test_seeded_2d(ctx, oracle, seed):
    img = synth.random_image(200, 300, kmax=5, seed=seed)
    for op, r in [("dilate", 8.0 / 200), ("dilate", 5.5 / 200), ("erode", 4.0), ("erode", 2.5), ("open", 3.0 / 200),
                  ("close", 3.0 / 200), ("negate", 0.0)]:
        d = image2d.DoubleCompressedImage.from_image(img, ctx)
        d.negate() if op == "negate" else getattr(d, op)(r)
        assert d.bit_equal(oracle.morph2d(img, op, r)), (op, r)


def test_config2_2048_rows(ctx, oracle):
    # BASELINE config 2: 2048^2 grid, effective R = 16 dilation (call dilate(16/2048)), erode R = 4
    img = synth.star_image(2048, 2048, 64)
    for op, r in [("dilate", 16.0 / 2048), ("erode", 4.0)]:
        d = image2d.DoubleCompressedImage.from_image(img, ctx)
        getattr(d, op)(r)
        assert d.bit_equal(oracle.morph2d(img, op, r)), op
    empty = image2d.DoubleCompressedImage.from_image(DexelImage.from_lists(64, [[] for _ in range(5)]), ctx)
    empty.erode(2.0)
    assert empty.numSegments() == 0
    empty.negate()
    assert empty.numSegments() == 5


# ---- BASELINE sizes ---------------------------------------------------------------------------------------
def test_config3_lattice_ours_vs_brute_force(ctx, oracle):
    # config 3: -n 512 -p 10 -r 5, dilation, 'ours' against the 'brute_force' oracle method, both on the GPU,
    # plus the CPU oracle on the same input (283 k columns)
    vol = synth.lattice(512, padding=10)
    assert (vol.nx, vol.ny) == (532, 532)
    a, _, _ = morpho.make_operator("ours", ctx).dilation(vol, 5.0)
    b, _, _ = morpho.make_operator("brute_force", ctx).dilation(vol, 5.0)
    assert a.same_topology(b)
    assert np.abs(a.spans - b.spans).max() <= 1e-12
    assert a.bit_equal(oracle.morph3d(vol, "dilation", 5.0, "ours"))


def test_config4_opening_closing_1024(ctx, oracle):
    # config 4: -n 1024 -r 16 opening and closing. Full-size oracle run on the CPU is ~10 s with threads.
    vol = synth.torus_z(1024, padding=18)
    op = morpho.make_operator("ours", ctx)
    for opn in ("opening", "closing"):
        got, _, _ = morpho.apply_operation(op, opn, vol, 16.0)
        util.assert_same(got, oracle.morph3d(vol, opn, 16.0, "ours"), opn, "ours", "config 4")
    # size-independent properties: closing is extensive, opening anti-extensive (single-interval columns)
    cl, _, _ = op.closing(vol, 16.0)
    opn_, _, _ = op.opening(vol, 16.0)
    ci, cc, co = vol.counts(), cl.counts(), opn_.counts()
    assert np.all(cc[ci > 0] >= 1) and np.all(ci[co > 0] >= 1)


def test_config5_dilation_2048_properties(ctx, oracle):
    # config 5 at full size (4.19 M columns, R = 32): the oracle checks a band of rows bit for bit, and
    # size-independent properties hold over the whole result.
    vol = synth.torus_z(2048)
    R = 32.0
    op = morpho.make_operator("ours", ctx)
    got, t1, t2 = op.dilation(vol, R)
    assert (got.nx, got.ny) == (2048, 2048)
    # (1) band parity: rows [y0-R, y1+R) of the input determine rows [y0, y1) of the output exactly
    y0, y1, J = 300, 316, 32
    band = CompressedVolume(vol.nx, (y1 + J) - (y0 - J), *_rows(vol, y0 - J, y1 + J))
    want = oracle.morph3d(band, "dilation", R, "ours")
    g_off, g_sp = _rows(got, y0, y1)
    w_off, w_sp = _rows(want, J, J + (y1 - y0))
    assert np.array_equal(g_off, w_off) and np.array_equal(g_sp.view(np.uint64), w_sp.view(np.uint64))
    # (2) extensive: every input interval is inside an output interval grown by exactly R at dx=dy=0
    ci, co = vol.counts(), got.counts()
    assert np.all(co[ci > 0] == 1)
    cols = np.nonzero(ci > 0)[0]
    lo_in, hi_in = vol.spans[vol.off[cols].astype(np.int64)].T
    lo_out, hi_out = got.spans[got.off[cols].astype(np.int64)].T
    assert np.all(lo_out <= lo_in - R) and np.all(hi_out >= hi_in + R)
    # (3) a bigger radius gives a superset; (4) determinism (checksum of a second run)
    again, _, _ = op.dilation(vol, R)
    assert util.checksum(again) == util.checksum(got)
    small, _, _ = op.dilation(vol, 20.0)
    cs = small.counts()
    assert np.all(co[cs > 0] >= 1)


def test_config5_full_grid_matches_the_reference(ctx):
    """BASELINE config 5 at full size against the reference's OWN result for all 4.19 M columns (tests/golden/
    full_c5_*.npz: per-column interval counts + per-row checksums of the endpoint bits, generated by
    tests/golden/make_golden_full.py from oracle/_ref, the reference's sources compiled in place)."""
    z = util.golden_full("c5_torus_z_n2048_r32")
    vol = synth.torus_z(2048)
    util.assert_digest(vol, z, "in", what="synthetic input (generator drift?)")
    op = morpho.make_operator("ours", ctx)
    got, _, _ = op.dilation(vol, float(z["radius"]))                       # host-buffer call (banded pipeline)
    util.assert_digest(got, z, "dilation", what="config 5, vo_morph3d")
    d = morpho.DeviceVolume.upload(ctx, vol)
    out, _, _ = op.morph_dev("dilation", d, float(z["radius"]))            # resident path (what bench.py times)
    util.assert_digest(out.download(), z, "dilation", what="config 5, vo_morph3d_dev")
    out.free(); d.free()


def test_config4_full_grid_matches_the_reference(ctx):
    """BASELINE config 4 (-n 1024 -p 18 -r 16): all four operations of 'ours' against the reference's own full-size
    result. Primitives bit for bit, composites with exact topology and endpoint sums within util.COMPOSITE_TOL."""
    z = util.golden_full("c4_torus_z_n1024_p18_r16")
    vol = synth.torus_z(1024, padding=18)
    util.assert_digest(vol, z, "in", what="synthetic input (generator drift?)")
    op = morpho.make_operator("ours", ctx)
    for name in ("dilation", "erosion", "opening", "closing"):
        got, _, _ = morpho.apply_operation(op, name, vol, float(z["radius"]))
        util.assert_digest(got, z, name, exact=name in ("dilation", "erosion"), what="config 4")


LARGE_R = [(64.5, 44, 36), (100.0, 40, 33), (255.0, 37, 30), (256.0, 37, 30), (300.25, 30, 28)]


@pytest.mark.parametrize("radius,nx,ny", LARGE_R, ids=[f"R{r:g}" for r, _, _ in LARGE_R])
def test_large_radii(ctx, oracle, radius, nx, ny):
    """floor(R) >= 64 takes the one-thread-per-slot pass 1 and the general pass 2; floor(R) + 1 >= 256 no longer fits the
    8-bit class-window bound (kernels.cuh: FLAG_ALL). Both methods, dilation and erosion bit for bit against the oracle."""
    vol = synth.random_volume(nx, ny, kmax=3, padding=2, seed=31, zrange=900.0)
    for method in ("ours", "brute_force"):
        op = morpho.make_operator(method, ctx)
        for opn in ("dilation", "erosion"):
            got, _, _ = morpho.apply_operation(op, opn, vol, radius)
            util.assert_same(got, oracle.morph3d(vol, opn, radius, method), opn, method, f"R = {radius}")
    assert got.numSegments() >= 0


def test_unsorted_offsets_are_rejected(ctx):
    """include/voroffset_b200.h promises VO_ERR_ARG for offsets that are not non-decreasing: the plain path checks on the
    host (vo_lib.cu: upload), the banded host-buffer call on the device (k_thresh, ThreshArgs::bad) - and the context
    stays usable."""
    op = morpho.make_operator("ours", ctx)
    small = synth.random_volume(40, 32, kmax=4, padding=2, seed=3)
    big = synth.torus_z(1536)
    for vol, c in ((small, small.nx * 5 + 7), (big, big.nx * 700 + 801), (big, big.nx * 192 + 3)):
        good, _, _ = op.dilation(vol, 20.0 if vol is big else 3.0)
        off = vol.off.copy()
        nz = np.nonzero(np.diff(off.astype(np.int64)) > 0)[0]
        k = int(nz[np.searchsorted(nz, c)])
        off[k + 1] = off[k] - 1 if off[k] > 0 else off[k + 2] + 5            # decreasing inside a band / a row
        bad = CompressedVolume(vol.nx, vol.ny, off, vol.spans)
        with pytest.raises(_lib.VoroffsetError) as e:
            op.dilation(bad, 20.0 if vol is big else 3.0)
        assert e.value.code == 1 and "non-decreasing" in str(e.value)
        again, _, _ = op.dilation(vol, 20.0 if vol is big else 3.0)
        assert again.bit_equal(good)


def test_2d_dilate_with_effective_radius_beyond_4096(ctx, oracle):
    """dilate(r) sweeps with R = r * rows (DoubleCompressedImage.cpp:685-686): r = 2.5 on 2048 rows is R = 5120."""
    img = synth.star_image(2048, 256, 12)
    d = image2d.DoubleCompressedImage.from_image(img, ctx)
    d.dilate(2.5)
    assert d.bit_equal(oracle.morph2d(img, "dilate", 2.5))
    small = synth.random_image(60, 90, kmax=4, seed=8)
    for op, r in (("dilate", 80.0), ("erode", 70.5), ("erode", 5000.0)):
        d = image2d.DoubleCompressedImage.from_image(small, ctx)
        getattr(d, op)(r)
        assert d.bit_equal(oracle.morph2d(small, op, r)), (op, r)


def _rows(vol, y0, y1):
    c0, c1 = y0 * vol.nx, y1 * vol.nx
    off = vol.off[c0:c1 + 1].astype(np.int64)
    return (off - off[0]).astype(np.uint32), vol.spans[off[0]:off[-1]]


# ---- the reference-side binding (integration/) behind the reference's own base-class pointer ------------
def test_dropin_subclasses_match_reference_operators():
    """oracle/_ref/dropin_check drives vor3d::VoronoiMorphoVorPower / BruteForce and the B200-backed
    subclasses of vor3d::VoronoiMorpho through the same std::unique_ptr<VoronoiMorpho>, on the reference's
    own CompressedVolume, with the call structure of offset3d.cpp:116-136 (built where /root/reference exists)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "dropin_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_check not built (needs /root/reference at build time)")
    res = subprocess.run([exe, "56", "5.5"], capture_output=True, text=True, timeout=600)
    print(res.stdout)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.count("OK") == 9


def test_pipelined_host_path_equals_plain_path(ctx):
    """Large host-buffer dilations overlap upload / passes / download band by band (vo_lib.cu:
    dilate_ours_pipelined); the result must be the plain path's, bit for bit."""
    vol = synth.torus_z(1792, padding=0)
    op = morpho.make_operator("ours", ctx)
    ctx.set_option("pipeline", "off")
    plain, _, _ = op.dilation(vol, 31.5)
    ctx.set_option("pipeline", "on")
    piped, t1, _ = op.dilation(vol, 31.5)
    assert piped.bit_equal(plain) and t1 > 0
    again, _, _ = op.dilation(vol, 31.5)          # second call reuses the size hints
    assert again.bit_equal(plain)


def _points_volume(nx, ny, points):
    """Volume with one short interval in each of the given (x, y) columns."""
    cnt = np.zeros(nx * ny, dtype=np.int64)
    pts = sorted(set((y * nx + x) for x, y in points))
    cnt[pts] = 1
    off = np.concatenate(([0], np.cumsum(cnt))).astype(np.uint32)
    spans = np.array([[100.0 + 0.37 * (c % 17), 101.5 + 0.37 * (c % 17)] for c in pts]).reshape(-1, 2)
    return CompressedVolume(nx, ny, off, spans)


def test_pipelined_host_path_edge_cases(ctx):
    """The banded host-buffer path on inputs that stress its bookkeeping: an empty volume, isolated points (the result
    has ~3000 times the intervals of the input, so it outgrows the result buffers sized from the input and the call
    must fall back, then succeed banded with the learnt size), and the same context going back to a small result."""
    nx, ny, R = 1536, 1056, 32.0                     # large enough for the banded path (vo_lib.cu: dilate_ours_pipelined)
    op = morpho.make_operator("ours", ctx)
    rng = np.random.RandomState(5)
    pts = [(int(rng.randint(40, nx - 40)), int(rng.randint(40, ny - 40))) for _ in range(120)]
    pts += [(0, 0), (nx - 1, ny - 1), (nx - 1, 0), (0, ny - 1), (700, 131), (700, 132), (701, 527), (701, 528)]   # corners, band seams
    for name, vol in (("empty", _points_volume(nx, ny, [])), ("points", _points_volume(nx, ny, pts)),
                      ("one point", _points_volume(nx, ny, [(5, 5)]))):
        ctx.set_option("pipeline", "off")
        plain, _, _ = op.dilation(vol, R)
        ctx.set_option("pipeline", "on")
        for attempt in range(3):
            piped, _, _ = op.dilation(vol, R)
            assert piped.bit_equal(plain), f"{name}, call {attempt}"
    assert plain.numSegments() > 0


def _with_extra_intervals(vol, rows, cols, extra, gap):
    """`vol` (at most one interval per column) with `extra` further intervals stacked above the first one, `gap` apart,
    in the columns of the given row / column ranges that hold one."""
    nx, ny = vol.nx, vol.ny
    cnt = np.diff(vol.off.astype(np.int64))
    assert cnt.max() <= 1
    region = np.zeros((ny, nx), dtype=bool)
    region[rows[0]:rows[1], cols[0]:cols[1]] = True
    sel = (cnt == 1) & region.reshape(-1)
    new_cnt = cnt + extra * sel
    off = np.concatenate(([0], np.cumsum(new_cnt))).astype(np.uint32)
    spans = np.empty((int(off[-1]), 2))
    has = cnt >= 1
    spans[off[:-1][has]] = vol.spans.reshape(-1, 2)[vol.off[:-1][has]]
    base = vol.spans.reshape(-1, 2)[vol.off[:-1][sel]]
    for k in range(1, extra + 1):
        spans[off[:-1][sel] + k] = np.stack((base[:, 1] + k * gap, base[:, 1] + k * gap + 0.4 * gap), axis=1)
    return CompressedVolume(nx, ny, off, spans), int(sel.sum())


def test_pipelined_host_path_with_launches_left_out(ctx):
    """The bands of the host-buffer call leave out the launches that are idle for height-field-like input (the tile
    kernel's list launches, the redo launches of both passes; vo_ctx::pipe_lean). Input that needs them - columns with
    several intervals, a running union beyond the fast capacity - must be noticed, done on the plain path, and the
    following calls (which make those launches) must give the same bits."""
    base = synth.torus_z(1536, padding=0)
    op = morpho.make_operator("ours", ctx)
    for name, (vol, n) in (("two layers", _with_extra_intervals(base, (500, 620), (0, base.nx), 1, 90.0)),
                           ("forty layers", _with_extra_intervals(base, (300, 303), (200, 204), 40, 70.0))):
        assert n > 0, name
        ctx.set_option("pipe_lean", "on")            # (forgets what an earlier call learnt)
        ctx.set_option("pipeline", "off")
        plain, _, _ = op.dilation(vol, 31.5)
        ctx.set_option("pipeline", "on")
        for attempt in range(3):
            piped, _, _ = op.dilation(vol, 31.5)
            assert piped.bit_equal(plain), f"{name}, call {attempt}"
    ctx.set_option("pipe_lean", "on")


def test_pipelined_host_path_copy_boundaries_and_sm_download(ctx):
    """The band copies of the host-buffer call start and end on 256-byte boundaries (vo_ctx::copy_align) and take a few
    bytes of the neighbouring band along; a grid width that is not a multiple of anything puts every band boundary at an
    odd offset, a volume without intervals in its first rows puts the first span copies at the very start of the
    array. Same bits as the exact copies and as the plain path - also with the spans downloaded by the SMs
    (copy_out: k_copy_out reads the band's range on the device, no host round trip)."""
    nx, ny, R = 1531, 1099, 32.0                     # odd on purpose; large enough for the banded path
    rng = np.random.RandomState(11)
    cnt = (rng.uniform(size=nx * ny) < 0.55).astype(np.int64)
    cnt[: nx * 150] = 0                              # empty first rows
    off = np.concatenate(([0], np.cumsum(cnt))).astype(np.uint32)
    n = int(off[-1])
    yy, xx = np.divmod(np.nonzero(cnt)[0], nx)
    a = 200.0 + 40.0 * np.sin(xx * 0.011) * np.cos(yy * 0.013) + rng.uniform(0.0, 0.3, size=n)
    spans = np.stack((a, a + 3.0 + rng.uniform(0.0, 2.0, size=n)), axis=1)
    vol = CompressedVolume(nx, ny, off, spans)
    op = morpho.make_operator("ours", ctx)
    ctx.set_option("pipeline", "off")
    plain, _, _ = op.dilation(vol, R)
    ctx.set_option("pipeline", "on")
    try:
        for opts in ({"copy_align": "256"}, {"copy_align": "0"}, {"copy_align": "4096"}, {"copy_align": "256", "copy_out": "32"}):
            for k, v in opts.items():
                ctx.set_option(k, v)
            for attempt in range(2):
                piped, _, _ = op.dilation(vol, R)
                assert piped.bit_equal(plain), f"{opts}, call {attempt}"
    finally:
        ctx.set_option("copy_align", "256")
        ctx.set_option("copy_out", "0")
