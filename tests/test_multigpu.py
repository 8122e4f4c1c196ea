"""Multi-GPU operators behind the C ABI (vo_mg_*, csrc/vo_mg.cuh): y-slabs, NCCL halo exchange inside the library.
Everything is compared with the single-GPU call, which the other tests pin against the oracle and the reference's
goldens. Tests that need two devices skip on a single-GPU box (scripts/mg_check.py runs the same checks under
`gpurun --gpus N`; its log is committed under profiles/)."""
import os
import subprocess

import numpy as np
import pytest

import util
from voroffset_b200 import _lib, morpho, multigpu, synth

pytestmark = pytest.mark.gpu

OPS = ("dilation", "erosion", "opening", "closing")


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_group_of_one_equals_the_single_context_call(ctx, oracle):
    """n_dev = 1 needs no NCCL; it still goes through the slab code (border rows owned by the only slab, ...)."""
    mg = multigpu.MultiGpu.single_process([0])
    assert mg.world == 1 and mg.local_count == 1
    vol = synth.random_volume(60, 44, kmax=4, padding=5, seed=17)
    for method in ("ours", "brute_force"):
        for opn in OPS:
            got, _, _ = mg.morph(opn, vol, 4.3, method)
            util.assert_same(got, oracle.morph3d(vol, opn, 4.3, method), opn, method, "vo_mg, one GPU")
    d = morpho.DeviceVolume.upload(mg.contexts[0], vol)
    outs, t1, t2 = mg.morph_dev("closing", [d], 4.3, vol.zmin, vol.zmax)
    want, _, _ = morpho.make_operator("ours", ctx).closing(vol, 4.3)
    assert outs[0].download().bit_equal(want)
    st = mg.stats(0)
    assert st["messages"] == 0 and st["halo_bytes"] == 0
    outs[0].free(); d.free()
    mg.close()


def test_group_rejects_bad_input():
    mg = multigpu.MultiGpu.single_process([0])
    vol = synth.random_volume(20, 16, kmax=3, padding=2, seed=1)
    off = vol.off.copy()
    off[5] = off[4] + 1000
    from voroffset_b200.volume import CompressedVolume
    with pytest.raises(_lib.VoroffsetError):
        mg.morph("dilation", CompressedVolume(vol.nx, vol.ny, off, vol.spans), 3.0)
    with pytest.raises(_lib.VoroffsetError):
        mg.morph("dilation", vol, -1.0)
    got, _, _ = mg.morph("dilation", vol, 3.0)
    assert got.numSegments() > 0
    mg.close()


CASES = [
    ("torus_z_n384_p20_R10", lambda: synth.torus_z(384, padding=20), 10.0),
    ("random_k8_150x300_R6.4", lambda: synth.random_volume(150, 300, kmax=8, padding=7, seed=12), 6.4),
    ("lattice_n160_R5", lambda: synth.lattice(160, padding=8), 5.0),
    ("thin_slabs_64x70_R16", lambda: synth.random_volume(64, 70, kmax=3, padding=17, seed=5), 16.5),    # fewer rows than GPUs x halo
]


@pytest.mark.parametrize("name,gen,radius", CASES, ids=[c[0] for c in CASES])
def test_slabs_over_all_gpus_equal_the_single_gpu_result(ctx, name, gen, radius):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    vol = gen()
    mg = multigpu.MultiGpu.single_process(list(range(n)))
    for method in ("ours", "brute_force"):
        op = morpho.make_operator(method, ctx)
        for opn in OPS if method == "ours" else ("dilation", "erosion"):
            want, _, _ = morpho.apply_operation(op, opn, vol, radius)
            for attempt in range(2):                  # the second call has link capacities: one message, overlapped where it applies
                got, _, _ = mg.morph(opn, vol, radius, method)
                assert got.bit_equal(want), f"{name}: {method} {opn} on {n} GPUs differs from one GPU (call {attempt})"
    mg.close()


def test_config5_on_all_gpus_bit_identical_and_overlapped(ctx):
    """BASELINE config 5 as written: ONE 2048 x 2048 grid, R = 32, cut over the GPUs of the box; every operation against
    the single-GPU result and the dilation against the reference's digest. The steady-state steps take the overlapped
    slab path (one NCCL group per step)."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    mg = multigpu.MultiGpu.single_process(list(range(n)))
    vol = synth.torus_z(2048)
    z = util.golden_full("c5_torus_z_n2048_r32")
    for attempt in range(3):
        got, _, _ = mg.morph("dilation", vol, 32.0)
        util.assert_digest(got, z, "dilation", what=f"config 5 on {n} GPUs")
    assert mg.stats(0)["overlapped"] == 1 and mg.stats(0)["messages"] == 1
    volp = synth.torus_z(2048, padding=34)
    op = morpho.make_operator("ours", ctx)
    for opn in ("erosion", "closing", "opening"):
        want, _, _ = morpho.apply_operation(op, opn, volp, 32.0)
        for attempt in range(2):
            got, _, _ = mg.morph(opn, volp, 32.0)
            assert got.bit_equal(want), f"config 5 {opn} on {n} GPUs (call {attempt})"
    mg.close()


def test_erosion_in_dual_form_on_all_gpus(ctx):
    """Volumes with one interval per column erode in dual form on every slab (the ranks agree first: one all-reduced
    word). "erosion" = "dual" on every context makes a fallback an error, so this is that path - and a grid in which
    only ONE slab holds a column with two intervals must make the whole group take the general path, same result."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    mg = multigpu.MultiGpu.single_process(list(range(n)))
    op = morpho.make_operator("ours", ctx)
    vol = synth.torus_z(512, padding=20)
    want, _, _ = op.erosion(vol, 14.0)
    for c in mg.contexts:
        c.set_option("pass1", "tile")        # (the dual form lives in the tile kernel, which slabs of 552 / 8 rows would be too small for)
        c.set_option("erosion", "dual")
    for attempt in range(2):
        got, _, _ = mg.morph("erosion", vol, 14.0)
        assert got.bit_equal(want), f"dual erosion on {n} GPUs differs from one GPU (call {attempt})"
    for c in mg.contexts:
        c.set_option("erosion", "auto")
    for opn in ("opening", "closing"):
        want, _, _ = morpho.apply_operation(op, opn, vol, 14.0)
        got, _, _ = mg.morph(opn, vol, 14.0)
        assert got.bit_equal(want), f"{opn} on {n} GPUs"
    # one column of the LAST slab gets a second interval
    from voroffset_b200.volume import CompressedVolume
    lists = [tuple(vol.at(x, y)) for y in range(vol.ny) for x in range(vol.nx)]
    c = vol.nx // 2 + vol.nx * (vol.ny - 60)
    a0, b0 = lists[c]
    lists[c] = (a0, a0 + 0.3 * (b0 - a0), a0 + 0.6 * (b0 - a0), b0)
    mixed = CompressedVolume.from_lists(vol.nx, vol.ny, lists, origin=vol.origin, extent=vol.extent, spacing=vol.spacing, padding=vol.padding)
    want, _, _ = op.erosion(mixed, 14.0)
    got, _, _ = mg.morph("erosion", mixed, 14.0)
    assert got.bit_equal(want), "mixed grid: the group must agree on the general path"
    for c in mg.contexts:
        c.set_option("erosion", "dual")
    with pytest.raises(_lib.VoroffsetError):
        mg.morph("erosion", mixed, 14.0)
    mg.close()


def test_halo_capacity_overflow_falls_back(ctx):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    mg = multigpu.MultiGpu.single_process(list(range(n)))
    op = morpho.make_operator("ours", ctx)
    sparse = synth.random_volume(300, 40 * n, kmax=1, padding=0, seed=8, fill=0.05)
    dense = synth.random_volume(300, 40 * n, kmax=8, padding=0, seed=9)
    for vol in (sparse, sparse, dense, dense, sparse):
        want, _, _ = op.dilation(vol, 7.5)
        got, _, _ = mg.morph("dilation", vol, 7.5)
        assert got.bit_equal(want)
    mg.close()


def test_offset3d_gpus_flag(tmp_path):
    """offset3d --gpus N -x closing is bit-identical to the single-GPU run of the same executable."""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    from test_cli import BIN, _build, run, torus_obj
    _build()
    mesh = tmp_path / "torus.obj"
    torus_obj(mesh)
    outs = {}
    for g in (1, n):
        for opn in ("closing", "erosion"):
            out = tmp_path / f"{opn}_{g}.vol"
            r = run("offset3d", mesh, out, "-n", 200, "-p", 12, "-r", 9.5, "-x", opn, "--gpus", g, "-j", tmp_path / f"{opn}_{g}.json")
            assert r.returncode == 0, r.stderr
            outs[(opn, g)] = open(out).read()
    for opn in ("closing", "erosion"):
        assert outs[(opn, 1)] == outs[(opn, n)]
