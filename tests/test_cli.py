"""The re-hosted offset3d / offset2d executables (voroffset_b200/cpp): flags, JSON keys, dexeliser and
mesh export on CPU; the full pipeline against the Python mirror on the GPU."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "voroffset_b200", "cpp", "bin")


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "voroffset_b200", "cpp"), "-s"], check=True)


def torus_obj(path, major=1.0, minor=0.35, nu=96, nv=48, shift=(0.013, 0.007, 0.003)):
    """Closed triangle mesh of a torus with axis z (slightly shifted so that no vertex sits on a column centre)."""
    with open(path, "w") as f:
        for i in range(nu):
            for j in range(nv):
                u, v = 2 * math.pi * i / nu, 2 * math.pi * j / nv
                r = major + minor * math.cos(v)
                f.write(f"v {r * math.cos(u) + shift[0]:.17g} {r * math.sin(u) + shift[1]:.17g} {minor * math.sin(v) + shift[2]:.17g}\n")
        idx = lambda i, j: (i % nu) * nv + (j % nv) + 1
        for i in range(nu):
            for j in range(nv):
                a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
                f.write(f"f {a} {b} {c}\nf {a} {c} {d}\n")


def run(exe, *args):
    return subprocess.run([os.path.join(BIN, exe), *map(str, args)], capture_output=True, text=True, timeout=600)


def test_help_and_flag_validation():
    _build()
    r = run("offset3d", "--help")
    assert r.returncode == 0
    for flag in ("-i", "-o", "-j", "-d", "-n", "-p", "-t", "-r", "-m", "-x", "-f", "-u", "--gpus"):   # offset3d.cpp:38-51 (+ ours)
        assert flag in r.stderr
    assert run("offset3d").returncode == 1
    assert run("offset3d", "/nonexistent.obj").returncode == 1
    assert run("offset2d", "--help").returncode == 0


def test_dexeliser_and_export_without_gpu(tmp_path):
    """-x noop: mesh -> dexels -> volume text file / hex OBJ / xyz points, no GPU needed."""
    _build()
    from voroffset_b200.volume import CompressedVolume
    mesh = tmp_path / "torus.obj"
    torus_obj(mesh)
    vol_path = tmp_path / "torus.vol"
    r = run("offset3d", mesh, vol_path, "-n", 64, "-p", 3, "-x", "noop", "-j", tmp_path / "o.json")
    assert r.returncode == 0, r.stderr
    with open(vol_path) as f:
        vol = CompressedVolume.load(f)
    # grid = ceil(extent/spacing) + 2 p (CompressedVolume.cpp:18-21); spacing = max extent / n (Dexelize.cpp:259-263)
    assert vol.padding == 3 and vol.nx == 64 + 6 and vol.ny == 64 + 6
    assert vol.spacing == pytest.approx(2.7 / 64, rel=1e-3)
    cnt = vol.counts()
    assert cnt.max() == 1 and 0.55 < cnt.sum() / (64 * 64) < 0.65        # k_in ~ 0.6 of the bounding square
    # interval half-length at a column matches the analytic torus within the polygonisation error
    x, y = vol.nx // 2 + 22, vol.ny // 2
    cx = (x + 0.5) * vol.spacing + vol.origin[0] - 0.013
    cy = (y + 0.5) * vol.spacing + vol.origin[1] - 0.007
    half = math.sqrt(max(0.35 ** 2 - (math.hypot(cx, cy) - 1.0) ** 2, 0.0)) / vol.spacing
    z = vol.at(x, y)
    assert z.size == 2 and (z[1] - z[0]) / 2 == pytest.approx(half, rel=0.02)
    js = json.load(open(tmp_path / "o.json"))
    for key in ("method", "num_threads", "model_name", "voxel_size", "padding", "num_dexels", "radius", "grid_size",
                "num_segments", "operation", "time", "time_first_pass", "time_second_pass"):     # offset3d.cpp:83-96,156-177
        assert key in js
    assert js["grid_size"] == [vol.nx, vol.ny] and js["num_segments"] == vol.numSegments()
    # existing outputs are not overwritten without -f
    r2 = run("offset3d", mesh, vol_path, "-n", 32, "-x", "noop")
    assert "already exists" in r2.stdout
    # hex mesh and point dump
    assert run("offset3d", mesh, tmp_path / "hex.obj", "-n", 24, "-x", "noop").returncode == 0
    lines = open(tmp_path / "hex.obj").read().splitlines()
    nv, nf = sum(l.startswith("v ") for l in lines), sum(l.startswith("f ") for l in lines)
    assert nv % 8 == 0 and nf == 6 * nv // 8 and nv > 0
    assert run("offset3d", mesh, tmp_path / "pts.xyz", "-n", 24, "-x", "noop").returncode == 0
    assert len(open(tmp_path / "pts.xyz").read().splitlines()) == 2 * nv // 8
    # the reference's hex mesh (Dexelize.cpp:312-351) in Medit format: eight vertices, six border quads and one cell per
    # interval; every cell has positive volume in Medit's corner order and its quads are the faces of its own box
    assert run("offset3d", mesh, tmp_path / "hex.mesh", "-n", 24, "-x", "noop").returncode == 0
    tok = open(tmp_path / "hex.mesh").read().split()
    iv, iq, ih = tok.index("Vertices"), tok.index("Quadrilaterals"), tok.index("Hexahedra")
    n_v, n_q, n_h = int(tok[iv + 1]), int(tok[iq + 1]), int(tok[ih + 1])
    assert n_v == nv and n_q == 6 * n_h and n_v == 8 * n_h and tok[-1] == "End"
    V = np.array(tok[iv + 2: iv + 2 + 4 * n_v], dtype=float).reshape(-1, 4)[:, :3]
    H = np.array(tok[ih + 2: ih + 2 + 9 * n_h], dtype=int).reshape(-1, 9)[:, :8] - 1
    Q = np.array(tok[iq + 2: iq + 2 + 5 * n_q], dtype=int).reshape(-1, 5)[:, :4] - 1
    for h in H[:: max(1, n_h // 50)]:
        p = V[h]
        e1, e2, e3 = p[1] - p[0], p[3] - p[0], p[4] - p[0]          # Medit: 0-1-2-3 bottom, 4-7 above them
        assert np.dot(np.cross(e1, e2), e3) > 0
        assert np.allclose(p[2], p[0] + e1 + e2) and np.allclose(p[6], p[0] + e1 + e2 + e3)
    c = 0
    box = V[8 * c: 8 * c + 8]
    centre = box.mean(axis=0)
    for q in Q[6 * c: 6 * c + 6]:
        assert set(q) <= set(range(8 * c, 8 * c + 8))
        a, b, d = V[q[0]], V[q[1]], V[q[3]]
        assert np.dot(np.cross(b - a, d - a), V[q].mean(axis=0) - centre) > 0      # outward


def test_no_cpu_fallback_in_cli(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _build()
    mesh = tmp_path / "torus.obj"
    torus_obj(mesh, nu=24, nv=12)
    r = run("offset3d", mesh, tmp_path / "o.vol", "-n", 16, "-r", 2, "-x", "dilation")
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["ours", "brute_force"])
def test_offset3d_pipeline_matches_the_oracle(tmp_path, ctx, oracle, method):
    """The executable end to end (mesh -> device dexeliser -> operator -> .vol) against the ORACLE run on the dexelised
    input the same executable wrote with -x noop (host loop), and against the Python mirror of the operator interface."""
    _build()
    import util
    from voroffset_b200 import morpho
    from voroffset_b200.volume import CompressedVolume
    mesh = tmp_path / "torus.obj"
    torus_obj(mesh)
    assert run("offset3d", mesh, tmp_path / "in.vol", "-n", 96, "-p", 7, "-x", "noop").returncode == 0
    with open(tmp_path / "in.vol") as f:
        vin = CompressedVolume.load(f)
    op = morpho.make_operator(method, ctx)
    for operation in ("dilation", "erosion", "closing", "opening"):      # scripts/fig19.sh:29-33 uses exactly these
        out = tmp_path / f"{operation}.vol"
        r = run("offset3d", mesh, out, "-n", 96, "-p", 7, "-r", 5.5, "-x", operation, "-m", method, "-j", tmp_path / f"{operation}.json", "-f")
        assert r.returncode == 0, r.stderr
        with open(out) as f:
            got = CompressedVolume.load(f)
        util.assert_same(got, oracle.morph3d(vin, operation, 5.5, method), operation, method, "offset3d vs oracle")
        want, _, _ = morpho.apply_operation(op, operation, vin, 5.5)
        if operation in ("closing", "opening"):
            # the CLI composes two calls through the host (offset3d.cpp:124-133), the mirror composes on the device
            assert got.same_topology(want) and np.abs(got.spans - want.spans).max() < 1e-9
        else:
            assert got.bit_equal(want)
        js = json.load(open(tmp_path / f"{operation}.json"))
        assert js["operation"] == operation and js["method"] == method and js["radius"] == 5.5
        assert js["time_first_pass"] >= 0


@pytest.mark.gpu
def test_offset2d_pipeline(tmp_path, ctx, oracle):
    _build()
    from voroffset_b200 import synth
    from voroffset_b200.volume import DexelImage
    img = synth.star_image(128, 160, 9, seed=3)
    src = tmp_path / "in.dex"
    with open(src, "w") as f:                               # DoubleCompressedImage::save format
        f.write(f"{img.width} {img.rows}\n")
        for i in range(img.rows):
            row = img.at(i)
            f.write(str(row.size) + "".join(f" {v:.17g}" for v in row) + "\n")

    def load(path):
        tok = open(path).read().split()
        w, h = int(tok[0]), int(tok[1])
        k, lists = 2, []
        for _ in range(h):
            n = int(tok[k]); lists.append([float(v) for v in tok[k + 1:k + 1 + n]]); k += 1 + n
        return DexelImage.from_lists(w, lists)

    assert run("offset2d", src, "-o", tmp_path / "d.dex", "-r", 6.0 / 128).returncode == 0
    assert load(tmp_path / "d.dex").bit_equal(oracle.morph2d(img, "dilate", 6.0 / 128))
    assert run("offset2d", src, "-o", tmp_path / "e.dex", "-r", 3.0, "-e").returncode == 0
    assert load(tmp_path / "e.dex").bit_equal(oracle.morph2d(img, "erode", 3.0))
    assert run("offset2d", src, "-o", tmp_path / "n.dex", "-n").returncode == 0
    assert load(tmp_path / "n.dex").bit_equal(oracle.morph2d(img, "negate", 0.0))


def test_mesh_reader_rejects_bad_indices(tmp_path):
    """A facet that names a vertex outside the file is an error, not an out-of-bounds read (host loop and device path
    share the reader)."""
    _build()
    for name, text in (("a.obj", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 99\n"), ("b.obj", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 -7\n"),
                       ("c.off", "OFF\n3 1 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 5\n")):
        mesh = tmp_path / name
        mesh.write_text(text)
        r = run("offset3d", mesh, "-n", 8, "-x", "noop")
        assert r.returncode == 1 and "Invalid input mesh" in r.stderr, (name, r.stderr)
