"""2D ingestion (the step before vor2d's morphology): DoubleCompressedImage::fromImage and the SVG front end of the
re-hosted offset2d.

  * oracle.from_image (plain C, oracle/oracle.c) is PINNED bit for bit against the reference's own fromImage compiled
    in place (oracle/_ref: ref2d_from_image) on seeded polygons, vertices exactly on scan lines included;
  * the host C++ restatement (voroffset_b200/cpp/vo_svg.cpp) is checked through `offset2d file.svg` with radius 0
    (no GPU involved) against the oracle and, where built, the reference;
  * the SVG reader follows nanosvg's conventions (float coordinates, 1/3 - 2/3 control points on straight segments,
    mm at 90 DPI); nanosvg itself is absent, so that part is unpinned and only checked against a numpy model of
    those conventions.
"""
import math
import os
import subprocess

import numpy as np
import pytest

from voroffset_b200.volume import DexelImage

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "voroffset_b200", "cpp", "bin")
f32 = np.float32


def star(cx, cy, rad, n, phase=0.1):
    a = np.linspace(0, 2 * np.pi, 2 * n, endpoint=False) + phase
    rr = np.where(np.arange(2 * n) % 2 == 0, rad, rad * 0.45)
    return np.stack([cx + rr * np.cos(a), cy + rr * np.sin(a)], 1)


def seeded_curves(seed, count, w, h):
    rng = np.random.default_rng(seed)
    g = int(math.ceil(math.sqrt(count)))
    out = []
    for k in range(count):
        cx = (k % g + 0.5) * h / g + rng.uniform(-1, 1)
        cy = (k // g + 0.5) * w / g + rng.uniform(-1, 1)
        out.append(star(cx, cy, 0.42 * min(w, h) / g, int(rng.integers(3, 9)), rng.uniform(0, 1)))
    return out


@pytest.mark.parametrize("seed,count", [(1, 4), (2, 9), (3, 25)])
def test_oracle_from_image_matches_the_reference(oracle, reference, seed, count):
    w, h = 120, 90
    curves = seeded_curves(seed, count, w, h)
    # vertices exactly on scan lines, edges along scan lines, a curve in front of the ray's content
    curves.append(np.array([[5, 5], [12, 5], [12, 9], [5, 9.5]], float))
    curves.append(np.array([[30, 1.0], [30, 2.5], [33, 2.5], [36, 4], [36, 1.0]], float))
    got = oracle.from_image(w, h, curves)
    want = reference.from_image(w, h, curves)
    assert got.numSegments() > 50
    assert got.bit_equal(want)
    # order matters (unionIntersections prepends / appends whole curves): reversed input, same check
    assert oracle.from_image(w, h, curves[::-1]).bit_equal(reference.from_image(w, h, curves[::-1]))


def _svg_polygons(path, polys_px, width_px, height_px):
    with open(path, "w") as f:
        f.write(f'<?xml version="1.0"?>\n<!-- test -->\n<svg xmlns="http://www.w3.org/2000/svg" width="{width_px}" height="{height_px}">\n')
        for k, p in enumerate(polys_px):
            pts = " ".join(f"{float(x)!r},{float(y)!r}" for x, y in p)
            if k % 2 == 0:
                f.write(f'  <polygon points="{pts}"/>\n')
            else:
                d = "M " + " L ".join(f"{float(x)!r} {float(y)!r}" for x, y in p) + " Z"
                f.write(f'  <path d="{d}" style="fill:#000"/>\n')
        f.write('  <defs><rect x="1" y="1" width="5" height="5"/></defs>\n  <g style="display: none"><rect x="1" y="1" width="50" height="50"/></g>\n</svg>\n')


def _nanosvg_model(polys_px, width_px, height_px):
    """Contours in mm the way nanosvg stores straight segments (float32; control points at 1/3 and 2/3) and
    Dexelize.cpp:33-37 reads them (all points but the last), plus the image size of Dexelize.cpp:42."""
    us = f32(1.0) / (f32(1.0) / f32(25.4) * f32(90.0))
    curves = []
    for p in polys_px:
        q = [(f32(x), f32(y)) for x, y in p]
        pts = [q[0]]
        for (x, y) in q[1:] + [q[0]]:
            px, py = pts[-1]
            dx, dy = f32(x - px), f32(y - py)
            pts += [(f32(px + f32(dx / f32(3))), f32(py + f32(dy / f32(3)))), (f32(x - f32(dx / f32(3))), f32(y - f32(dy / f32(3)))), (x, y)]
        pts = pts[:-1]
        curves.append(np.array([[float(f32(x * us)), float(f32(y * us))] for x, y in pts]))
    w_mm, h_mm = float(f32(f32(width_px) * us)), float(f32(f32(height_px) * us))
    return curves, int(math.ceil(h_mm * 25.4 / 90)), int(math.ceil(w_mm * 25.4 / 90))


def _offset2d(*args):
    subprocess.run(["make", "-C", os.path.join(ROOT, "voroffset_b200", "cpp"), "-s"], check=True)
    return subprocess.run([os.path.join(BIN, "offset2d"), *map(str, args)], capture_output=True, text=True, timeout=120)


def _load_dex(path):
    with open(path) as f:
        w, n = map(int, f.readline().split())
        lists = []
        for _ in range(n):
            t = f.readline().split()
            lists.append([float(v) for v in t[1:1 + int(t[0])]])
    return DexelImage.from_lists(w, lists)


def test_offset2d_reads_svg_like_the_reference(oracle, tmp_path):
    # pixels; after the reference's unit handling only x < ~150 mm (530 px) is ever scanned
    polys = [star(150 + 170 * (k % 3), 200 + 260 * (k // 3), 70 + 5 * k, 4 + k, 0.2 * k) for k in range(6)]
    svg, out = tmp_path / "in.svg", tmp_path / "out.dex"
    _svg_polygons(svg, polys, 1800, 1400)
    r = _offset2d(svg, "-o", out)
    assert r.returncode == 0, r.stderr
    assert "(6 polygones)" in r.stderr                      # Dexelize.cpp:39-40; defs and hidden groups are skipped
    got = _load_dex(out)
    curves, ray_len, n_rays = _nanosvg_model(polys, 1800, 1400)
    assert (got.rows, got.width) == (n_rays, ray_len)
    want = oracle.from_image(ray_len, n_rays, curves)
    assert want.numSegments() > 100
    assert got.bit_equal(want)
    from oracle.cpu import Reference, reference_available
    if reference_available():
        assert got.bit_equal(Reference().from_image(ray_len, n_rays, curves))


def test_svg_reader_units_viewbox_transforms_and_errors(tmp_path):
    def run(body, head='width="200mm" height="100mm" viewBox="0 0 400 200"'):
        svg, out = tmp_path / "t.svg", tmp_path / "t.dex"
        svg.write_text(f'<svg {head}>{body}</svg>')
        r = _offset2d(svg, "-o", out, "-f")
        return r, (_load_dex(out) if r.returncode == 0 else None)

    # 200 mm = 708.66 px wide; viewBox 400 wide -> 1 user unit = 0.5 mm; rays = ceil(200 * 25.4 / 90) = 57
    r, img = run('<rect x="20" y="40" width="60" height="100"/>')
    assert r.returncode == 0, r.stderr
    assert img.rows == 57 and img.width == math.ceil(100 * 25.4 / 90)
    cnt = np.diff(img.off.astype(np.int64))
    assert list(np.nonzero(cnt)[0]) == list(range(11, 40))           # 10 mm < x < 40 mm (a vertex ON a line does not count twice)
    assert np.allclose(img.spans[0], [20.0, 70.0], atol=1e-4)
    # quad-mesh export (src/vor2d/Dexelize.cpp:48-89): four vertices and one quad per interval
    obj = tmp_path / "t.obj"
    assert _offset2d(tmp_path / "t.svg", "-o", obj, "-f").returncode == 0
    lines = obj.read_text().split("\n")
    assert sum(l.startswith("v ") for l in lines) == 4 * img.numSegments() and sum(l.startswith("f ") for l in lines) == img.numSegments()
    assert lines[0].split()[:2] == ["v", "11"] and float(lines[0].split()[2]) == float(img.spans[0, 0])
    # the same rectangle through a group transform and a relative path
    r2, img2 = run('<g transform="translate(20 40) scale(2)"><path d="m 0 0 h 30 v 50 h -30 z"/></g>')
    assert r2.returncode == 0, r2.stderr
    assert np.array_equal(img2.off, img.off) and np.allclose(img2.spans, img.spans, atol=1e-4)
    # circle: four cubics whose control points become vertices (Dexelize.cpp:33-37) -> an octagon-like outline
    r3, img3 = run('<circle cx="60" cy="100" r="40"/>')
    assert r3.returncode == 0 and np.diff(img3.off.astype(np.int64)).max() == 1
    for bad in ('<path d="M 0 0 A 5 5 0 0 1 10 10 Z"/>', '<rect x="1" y="1" width="5" height="5" rx="1"/>', '<use href="#a"/>'):
        rb, _ = run(bad)
        assert rb.returncode == 1 and "not supported" in rb.stderr


def test_offset2d_transpose_matches_the_reference_fixture(tmp_path):
    """The same without oracle/_ref: tests/golden/ingest2d.npz holds the reference's transposeInPlace output."""
    import util
    z = np.load(os.path.join(util.GOLDEN, "ingest2d.npz"))
    img = DexelImage(int(z["t_rows"]), int(z["t_width"]), z["t_in_off"], z["t_in_spans"])
    src, out = tmp_path / "in.dex", tmp_path / "out.dex"
    with open(src, "w") as f:
        f.write(f"{img.width} {img.rows}\n")
        for i in range(img.rows):
            row = img.spans[int(img.off[i]):int(img.off[i + 1])].reshape(-1)
            f.write(str(row.size) + "".join(f" {v!r}" for v in row.tolist()) + "\n")
    r = _offset2d(src, "-o", out, "-t")
    assert r.returncode == 0, r.stderr
    got = _load_dex(out)
    assert got.off.tolist() == z["t_off"].tolist() and (got.spans.view("u8") == z["t_spans"].view("u8")).all()


@pytest.mark.parametrize("seed", [4, 5, -1])
def test_offset2d_transpose_matches_the_reference(reference, tmp_path, seed):
    """offset2d -t: DoubleCompressedImage::transposeInPlace (DoubleCompressedImage.cpp:478-584, events truncated to
    int) restated in voroffset_b200/cpp/vo_svg.cpp, against the reference's own routine."""
    from voroffset_b200 import synth
    if seed >= 0:
        img = synth.random_image(37, 53, kmax=4, seed=seed)
    else:
        # sub-pixel intervals (both ends truncate to the same line: ONE event), intervals ending on the same line in
        # neighbouring rays, an empty ray in between, events on the first and on the last line
        img = DexelImage.from_lists(12, [[3.2, 3.7, 5.0, 9.9], [3.9, 6.1], [], [0.0, 2.5, 2.6, 11.0], [0.4, 11.9], [7.5, 7.6]])
    src, out = tmp_path / "in.dex", tmp_path / "out.dex"
    with open(src, "w") as f:
        f.write(f"{img.width} {img.rows}\n")
        for i in range(img.rows):
            row = img.spans[int(img.off[i]):int(img.off[i + 1])].reshape(-1)
            f.write(str(row.size) + "".join(f" {v!r}" for v in row.tolist()) + "\n")
    r = _offset2d(src, "-o", out, "-t")
    assert r.returncode == 0, r.stderr
    got = _load_dex(out)
    want = reference.transposed(img)
    assert (got.rows, got.width) == (want.rows, want.width) == (img.width, img.rows)
    assert got.numSegments() > (20 if seed >= 0 else 3) and got.bit_equal(want)


def test_svg_reader_survives_malformed_input(tmp_path):
    """Seeded fragments glued together at random: the executable must finish with exit code 0 or 1 (error reported),
    never crash or hang."""
    import random
    rnd = random.Random(5)
    frags = ['<svg width="100" height="80">', '<svg viewBox="0 0 10 10" width="50mm" height="40mm">', '<svg>', '</svg>',
             '<g transform="translate(3,4) rotate(30)">', '</g>', '<path d="M 1 1 L 5 1 L 5 5 Z"/>',
             '<path d="m1,1 5,0 0,5z M 20 20 h 5 v 5 h -5 z"/>', '<path d="M 1 1 C 2 2 3 3 4 1 S 6 0 7 1 Q 8 8 9 1 T 10 10 Z"/>',
             '<path d="M"/>', '<path d="L 1 1 2"/>', '<path d="M 1e400 1 L nan 2 Z"/>', '<polygon points="1,1 9,1 5,9"/>',
             '<polygon points="1,1 9"/>', '<polyline points=""/>', '<rect x="1" y="1" width="30" height="20"/>',
             '<rect width="-1" height="5"/>', '<circle cx="20" cy="20" r="10"/>', '<ellipse cx="5" cy="5" rx="3" ry="0"/>',
             '<line x1="0" y1="0" x2="10" y2="10"/>', '<defs>', '</defs>', '<!-- c -->', '<?xml?>', '<svg width="abc">',
             "<path d='M 1 1 L 2 2", '<g transform="matrix(1 0 0 1)">', '<g transform="scale(0)">',
             '<path transform="skewX(89.999)" d="M0 0 L 1 1 L 0 1 z"/>', '<text>hi</text>', '<<<>>>']
    svg, out = tmp_path / "f.svg", tmp_path / "f.dex"
    for _ in range(40):
        body = "".join(rnd.choice(frags) for _ in range(rnd.randint(1, 8)))
        if rnd.random() < 0.7:
            body = f'<svg width="{rnd.randint(1, 400)}" height="{rnd.randint(1, 400)}">' + body + "</svg>"
        svg.write_text(body)
        r = subprocess.run([os.path.join(BIN, "offset2d"), str(svg), "-o", str(out), "-f"], capture_output=True, text=True, timeout=30)
        assert r.returncode in (0, 1), (r.returncode, body, r.stderr[-200:])
