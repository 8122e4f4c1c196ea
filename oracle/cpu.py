"""ctypes wrappers over the two CPU checkers (TEST INFRASTRUCTURE ONLY; see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from voroffset_b200.volume import CompressedVolume, DexelImage

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libvoroffset_ref.so")

OPS3D = {"dilation": 0, "erosion": 1, "opening": 2, "closing": 3}
METHODS = {"ours": 0, "brute_force": 1}
OPS2D = {"dilate": 0, "erode": 1, "open": 2, "close": 3, "negate": 4}

_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> None:
    """Compile liboracle.so (always possible) and oracle/_ref (only where the reference exists)."""
    args = ["make", "-C", _HERE, "all"] + (["FORCE=1"] if force else [])
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)


def _in_arrays(off, spans):
    off64 = np.ascontiguousarray(off, dtype=np.uint64)
    ev = np.ascontiguousarray(spans, dtype=np.float64).reshape(-1)
    if ev.size == 0:
        ev = np.zeros(2)
    return off64, ev


def _take(lib_free, n_lists, poff, pev):
    off = np.ctypeslib.as_array(poff, shape=(n_lists + 1,)).copy()
    m = int(off[-1])
    spans = np.ctypeslib.as_array(pev, shape=(max(2 * m, 1),)).copy()[:2 * m].reshape(-1, 2)
    lib_free(poff)
    lib_free(pev)
    return off, spans


def _curves(curves):
    coff = np.zeros(len(curves) + 1, dtype=np.uint64)
    coff[1:] = np.cumsum([len(c) for c in curves])
    pts = np.ascontiguousarray(np.concatenate([np.asarray(c, dtype=np.float64).reshape(-1, 2) for c in curves])
                               if len(curves) else np.zeros((1, 2))).reshape(-1)
    return coff, pts


class Oracle:
    """Plain-C restatement (oracle.c)."""

    def __init__(self, threads: int = 1):
        if not os.path.exists(ORACLE_SO):
            build()
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_morph3d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _u64p, _f64p, C.c_double,
                                     C.c_double, C.c_double, C.POINTER(_u64p), C.POINTER(_f64p)]
        L.oracle_negate3d.argtypes = [C.c_int, C.c_int, _u64p, _f64p, C.c_double, C.c_double,
                                      C.POINTER(_u64p), C.POINTER(_f64p)]
        L.oracle_negate_inv3d.argtypes = L.oracle_negate3d.argtypes
        L.oracle_xor3d.argtypes = [C.c_int, C.c_int, _u64p, _f64p, _u64p, _f64p, C.c_double, C.c_double,
                                   C.c_double, _f64p, C.POINTER(_u64p), C.POINTER(_f64p)]
        L.oracle_morph2d.argtypes = [C.c_int, C.c_int, C.c_int, _u64p, _f64p, C.c_double,
                                     C.POINTER(_u64p), C.POINTER(_f64p)]
        L.oracle_cap_table_ours.argtypes = [C.c_double, C.POINTER(C.c_int), C.POINTER(_f64p)]
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_from_image2d.argtypes = [C.c_int, C.c_int, _u64p, _f64p, C.POINTER(_u64p), C.POINTER(_f64p)]
        L.oracle_dexelize.argtypes = [C.c_uint64, _f64p, C.c_uint64, C.POINTER(C.c_int32), C.c_double, C.c_double,
                                      C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_u64p), C.POINTER(_f64p)]
        self.set_threads(threads)

    def set_threads(self, n: int):
        self.threads = max(1, int(n))
        self.lib.oracle_set_threads(self.threads)

    def _free(self, p):
        self.lib.oracle_free(C.cast(p, C.c_void_p))

    def from_image(self, width: int, rows: int, curves) -> DexelImage:
        """DoubleCompressedImage(width, rows).fromImage(curves) (DoubleCompressedImage.cpp:25-111); a curve is an
        (n, 2) array of (x, y) = (real, imag) points."""
        coff, pts = _curves(curves)
        poff, pev = _u64p(), _f64p()
        if self.lib.oracle_from_image2d(rows, len(curves), coff.ctypes.data_as(_u64p), pts.ctypes.data_as(_f64p),
                                        C.byref(poff), C.byref(pev)):
            raise RuntimeError("oracle_from_image2d failed")
        o, s = _take(self._free, rows, poff, pev)
        return DexelImage(rows, width, o, s)

    def dexelize(self, V, F, grid: CompressedVolume, window=None) -> CompressedVolume:
        """compute_sign (Dexelize.cpp:166-225) for the columns x0 <= x < x1, y0 <= y < y1 of `grid` (default: all)."""
        V = np.ascontiguousarray(V, dtype=np.float64).reshape(-1, 3)
        F = np.ascontiguousarray(F, dtype=np.int32).reshape(-1, 3)
        x0, x1, y0, y1 = window or (0, grid.nx, 0, grid.ny)
        poff, pev = _u64p(), _f64p()
        rc = self.lib.oracle_dexelize(V.shape[0], V.ctypes.data_as(_f64p), F.shape[0], F.ctypes.data_as(C.POINTER(C.c_int32)),
                                      grid.origin[0], grid.origin[1], grid.spacing, x0, x1, y0, y1,
                                      C.byref(poff), C.byref(pev))
        if rc:
            raise RuntimeError("oracle_dexelize failed")
        off, spans = _take(self._free, (x1 - x0) * (y1 - y0), poff, pev)
        return grid.like(x1 - x0, y1 - y0, off, spans)

    def morph3d(self, vol: CompressedVolume, op: str, radius: float, method: str = "ours") -> CompressedVolume:
        off, ev = _in_arrays(vol.off, vol.spans)
        poff, pev = _u64p(), _f64p()
        rc = self.lib.oracle_morph3d(OPS3D[op], METHODS[method], vol.nx, vol.ny,
                                     off.ctypes.data_as(_u64p), ev.ctypes.data_as(_f64p), float(radius),
                                     vol.zmin, vol.zmax, C.byref(poff), C.byref(pev))
        if rc:
            raise RuntimeError(f"oracle_morph3d failed rc={rc}")
        o, s = _take(self._free, vol.nx * vol.ny, poff, pev)
        return vol.like(vol.nx, vol.ny, o, s)

    def negate3d(self, vol, zlo, zhi):
        off, ev = _in_arrays(vol.off, vol.spans)
        poff, pev = _u64p(), _f64p()
        rc = self.lib.oracle_negate3d(vol.nx, vol.ny, off.ctypes.data_as(_u64p), ev.ctypes.data_as(_f64p),
                                      zlo, zhi, C.byref(poff), C.byref(pev))
        if rc:
            raise RuntimeError("oracle_negate3d failed")
        o, s = _take(self._free, (vol.nx + 2) * (vol.ny + 2), poff, pev)
        return vol.like(vol.nx + 2, vol.ny + 2, o, s)

    def negate_inv3d(self, vol, zlo, zhi):
        off, ev = _in_arrays(vol.off, vol.spans)
        poff, pev = _u64p(), _f64p()
        rc = self.lib.oracle_negate_inv3d(vol.nx, vol.ny, off.ctypes.data_as(_u64p), ev.ctypes.data_as(_f64p),
                                          zlo, zhi, C.byref(poff), C.byref(pev))
        if rc:
            raise RuntimeError("oracle_negate_inv3d failed")
        o, s = _take(self._free, (vol.nx - 2) * (vol.ny - 2), poff, pev)
        return vol.like(vol.nx - 2, vol.ny - 2, o, s)

    def xor3d(self, a: CompressedVolume, b: CompressedVolume):
        oa, ea = _in_arrays(a.off, a.spans)
        ob, eb = _in_arrays(b.off, b.spans)
        poff, pev = _u64p(), _f64p()
        vol = C.c_double(0)
        rc = self.lib.oracle_xor3d(a.nx, a.ny, oa.ctypes.data_as(_u64p), ea.ctypes.data_as(_f64p),
                                   ob.ctypes.data_as(_u64p), eb.ctypes.data_as(_f64p), a.zmin, a.zmax,
                                   a.spacing, C.byref(vol), C.byref(poff), C.byref(pev))
        if rc:
            raise RuntimeError("oracle_xor3d failed")
        o, s = _take(self._free, a.nx * a.ny, poff, pev)
        return vol.value, a.like(a.nx, a.ny, o, s)

    def morph2d(self, img: DexelImage, op: str, r: float) -> DexelImage:
        off, ev = _in_arrays(img.off, img.spans)
        poff, pev = _u64p(), _f64p()
        rc = self.lib.oracle_morph2d(OPS2D[op], img.rows, img.width, off.ctypes.data_as(_u64p),
                                     ev.ctypes.data_as(_f64p), float(r), C.byref(poff), C.byref(pev))
        if rc:
            raise RuntimeError(f"oracle_morph2d failed rc={rc}")
        o, s = _take(self._free, img.rows, poff, pev)
        return DexelImage(img.rows, img.width, o, s)

    def cap_table_ours(self, radius: float) -> np.ndarray:
        J = C.c_int(0)
        p = _f64p()
        if self.lib.oracle_cap_table_ours(float(radius), C.byref(J), C.byref(p)):
            raise RuntimeError("oracle_cap_table_ours failed")
        n = J.value + 1
        t = np.ctypeslib.as_array(p, shape=(n * n,)).copy().reshape(n, n)
        self._free(p)
        return t


def reference_available() -> bool:
    return os.path.exists(REF_SO)


class Reference:
    """The reference's own code (oracle/_ref), driven through ref_driver.cpp."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(f"{REF_SO} not built (needs /root/reference at build time)")
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.ref_free.argtypes = [C.c_void_p]
        L.ref3d_morph.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, _f64p, C.c_double, C.c_int,
                                  _u64p, _f64p, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(_u64p), C.POINTER(_f64p), _f64p, _f64p, C.c_char_p, C.c_int]
        L.ref3d_xor.argtypes = [C.c_int, C.c_int, _f64p, _f64p, C.c_double, C.c_int, _u64p, _f64p, _u64p, _f64p,
                                _f64p, C.POINTER(_u64p), C.POINTER(_f64p), C.c_char_p, C.c_int]
        L.ref2d_from_image.argtypes = [C.c_int, C.c_int, C.c_int, _u64p, _f64p, C.POINTER(_u64p), C.POINTER(_f64p),
                                       C.c_char_p, C.c_int]
        L.ref2d_transpose.argtypes = [C.c_int, C.c_int, _u64p, _f64p, C.POINTER(_u64p), C.POINTER(_f64p), C.c_char_p, C.c_int]
        L.ref2d_morph.argtypes = [C.c_int, C.c_int, C.c_int, _u64p, _f64p, C.c_double,
                                  C.POINTER(_u64p), C.POINTER(_f64p), C.c_char_p, C.c_int]

    def _free(self, p):
        self.lib.ref_free(C.cast(p, C.c_void_p))

    def morph3d(self, vol: CompressedVolume, op: str, radius: float, method: str = "ours", threads: int = 1,
                times: dict | None = None) -> CompressedVolume:
        off, ev = _in_arrays(vol.off, vol.spans)
        origin = (C.c_double * 3)(*vol.origin)
        extent = (C.c_double * 3)(*vol.extent)
        poff, pev = _u64p(), _f64p()
        onx, ony = C.c_int(0), C.c_int(0)
        t1, t2 = C.c_double(0), C.c_double(0)
        err = C.create_string_buffer(1024)
        rc = self.lib.ref3d_morph(OPS3D[op], METHODS[method], int(threads), vol.nx, vol.ny, origin, extent,
                                  vol.spacing, vol.padding, off.ctypes.data_as(_u64p), ev.ctypes.data_as(_f64p),
                                  float(radius), C.byref(onx), C.byref(ony), C.byref(poff), C.byref(pev),
                                  C.byref(t1), C.byref(t2), err, 1024)
        if rc:
            raise RuntimeError(err.value.decode(errors="replace"))
        if times is not None:
            times["time_1"], times["time_2"] = t1.value, t2.value
        o, s = _take(self._free, onx.value * ony.value, poff, pev)
        return vol.like(onx.value, ony.value, o, s)

    def mid_count(self, vol: CompressedVolume, radius: float) -> int:
        """Pieces in the reference's own intermediate volume (VoronoiVorPower.cpp:50-65) = N * k_mid."""
        off, ev = _in_arrays(vol.off, vol.spans)
        origin = (C.c_double * 3)(*vol.origin)
        extent = (C.c_double * 3)(*vol.extent)
        n = C.c_uint64(0)
        err = C.create_string_buffer(1024)
        self.lib.ref3d_mid_count.argtypes = [C.c_int, C.c_int, _f64p, _f64p, C.c_double, C.c_int, _u64p, _f64p,
                                             C.c_double, C.POINTER(C.c_uint64), C.c_char_p, C.c_int]
        rc = self.lib.ref3d_mid_count(vol.nx, vol.ny, origin, extent, vol.spacing, vol.padding,
                                      off.ctypes.data_as(_u64p), ev.ctypes.data_as(_f64p), float(radius),
                                      C.byref(n), err, 1024)
        if rc:
            raise RuntimeError(err.value.decode(errors="replace"))
        return int(n.value)

    def xor3d(self, a: CompressedVolume, b: CompressedVolume):
        oa, ea = _in_arrays(a.off, a.spans)
        ob, eb = _in_arrays(b.off, b.spans)
        origin = (C.c_double * 3)(*a.origin)
        extent = (C.c_double * 3)(*a.extent)
        poff, pev = _u64p(), _f64p()
        vol = C.c_double(0)
        err = C.create_string_buffer(1024)
        rc = self.lib.ref3d_xor(a.nx, a.ny, origin, extent, a.spacing, a.padding,
                                oa.ctypes.data_as(_u64p), ea.ctypes.data_as(_f64p),
                                ob.ctypes.data_as(_u64p), eb.ctypes.data_as(_f64p),
                                C.byref(vol), C.byref(poff), C.byref(pev), err, 1024)
        if rc:
            raise RuntimeError(err.value.decode(errors="replace"))
        o, s = _take(self._free, a.nx * a.ny, poff, pev)
        return vol.value, a.like(a.nx, a.ny, o, s)

    def from_image(self, width: int, rows: int, curves) -> DexelImage:
        """The reference's own DoubleCompressedImage::fromImage (DoubleCompressedImage.cpp:25-40)."""
        coff, pts = _curves(curves)
        poff, pev = _u64p(), _f64p()
        err = C.create_string_buffer(1024)
        rc = self.lib.ref2d_from_image(width, rows, len(curves), coff.ctypes.data_as(_u64p), pts.ctypes.data_as(_f64p),
                                       C.byref(poff), C.byref(pev), err, 1024)
        if rc:
            raise RuntimeError(err.value.decode(errors="replace"))
        o, s = _take(self._free, rows, poff, pev)
        return DexelImage(rows, width, o, s)

    def transposed(self, img: DexelImage) -> DexelImage:
        """The reference's own DoubleCompressedImage::transposeInPlace (DoubleCompressedImage.cpp:478-584)."""
        off, ev = _in_arrays(img.off, img.spans)
        poff, pev = _u64p(), _f64p()
        err = C.create_string_buffer(1024)
        rc = self.lib.ref2d_transpose(img.rows, img.width, off.ctypes.data_as(_u64p), ev.ctypes.data_as(_f64p),
                                      C.byref(poff), C.byref(pev), err, 1024)
        if rc:
            raise RuntimeError(err.value.decode(errors="replace"))
        o, s = _take(self._free, img.width, poff, pev)
        return DexelImage(img.width, img.rows, o, s)

    def morph2d(self, img: DexelImage, op: str, r: float) -> DexelImage:
        off, ev = _in_arrays(img.off, img.spans)
        poff, pev = _u64p(), _f64p()
        err = C.create_string_buffer(1024)
        rc = self.lib.ref2d_morph(OPS2D[op], img.rows, img.width, off.ctypes.data_as(_u64p),
                                  ev.ctypes.data_as(_f64p), float(r), C.byref(poff), C.byref(pev), err, 1024)
        if rc:
            raise RuntimeError(err.value.decode(errors="replace"))
        o, s = _take(self._free, img.rows, poff, pev)
        return DexelImage(img.rows, img.width, o, s)
