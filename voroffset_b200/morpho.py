"""Host-side mirror of the reference's operator interface for the 3D path.

`VoronoiMorpho` mirrors `voroffset3d::VoronoiMorpho` (src/vor3d/Voronoi.h:13-44): `dilation`,
`erosion`, `calculateXor`; the concrete classes are picked by method name exactly like
app/cli3d/offset3d.cpp:104-112 ("ours" -> VoronoiMorphoVorPower, "brute_force" ->
VoronoiMorphoBruteForce) and `apply_operation` is the -x switch of offset3d.cpp:116-136.

Differences forced by Python: out-parameters become return values
(`result, time_1, time_2 = op.dilation(input, radius)`), `vor_assert` failures surface as
`VoroffsetError` (a RuntimeError, like the reference's std::runtime_error, Common.cpp:7-17).
Every call goes through the C ABI; there is no CPU implementation behind these classes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .volume import CompressedVolume


class DeviceVolume:
    """A dexel volume resident in HBM (vo_dvol)."""

    def __init__(self, ctx: _lib.Context, handle, meta: CompressedVolume | None = None):
        self.ctx, self.handle, self.meta = ctx, handle, meta

    @classmethod
    def upload(cls, ctx: _lib.Context, vol: CompressedVolume) -> "DeviceVolume":
        h = C.c_void_p()
        spans = vol.spans if vol.spans.size else np.zeros((1, 2))
        ctx.check(ctx.lib.vo_dvol_upload(ctx.handle, vol.nx, vol.ny, _lib.ptr(vol.off), _lib.ptr(spans), C.byref(h)))
        return cls(ctx, h, vol)

    def info(self):
        nx, ny, n = C.c_int(), C.c_int(), C.c_uint64()
        po, ps = C.c_void_p(), C.c_void_p()
        self.ctx.check(self.ctx.lib.vo_dvol_info(self.handle, C.byref(nx), C.byref(ny), C.byref(n), C.byref(po), C.byref(ps)))
        return nx.value, ny.value, int(n.value), po.value, ps.value

    def download(self, like: CompressedVolume | None = None) -> CompressedVolume:
        nx, ny, n, _, _ = self.info()
        off = np.empty(nx * ny + 1, dtype=np.uint32)
        spans = np.empty((max(n, 1), 2), dtype=np.float64)
        self.ctx.check(self.ctx.lib.vo_dvol_download(self.ctx.handle, self.handle, _lib.ptr(off), _lib.ptr(spans)))
        meta = like or self.meta
        if meta is None:
            return CompressedVolume(nx, ny, off, spans[:n])
        return meta.like(nx, ny, off, spans[:n])

    def rows(self, y0: int, y1: int) -> "DeviceVolume":
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.vo_dvol_rows(self.ctx.handle, self.handle, y0, y1, C.byref(h)))
        return DeviceVolume(self.ctx, h, self.meta)

    def free(self):
        if self.handle:
            self.ctx.lib.vo_dvol_free(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def concat_rows(ctx, parts) -> DeviceVolume:
    hs = [p.handle if p is not None else None for p in parts] + [None, None, None]
    h = C.c_void_p()
    ctx.check(ctx.lib.vo_dvol_concat_rows(ctx.handle, hs[0], hs[1], hs[2], C.byref(h)))
    return DeviceVolume(ctx, h, next(p.meta for p in parts if p is not None))


class VoronoiMorpho:
    """src/vor3d/Voronoi.h:13-44."""

    method: str = ""

    def __init__(self, ctx: _lib.Context | None = None, device: int = 0):
        self.ctx = ctx or _lib.default_context(device)

    # -- host containers in, host containers out (the drop-in call) -----------------------------
    def _morph(self, op: str, input: CompressedVolume, radius: float):
        ctx = self.ctx
        poff, pspans = _lib._u32p(), _lib._f64p()
        n = C.c_uint64()
        t1, t2 = C.c_double(0), C.c_double(0)
        spans = input.spans if input.spans.size else np.zeros((1, 2))
        ctx.check(ctx.lib.vo_morph3d(ctx.handle, _lib.OPS3D[op], _lib.METHODS[self.method], input.nx, input.ny,
                                     input.zmin, input.zmax, _lib.ptr(input.off), _lib.ptr(spans), float(radius),
                                     C.byref(poff), C.byref(pspans), C.byref(n), C.byref(t1), C.byref(t2)))
        off, sp = ctx.take_host(poff, pspans, input.nx * input.ny, int(n.value))
        return input.like(input.nx, input.ny, off, sp), t1.value, t2.value

    def morph_rows(self, op: str, input: CompressedVolume, radius: float, y0: int, y1: int, row0: int, row1: int):
        """vo_morph3d_rows on the rows [y0, y1) of `input` (a y-slab with its ghost rows, cut from the host CSR without
        copying): the operator on those rows, rows [row0, row1) of that window returned. -> (result, time_1, time_2)"""
        ctx = self.ctx
        poff, pspans = _lib._u32p(), _lib._f64p()
        n = C.c_uint64()
        t1, t2 = C.c_double(0), C.c_double(0)
        spans = input.spans if input.spans.size else np.zeros((1, 2))
        off = input.off[y0 * input.nx:y1 * input.nx + 1]
        ctx.check(ctx.lib.vo_morph3d_rows(ctx.handle, _lib.OPS3D[op], _lib.METHODS[self.method], input.nx, y1 - y0,
                                          input.zmin, input.zmax, _lib.ptr(off), _lib.ptr(spans), float(radius), row0, row1,
                                          C.byref(poff), C.byref(pspans), C.byref(n), C.byref(t1), C.byref(t2)))
        o, sp = ctx.take_host(poff, pspans, input.nx * (row1 - row0), int(n.value))
        return input.like(input.nx, row1 - row0, o, sp), t1.value, t2.value

    def dilation(self, input: CompressedVolume, radius: float):
        """Voronoi.h:18. Returns (result, time_1, time_2) with the times in ms."""
        return self._morph("dilation", input, radius)

    def erosion(self, input: CompressedVolume, radius: float):
        """Voronoi.h:29 / Voronoi.cpp:8-17."""
        return self._morph("erosion", input, radius)

    def opening(self, input: CompressedVolume, radius: float):
        """offset3d.cpp:129-133, composed on the device (the intermediate never leaves HBM)."""
        return self._morph("opening", input, radius)

    def closing(self, input: CompressedVolume, radius: float):
        """offset3d.cpp:124-128."""
        return self._morph("closing", input, radius)

    def calculateXor(self, voxel_1: CompressedVolume, voxel_2: CompressedVolume):
        """Voronoi.cpp:91-111. Returns (volume, result)."""
        ctx = self.ctx
        if voxel_1.gridSize() != voxel_2.gridSize():
            raise ValueError("calculateXor assumes the two voxels have the same grid size")
        poff, pspans = _lib._u32p(), _lib._f64p()
        n = C.c_uint64()
        vol = C.c_double(0)
        sa = voxel_1.spans if voxel_1.spans.size else np.zeros((1, 2))
        sb = voxel_2.spans if voxel_2.spans.size else np.zeros((1, 2))
        ctx.check(ctx.lib.vo_xor3d(ctx.handle, voxel_1.nx, voxel_1.ny, voxel_1.zmin, voxel_1.zmax, voxel_1.spacing,
                                   _lib.ptr(voxel_1.off), _lib.ptr(sa), _lib.ptr(voxel_2.off), _lib.ptr(sb),
                                   C.byref(poff), C.byref(pspans), C.byref(n), C.byref(vol)))
        off, sp = ctx.take_host(poff, pspans, voxel_1.nx * voxel_1.ny, int(n.value))
        return vol.value, voxel_1.like(voxel_1.nx, voxel_1.ny, off, sp)

    # -- resident data ----------------------------------------------------------------------------
    def morph_dev(self, op: str, input: DeviceVolume, radius: float, zmin: float = 0.0, zmax: float = 0.0):
        ctx = self.ctx
        h = C.c_void_p()
        t1, t2 = C.c_double(0), C.c_double(0)
        if input.meta is not None:
            zmin, zmax = input.meta.zmin, input.meta.zmax
        ctx.check(ctx.lib.vo_morph3d_dev(ctx.handle, _lib.OPS3D[op], _lib.METHODS[self.method], input.handle,
                                         zmin, zmax, float(radius), C.byref(h), C.byref(t1), C.byref(t2)))
        return DeviceVolume(ctx, h, input.meta), t1.value, t2.value


class VoronoiMorphoVorPower(VoronoiMorpho):
    """'ours' (src/vor3d/VoronoiVorPower.h:5-12): two separable passes."""
    method = "ours"


class VoronoiMorphoBruteForce(VoronoiMorpho):
    """'brute_force' (src/vor3d/VoronoiBruteForce.h:5-15): sphere union."""
    method = "brute_force"


def make_operator(method: str, ctx: _lib.Context | None = None, device: int = 0) -> VoronoiMorpho:
    """offset3d.cpp:104-112."""
    if method == "ours":
        return VoronoiMorphoVorPower(ctx, device)
    if method == "brute_force":
        return VoronoiMorphoBruteForce(ctx, device)
    raise ValueError(f"Invalid method: {method}")


def apply_operation(op: VoronoiMorpho, operation: str, input: CompressedVolume, radius: float):
    """The -x switch of offset3d.cpp:116-136. Returns (output, time_1, time_2)."""
    if operation == "noop":
        return input, 0.0, 0.0
    if operation in ("erosion", "dilation", "closing", "opening"):
        return getattr(op, operation)(input, radius)
    raise ValueError("Operation")
