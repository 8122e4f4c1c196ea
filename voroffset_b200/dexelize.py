"""Mesh -> dexel volume on the device: the step before the morphology path.

`create_dexels` mirrors `voroffset3d::create_dexels(filename, voxel_size, padding, num_voxels)`
(src/vor3d/Dexelize.cpp:231-274) on in-memory arrays: bounding box, `voxel_size = max_extent / num_voxels`
when `num_voxels > 0` (:259-263), the grid of the `CompressedVolume` constructor (CompressedVolume.cpp:11-23),
then the ray-marching loop `compute_sign` (:166-225), which is what runs on the GPU (`vo_dexelize_dev`,
csrc/dexelize.cuh). There is no CPU implementation behind it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .morpho import DeviceVolume
from .volume import CompressedVolume


def load_obj(path: str):
    """Vertices and (fan-triangulated) faces of a Wavefront OBJ file."""
    V, F = [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                V.append([float(t[1]), float(t[2]), float(t[3])])
            elif t[0] == "f":
                idx = [int(s.split("/")[0]) for s in t[1:]]
                idx = [i - 1 if i > 0 else len(V) + i for i in idx]
                F.extend([idx[0], idx[k], idx[k + 1]] for k in range(1, len(idx) - 1))
    return np.asarray(V, dtype=np.float64).reshape(-1, 3), np.asarray(F, dtype=np.int32).reshape(-1, 3)


def save_obj(path: str, V, F) -> None:
    with open(path, "w") as f:
        for p in np.asarray(V):
            f.write(f"v {p[0]:.17g} {p[1]:.17g} {p[2]:.17g}\n")
        for t in np.asarray(F):
            f.write(f"f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n")


def grid_for(V, voxel_size: float | None, padding: int, num_voxels: int) -> CompressedVolume:
    """The empty volume `create_dexels` builds before ray marching (Dexelize.cpp:255-272)."""
    V = np.asarray(V, dtype=np.float64).reshape(-1, 3)
    if V.shape[0] == 0:
        raise ValueError("Invalid input mesh.")
    lo, hi = V.min(axis=0), V.max(axis=0)
    extent = hi - lo
    if num_voxels > 0:
        voxel_size = float(max(extent[0], max(extent[1], extent[2])) / num_voxels)
    if voxel_size is None or not voxel_size > 0:
        raise ValueError("voxel_size or num_voxels must be given")
    return CompressedVolume.from_box(tuple(lo), tuple(extent), float(voxel_size), int(padding))


def dexelize_dev(ctx: _lib.Context, V, F, grid: CompressedVolume):
    """compute_sign (Dexelize.cpp:166-225) for the columns of `grid`; returns (DeviceVolume, device ms)."""
    V = np.ascontiguousarray(V, dtype=np.float64).reshape(-1, 3)
    F = np.ascontiguousarray(F, dtype=np.int32).reshape(-1, 3)
    h = C.c_void_p()
    ms = C.c_double(0)
    ctx.check(ctx.lib.vo_dexelize_dev(ctx.handle, V.shape[0], _lib.ptr(V) if V.size else None, F.shape[0],
                                      _lib.ptr(F) if F.size else None, grid.origin[0], grid.origin[1], grid.spacing,
                                      grid.nx, grid.ny, C.byref(h), C.byref(ms)))
    return DeviceVolume(ctx, h, grid), ms.value


def create_dexels(V, F, voxel_size: float | None = None, padding: int = 0, num_voxels: int = -1,
                  ctx: _lib.Context | None = None, device: int = 0) -> CompressedVolume:
    """Dexelize.cpp:231-274 on arrays; the volume comes back to the host like the reference's return value."""
    ctx = ctx or _lib.default_context(device)
    grid = grid_for(V, voxel_size, padding, num_voxels)
    dv, _ = dexelize_dev(ctx, V, F, grid)
    out = dv.download(grid)
    dv.free()
    return out
