"""Host-side mirror of the reference's dexel containers, held as flat CSR.

``CompressedVolume`` mirrors ``voroffset3d::CompressedVolume`` (src/vor3d/CompressedVolume.h:10-41,
CompressedVolume.cpp:11-152, base class CompressedVolumeBase.h:14-49): column (x, y) is list number
``x + nx*y`` (CompressedVolume.h:28-29) and holds ascending disjoint z-intervals in dexel units
(world z / spacing, Dexelize.cpp:204). Instead of ``vector<vector<double>>`` the data is CSR:
``off`` (uint32, nx*ny+1) and ``spans`` (float64, (M, 2)) - exactly the layout the C-ABI takes
(include/voroffset_b200.h) and the layout the kernels read from HBM.

``DexelImage`` mirrors ``voroffset::DoubleCompressedImage`` storage (src/vor2d/DoubleCompressedImage.h:
17-27): ``rows`` lists along the sweep axis, each with intervals inside [0, width].
"""
from __future__ import annotations

import dataclasses
import io
import math
from typing import Iterable, Sequence

import numpy as np

OFF_DTYPE = np.uint32


def _as_off(off) -> np.ndarray:
    off = np.ascontiguousarray(off)
    if off.dtype != OFF_DTYPE:
        if off.size and int(off.max()) >= 2 ** 32:
            raise OverflowError("interval count does not fit the uint32 CSR offsets")
        off = off.astype(OFF_DTYPE)
    return off


def _as_spans(spans) -> np.ndarray:
    spans = np.ascontiguousarray(spans, dtype=np.float64)
    return spans.reshape(-1, 2)


def csr_from_lists(lists: Iterable[Sequence[float]]):
    """Build (off, spans) from per-list flat event sequences [z1, z2, z1', z2', ...]."""
    counts = []
    chunks = []
    for ev in lists:
        ev = np.asarray(ev, dtype=np.float64).reshape(-1)
        if ev.size % 2:
            raise ValueError("odd number of events in a dexel list")
        counts.append(ev.size // 2)
        chunks.append(ev)
    off = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    spans = np.concatenate(chunks) if chunks else np.zeros(0)
    return _as_off(off), _as_spans(spans)


@dataclasses.dataclass
class CompressedVolume:
    nx: int
    ny: int
    off: np.ndarray
    spans: np.ndarray
    origin: tuple = (0.0, 0.0, 0.0)
    extent: tuple = (0.0, 0.0, 0.0)
    spacing: float = 1.0
    padding: int = 0

    def __post_init__(self):
        self.off = _as_off(self.off)
        self.spans = _as_spans(self.spans)
        if self.off.shape[0] != self.nx * self.ny + 1:
            raise ValueError("off must have nx*ny+1 entries")
        if int(self.off[-1]) != self.spans.shape[0]:
            raise ValueError("off[-1] must equal the number of spans")
        self.origin = tuple(float(v) for v in self.origin)
        self.extent = tuple(float(v) for v in self.extent)

    # -- construction ------------------------------------------------------------------------
    @classmethod
    def from_box(cls, origin, extent, voxel_size, padding):
        """Empty volume with the reference constructor's grid (CompressedVolume.cpp:11-23)."""
        origin = tuple(float(o) - padding * voxel_size for o in origin)
        nx = int(math.ceil(extent[0] / voxel_size) + 2 * padding)
        ny = int(math.ceil(extent[1] / voxel_size) + 2 * padding)
        return cls(nx, ny, np.zeros(nx * ny + 1, OFF_DTYPE), np.zeros((0, 2)), origin, extent, voxel_size, padding)

    @classmethod
    def from_lists(cls, nx, ny, lists, **meta):
        off, spans = csr_from_lists(lists)
        return cls(nx, ny, off, spans, **meta)

    def like(self, nx, ny, off, spans) -> "CompressedVolume":
        """`result.reset(origin, extent, spacing, padding, nx, ny)` (CompressedVolumeBase.cpp:13-21)."""
        return CompressedVolume(nx, ny, off, spans, self.origin, self.extent, self.spacing, self.padding)

    # -- reference accessors -----------------------------------------------------------------
    def gridSize(self):
        return (self.nx, self.ny)

    def numDexels(self):
        return self.nx * self.ny

    def at(self, x: int, y: int) -> np.ndarray:
        c = x + self.nx * y
        return self.spans[int(self.off[c]):int(self.off[c + 1])].reshape(-1)

    def numSegments(self) -> int:
        return int(self.off[-1])

    def get_volume(self) -> float:
        """CompressedVolume.cpp:61-73 (value only; the summation order is numpy's)."""
        return float(self.spacing ** 3 * np.sum(self.spans[:, 1] - self.spans[:, 0]))

    @property
    def zmin(self) -> float:
        """VoronoiVorPower.cpp:28 / Voronoi.cpp:10."""
        return self.origin[2] / self.spacing

    @property
    def zmax(self) -> float:
        """VoronoiVorPower.cpp:29 / Voronoi.cpp:11 (same association order)."""
        return self.origin[2] / self.spacing + 2 * self.padding + self.extent[2] / self.spacing

    def counts(self) -> np.ndarray:
        return np.diff(self.off.astype(np.int64))

    # -- text format of CompressedVolume::save / load (CompressedVolume.cpp:116-152) ----------
    def save(self, out) -> None:
        w = out.write
        w(f"{self.origin[0]:.17g} {self.origin[1]:.17g} {self.origin[2]:.17g}\n")
        w(f"{self.extent[0]:.17g} {self.extent[1]:.17g} {self.extent[2]:.17g}\n")
        w(f"{self.nx} {self.ny}\n{self.padding}\n{self.spacing:.17g}\n")
        flat = self.spans.reshape(-1)
        for c in range(self.nx * self.ny):
            row = flat[2 * int(self.off[c]):2 * int(self.off[c + 1])]
            w(str(row.size) + "".join(f" {v:.17g}" for v in row) + "\n")

    @classmethod
    def load(cls, inp) -> "CompressedVolume":
        tok = iter(inp.read().split())
        nxt = lambda: next(tok)
        origin = tuple(float(nxt()) for _ in range(3))
        extent = tuple(float(nxt()) for _ in range(3))
        nx, ny = int(nxt()), int(nxt())
        padding = int(nxt())
        spacing = float(nxt())
        lists = []
        for _ in range(nx * ny):
            n = int(nxt())
            lists.append([float(nxt()) for _ in range(n)])
        off, spans = csr_from_lists(lists)
        return cls(nx, ny, off, spans, origin, extent, spacing, padding)

    def dumps(self) -> str:
        s = io.StringIO()
        self.save(s)
        return s.getvalue()

    def same_topology(self, other: "CompressedVolume") -> bool:
        return self.nx == other.nx and self.ny == other.ny and np.array_equal(self.off, other.off)

    def bit_equal(self, other: "CompressedVolume") -> bool:
        return self.same_topology(other) and np.array_equal(
            self.spans.view(np.uint64), other.spans.view(np.uint64))


@dataclasses.dataclass
class DexelImage:
    """CSR mirror of DoubleCompressedImage: `rows` lists (sweep axis), intervals within [0, width]."""
    rows: int
    width: int
    off: np.ndarray
    spans: np.ndarray

    def __post_init__(self):
        self.off = _as_off(self.off)
        self.spans = _as_spans(self.spans)
        if self.off.shape[0] != self.rows + 1:
            raise ValueError("off must have rows+1 entries")
        if int(self.off[-1]) != self.spans.shape[0]:
            raise ValueError("off[-1] must equal the number of spans")

    @classmethod
    def from_lists(cls, width, lists):
        lists = list(lists)
        off, spans = csr_from_lists(lists)
        return cls(len(lists), width, off, spans)

    def height(self):
        return self.rows

    def at(self, i: int) -> np.ndarray:
        return self.spans[int(self.off[i]):int(self.off[i + 1])].reshape(-1)

    def numSegments(self) -> int:
        return int(self.off[-1])

    def isValid(self) -> bool:
        """DoubleCompressedImage.cpp:197-223: per row the events are non-decreasing starting from
        -1, and the last event is <= width."""
        flat = self.spans.reshape(-1)
        if flat.size == 0:
            return True
        if flat.min() < -1 or np.any(np.isnan(flat)):
            return False
        if np.any(self.spans[:, 1] < self.spans[:, 0]):
            return False
        row_of = np.repeat(np.arange(self.rows), np.diff(self.off.astype(np.int64)))
        same = row_of[1:] == row_of[:-1]
        if np.any(self.spans[1:, 0][same] < self.spans[:-1, 1][same]):
            return False
        last = np.diff(self.off.astype(np.int64)) > 0
        ends = self.spans[self.off[1:][last].astype(np.int64) - 1, 1]
        return bool(np.all(ends <= self.width))

    def bit_equal(self, other: "DexelImage") -> bool:
        return (self.rows == other.rows and self.width == other.width
                and np.array_equal(self.off, other.off)
                and np.array_equal(self.spans.view(np.uint64), other.spans.view(np.uint64)))
