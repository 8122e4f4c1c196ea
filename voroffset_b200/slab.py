"""Multi-GPU dilation: one process per GPU, the grid cut into y-slabs, one neighbour halo exchange.

Why y-slabs: on the device the first pass gathers along x inside each row (slab-local) and the second
pass unions radius classes along y, reaching at most J = floor(R) rows (the reference's locality:
Voronoi2D.cpp:651,658 / SeparatePower2D.cpp:241,266 - a seed influences at most floor(R) lines). With the
x-fastest layout (CompressedVolume.h:28-29) the J boundary rows of a slab are one contiguous CSR range,
so the halo is a plain (offsets, spans) pair: no packing kernel, just two point-to-point messages per
neighbour (sizes first, payload second) over torch.distributed - NCCL/NVLink on GPUs, gloo in the CPU
tests. This is the reference's dormant TBB decomposition (VoronoiVorPower.cpp:41-63,70-92: tasks own
disjoint slices) stretched across devices; SURVEY.md 8(e) describes the same scheme for x-slabs, the
axis differs only because our pass order is x then y.

The halo carries INPUT rows (exchanged before pass 1; each rank then runs pass 1 on its J halo rows too).
That costs 2J/ny_local extra pass-1 work and moves ~J*nx*(4 + 16 k_in) bytes per neighbour instead of
the (J+1)x larger mid rows.

The driver is written against a small backend interface so that the exchange / cropping logic runs in
the world_size-2 gloo tests on CPU (tests inject a checker-backed backend); the product backend below
is CUDA-only.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .morpho import DeviceVolume, concat_rows
from .volume import CompressedVolume


def slab_bounds(ny: int, world: int):
    """Rows [y0, y1) owned by each rank: as even as possible, earlier ranks take the remainder."""
    base, rem = divmod(ny, world)
    out, y = [], 0
    for r in range(world):
        h = base + (1 if r < rem else 0)
        out.append((y, y + h))
        y += h
    return out


class CudaSlabBackend:
    """Product backend: volumes are `DeviceVolume`s, halos travel as CUDA tensors."""

    def __init__(self, ctx: _lib.Context):
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device)

    def shape(self, vol: DeviceVolume):
        nx, ny, n, _, _ = vol.info()
        return nx, ny

    def rows_as_tensors(self, vol: DeviceVolume, y0: int, y1: int):
        part = vol.rows(y0, y1)
        nx, ny, n, _, _ = part.info()
        off = torch.empty(nx * ny + 1, dtype=torch.int32, device=self.device)
        spans = torch.empty(max(n, 1) * 2, dtype=torch.float64, device=self.device)
        torch.cuda.synchronize(self.device)
        # vo_dvol_download accepts device destinations (cudaMemcpyDefault) and synchronises its stream
        self.ctx.check(self.ctx.lib.vo_dvol_download(self.ctx.handle, part.handle, off.data_ptr(), spans.data_ptr()))
        part.free()
        return off, spans[: 2 * n]

    def empty_tensors(self, n_off: int, n_spans: int):
        return (torch.empty(n_off, dtype=torch.int32, device=self.device),
                torch.empty(max(2 * n_spans, 2), dtype=torch.float64, device=self.device))

    def from_tensors(self, nx: int, ny: int, off: torch.Tensor, spans: torch.Tensor, n_spans: int, like: DeviceVolume):
        torch.cuda.synchronize(self.device)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.vo_dvol_from_device(self.ctx.handle, nx, ny, off.data_ptr(), spans.data_ptr(),
                                                        n_spans, C.byref(h)))
        return DeviceVolume(self.ctx, h, like.meta)

    def concat(self, parts):
        return concat_rows(self.ctx, [p for p in parts if p is not None])

    def dilate_rows(self, vol: DeviceVolume, radius: float, y0: int, y1: int):
        ctx = self.ctx
        mid, out = C.c_void_p(), C.c_void_p()
        ms1, ms2 = C.c_double(0), C.c_double(0)
        ctx.check(ctx.lib.vo_pass1_dev(ctx.handle, vol.handle, float(radius), C.byref(mid), C.byref(ms1)))
        try:
            ctx.check(ctx.lib.vo_pass2_dev(ctx.handle, mid, y0, y1, C.byref(out), C.byref(ms2)))
        finally:
            ctx.lib.vo_dmid_free(ctx.handle, mid)
        return DeviceVolume(ctx, out, vol.meta), ms1.value, ms2.value

    def release(self, vol):
        if vol is not None:
            vol.free()


class SlabDilation:
    """Dilation of a y-slab-sharded volume. Every rank calls `dilate` with its own rows; the result is
    the rank's rows of the global dilation (bit-identical to the single-GPU result)."""

    def __init__(self, backend, rank: Optional[int] = None, world: Optional[int] = None, group=None):
        self.backend = backend
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.last_ms = (0.0, 0.0)
        self.last_halo_bytes = 0

    def _exchange(self, vol, nx: int, ny: int, J: int):
        """Send my first J rows to rank-1 and my last J rows to rank+1; receive theirs."""
        be = self.backend
        prev_r = self.rank - 1 if self.rank > 0 else None
        next_r = self.rank + 1 if self.rank < self.world - 1 else None
        send = {}
        if prev_r is not None:
            send[prev_r] = be.rows_as_tensors(vol, 0, J)
        if next_r is not None:
            send[next_r] = be.rows_as_tensors(vol, ny - J, ny)
        # 1st message: span counts
        dev = next(iter(send.values()))[0].device if send else None
        counts_out = {r: torch.tensor([t[1].numel() // 2], dtype=torch.int64, device=dev) for r, t in send.items()}
        counts_in = {r: torch.zeros(1, dtype=torch.int64, device=dev) for r in send}
        ops = []
        for r in send:
            ops.append(dist.P2POp(dist.isend, counts_out[r], r, self.group))
            ops.append(dist.P2POp(dist.irecv, counts_in[r], r, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        # 2nd message: offsets + spans
        recv = {}
        ops = []
        nbytes = 0
        for r in send:
            n_in = int(counts_in[r].item())
            off_in, sp_in = be.empty_tensors(J * nx + 1, n_in)
            recv[r] = (off_in, sp_in, n_in)
            off_out, sp_out = send[r]
            ops.append(dist.P2POp(dist.isend, off_out, r, self.group))
            ops.append(dist.P2POp(dist.irecv, off_in, r, self.group))
            if sp_out.numel():
                ops.append(dist.P2POp(dist.isend, sp_out, r, self.group))
            if n_in:
                ops.append(dist.P2POp(dist.irecv, sp_in[: 2 * n_in], r, self.group))
            nbytes += off_out.numel() * 4 + sp_out.numel() * 8
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        self.last_halo_bytes = nbytes
        halo_prev = be.from_tensors(nx, J, *recv[prev_r], like=vol) if prev_r is not None else None
        halo_next = be.from_tensors(nx, J, *recv[next_r], like=vol) if next_r is not None else None
        return halo_prev, halo_next

    def dilate(self, vol, radius: float):
        be = self.backend
        nx, ny = be.shape(vol)
        J = int(math.floor(radius))
        if self.world > 1 and ny < J:
            raise ValueError(f"slab of {ny} rows is thinner than the halo ({J} rows): use fewer ranks")
        halo_prev = halo_next = None
        if self.world > 1 and J > 0:
            halo_prev, halo_next = self._exchange(vol, nx, ny, J)
        if halo_prev is None and halo_next is None:
            out, ms1, ms2 = be.dilate_rows(vol, radius, 0, ny)
        else:
            ext = be.concat([halo_prev, vol, halo_next])
            y0 = J if halo_prev is not None else 0
            out, ms1, ms2 = be.dilate_rows(ext, radius, y0, y0 + ny)
            be.release(ext)
            be.release(halo_prev)
            be.release(halo_next)
        self.last_ms = (ms1, ms2)
        return out


def shard_rows(vol: CompressedVolume, rank: int, world: int) -> CompressedVolume:
    """Host-side helper: the rows of `vol` owned by `rank`."""
    y0, y1 = slab_bounds(vol.ny, world)[rank]
    c0, c1 = y0 * vol.nx, y1 * vol.nx
    off = vol.off[c0:c1 + 1].astype(np.int64)
    return vol.like(vol.nx, y1 - y0, (off - off[0]).astype(np.uint32), vol.spans[off[0]:off[-1]])
