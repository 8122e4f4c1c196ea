"""Multi-GPU dilation: one process per GPU, the grid cut into y-slabs, one neighbour halo exchange.

Why y-slabs: on the device the first pass gathers along x inside each row (slab-local) and the second
pass unions radius classes along y, reaching at most J = floor(R) rows (the reference's locality:
Voronoi2D.cpp:651,658 / SeparatePower2D.cpp:241,266 - a seed influences at most floor(R) lines). With the
x-fastest layout (CompressedVolume.h:28-29) the J boundary rows of a slab are one contiguous CSR range,
so the halo is a plain (offsets, spans) pair: no packing kernel, point-to-point messages to the two
neighbours over torch.distributed - NCCL/NVLink on GPUs, gloo in the CPU tests. This is the reference's
dormant TBB decomposition (VoronoiVorPower.cpp:41-63,70-92: tasks own disjoint slices) stretched across
devices; SURVEY.md 8(e) describes the same scheme for x-slabs, the axis differs only because our pass
order is x then y.

The halo carries INPUT rows (exchanged before pass 1; each rank then runs pass 1 on its J halo rows too).
That costs 2J/ny_local extra pass-1 work and moves ~J*nx*(4 + 16 k_in) bytes per neighbour instead of
the much larger mid rows.

Message protocol (per neighbour and direction). The number of intervals in a halo is data dependent, the
receive buffer must be posted with a size. First call: the two sides swap their interval counts, then the
payload. Afterwards both sides remember a capacity for each direction (1.5x the last count, a pure
function of what was sent, so sender and receiver always agree) and a step is ONE batch of messages:
offsets + [count, overflow flag] in a fixed-size int32 tensor, spans padded to the agreed capacity. If a
halo ever outgrows its capacity the flag is set and the exact-size payload follows in a second message.

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


def _grow(n: int) -> int:
    """Capacity both sides derive from a transmitted count."""
    return max(1024, int(n * 1.5) + 16)


class CudaSlabBackend:
    """Product backend: volumes are `DeviceVolume`s, halos travel as CUDA tensors."""

    def __init__(self, ctx: _lib.Context):
        self.ctx = ctx
        self.device = torch.device("cuda", ctx.device)

    def shape(self, vol: DeviceVolume):
        nx, ny, n, _, _ = vol.info()
        return nx, ny

    def new_tensors(self, n_off: int, n_spans: int):
        return (torch.empty(n_off, dtype=torch.int32, device=self.device),
                torch.empty(max(2 * n_spans, 2), dtype=torch.float64, device=self.device))

    def rows_into(self, vol: DeviceVolume, y0: int, y1: int, off: torch.Tensor, spans: torch.Tensor) -> int:
        """Rows [y0, y1) into (off[: (y1-y0)*nx+1], spans); returns the interval count (spans untouched if too small)."""
        n = C.c_uint64(0)
        self.ctx.check(self.ctx.lib.vo_dvol_rows_to(self.ctx.handle, vol.handle, y0, y1, off.data_ptr(), spans.data_ptr(),
                                                    spans.numel() // 2, C.byref(n)))
        return int(n.value)

    def from_tensors(self, nx: int, ny: int, off: torch.Tensor, spans: torch.Tensor, n_spans: int, like: DeviceVolume):
        torch.cuda.current_stream(self.device).synchronize()
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

    # -- overlapped form: pass 1 of the halo-independent rows runs while the halos travel ----------------
    def slab_begin(self, vol: DeviceVolume, radius: float, prev_link, next_link):
        """Packs the two outgoing halos into the links' send buffers (on the device) and starts pass 1 of the rows
        that need no halo. Returns an opaque handle, or None when the library declines (small grids, radius < 1, ...)."""
        h = C.c_void_p()
        send = []
        for lk in (prev_link, next_link):
            send += [None, None, 0] if lk is None else [lk.off_out.data_ptr(), lk.sp_out.data_ptr(), int(lk.cap_out)]
        rc = self.ctx.lib.vo_slab_begin(self.ctx.handle, vol.handle, float(radius), int(prev_link is not None), int(next_link is not None),
                                        int(prev_link.cap_in) if prev_link else 0, int(next_link.cap_in) if next_link else 0,
                                        *send, torch.cuda.current_stream(self.device).cuda_stream, C.byref(h))
        if rc == 1:                                   # VO_ERR_ARG: not a case for the overlapped path
            return None
        self.ctx.check(rc)
        return h

    def slab_finish(self, slab, prev, nxt, like: DeviceVolume):
        """prev / nxt: (offsets tensor, spans tensor, interval count) or None. Returns (result, ms1, ms2)."""
        torch.cuda.current_stream(self.device).synchronize()
        args = []
        for part in (prev, nxt):
            args += [None, None, 0] if part is None else [part[0].data_ptr(), part[1].data_ptr(), int(part[2])]
        out = C.c_void_p()
        ms1, ms2 = C.c_double(0), C.c_double(0)
        self.ctx.check(self.ctx.lib.vo_slab_finish(self.ctx.handle, slab, *args, C.byref(out), C.byref(ms1), C.byref(ms2)))
        return DeviceVolume(self.ctx, out, like.meta), ms1.value, ms2.value

    def slab_abort(self, slab):
        self.ctx.lib.vo_slab_abort(self.ctx.handle, slab)


class _Link:
    """Per-neighbour message buffers and the capacities agreed so far."""

    def __init__(self):
        self.cap_out = None      # capacity (intervals) the neighbour expects from me
        self.cap_in = None       # capacity I expect from the neighbour
        self.off_out = self.sp_out = self.off_in = self.sp_in = None


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
        self.last_messages = 0
        self._links = {}
        self._shape = None

    # -- messaging --------------------------------------------------------------------------------
    def _batch(self, ops):
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            self.last_messages += 1

    def _exchange(self, vol, nx: int, ny: int, J: int):
        """Send my first J rows to rank-1 and my last J rows to rank+1; receive theirs."""
        be = self.backend
        if self._shape != (nx, J):               # geometry changed: forget the agreed capacities
            self._links, self._shape = {}, (nx, J)
        L = J * nx + 1                           # offsets of a halo
        nbrs = [(r, rows) for r, rows in ((self.rank - 1, (0, J)), (self.rank + 1, (ny - J, ny))) if 0 <= r < self.world]
        links = {r: self._links.setdefault(r, _Link()) for r, _ in nbrs}
        self.last_messages, nbytes = 0, 0

        # fill the send buffers (offsets + header [count, overflow]; spans if they fit the agreed capacity)
        n_out = {}
        for r, (y0, y1) in nbrs:
            lk = links[r]
            if lk.off_out is None or lk.sp_out.numel() // 2 < (lk.cap_out or 1024):
                lk.off_out, lk.sp_out = be.new_tensors(L + 2, lk.cap_out or 1024)
            n = be.rows_into(vol, y0, y1, lk.off_out, lk.sp_out)
            if lk.cap_out is not None and n > lk.cap_out or lk.cap_out is None and n > lk.sp_out.numel() // 2:
                # does not fit: exact-size payload (second message / first call)
                _, exact = be.new_tensors(1, n)
                n = be.rows_into(vol, y0, y1, lk.off_out, exact)
                lk.exact_out = exact
            else:
                lk.exact_out = None
            n_out[r] = n

        first = [r for r, _ in nbrs if links[r].cap_out is None]
        if first:
            # first call on this link: swap the counts, then exact-size payloads
            dev = links[first[0]].off_out.device
            c_out = {r: torch.tensor([n_out[r]], dtype=torch.int64, device=dev) for r in first}
            c_in = {r: torch.zeros(1, dtype=torch.int64, device=dev) for r in first}
            ops = []
            for r in first:
                ops += [dist.P2POp(dist.isend, c_out[r], r, self.group), dist.P2POp(dist.irecv, c_in[r], r, self.group)]
            self._batch(ops)
            for r in first:
                lk = links[r]
                lk.cap_in, lk.cap_out = _grow(int(c_in[r].item())), _grow(n_out[r])
                lk.n_in_first = int(c_in[r].item())

        # main message: fixed-size offsets (+ header) and capacity-padded spans
        ops = []
        for r, _ in nbrs:
            lk = links[r]
            overflow = 1 if (lk.exact_out is not None and r not in first) else 0
            hdr = torch.tensor([n_out[r] & 0x7fffffff, overflow], dtype=torch.int32, device=lk.off_out.device)
            lk.off_out[L:L + 2] = hdr
            if lk.off_in is None or lk.sp_in.numel() // 2 < lk.cap_in:
                lk.off_in, lk.sp_in = be.new_tensors(L + 2, lk.cap_in)
            ops += [dist.P2POp(dist.isend, lk.off_out, r, self.group), dist.P2POp(dist.irecv, lk.off_in, r, self.group)]
            if r in first:
                # exact sizes known from the count swap
                src = lk.exact_out if lk.exact_out is not None else lk.sp_out
                if n_out[r]:
                    ops.append(dist.P2POp(dist.isend, src[: 2 * n_out[r]], r, self.group))
                if lk.n_in_first:
                    ops.append(dist.P2POp(dist.irecv, lk.sp_in[: 2 * lk.n_in_first], r, self.group))
                nbytes += 2 * n_out[r] * 8
            else:
                ops += [dist.P2POp(dist.isend, lk.sp_out[: 2 * lk.cap_out], r, self.group),
                        dist.P2POp(dist.irecv, lk.sp_in[: 2 * lk.cap_in], r, self.group)]
                nbytes += 2 * lk.cap_out * 8
            nbytes += (L + 2) * 4
        self._batch(ops)

        # headers, overflow follow-ups, new capacities
        n_in, ops, exact_in = {}, [], {}
        for r, _ in nbrs:
            lk = links[r]
            hdr = lk.off_in[L:L + 2].tolist()
            n_in[r] = int(hdr[0])
            if r not in first:
                if hdr[1]:                                      # neighbour's halo outgrew the capacity
                    _, exact_in[r] = be.new_tensors(1, n_in[r])
                    ops.append(dist.P2POp(dist.irecv, exact_in[r][: 2 * n_in[r]], r, self.group))
                if lk.exact_out is not None:
                    ops.append(dist.P2POp(dist.isend, lk.exact_out[: 2 * n_out[r]], r, self.group))
                    nbytes += 2 * n_out[r] * 8
        self._batch(ops)
        halos = {}
        for r, _ in nbrs:
            lk = links[r]
            spans_in = exact_in.get(r, lk.sp_in)
            halos[r] = be.from_tensors(nx, J, lk.off_in, spans_in, n_in[r], like=vol)
            lk.cap_in = max(lk.cap_in, _grow(n_in[r]))
            lk.cap_out = max(lk.cap_out, _grow(n_out[r]))
            lk.exact_out = None
        self.last_halo_bytes = nbytes
        return halos.get(self.rank - 1), halos.get(self.rank + 1)

    def _dilate_overlapped(self, vol, radius: float, nx: int, ny: int, J: int):
        """Steady-state step on a backend with slab_begin / slab_finish: the outgoing halos are packed on the device,
        the single message batch travels while pass 1 of the rows that do not depend on a halo runs. Returns None
        when this step has to take the plain path from the start (first call on a link, library declines)."""
        be = self.backend
        if self._shape != (nx, J):
            return None
        L = J * nx + 1
        nbrs = [(r, rows) for r, rows in ((self.rank - 1, (0, J)), (self.rank + 1, (ny - J, ny))) if 0 <= r < self.world]
        links = {r: self._links.get(r) for r, _ in nbrs}
        for lk in links.values():
            if lk is None or lk.cap_out is None or lk.cap_in is None or lk.off_in is None or lk.off_out is None:
                return None
            if lk.sp_out.numel() // 2 < lk.cap_out:
                lk.off_out, lk.sp_out = be.new_tensors(L + 2, lk.cap_out)
            if lk.sp_in.numel() // 2 < lk.cap_in:
                lk.off_in, lk.sp_in = be.new_tensors(L + 2, lk.cap_in)
        slab = be.slab_begin(vol, radius, links.get(self.rank - 1), links.get(self.rank + 1))
        if slab is None:
            return None
        ops, nbytes = [], 0
        for r, _ in nbrs:
            lk = links[r]
            ops += [dist.P2POp(dist.isend, lk.off_out, r, self.group), dist.P2POp(dist.irecv, lk.off_in, r, self.group),
                    dist.P2POp(dist.isend, lk.sp_out[: 2 * lk.cap_out], r, self.group),
                    dist.P2POp(dist.irecv, lk.sp_in[: 2 * lk.cap_in], r, self.group)]
            nbytes += 2 * lk.cap_out * 8 + (L + 2) * 4
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        self.last_messages, self.last_halo_bytes = 1, nbytes
        order = [r for r, _ in nbrs]
        hdrs = torch.stack([links[r].off_in[L:L + 2] for r in order] + [links[r].off_out[L:L + 2] for r in order]).tolist()
        hdr_in = {r: hdrs[i] for i, r in enumerate(order)}
        hdr_out = {r: hdrs[len(order) + i] for i, r in enumerate(order)}
        overflow = any(h[1] for h in hdr_in.values()) or any(h[1] for h in hdr_out.values())
        exact_in = {}
        if overflow:
            # a halo outgrew its agreed capacity: the exact payload follows, this step finishes on the plain path
            ops, keep = [], []
            for r, (y0, y1) in nbrs:
                if hdr_in[r][1]:
                    _, exact_in[r] = be.new_tensors(1, int(hdr_in[r][0]))
                    ops.append(dist.P2POp(dist.irecv, exact_in[r][: 2 * int(hdr_in[r][0])], r, self.group))
                if hdr_out[r][1]:
                    off_tmp, exact = be.new_tensors(L + 2, int(hdr_out[r][0]))
                    be.rows_into(vol, y0, y1, off_tmp, exact)
                    keep.append(exact)
                    ops.append(dist.P2POp(dist.isend, exact[: 2 * int(hdr_out[r][0])], r, self.group))
            self._batch(ops)
        parts = {}
        for r, _ in nbrs:
            lk = links[r]
            n_in = int(hdr_in[r][0])
            parts[r] = (lk.off_in, exact_in.get(r, lk.sp_in), n_in)
            lk.cap_in = max(lk.cap_in, _grow(n_in))
            lk.cap_out = max(lk.cap_out, _grow(int(hdr_out[r][0])))
        if not overflow:
            out, ms1, ms2 = be.slab_finish(slab, parts.get(self.rank - 1), parts.get(self.rank + 1), like=vol)
            self.last_ms = (ms1, ms2)
            return out
        be.slab_abort(slab)
        halos = {r: be.from_tensors(nx, J, p[0], p[1], p[2], like=vol) for r, p in parts.items()}
        return self._finish_plain(vol, radius, ny, J, halos.get(self.rank - 1), halos.get(self.rank + 1))

    def _finish_plain(self, vol, radius, ny, J, halo_prev, halo_next):
        be = self.backend
        ext = be.concat([halo_prev, vol, halo_next])
        y0 = J if halo_prev is not None else 0
        out, ms1, ms2 = be.dilate_rows(ext, radius, y0, y0 + ny)
        be.release(ext)
        be.release(halo_prev)
        be.release(halo_next)
        self.last_ms = (ms1, ms2)
        return out

    def dilate(self, vol, radius: float):
        be = self.backend
        nx, ny = be.shape(vol)
        J = int(math.floor(radius))
        if self.world > 1 and ny < J:
            raise ValueError(f"slab of {ny} rows is thinner than the halo ({J} rows): use fewer ranks")
        if self.world > 1 and J > 0 and hasattr(be, "slab_begin"):
            out = self._dilate_overlapped(vol, radius, nx, ny, J)
            if out is not None:
                return out
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
