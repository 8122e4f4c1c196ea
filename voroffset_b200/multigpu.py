"""Multi-GPU form of the 3D operators: thin ctypes mirror of the vo_mg_* entry points (include/voroffset_b200.h,
csrc/vo_mg.cuh). The grid is cut into y-slabs, one per GPU; the floor(R) boundary rows travel with NCCL inside the
library. Two ways to build a group, like the C ABI:

    MultiGpu.single_process([0, 1, 2, 3])      one process, one host thread per GPU (offset3d --gpus N)
    MultiGpu.from_torch_distributed(device)    one process per GPU under torchrun (bench.py): rank 0's NCCL id is
                                               broadcast through the already initialised process group

There is no CPU path behind this class.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .morpho import DeviceVolume
from .volume import CompressedVolume


def _prefer_torch_nccl():
    """The library dlopens "libnccl.so.2" on first use. Inside a Python process that must be the copy PyTorch links
    against (nvidia/nccl/lib/libnccl.so.2 of the wheel): two NCCL builds under one soname cannot coexist, whichever
    is loaded first wins - so name torch's copy explicitly unless the caller has chosen one (VO_NCCL_LIB)."""
    import os
    if os.environ.get("VO_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["VO_NCCL_LIB"] = cand
                return
    except Exception:
        pass


class MultiGpu:
    def __init__(self, handle, lib):
        self.handle, self.lib = handle, lib
        self.world = int(lib.vo_mg_world(handle))
        self.local_count = int(lib.vo_mg_local_count(handle))
        self.ranks = [int(lib.vo_mg_rank(handle, i)) for i in range(self.local_count)]
        self.contexts = []
        for i in range(self.local_count):
            ctx = _lib.Context(0, _borrowed=lib.vo_mg_ctx(handle, i))
            self.contexts.append(ctx)

    # -- construction ---------------------------------------------------------------------------------
    @classmethod
    def single_process(cls, devices) -> "MultiGpu":
        _prefer_torch_nccl()
        lib = _lib.load()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = lib.vo_mg_create(devs, len(devices), C.byref(h))
        if rc != _lib.VO_OK:
            raise _lib.VoroffsetError(rc, f"vo_mg_create({list(devices)}) failed (CUDA devices / NCCL unavailable)")
        mg = cls(h, lib)
        for ctx, d in zip(mg.contexts, devices):
            ctx.device = int(d)
        return mg

    @classmethod
    def from_torch_distributed(cls, device: int, group=None) -> "MultiGpu":
        import torch
        import torch.distributed as dist
        _prefer_torch_nccl()
        lib = _lib.load()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0 and world > 1:
            buf = (C.c_uint8 * 128)()
            rc = lib.vo_mg_unique_id(buf)
            if rc != _lib.VO_OK:
                raise _lib.VoroffsetError(rc, "vo_mg_unique_id failed (libnccl.so.2 not loadable)")
            ident = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        if world > 1:
            t = ident.to(torch.device("cuda", device))
            dist.broadcast(t, 0, group=group)
            ident = t.cpu()
        raw = (C.c_uint8 * 128)(*ident.tolist())
        h = C.c_void_p()
        rc = lib.vo_mg_create_rank(int(device), rank, world, raw, C.byref(h))
        if rc != _lib.VO_OK:
            raise _lib.VoroffsetError(rc, f"vo_mg_create_rank(device={device}, rank={rank}/{world}) failed")
        mg = cls(h, lib)
        mg.contexts[0].device = int(device)
        return mg

    def close(self):
        if getattr(self, "handle", None):
            for c in self.contexts:
                c.handle = None
            self.lib.vo_mg_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != _lib.VO_OK:
            raise _lib.VoroffsetError(rc, self.lib.vo_mg_last_error(self.handle).decode(errors="replace"))

    # -- host buffers in, host buffers out (single-process groups) -------------------------------------------
    def morph(self, op: str, vol: CompressedVolume, radius: float, method: str = "ours"):
        """The drop-in call on N GPUs. Returns (result, time_1, time_2)."""
        poff, pspans, n = _lib._u32p(), _lib._f64p(), C.c_uint64()
        t1, t2 = C.c_double(0), C.c_double(0)
        spans = vol.spans if vol.spans.size else np.zeros((1, 2))
        self.check(self.lib.vo_mg_morph3d(self.handle, _lib.OPS3D[op], _lib.METHODS[method], vol.nx, vol.ny, vol.zmin, vol.zmax,
                                          _lib.ptr(vol.off), _lib.ptr(spans), float(radius), C.byref(poff), C.byref(pspans),
                                          C.byref(n), C.byref(t1), C.byref(t2)))
        off, sp = self.contexts[0].take_host(poff, pspans, vol.nx * vol.ny, int(n.value))
        return vol.like(vol.nx, vol.ny, off, sp), t1.value, t2.value

    # -- resident slabs ---------------------------------------------------------------------------------
    def morph_dev(self, op: str, slabs, radius: float, zmin: float, zmax: float, method: str = "ours"):
        """slabs[i]: DeviceVolume of local rank i (uploaded through self.contexts[i]). Collective over the group.
        Returns ([DeviceVolume per local rank], time_1, time_2)."""
        n = self.local_count
        ins = (C.c_void_p * n)(*[s.handle for s in slabs])
        outs = (C.c_void_p * n)()
        t1, t2 = C.c_double(0), C.c_double(0)
        self.check(self.lib.vo_mg_morph3d_dev(self.handle, _lib.OPS3D[op], _lib.METHODS[method], ins, float(zmin), float(zmax),
                                              float(radius), outs, C.byref(t1), C.byref(t2)))
        return [DeviceVolume(self.contexts[i], C.c_void_p(outs[i]), slabs[i].meta) for i in range(n)], t1.value, t2.value

    def stats(self, local: int = 0) -> dict:
        a, b = C.c_double(0), C.c_double(0)
        nb = C.c_uint64(0)
        m, o, p = C.c_int(0), C.c_int(0), C.c_int(0)
        self.check(self.lib.vo_mg_stats(self.handle, local, C.byref(a), C.byref(b), C.byref(nb), C.byref(m), C.byref(o), C.byref(p)))
        return {"halo_ms": a.value, "halo_wait_ms": b.value, "halo_bytes": int(nb.value), "messages": m.value,
                "overlapped": o.value, "plain": p.value}
