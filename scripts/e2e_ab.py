"""Interleaved A/B of option sets for the host-buffer C5 dilation (one context per set, the sets alternate round by
round so that drifts of the box hit all of them alike). Usage: e2e_ab.py 'k=v,k=v' 'k=v' ... ('-' = defaults)"""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from voroffset_b200 import synth, _lib
vol = synth.torus_z(2048); R = 32.0
off_pin = torch.from_numpy(vol.off.view(np.int32)).pin_memory(); sp_pin = torch.from_numpy(vol.spans).pin_memory()
specs = sys.argv[1:] or ["-"]
ctxs = []
for spec in specs:
    ctx = _lib.Context(0)
    for kv in filter(None, spec.replace("-", "").split(",")):
        k, v = kv.split("="); ctx.set_option(k, v)
    ctxs.append(ctx)
def call(ctx):
    poff, pspans, n = _lib._u32p(), _lib._f64p(), C.c_uint64()
    ctx.check(ctx.lib.vo_morph3d(ctx.handle, 0, 0, vol.nx, vol.ny, vol.zmin, vol.zmax, off_pin.data_ptr(), sp_pin.data_ptr(), R,
                                 C.byref(poff), C.byref(pspans), C.byref(n), None, None))
    ctx.lib.vo_free(C.cast(poff, C.c_void_p)); ctx.lib.vo_free(C.cast(pspans, C.c_void_p))
for ctx in ctxs:
    for _ in range(4): call(ctx)
ts = [[] for _ in specs]
for rnd in range(8):
    for i, ctx in enumerate(ctxs):
        call(ctx)
        t = time.perf_counter()
        for _ in range(5): call(ctx)
        ts[i].append((time.perf_counter() - t) * 200)
for spec, t in zip(specs, ts):
    print(f"{spec:50s} e2e ms min {min(t):.3f} median {sorted(t)[len(t) // 2]:.3f}", flush=True)
