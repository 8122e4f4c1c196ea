"""Floor of the host-buffer call's transfer pattern: the same call with its kernels left out (library built with
-DVO_KTRACE, VO_LIB=...; pipe_dry). Usage: e2e_dry.py [key=value ...]"""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from voroffset_b200 import synth, _lib
vol = synth.torus_z(2048); R = 32.0
off_pin = torch.from_numpy(vol.off.view(np.int32)).pin_memory(); sp_pin = torch.from_numpy(vol.spans).pin_memory()
ctx = _lib.Context(0)
for kv in sys.argv[1:]:
    k, v = kv.split("="); ctx.set_option(k, v)
def call():
    poff, pspans, n = _lib._u32p(), _lib._f64p(), C.c_uint64()
    ctx.check(ctx.lib.vo_morph3d(ctx.handle, 0, 0, vol.nx, vol.ny, vol.zmin, vol.zmax, off_pin.data_ptr(), sp_pin.data_ptr(), R,
                                 C.byref(poff), C.byref(pspans), C.byref(n), None, None))
    ctx.lib.vo_free(C.cast(poff, C.c_void_p)); ctx.lib.vo_free(C.cast(pspans, C.c_void_p))
    return n.value
def timed(tag):
    for _ in range(3): call()
    ts = []
    for _ in range(7):
        t = time.perf_counter()
        for _ in range(5): n = call()
        ts.append((time.perf_counter() - t) * 200)
    print(f"{tag:10s} {' '.join(sys.argv[1:]):40s} e2e ms min {min(ts):.3f} median {sorted(ts)[3]:.3f}  intervals {n}", flush=True)
timed("real")
ctx.set_option("pipe_dry", str(call()))
timed("dry")
