"""Kernel trace of one host-buffer C5 dilation (library built with -DVO_KTRACE, passed as VO_LIB). Usage: ktrace_e2e.py out.csv [key=value ...]"""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from voroffset_b200 import synth, _lib
ctx = _lib.Context(0)
vol = synth.torus_z(2048); R = 32.0
off_pin = torch.from_numpy(vol.off.view(np.int32)).pin_memory(); sp_pin = torch.from_numpy(vol.spans).pin_memory()
for kv in sys.argv[2:]:
    k, v = kv.split("="); ctx.set_option(k, v)
def call():
    poff, pspans, n = _lib._u32p(), _lib._f64p(), C.c_uint64()
    ctx.check(ctx.lib.vo_morph3d(ctx.handle, 0, 0, vol.nx, vol.ny, vol.zmin, vol.zmax, off_pin.data_ptr(), sp_pin.data_ptr(), R,
                                 C.byref(poff), C.byref(pspans), C.byref(n), None, None))
    ctx.lib.vo_free(C.cast(poff, C.c_void_p)); ctx.lib.vo_free(C.cast(pspans, C.c_void_p))
for _ in range(4): call()
ts = []
for _ in range(5):
    t = time.perf_counter(); call(); ts.append((time.perf_counter() - t) * 1e3)
print("untraced calls ms", [round(x, 3) for x in ts], flush=True)
ctx.set_option("ktrace", "400000")
t = time.perf_counter(); call(); print("traced call ms", round((time.perf_counter() - t) * 1e3, 3))
ctx.set_option("ktrace_dump", sys.argv[1])
