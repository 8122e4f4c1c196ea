"""e2e time of the host-buffer dilation (vo_morph3d, pinned buffers) against the number of pipeline bands."""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from voroffset_b200 import synth, morpho, _lib
ctx = _lib.Context(0)
vol = synth.torus_z(2048); R = 32.0
off_pin = torch.from_numpy(vol.off.view(np.int32)).pin_memory(); sp_pin = torch.from_numpy(vol.spans).pin_memory()
def call():
    poff, pspans, n = _lib._u32p(), _lib._f64p(), C.c_uint64()
    ctx.check(ctx.lib.vo_morph3d(ctx.handle, 0, 0, vol.nx, vol.ny, vol.zmin, vol.zmax, off_pin.data_ptr(), sp_pin.data_ptr(), R,
                                 C.byref(poff), C.byref(pspans), C.byref(n), None, None))
    ctx.lib.vo_free(C.cast(poff, C.c_void_p)); ctx.lib.vo_free(C.cast(pspans, C.c_void_p))
for bands in (sys.argv[1:] or ["3", "4", "6", "8", "12"]):
    ctx.set_option("bands", bands)
    for _ in range(3): call()
    t = time.perf_counter()
    for _ in range(10): call()
    print("bands", bands, "e2e ms", round((time.perf_counter() - t) * 100, 3), flush=True)
ctx.set_option("pipeline", "off")
for _ in range(3): call()
t = time.perf_counter()
for _ in range(10): call()
print("plain (upload, passes, download in sequence) e2e ms", round((time.perf_counter() - t) * 100, 3))
