"""Two host-buffer dilations of the C5 torus (for ncu launch lists of the banded pipeline). Usage: e2e_once.py [bands] [calls]"""
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np, torch
from voroffset_b200 import synth, _lib
ctx = _lib.Context(0)
vol = synth.torus_z(2048); R = 32.0
off_pin = torch.from_numpy(vol.off.view(np.int32)).pin_memory(); sp_pin = torch.from_numpy(vol.spans).pin_memory()
ctx.set_option("bands", sys.argv[1] if len(sys.argv) > 1 else "8")
for kv in sys.argv[3:]:            # further options as key=value (band_free=32 band_split=1 ...)
    k, v = kv.split("=")
    ctx.set_option(k, v)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    poff, pspans, n = _lib._u32p(), _lib._f64p(), C.c_uint64()
    ctx.check(ctx.lib.vo_morph3d(ctx.handle, 0, 0, vol.nx, vol.ny, vol.zmin, vol.zmax, off_pin.data_ptr(), sp_pin.data_ptr(), R,
                                 C.byref(poff), C.byref(pspans), C.byref(n), None, None))
    ctx.lib.vo_free(C.cast(poff, C.c_void_p)); ctx.lib.vo_free(C.cast(pspans, C.c_void_p))
print("ok", n.value)
