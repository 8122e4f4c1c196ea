"""A/B of band_split (share of the SMs a band's pass-1 launch set takes) and bands on the host-buffer call."""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from voroffset_b200 import synth, _lib
ctx = _lib.Context(0)
vol = synth.torus_z(2048); R = 32.0
off_pin = torch.from_numpy(vol.off.view(np.int32)).pin_memory(); sp_pin = torch.from_numpy(vol.spans).pin_memory()
def call():
    poff, pspans, n = _lib._u32p(), _lib._f64p(), C.c_uint64()
    ctx.check(ctx.lib.vo_morph3d(ctx.handle, 0, 0, vol.nx, vol.ny, vol.zmin, vol.zmax, off_pin.data_ptr(), sp_pin.data_ptr(), R,
                                 C.byref(poff), C.byref(pspans), C.byref(n), None, None))
    ctx.lib.vo_free(C.cast(poff, C.c_void_p)); ctx.lib.vo_free(C.cast(pspans, C.c_void_p))
for rep in range(2):
    for split in ("1", "2", "3"):
        ctx.set_option("band_split", split)
        res = []
        for bands in ("6", "8", "12"):
            ctx.set_option("bands", bands)
            for _ in range(3): call()
            t = time.perf_counter()
            for _ in range(10): call()
            res.append(f"bands {bands}: {(time.perf_counter() - t) * 100:.3f}")
        print("band_split", split, "|", " | ".join(res), flush=True)
