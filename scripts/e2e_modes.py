"""e2e time of the host-buffer dilation against the pipeline's knobs: bands x tile_ctas x band_split."""
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
for free in ("0", "8", "16", "24", "32", "48"):
    for split in ("1", "2"):
        ctx.set_option("band_free", free); ctx.set_option("band_split", split)
        for _ in range(3): call()
        ts = []
        for _ in range(4):
            t = time.perf_counter()
            for _ in range(5): call()
            ts.append((time.perf_counter() - t) * 200)
        print("band_free", free, "band_split", split, "e2e ms", round(min(ts), 3), "median", round(sorted(ts)[len(ts) // 2], 3), flush=True)
