"""A/B of the CTAs per SM of the pass-1 tile kernel: resident dilation / erosion times and the host-buffer call."""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from voroffset_b200 import synth, morpho, _lib
ctx = _lib.Context(0); op = morpho.make_operator("ours", ctx)
vol = synth.torus_z(2048); R = 32.0
d = morpho.DeviceVolume.upload(ctx, vol)
volp = synth.torus_z(2048, padding=34); dp = morpho.DeviceVolume.upload(ctx, volp)
off_pin = torch.from_numpy(vol.off.view(np.int32)).pin_memory(); sp_pin = torch.from_numpy(vol.spans).pin_memory()
def call():
    poff, pspans, n = _lib._u32p(), _lib._f64p(), C.c_uint64()
    ctx.check(ctx.lib.vo_morph3d(ctx.handle, 0, 0, vol.nx, vol.ny, vol.zmin, vol.zmax, off_pin.data_ptr(), sp_pin.data_ptr(), R,
                                 C.byref(poff), C.byref(pspans), C.byref(n), None, None))
    ctx.lib.vo_free(C.cast(poff, C.c_void_p)); ctx.lib.vo_free(C.cast(pspans, C.c_void_p))
for ctas in (sys.argv[1:] or ["1", "2", "4", "1", "2", "4"]):
    ctx.set_option("tile_ctas", ctas)
    res = []
    for opn, dv in (("dilation", d), ("erosion", dp)):
        ts, k1 = [], []
        for i in range(12):
            ctx.mark(0); out, t1, t2 = op.morph_dev(opn, dv, R); ctx.mark(1); out.free()
            ts.append(ctx.elapsed_ms(0, 1)); k1.append(ctx.last_profile()[0])
        res.append(f"{opn} {np.median(ts[3:]):.4f} (tile kernel {np.median(k1[3:]):.4f})")
    for bands in ("6", "8", "12"):
        ctx.set_option("bands", bands)
        for _ in range(3): call()
        t = time.perf_counter()
        for _ in range(10): call()
        res.append(f"e2e[{bands}] {(time.perf_counter() - t) * 100:.3f}")
    print("tile_ctas", ctas, "|", " | ".join(res), flush=True)
