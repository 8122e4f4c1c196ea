"""Single-GPU timing of the overlapped slab step (vo_slab_begin / vo_slab_finish with halos taken from the same
device) against the plain call on the concatenated rows: what the split itself costs, without NCCL."""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from voroffset_b200 import synth, morpho, _lib
ctx = _lib.Context(0); dev = torch.device("cuda", 0)
n, R, J = 2048, 32.0, 32
big = synth.torus_z(n)                               # rows [0, n): "own"; halo = rows of a second copy
d = morpho.DeviceVolume.upload(ctx, big)
own = d
halo = d.rows(0, J)                                  # pretend the next neighbour's first J rows
hn = halo.info()[2]
off = torch.empty(J * n + 1, dtype=torch.int32, device=dev); sp = torch.empty(2 * (hn + 10), dtype=torch.float64, device=dev)
cnt = C.c_uint64(0)
ctx.check(ctx.lib.vo_dvol_rows_to(ctx.handle, d.handle, 0, J, off.data_ptr(), sp.data_ptr(), hn + 10, C.byref(cnt)))
ext = morpho.concat_rows(ctx, [own, halo])
op = morpho.make_operator("ours", ctx)
for it in range(5):
    ctx.mark(0)
    mid, out = C.c_void_p(), C.c_void_p()
    ctx.check(ctx.lib.vo_pass1_dev(ctx.handle, ext.handle, R, C.byref(mid), None))
    ctx.check(ctx.lib.vo_pass2_dev(ctx.handle, mid, 0, n, C.byref(out), None))
    ctx.lib.vo_dmid_free(ctx.handle, mid)
    ctx.mark(1)
    a = morpho.DeviceVolume(ctx, out, big)
    t_plain = ctx.elapsed_ms(0, 1)
    ctx.mark(2)
    slab, out2 = C.c_void_p(), C.c_void_p()
    ms1, ms2 = C.c_double(0), C.c_double(0)
    assert ctx.lib.vo_slab_begin(ctx.handle, own.handle, R, 0, 1, 0, hn + 10, None, None, 0, None, None, 0, None, C.byref(slab)) == 0
    ctx.check(ctx.lib.vo_slab_finish(ctx.handle, slab, None, None, 0, off.data_ptr(), sp.data_ptr(), hn, C.byref(out2), C.byref(ms1), C.byref(ms2)))
    ctx.mark(3)
    b = morpho.DeviceVolume(ctx, out2, big)
    print(it, "plain (concatenated rows)", round(t_plain, 3), "ms | slab begin+finish", round(ctx.elapsed_ms(2, 3), 3), "ms (pass1", round(ms1.value, 3), "pass2", round(ms2.value, 3), ")",
          "same:", a.download().bit_equal(b.download()) if it == 0 else "-", flush=True)
    a.free(); b.free()
