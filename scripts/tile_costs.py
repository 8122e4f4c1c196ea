"""Development aid: per-tile cycle counts of the pass-1 tile kernel (vo_set_option tile_debug) on the C5 torus."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from voroffset_b200 import synth, morpho, _lib
n, R = 2048, 32.0
ctx = _lib.Context(0); op = morpho.make_operator("ours", ctx)
vol = synth.torus_z(n); d = morpho.DeviceVolume.upload(ctx, vol)
tiles_xw = (vol.nx + 31) // 32
dbg = torch.zeros(tiles_xw * vol.ny * 4, dtype=torch.int64, device="cuda")
for i in range(3):
    out, t1, t2 = op.morph_dev("dilation", d, R); out.free()
ctx.set_option("tile_debug", str(dbg.data_ptr()))
out, t1, t2 = op.morph_dev("dilation", d, R); out.free()
ctx.set_option("tile_debug", "0")
a = dbg.cpu().numpy().reshape(-1, 4)
cyc, ncand, ent, p1 = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
nz = cyc > 0
print("tiles", len(cyc), "processed (non-empty)", nz.sum(), "pass1 ms", t1)
print("cycles: mean %.0f median %.0f p99 %.0f max %.0f  sum %.3e" % (cyc[nz].mean(), np.median(cyc[nz]), np.percentile(cyc[nz], 99), cyc.max(), cyc.sum()))
order = np.argsort(-cyc)[:15]
for t in order:
    print("tile y=%d x0=%d cycles=%d (%.1f us; phase 1 + staging of the next tile: %d) cand=%d entries=%d" % (t // tiles_xw, (t % tiles_xw) * 32, cyc[t], cyc[t] / 1.9e3, p1[t], ncand[t], ent[t]))
h, edges = np.histogram(cyc[nz], bins=[0, 5e3, 1e4, 2e4, 4e4, 8e4, 1.6e5, 3.2e5, 1e9])
print("histogram (cycles):", list(zip(edges[:-1].astype(int), h)))
print("share of total cycles in tiles > 80k cycles: %.3f" % (cyc[cyc > 8e4].sum() / cyc.sum()))
print("phase-1 share of all cycles: %.3f; in the heavy tiles: %.3f" % (p1.sum() / cyc.sum(), p1[cyc > 8e4].sum() / cyc[cyc > 8e4].sum()))
