"""Timeline of a kernel trace (build with -DVO_KTRACE, vo_set_option("ktrace_dump", path)): per launch (records of one
kernel id clustered in time) first start / last end / CTAs / SMs used, the copy marks, and the share of SM time the tile
kernel's warps hold. Usage: ktrace_view.py trace.csv [gap_us]"""
import sys
import numpy as np
NAMES = {1: "thresh", 2: "order_count", 3: "order_place", 4: "tile", 5: "tile_list", 6: "pass1_redo", 7: "pass2_rows", 8: "pass2_redo",
         9: "scan_compact", 10: "mark", 11: "copy_out"}
a = np.loadtxt(sys.argv[1], delimiter=",", skiprows=1, dtype=np.int64, ndmin=2)
gap = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
t_base = a[:, 4].min()
a[:, 4] -= t_base
a[:, 5] = np.where(a[:, 5] > 0, a[:, 5] - t_base, a[:, 4])
ev = []
for kid in np.unique(a[:, 0]):
    r = a[a[:, 0] == kid]
    if kid == 10:
        for row in r:
            aux = int(row[2]); kind = {1: "upload done", 2: "download begin", 3: "download end"}.get(aux // 1000, "mark")
            ev.append((row[4] / 1e3, row[4] / 1e3, f"{kind} {aux % 1000}", 1, 1, 0.0))
        continue
    # a launch = the records that share `aux` (band row / first tile) for the kernels that carry one, else clustered by start time
    keys = np.unique(r[:, 2]) if kid in (1, 2, 3, 4, 5, 7, 8) else [None]
    for k in keys:
        rr = r if k is None else r[r[:, 2] == k]
        rr = rr[np.argsort(rr[:, 4])]
        cuts = np.where(np.diff(rr[:, 4]) > 50e3)[0] + 1 if k is None else []
        for part in np.split(rr, cuts):
            if len(part) == 0: continue
            busy = (part[:, 5] - part[:, 4]).sum() / 1e3
            ev.append((part[:, 4].min() / 1e3, part[:, 5].max() / 1e3, f"{NAMES.get(int(kid), kid)} aux={k}", len(part), len(np.unique(part[:, 1])), busy))
print(f"{'start us':>9} {'end us':>9} {'dur':>8}  {'what':28s} {'recs':>6} {'SMs':>4} {'sum of record times us':>12}")
for s, e, n, c, sm, busy in sorted(ev):
    if e - s < gap and c > 1 and False: continue
    print(f"{s:9.1f} {e:9.1f} {e - s:8.1f}  {n:28s} {c:6d} {sm:4d} {busy:12.1f}")
# SM time held by tile warps (16 warp slots per SM)
tile = a[(a[:, 0] == 4)]
if len(tile):
    span = (a[:, 5].max() - a[:, 4].min()) / 1e3
    nsm = len(np.unique(a[:, 1]))
    held = (tile[:, 5] - tile[:, 4]).sum() / 1e3
    print(f"span {span:.1f} us, {nsm} SMs; tile warps hold {held / (16 * nsm * span) * 100:.1f} % of the warp slots over the span")
    # per 20 us bin: number of tile warps resident
    edges = np.arange(0, span + 20, 20.0)
    occ = np.zeros(len(edges) - 1)
    for t0, t1 in zip(tile[:, 4] / 1e3, tile[:, 5] / 1e3):
        i0, i1 = int(t0 // 20), int(min(t1, span) // 20)
        for i in range(i0, min(i1, len(occ) - 1) + 1):
            occ[i] += max(0.0, min(t1, edges[i + 1]) - max(t0, edges[i])) / 20.0
    print("tile warps resident per 20 us bin (of %d):" % (16 * nsm))
    print(" ".join(f"{int(o):4d}" for o in occ))
