#!/usr/bin/env python
"""bench.py - dexel-columns/s of the 3D dilation hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU implementation

A "step" is one dilation ('ours', radius 32) of one 2048 x 2048-column synthetic torus volume per GPU
(BASELINE.json configs[4] / SURVEY.md 8(d) C5). At N > 1 every multi-GPU number goes through the library's own
multi-GPU entry points (vo_mg_*: y-slabs, floor(R)-row input halo exchanged with NCCL inside libvoroffset_b200.so).

One JSON line on stdout (rank 0):
  value        WEAK scaling, whole-job columns/s with inputs resident in HBM: every rank holds one 2048 x 2048 slab of
               a 2048 x (2048 N) grid (CUDA events on the library's stream, max over ranks)
  strong       BASELINE config 5 AS WRITTEN: ONE 2048 x 2048 grid cut into N slabs (same timing rules)
  strong_large the same for one 8192 x 8192 grid (67 M columns), where N GPUs have enough rows each
  ops          erosion / opening / closing of the config-5 grid (padding 34) on the N GPUs
  slab_parity  every multi-GPU result above was compared, on every rank and before timing, with the single-GPU result
               of the same rows, bit for bit (the run fails on a mismatch)
  halo         per step: device time of the NCCL groups and host time stalled waiting for the halos (max over ranks)
  e2e          the same metric through the host-buffer C-ABI call: H2D of the CSR input from pinned memory, both passes,
               D2H of the CSR result, all inside the timed region (N = 1: the banded pipeline of vo_morph3d; N > 1:
               vo_morph3d_rows on every rank - its slab with the floor(R) ghost rows of its neighbours cut from host
               memory, the same banded pipeline, the slab's rows back; checked against the resident multi-GPU rows)
  roofline     dominant kernel (k_pass1_tile): algorithmic bytes / its event-timed duration
  cpu_baseline the reference's own code (oracle/_ref) on a bounded sample with all host threads
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import socket
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "dexel-columns/sec (3D dilation, method ours)"
UNIT = "columns/s"

# reference-equivalent mid pieces per column (k_mid of SURVEY.md 8(d)): counted by running the reference's
# own first pass (oracle/_ref ref3d_mid_count = VoronoiVorPower.cpp:50-65) on these exact volumes.
K_MID_REF = {(256, 8.0): 10.027, (512, 16.0): 19.461, (1024, 16.0): 19.715, (2048, 32.0): 38.825}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--n", type=int, default=2048, help="dexels per side of one slab (configs: 1024 / 2048)")
    p.add_argument("--radius", type=float, default=32.0)
    p.add_argument("--large-n", type=int, default=8192, help="side of the large strong-scaling grid (0: skip)")
    p.add_argument("--cpu-sample-cols", type=int, default=128, help="x-width of the CPU sample band")
    p.add_argument("--cpu-full", default="auto", choices=["auto", "yes", "no"],
                   help="--impl reference: run the reference on the FULL grid (auto: when the host has the cores and the memory)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-extras", action="store_true", help="skip the strong / large / ops blocks")
    return p.parse_args()


def workload_name(a):
    return f"offset3d synthetic torus -n {a.n} -r {a.radius:g} -x dilation -m ours (grid {a.n}x{a.n} per GPU)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(n, radius):
    """dram bytes per k_pass1 launch from the committed ncu --set full capture, if one matches."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(path))
        key = f"k_pass1@n{n}_r{radius:g}"
        return t.get(key)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks line")
# ---------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own code, on the full grid where the host allows it
# ---------------------------------------------------------------------------------------------------
def cpu_sample(a):
    """A 128-column-wide x-band (all rows) through the tube of the same torus volume."""
    from voroffset_b200 import synth
    from voroffset_b200.volume import CompressedVolume
    vol = synth.torus_z(a.n)
    w = min(a.cpu_sample_cols, vol.nx)
    x0 = vol.nx // 4
    cnt = vol.counts().reshape(vol.ny, vol.nx)[:, x0:x0 + w]
    starts = vol.off[:-1].astype(np.int64).reshape(vol.ny, vol.nx)[:, x0:x0 + w]
    idx = np.concatenate([np.arange(s, s + c) for s, c in zip(starts.reshape(-1), cnt.reshape(-1))]) if cnt.sum() else np.zeros(0, np.int64)
    off = np.zeros(w * vol.ny + 1, dtype=np.int64)
    np.cumsum(cnt.reshape(-1), out=off[1:])
    band = CompressedVolume(w, vol.ny, off, vol.spans[idx], vol.origin, vol.extent, vol.spacing, vol.padding)
    desc = (f"x-band [{x0},{x0 + w}) x all {vol.ny} rows of the {vol.nx}x{vol.ny} torus (k_in {band.numSegments() / (w * vol.ny):.2f}), "
            f"R={a.radius:g}, every column of the band counted")
    return band, desc


def _mem_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def _cache_path(a):
    return os.path.join("/tmp", f"voroffset_b200_ref_{socket.gethostname()}_{os.cpu_count()}c_n{a.n}_r{a.radius:g}.json")


def run_cpu(a, steps, warmup, full=False):
    from oracle.cpu import Oracle, Reference, reference_available
    cores = os.cpu_count() or 1
    if full:
        from voroffset_b200 import synth
        band = synth.torus_z(a.n)
        desc = (f"the FULL {band.nx}x{band.ny} grid (same config as the GPU arm), R={a.radius:g}, reference 'ours' with {cores} threads "
                f"through its own (dormant) TBB regions (VoronoiVorPower.cpp:41-63,70-92); marshalling into / out of its containers included")
    else:
        band, desc = cpu_sample(a)
    ncols = band.nx * band.ny
    if reference_available():
        ref = Reference()
        kind = "reference"
        run = lambda: ref.morph3d(band, "dilation", a.radius, "ours", threads=cores)
    else:
        orc = Oracle(threads=cores)
        kind = "port"
        run = lambda: orc.morph3d(band, "dilation", a.radius, "ours")
    for _ in range(warmup):
        run()
    ts = []
    for _ in range(steps):
        t = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t)
    sec = sum(ts) / len(ts)
    return {"value": ncols / sec, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc,
            "sec_per_sample": sec, "sample_columns": ncols, "same_config": bool(full)}


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # the full 2048^2 grid needs ~25 GB and (577 s of sweeps) / cores: one step, no warm-up, cached per host
    full = a.cpu_full == "yes" or (a.cpu_full == "auto" and cores >= 8 and _mem_gb() >= 40.0 and a.n <= 2048)
    cb, cached = None, False
    if full and os.path.exists(_cache_path(a)):
        try:
            cb, cached = json.load(open(_cache_path(a))), True
        except Exception:
            cb = None
    if cb is None:
        if full:
            steps, warmup = 1, 0
        else:
            steps, warmup = max(1, min(a.steps, 3)), min(a.warmup, 1)
        cb = run_cpu(a, steps, warmup, full=full)
        cb["steps"], cb["warmup"] = steps, warmup
        if full:
            try:
                json.dump(cb, open(_cache_path(a), "w"))
            except Exception:
                pass
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": cb["steps"],
            "warmup": cb["warmup"], "ms_per_step": cb["sec_per_sample"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "sample": cb["sample"], "same_config": cb.get("same_config", False),
                       "cached": cached},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def _rows(vol, y0, y1):
    """Rows [y0, y1) of a host volume."""
    c0, c1 = y0 * vol.nx, y1 * vol.nx
    off = vol.off[c0:c1 + 1].astype(np.int64)
    return vol.like(vol.nx, y1 - y0, (off - off[0]).astype(np.uint32), vol.spans[off[0]:off[-1]])


def _stack(parts):
    """Host volumes of equal nx stacked along y."""
    parts = [p for p in parts if p is not None]
    offs, base = [np.zeros(1, dtype=np.int64)], 0
    for p in parts:
        offs.append(p.off[1:].astype(np.int64) + base)
        base += int(p.off[-1])
    spans = np.concatenate([p.spans.reshape(-1, 2) for p in parts]) if base else np.zeros((0, 2))
    return parts[0].like(parts[0].nx, sum(p.ny for p in parts), np.concatenate(offs).astype(np.uint32), spans)


def pin_near_gpu(local):
    """Best effort: run this process (and first-touch its pinned buffers) on the NUMA node of its GPU."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        if bus is None:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)], capture_output=True, text=True).stdout.strip()
            bus = out[-12:].lower() if out else None
        else:
            bus = f"0000:{bus:02x}:00.0" if isinstance(bus, int) else str(bus).lower()[-12:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def main():
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
        return
    import torch
    import torch.distributed as dist
    from voroffset_b200 import _lib, morpho, multigpu, slab, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_near_gpu(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        mg = multigpu.MultiGpu.from_torch_distributed(local)
        ctx = mg.contexts[0]
    else:
        mg = None
        ctx = _lib.Context(local)
    R = a.radius
    J = int(math.floor(R))
    op = morpho.make_operator("ours", ctx)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def all_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_min(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    def all_ok(flag):
        t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def run_op(opn, d_in, meta):
        """One operator on this rank's resident rows: (result, time_1, time_2)."""
        if mg is None:
            return op.morph_dev(opn, d_in, R)
        outs, t1, t2 = mg.morph_dev(opn, [d_in], R, meta.zmin, meta.zmax)
        return outs[0], t1, t2

    def timed(opn, d_in, meta, steps, warmup):
        """steps timed calls (CUDA events on the library stream, L2 flushed in between). Returns a dict."""
        for _ in range(warmup):
            r, _, _ = run_op(opn, d_in, meta)
            r.free()
        barrier()
        ms, k1s, k2s, p1s, p2s, hw, hm, nseg = [], [], [], [], [], [], [], 0
        for _ in range(steps):
            flush.zero_()                          # L2 flush between timed iterations (untimed)
            torch.cuda.synchronize(dev)
            ctx.mark(0)
            r, t1, t2 = run_op(opn, d_in, meta)
            ctx.mark(1)
            ms.append(ctx.elapsed_ms(0, 1))
            k1, k2 = ctx.last_profile()
            k1s.append(k1); k2s.append(k2); p1s.append(t1); p2s.append(t2)
            if mg is not None:
                st = mg.stats(0)
                hw.append(st["halo_wait_ms"]); hm.append(st["halo_ms"])
            nseg = r.info()[2]
            r.free()
        barrier()
        total = all_max(sum(ms))
        return {"total_ms": total, "ms_per_step": total / steps, "k1": statistics.mean(k1s), "k2": statistics.mean(k2s),
                "p1": statistics.mean(p1s), "p2": statistics.mean(p2s), "nseg": nseg,
                # (the ranks are not re-synchronised between steps: a rank with one neighbour runs ahead and then waits for
                # the slower ones, so the MAX over ranks mostly shows that skew and the MIN what the exchange itself costs)
                "halo_wait_ms": all_max(statistics.mean(hw)) if hw else 0.0, "halo_ms": all_max(statistics.mean(hm)) if hm else 0.0,
                "halo_wait_ms_min": all_min(statistics.mean(hw)) if hw else 0.0}

    def parity_rows(got_dev, ext_host, y_lo, y_hi, opn="dilation"):
        """This rank's multi-GPU rows against the single-GPU operator on `ext_host`, rows [y_lo, y_hi), bit for bit."""
        d = morpho.DeviceVolume.upload(ctx, ext_host)
        full, _, _ = op.morph_dev(opn, d, R)
        want = full.rows(y_lo, y_hi)
        a_, b_ = got_dev.download(), want.download()
        ok = a_.bit_equal(b_)
        for v in (want, full, d):
            v.free()
        return ok

    # ---- weak: one n x n torus per rank = rows [rank n, (rank+1) n) of an n x (n world) grid ---------------------
    vol = synth.torus_z(a.n)
    ncols = vol.nx * vol.ny
    k_in = vol.numSegments() / ncols
    d_in = morpho.DeviceVolume.upload(ctx, vol)
    parity = {}
    if world > 1:
        r0, _, _ = run_op("dilation", d_in, vol)
        ext = _stack([_rows(vol, vol.ny - J, vol.ny) if rank > 0 else None, vol, _rows(vol, 0, J) if rank + 1 < world else None])
        y_lo = J if rank > 0 else 0
        parity["weak"] = all_ok(parity_rows(r0, ext, y_lo, y_lo + vol.ny))
        r0.free()
        if not parity["weak"]:
            raise SystemExit("slab parity FAILED (weak workload): multi-GPU rows differ from the single-GPU result")
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    launches0 = ctx.launches
    t_wall0 = time.time()
    w = timed("dilation", d_in, vol, a.steps, a.warmup)
    t_wall1 = time.time()
    launches = ctx.launches - launches0
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None
    nseg = w["nseg"]

    def drop_scratch_cache():
        """Scratch blocks the library keeps for reuse: drop them between workloads of very different size,
        so that the multi-GB blocks of the 8192^2 grid are not released in the middle of a later timed step."""
        ctx.set_option("block_cache", "off")
        ctx.set_option("block_cache", "on")

    extras = {}
    if not a.no_extras:
        xs, xw = max(3, a.steps // 2), 3
        # ---- strong: BASELINE config 5 as written - ONE n x n grid over the N GPUs --------------------------------
        if world > 1 and a.n // world >= J:
            y0, y1 = slab.slab_bounds(a.n, world)[rank]
            own = synth.torus_z_rows(a.n, y0, y1)
            d_own = morpho.DeviceVolume.upload(ctx, own)
            r0, _, _ = run_op("dilation", d_own, own)
            e0, e1 = max(0, y0 - J), min(a.n, y1 + J)
            parity["strong"] = all_ok(parity_rows(r0, synth.torus_z_rows(a.n, e0, e1), y0 - e0, y1 - e0))
            r0.free()
            if not parity["strong"]:
                raise SystemExit("slab parity FAILED (strong workload)")
            s = timed("dilation", d_own, own, a.steps, a.warmup)
            extras["strong"] = {"grid": [a.n, a.n], "slabs": world, "rows_per_gpu": a.n // world, "ms_per_step": s["ms_per_step"],
                                "value": a.n * a.n * a.steps / (s["total_ms"] * 1e-3), "unit": UNIT, "steps": a.steps,
                                "halo_wait_ms": s["halo_wait_ms"], "halo_wait_ms_min": s["halo_wait_ms_min"], "halo_ms": s["halo_ms"],
                                "pass_ms": {"pass1": s["p1"], "pass2": s["p2"]}, "slab_parity": parity["strong"]}
            d_own.free()
        elif world == 1:
            extras["strong"] = {"grid": [a.n, a.n], "slabs": 1, "rows_per_gpu": a.n, "ms_per_step": w["ms_per_step"],
                                "value": ncols * a.steps / (w["total_ms"] * 1e-3), "unit": UNIT, "steps": a.steps, "note": "N = 1: same run as value"}
        # ---- strong, large grid ---------------------------------------------------------------------------------
        if a.large_n and a.large_n // world >= J:
            nl = a.large_n
            y0, y1 = slab.slab_bounds(nl, world)[rank]
            own = synth.torus_z_rows(nl, y0, y1)
            d_own = morpho.DeviceVolume.upload(ctx, own)
            ok = True
            if world > 1:
                r0, _, _ = run_op("dilation", d_own, own)
                e0, e1 = max(0, y0 - J), min(nl, y1 + J)
                ok = parity_rows(r0, synth.torus_z_rows(nl, e0, e1), y0 - e0, y1 - e0)
                r0.free()
                parity["strong_large"] = all_ok(ok)
                if not parity["strong_large"]:
                    raise SystemExit("slab parity FAILED (large strong workload)")
            s = timed("dilation", d_own, own, xs, xw)
            extras["strong_large"] = {"grid": [nl, nl], "slabs": world, "rows_per_gpu": nl // world, "ms_per_step": s["ms_per_step"],
                                      "value": nl * nl * xs / (s["total_ms"] * 1e-3), "unit": UNIT, "steps": xs, "warmup": xw,
                                      "halo_wait_ms": s["halo_wait_ms"], "halo_wait_ms_min": s["halo_wait_ms_min"], "halo_ms": s["halo_ms"],
                                      "pass_ms": {"pass1": s["p1"], "pass2": s["p2"]}}
            if world > 1:
                extras["strong_large"]["slab_parity"] = parity["strong_large"]
            d_own.free()
            del own
            drop_scratch_cache()
        # ---- the other operations of the target on the config-5 grid (padding 34 = head-room for the composites) ----
        pad = J + 2
        full = synth.torus_z(a.n, padding=pad)
        if full.ny // world >= J:
            y0, y1 = slab.slab_bounds(full.ny, world)[rank]
            own = _rows(full, y0, y1) if world > 1 else full
            d_own = morpho.DeviceVolume.upload(ctx, own)
            d_full = morpho.DeviceVolume.upload(ctx, full) if world > 1 else None
            ops = {}
            for opn in ("erosion", "opening", "closing"):
                if world > 1:
                    r0, _, _ = run_op(opn, d_own, full)
                    ref_full, _, _ = op.morph_dev(opn, d_full, R)
                    ref_rows = ref_full.rows(y0, y1)
                    ok = r0.download().bit_equal(ref_rows.download())
                    for v in (r0, ref_rows, ref_full):
                        v.free()
                    parity[opn] = all_ok(ok)
                    if not parity[opn]:
                        raise SystemExit(f"slab parity FAILED ({opn})")
                s = timed(opn, d_own, full, xs, xw)
                prim = 2 if opn in ("opening", "closing") else 1
                ops[opn] = {"ms": s["ms_per_step"], "value": full.nx * full.ny * prim * xs / (s["total_ms"] * 1e-3), "unit": UNIT,
                            "halo_wait_ms": s["halo_wait_ms"]}
                if world > 1:
                    ops[opn]["slab_parity"] = parity[opn]
            extras["ops"] = {"grid": [full.nx, full.ny], "padding": pad, "slabs": world, "steps": xs, "warmup": xw,
                             "primitives_counted": "columns x 1 for erosion, x 2 for opening / closing (SURVEY.md 8(d))", **ops}
            d_own.free()
            if d_full is not None:
                d_full.free()

    # ---- e2e: host buffers in, host buffers out, through the drop-in call ------------------------
    e2e = None
    if not a.no_e2e:
        # N = 1: the grid as it is. N > 1: this rank's slab WITH the floor(R) ghost rows of its neighbours in the host buffer
        # (the host-side halo of vo_morph3d_rows: a host application that holds the grid has them anyway), rows of the slab back
        jp, jn = (J if rank > 0 else 0), (J if rank + 1 < world else 0)
        ext = vol if world == 1 else _stack([_rows(vol, vol.ny - J, vol.ny) if jp else None, vol, _rows(vol, 0, J) if jn else None])
        off_pin = torch.from_numpy(ext.off.view(np.int32)).pin_memory()
        sp_pin = torch.from_numpy(ext.spans).pin_memory()
        h2d = off_pin.numel() * 4 + sp_pin.numel() * 8
        d2h = 0
        e2e_first = []

        def step_e2e():
            nonlocal d2h
            poff, pspans, n = _lib._u32p(), _lib._f64p(), C.c_uint64()
            if world == 1:
                ctx.check(ctx.lib.vo_morph3d(ctx.handle, 0, 0, vol.nx, vol.ny, vol.zmin, vol.zmax, off_pin.data_ptr(),
                                             sp_pin.data_ptr(), R, C.byref(poff), C.byref(pspans), C.byref(n), None, None))
            else:
                ctx.check(ctx.lib.vo_morph3d_rows(ctx.handle, 0, 0, ext.nx, ext.ny, vol.zmin, vol.zmax, off_pin.data_ptr(),
                                                  sp_pin.data_ptr(), R, jp, jp + vol.ny, C.byref(poff), C.byref(pspans), C.byref(n), None, None))
            d2h = (ncols + 1) * 4 + int(n.value) * 16
            if not e2e_first:                                  # checksum of the first result: the same rows as the resident arm's
                o = np.ctypeslib.as_array(poff, shape=(ncols + 1,)).copy()
                sp = np.ctypeslib.as_array(pspans, shape=(max(int(n.value), 1) * 2,))[:2 * int(n.value)].copy()
                e2e_first.append((o, sp))
            ctx.lib.vo_free(C.cast(poff, C.c_void_p)); ctx.lib.vo_free(C.cast(pspans, C.c_void_p))

        for _ in range(max(1, min(a.warmup, 2))):
            step_e2e()
        if world > 1:
            # the rows of the host-buffer call against the resident multi-GPU rows of the same slab, bit for bit
            r0, _, _ = run_op("dilation", d_in, vol)
            want = r0.download()
            r0.free()
            same = np.array_equal(want.off, e2e_first[0][0]) and np.array_equal(want.spans.reshape(-1).view(np.uint64), e2e_first[0][1].view(np.uint64))
            parity["e2e_rows"] = all_ok(same)
            if not parity["e2e_rows"]:
                raise SystemExit("slab parity FAILED (host-buffer call with ghost rows)")
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step_e2e()
        barrier()
        e2e_s = all_max(time.perf_counter() - t0)
        e2e = {"value": ncols * world * a.steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3 / a.steps,
               "timing": "wall clock around the synchronous call(s) with pinned host buffers, max over ranks", "numa_node": numa,
               "call": "vo_morph3d (banded pipeline)" if world == 1 else
                       "vo_morph3d_rows per rank: the slab with its floor(R) ghost rows cut from host memory (no device-to-device exchange), "
                       "banded pipeline, the slab's rows back"}

    if rank == 0:
        peak, peak_src = peaks()
        k_mid = K_MID_REF.get((a.n, float(R)))
        k_mid_src = "reference first pass (oracle/_ref ref3d_mid_count)"
        if k_mid is None:
            k_mid, k_mid_src = (2 * math.floor(R) + 1) * k_in, "estimate (2 floor(R) + 1) * k_in"
        k_out = nseg / ncols
        b1 = (4 + 16 * k_in) + (4 + 24 * k_mid)          # SURVEY.md 8(d) pass 1, fp64
        b2 = (4 + 24 * k_mid) + (4 + 16 * k_out)         # pass 2
        k1 = w["k1"]
        ach = b1 * ncols / (k1 * 1e-3) / 1e9 if k1 > 0 else 0.0
        total_ms = w["total_ms"]
        line = {
            "metric": METRIC, "value": ncols * world * a.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "grid_per_gpu": [vol.nx, vol.ny], "radius": R, "method": "ours",
                       "operation": "dilation", "k_in": round(k_in, 4), "k_out": round(k_out, 4),
                       "parallelism": f"y-slabs x{world}, floor(R)-row input halo over NCCL inside libvoroffset_b200.so (vo_mg_*)" if world > 1 else "single GPU",
                       "l2": "flushed between timed steps (256 MiB memset, untimed); the touched part of the mid volume (~0.5 GB) also exceeds L2",
                       "timing": "CUDA events on the library stream around each step, summed, max over ranks"},
            "e2e": e2e, "gpu_launches": int(launches),
            "pass_ms": {"pass1": w["p1"], "pass2": w["p2"], "k_pass1": k1, "k_pass2": w["k2"]},
            "halo": {"halo_wait_ms": w["halo_wait_ms"], "halo_wait_ms_min": w["halo_wait_ms_min"], "halo_ms": w["halo_ms"],
                     "note": "per step: max / min over ranks of the host time stalled until the halos had landed, max of the device time of the NCCL groups"},
            "slab_parity": (all(parity.values()) if parity else None), "slab_parity_checks": parity,
            **extras,
            "roofline": {"bound": "hbm", "kernel": "k_pass1_tile (the three tile launches of pass 1)", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": ncu_traffic(a.n, R), "peak_source": peak_src,
                         "algorithmic_bytes_per_column": b1, "k_mid": k_mid, "k_mid_source": k_mid_src,
                         "definition": "SURVEY.md 8(d): B1 = (4 + 16 k_in) + (4 + 24 k_mid) bytes per column, k_mid = pieces per "
                                       "column of the REFERENCE's mid volume; our kernel never materialises those pieces "
                                       "(`traffic` = what it really moves, from ncu), so this fraction measures speed against the "
                                       "reference's data flow, not DRAM utilisation",
                         "dram": (None if not ncu_traffic(a.n, R) or k1 <= 0 else
                                  {"gbs": ncu_traffic(a.n, R) / (k1 * 1e-3) / 1e9, "frac": ncu_traffic(a.n, R) / (k1 * 1e-3) / 1e9 / peak,
                                   "note": "what the kernel physically moves (ncu dram bytes of the committed capture / its live duration): "
                                           "it is bound by instruction issue, not by HBM"}),
                         "whole_dilation": {"bytes_per_column": b1 + b2,
                                            "achieved": (b1 + b2) * ncols / ((total_ms / a.steps) * 1e-3) / 1e9,
                                            "frac": (b1 + b2) * ncols / ((total_ms / a.steps) * 1e-3) / 1e9 / peak}},
            "clocks": clk,
        }
        if world == 1 and not a.no_cpu_baseline:
            cb = None
            if os.path.exists(_cache_path(a)):      # the full-grid run of `--impl reference` on this host, if it has been made
                try:
                    cb = json.load(open(_cache_path(a)))
                    cb["sample"] += " [cached from bench.py --impl reference on this host]"
                except Exception:
                    cb = None
            if cb is None:
                cb = run_cpu(a, 1, 0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    d_in.free()
    if mg is not None:
        mg.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
