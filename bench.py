#!/usr/bin/env python
"""bench.py - dexel-columns/s of the 3D dilation hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU implementation

A "step" is one dilation ('ours', radius 32) of one 2048 x 2048-column synthetic torus volume per GPU
(BASELINE.json configs[4] / SURVEY.md 8(d) C5; at N > 1 the global grid is 2048 x (2048 N) cut into y-slabs,
one per rank, with a floor(R)-row halo exchanged over NCCL every step: weak scaling).

One JSON line on stdout (rank 0):
  value        whole-job columns/s with inputs resident in HBM (CUDA events on the library's stream)
  e2e          the same through the host-buffer C-ABI call (vo_morph3d): H2D of the CSR input from pinned
               memory, both passes, D2H of the CSR result, all inside the timed region
               (N = 1: the banded pipeline of vo_morph3d; N > 1: vo_dvol_upload, the slab step with its halo
               exchange, vo_dvol_download, one after the other)
  roofline     dominant kernel (k_pass1_tile): algorithmic bytes / its event-timed duration
  cpu_baseline the reference's own code (oracle/_ref) on a bounded sample with all host threads
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
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
    p.add_argument("--cpu-sample-cols", type=int, default=128, help="x-width of the CPU sample band")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
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
# reference arm / cpu baseline: the reference's own code on a bounded sample
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


def run_cpu(a, steps, warmup):
    from oracle.cpu import Oracle, Reference, reference_available
    cores = os.cpu_count() or 1
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
            "sec_per_sample": sec, "sample_columns": ncols}


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(a.steps, 3)), min(a.warmup, 1)
    cb = run_cpu(a, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": cb["sec_per_sample"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "sample": cb["sample"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
        return
    import torch
    import torch.distributed as dist
    from voroffset_b200 import _lib, morpho, slab, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    ctx = _lib.Context(local)
    if os.environ.get("VO_SLAB"):                   # development switch: "overlap" / "serial" halo exchange
        ctx.set_option("slab", os.environ["VO_SLAB"])
    vol = synth.torus_z(a.n)                       # this rank's slab: rows [rank*n, (rank+1)*n) of the global grid
    R = a.radius
    ncols = vol.nx * vol.ny
    k_in = vol.numSegments() / ncols
    op = morpho.make_operator("ours", ctx)
    d_in = morpho.DeviceVolume.upload(ctx, vol)
    sd = slab.SlabDilation(slab.CudaSlabBackend(ctx), rank, world) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        if sd is None:
            out, t1, t2 = op.morph_dev("dilation", d_in, R)
        else:
            out = sd.dilate(d_in, R)
            t1, t2 = sd.last_ms
        nseg = out.info()[2]
        out.free()
        return nseg, t1, t2

    for _ in range(a.warmup):
        step_resident()
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    barrier()
    launches0 = ctx.launches
    t_wall0 = time.time()
    step_ms, k1_ms, k2_ms, p1_ms, p2_ms = [], [], [], [], []
    nseg = 0
    for _ in range(a.steps):
        flush.zero_()                              # L2 flush between timed iterations (untimed)
        torch.cuda.synchronize(dev)
        ctx.mark(0)
        nseg, t1, t2 = step_resident()
        ctx.mark(1)
        step_ms.append(ctx.elapsed_ms(0, 1))
        k1, k2 = ctx.last_profile()
        k1_ms.append(k1); k2_ms.append(k2); p1_ms.append(t1); p2_ms.append(t2)
    barrier()
    t_wall1 = time.time()
    launches = ctx.launches - launches0
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- e2e: host buffers in, host buffers out, through the drop-in call ------------------------
    e2e = None
    if not a.no_e2e:
        off_pin = torch.from_numpy(vol.off.view(np.int32)).pin_memory()
        sp_pin = torch.from_numpy(vol.spans).pin_memory()
        h2d = off_pin.numel() * 4 + sp_pin.numel() * 8
        d2h = 0
        # pinned result buffers of the N > 1 path (the single-GPU call returns library-owned pinned blocks)
        out_pin = [torch.empty(ncols + 1, dtype=torch.int32).pin_memory(), torch.empty(4 * sp_pin.numel(), dtype=torch.float64).pin_memory()]

        def step_e2e():
            nonlocal d2h
            poff, pspans, n = _lib._u32p(), _lib._f64p(), C.c_uint64()
            if sd is None:
                ctx.check(ctx.lib.vo_morph3d(ctx.handle, 0, 0, vol.nx, vol.ny, vol.zmin, vol.zmax, off_pin.data_ptr(),
                                             sp_pin.data_ptr(), R, C.byref(poff), C.byref(pspans), C.byref(n), None, None))
                d2h = (ncols + 1) * 4 + int(n.value) * 16
                ctx.lib.vo_free(C.cast(poff, C.c_void_p)); ctx.lib.vo_free(C.cast(pspans, C.c_void_p))
            else:
                h = C.c_void_p()
                ctx.check(ctx.lib.vo_dvol_upload(ctx.handle, vol.nx, vol.ny, off_pin.data_ptr(), sp_pin.data_ptr(), C.byref(h)))
                d = morpho.DeviceVolume(ctx, h, vol)
                out = sd.dilate(d, R)
                nseg_out = out.info()[2]
                if out_pin[1].numel() < 2 * nseg_out:
                    out_pin[1] = torch.empty(int(2.2 * nseg_out), dtype=torch.float64).pin_memory()
                ctx.check(ctx.lib.vo_dvol_download(ctx.handle, out.handle, out_pin[0].data_ptr(), out_pin[1].data_ptr()))
                d2h = (ncols + 1) * 4 + nseg_out * 16
                out.free(); d.free()

        for _ in range(max(1, min(a.warmup, 2))):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step_e2e()
        barrier()
        e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e = {"value": ncols * world * a.steps / float(e2e_s.item()), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": float(e2e_s.item()) * 1e3 / a.steps,
               "timing": "wall clock around the synchronous vo_morph3d call (pinned host buffers), max over ranks"}

    if rank == 0:
        peak, peak_src = peaks()
        k_mid = K_MID_REF.get((a.n, float(R)))
        k_mid_src = "reference first pass (oracle/_ref ref3d_mid_count)"
        if k_mid is None:
            k_mid, k_mid_src = (2 * math.floor(R) + 1) * k_in, "estimate (2 floor(R) + 1) * k_in"
        k_out = nseg / ncols
        rows_p1 = ncols if world == 1 else ncols + 2 * math.floor(R) * vol.nx * (1 if world > 1 else 0)
        b1 = (4 + 16 * k_in) + (4 + 24 * k_mid)          # SURVEY.md 8(d) pass 1, fp64
        b2 = (4 + 24 * k_mid) + (4 + 16 * k_out)         # pass 2
        k1 = statistics.mean(k1_ms)
        ach = b1 * ncols / (k1 * 1e-3) / 1e9 if k1 > 0 else 0.0
        line = {
            "metric": METRIC, "value": ncols * world * a.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "grid_per_gpu": [vol.nx, vol.ny], "radius": R, "method": "ours",
                       "operation": "dilation", "k_in": round(k_in, 4), "k_out": round(k_out, 4),
                       "parallelism": f"y-slabs x{world}, floor(R)-row input halo over NCCL" if world > 1 else "single GPU",
                       "l2": "flushed between timed steps (256 MiB memset, untimed); the touched part of the mid volume (~0.5 GB) also exceeds L2",
                       "timing": "CUDA events on the library stream around each step, summed, max over ranks"},
            "e2e": e2e, "gpu_launches": int(launches),
            "pass_ms": {"pass1": statistics.mean(p1_ms), "pass2": statistics.mean(p2_ms),
                        "k_pass1": k1, "k_pass2": statistics.mean(k2_ms)},
            "roofline": {"bound": "hbm", "kernel": "k_pass1_tile (the three tile launches of pass 1)", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak, "traffic": ncu_traffic(a.n, R), "peak_source": peak_src,
                         "algorithmic_bytes_per_column": b1, "k_mid": k_mid, "k_mid_source": k_mid_src,
                         "definition": "SURVEY.md 8(d): B1 = (4 + 16 k_in) + (4 + 24 k_mid) bytes per column, k_mid = pieces per "
                                       "column of the REFERENCE's mid volume; our kernel never materialises those pieces "
                                       "(`traffic` = what it really moves, from ncu), so this fraction measures speed against the "
                                       "reference's data flow, not DRAM utilisation",
                         "whole_dilation": {"bytes_per_column": b1 + b2,
                                            "achieved": (b1 + b2) * ncols / ((total_ms / a.steps) * 1e-3) / 1e9,
                                            "frac": (b1 + b2) * ncols / ((total_ms / a.steps) * 1e-3) / 1e9 / peak}},
            "clocks": clk,
        }
        if world == 1 and not a.no_cpu_baseline:
            cb = run_cpu(a, 1, 0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
