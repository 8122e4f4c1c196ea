"""PCIe behaviour under the copy pattern of the banded host-buffer call, without any compute: B bands, per band two H2D
copies (offsets, spans) on one stream and two D2H copies on another, the download of band b released by an event when
the upload of band b + lag is done. Prints per-band completion times of both streams (CUDA events) and the total.
Usage: pcie_bands.py [bands] [lag]"""
import sys, time
import torch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
lag = int(sys.argv[2]) if len(sys.argv) > 2 else 2
up_off, up_sp, dn_off, dn_sp = 16777220, 40489552, 16777220, 44647152
h_uo = torch.empty(up_off, dtype=torch.uint8).pin_memory(); h_us = torch.empty(up_sp, dtype=torch.uint8).pin_memory()
h_do = torch.empty(dn_off, dtype=torch.uint8).pin_memory(); h_ds = torch.empty(dn_sp, dtype=torch.uint8).pin_memory()
d_uo = torch.empty(up_off, dtype=torch.uint8, device="cuda"); d_us = torch.empty(up_sp, dtype=torch.uint8, device="cuda")
d_do = torch.empty(dn_off, dtype=torch.uint8, device="cuda"); d_ds = torch.empty(dn_sp, dtype=torch.uint8, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
def cut(n, b): return (n * b // B) & ~15, (n * (b + 1) // B) & ~15 if b + 1 < B else n
def run(do_down=True, trace=False):
    torch.cuda.synchronize()
    ev_in = [torch.cuda.Event(enable_timing=True) for _ in range(B)]
    ev_out = [torch.cuda.Event(enable_timing=True) for _ in range(B)]
    ev_out0 = [torch.cuda.Event(enable_timing=True) for _ in range(B)]
    t0 = torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    with torch.cuda.stream(s_in):
        t0.record()
        for b in range(B):
            a0, a1 = cut(up_off, b); d_uo[a0:a1].copy_(h_uo[a0:a1], non_blocking=True)
            a0, a1 = cut(up_sp, b); d_us[a0:a1].copy_(h_us[a0:a1], non_blocking=True)
            ev_in[b].record()
    if do_down:
        with torch.cuda.stream(s_out):
            for b in range(B):
                s_out.wait_event(ev_in[min(B - 1, b + lag)])
                ev_out0[b].record()
                a0, a1 = cut(dn_off, b); h_do[a0:a1].copy_(d_do[a0:a1], non_blocking=True)
                a0, a1 = cut(dn_sp, b); h_ds[a0:a1].copy_(d_ds[a0:a1], non_blocking=True)
                ev_out[b].record()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t) * 1e3
    if trace:
        print("upload done   ", " ".join(f"{t0.elapsed_time(e):6.3f}" for e in ev_in))
        if do_down:
            print("download begin", " ".join(f"{t0.elapsed_time(e):6.3f}" for e in ev_out0))
            print("download done ", " ".join(f"{t0.elapsed_time(e):6.3f}" for e in ev_out))
    return ms
for _ in range(3): run()
print("uploads only: ms", round(min(run(False) for _ in range(10)), 3)); run(False, True)
print(f"uploads + downloads ({B} bands, lag {lag}): ms", round(min(run() for _ in range(10)), 3)); run(True, True)
