"""PCIe copy rates of the box: H2D alone, D2H alone, both at once (pinned buffers of the sizes the C5 e2e step moves).
The host-buffer call (bench.py's e2e) cannot be faster than its bytes over the concurrent rate."""
import json, sys, time
import torch
up, down = 57266772, 61424372
h_in = torch.empty(up, dtype=torch.uint8).pin_memory(); h_out = torch.empty(down, dtype=torch.uint8).pin_memory()
d_in = torch.empty(up, dtype=torch.uint8, device="cuda"); d_out = torch.empty(down, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(do_up, do_down, reps=20):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if do_up:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if do_down:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3
for _ in range(3): run(True, True, 2)
a, b, c = run(True, False), run(False, True), run(True, True)
print(json.dumps({"h2d_ms": round(a, 3), "h2d_GBps": round(up / a / 1e6, 1), "d2h_ms": round(b, 3), "d2h_GBps": round(down / b / 1e6, 1),
                  "both_ms": round(c, 3), "both_aggregate_GBps": round((up + down) / c / 1e6, 1),
                  "note": "both_ms = floor of the C5 host-buffer step if compute were free and the directions overlapped completely"}))
