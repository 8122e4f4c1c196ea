"""Does a saturated PCIe link slow down kernel launches? Times a chain of 200 tiny dependent kernels on one stream
alone, beside a large H2D copy, beside a large D2H copy (pinned memory, other streams)."""
import json, time
import torch
x = torch.zeros(1024, device="cuda")
big_h = torch.empty(512 << 20, dtype=torch.uint8).pin_memory(); big_d = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
s_k, s_c = torch.cuda.Stream(), torch.cuda.Stream()
def chain(copy):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if copy == "h2d":
        with torch.cuda.stream(s_c): big_d.copy_(big_h, non_blocking=True)
    elif copy == "d2h":
        with torch.cuda.stream(s_c): big_h.copy_(big_d, non_blocking=True)
    time.sleep(0.001)                      # the copy is under way (512 MB take ~10 ms)
    with torch.cuda.stream(s_k):
        e0.record()
        for _ in range(200): x.add_(1.0)
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / 200
for _ in range(3): chain(None)
out = {k or "alone": round(min(chain(k) for _ in range(5)), 2) for k in (None, "h2d", "d2h")}
out["unit"] = "us per dependent tiny kernel (host enqueue included: the chain is launch-bound)"
print(json.dumps(out))
