"""Calibration on the GPU box: plain fill / copy bandwidth (write-only and read+write ceilings)."""
import json, torch
dev = torch.device("cuda", 0)
n = 9 * 1024**3
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms_fill = t(lambda: a.fill_(7))
ms_copy = t(lambda: b.copy_(a))
ms_zero = t(lambda: a.zero_())
print(json.dumps({"fill_GBps": n / ms_fill / 1e6, "zero_GBps": n / ms_zero / 1e6, "copy_rw_GBps": 2 * n / ms_copy / 1e6, "bytes": n}))
