"""What the HBM does for a write-only stream (the regime of fold3_kernel) vs a copy: torch fill_ / copy_ of 3 GB."""
import json
import torch

n = 3 * 1024 ** 3 // 4
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")


def t(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms_fill = t(lambda: a.fill_(1.0))
ms_zero = t(lambda: torch.cuda.current_stream().synchronize() or a.zero_())
ms_copy = t(lambda: b.copy_(a))
ms_read = t(lambda: a.sum())
gb = n * 4 / 1e9
print(json.dumps({"bytes": n * 4, "fill_GBs": gb / ms_fill * 1e3, "zero_GBs": gb / ms_zero * 1e3,
                  "copy_GBs_rw": 2 * gb / ms_copy * 1e3, "read_sum_GBs": gb / ms_read * 1e3}))
