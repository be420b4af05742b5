"""Host <-> device copy bandwidth of this box with pinned memory (what bounds `e2e` of the small kits)."""
import torch
n = 332 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print("%s %.1f GB/s (332 MB copies, pinned)" % (name, 10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9))
