"""Context for the BA Jacobian kernel's roofline: it is ~90 % writes.  Measures write-only (fill_) and
copy (read+write) bandwidth with torch on this GPU for a 208 MB and a 2 GB buffer."""
import torch
for nbytes in (208_000_000, 2_000_000_000):
    n = nbytes // 8
    a = torch.empty(n, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
    for name, fn, moved in (("fill (write only)", lambda: a.fill_(1.0), nbytes), ("copy (read+write)", lambda: b.copy_(a), 2 * nbytes)):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{nbytes/1e6:.0f} MB {name}: {best*1e3:.1f} us -> {moved/best/1e6:.0f} GB/s")
