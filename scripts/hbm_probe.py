"""HBM probes on the pool's B200: write-only (fill), read-only (sum), copy — GB/s of bytes moved (torch kernels; denominators only)."""
import torch, json
n = 2 * 1024**3   # 16 GiB of float64
x = torch.empty(n, dtype=torch.float64, device="cuda")
y = torch.empty(n, dtype=torch.float64, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
out = {}
ms = t(lambda: x.fill_(1.5)); out["fill_write_only_GBps"] = n * 8 / ms / 1e6
ms = t(lambda: x.zero_()); out["memset_GBps"] = n * 8 / ms / 1e6
ms = t(lambda: x.sum()); out["sum_read_only_GBps"] = n * 8 / ms / 1e6
ms = t(lambda: y.copy_(x)); out["copy_GBps_read_plus_write"] = 2 * n * 8 / ms / 1e6
print(json.dumps(out))
