"""Dev probe: what does a pure write stream achieve on this GPU? (memset / fill / copy) -- context for the
assembly kernel's roofline, which is write-only."""
import torch, json
dev = torch.device("cuda")
n = 17 * (1 << 30) // 16
x = torch.empty(n, dtype=torch.complex128, device=dev)
y = torch.empty(n // 2, dtype=torch.complex128, device=dev)
def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
out = {}
ms = timeit(lambda: x.zero_());              out["memset_zero_GBs"] = x.numel() * 16 / ms / 1e6
ms = timeit(lambda: x.fill_(1.0 + 2.0j));    out["fill_kernel_GBs"] = x.numel() * 16 / ms / 1e6
xr = torch.view_as_real(x)
ms = timeit(lambda: xr.fill_(3.0));          out["fill_f64_GBs"] = x.numel() * 16 / ms / 1e6
ms = timeit(lambda: y.copy_(x[: n // 2]));   out["copy_rw_GBs"] = 2 * y.numel() * 16 / ms / 1e6
ms = timeit(lambda: torch.sum(xr));          out["read_sum_GBs"] = x.numel() * 16 / ms / 1e6
print(json.dumps(out))
