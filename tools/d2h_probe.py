"""D2H bandwidth of this box for the solution bytes of one sweep (4.3 GB) from one GPU: one stream against two / four
streams (chunks alternating), chunk sizes 64 MB .. 1 GB; pinned host memory from torch.  What `e2e.ceiling_gbs` of the
bench can and cannot reach.    python tools/d2h_probe.py"""
import time

import torch

n = 4_296_015_872 // 16
dev = torch.empty(n, dtype=torch.complex128, device="cuda")
host = torch.empty(n, dtype=torch.complex128).pin_memory()
dev.zero_()
torch.cuda.synchronize()


def run(n_streams, chunk_elems):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i, lo in enumerate(range(0, n, chunk_elems)):
        hi = min(n, lo + chunk_elems)
        with torch.cuda.stream(streams[i % n_streams]):
            host[lo:hi].copy_(dev[lo:hi], non_blocking=True)
    torch.cuda.synchronize()
    return n * 16 / (time.perf_counter() - t0) / 1e9


for chunk_mb in (64, 256, 1024):
    ce = chunk_mb * (1 << 20) // 16
    for ns in (1, 2, 4):
        run(ns, ce)
        print("chunk %4d MB, %d stream(s): %.1f GB/s" % (chunk_mb, ns, max(run(ns, ce) for _ in range(3))))
