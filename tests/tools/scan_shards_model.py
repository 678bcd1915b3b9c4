"""CPU scan of a wavenumber shard of the bench sweep with the numpy model of the batched kernel
(tests/models/partition_pcr_model.py): which systems would the unpivoted elimination leave with a raw backward error
above 1e-10, and does ONE refinement step bring them to LAPACK quality (the acceptance check of
fem1d.repair_flagged)?  No GPU.  ~1.5 ms per system: split the 65,536 systems of a shard over the cores.

    python tests/tools/scan_shards_model.py WORLD RANK [first] [last]        # rank RANK of a WORLD-rank job: ks[RANK::WORLD]

Round-1 result (WORLD = 1, 2, 4, every rank): the only such wavenumbers sit at k = 806.84 (one resonance of the
8-row / 2048-row partition); worst raw backward error 4.6e-8 (WORLD = 4, RANK = 1), 9.2e-9 (WORLD = 2, RANK = 0: the
system the GPU kernel flags as a breakdown); after one refinement step every one of them is at 4e-17 .. 9e-17.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from models.partition_pcr_model import solve  # noqa: E402
from oracle import banded1d  # noqa: E402  (test infrastructure: lives under tests/ because it uses the oracle)

N, N_K = 4096, 65536
world, rank = int(sys.argv[1]), int(sys.argv[2])
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
last = int(sys.argv[4]) if len(sys.argv) > 4 else N_K
ks = np.linspace(1.0, 1000.0, N_K * world)[rank::world][first:last]
t0, found = time.time(), 0
for i, k in enumerate(ks):
    dl, d, du, b = banded1d.assemble_bands(N, float(k), 1.0 / N)
    with np.errstate(all="ignore"):
        u = solve(dl, d, du, b)
    be = banded1d.residual_metrics(dl, d, du, b, u)[0]
    if not np.isfinite(be) or be > 1e-10:
        r = b - d * u
        r[1:] -= dl[1:] * u[:-1]
        r[:-1] -= du[:-1] * u[1:]
        with np.errstate(all="ignore"):
            e = solve(dl, d, du, r)
        be2 = banded1d.residual_metrics(dl, d, du, b, u + e)[0]
        found += 1
        print("system %d  k = %r  raw backward error %.2e  after one refinement step %.2e" % (first + i, float(k), be, be2),
              flush=True)
print("systems [%d, %d) of rank %d / %d: %d above 1e-10, %.1f s" % (first, last, rank, world, found, time.time() - t0))
