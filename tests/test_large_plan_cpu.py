"""Host logic of the periodic recursion (csrc/fh_tridiag_large.cu, generated rows on a uniform mesh): the planner's claim
"these rows / chunks of level l are plain" is checked against a brute-force symbolic reduction that tracks, for every row
of every level, WHICH rows it was computed from.  No GPU: fh_large_uniform_plan does no device work."""
import ctypes as C

import numpy as np
import pytest

from fem_helmholtz_eqn_b200 import _lib
from models.recursive_partition_model import chunk_bounds


def _plan(n_elem, begin, n_rows):
    lib = _lib.load()
    out = np.zeros((20, 6), dtype=np.int64)
    rc = lib.fh_large_uniform_plan(n_elem, begin, n_rows, out.ctypes.data_as(C.c_void_p), 20)
    assert rc > 0, (rc, lib.fh_last_error())
    return out[:rc]


def _symbolic_levels(n_elem, begin, n_rows):
    """Label of a row = identity of the arithmetic that produced it.  Level 0: interior rows are all alike, the two
    mesh-end rows are their own; an interface row is labelled by (which end, labels of the chunk it came from)."""
    ids = {}

    def intern(key):
        return ids.setdefault(key, len(ids))

    labels = [intern("interior")] * n_rows
    if begin == 0:
        labels[0] = intern("mesh row 0")
    if begin + n_rows - 1 == n_elem:
        labels[-1] = intern("mesh row N")
    levels = [labels]
    while len(labels) > 2:
        nxt = []
        for r0, ln in chunk_bounds(len(labels)):
            key = tuple(labels[r0:r0 + ln])
            nxt += [intern(("first", key)), intern(("last", key))]
        labels = nxt
        levels.append(labels)
    return levels


CASES = [  # (n_elem, begin, n_rows): whole meshes and slabs, ragged ends in every position
    (2, 0, 3), (16, 0, 17), (17, 0, 18), (40, 0, 41), (1023, 0, 1024), (1024, 0, 1025), (1040, 0, 1041),
    (4111, 0, 4112), (16384, 0, 16385), (131088, 0, 131089), (200_000, 0, 200_001), (300_017, 0, 300_018),
    (1 << 18, 0, (1 << 18) + 1), (1 << 18, 0, 1 << 15), (1 << 18, 1 << 15, 1 << 15), (1 << 18, (1 << 18) - 4999, 5000),
    (1_000_000, 333_334, 333_333), (1_000_000, 666_667, 333_334), (1_000_000, 0, 125_001), (999_999, 500_000, 123_457),
]


@pytest.mark.parametrize("n_elem,begin,n_rows", CASES)
def test_plain_rows_and_chunks_are_what_the_planner_says(n_elem, begin, n_rows):
    plan = _plan(n_elem, begin, n_rows)
    levels = _symbolic_levels(n_elem, begin, n_rows)
    assert len(plan) == len(levels) and plan[-1][0] == 2
    for l, (rows, chunks, plo, phi, clo, chi) in enumerate(plan):
        labels = levels[l]
        assert rows == len(labels)
        if l == len(plan) - 1:
            break
        bounds = chunk_bounds(rows)
        assert chunks == len(bounds) and 0 <= clo <= chi <= chunks
        # plain rows: one label per parity
        if phi > plo:
            for par in (0, 1):
                seen = {labels[r] for r in range(plo, phi) if (r & 1) == par}
                assert len(seen) <= 1, (l, par)
        # plain chunks: 16 rows at 16 c, all of them plain -> all reduce to the same two rows
        for c in range(clo, chi):
            r0, ln = bounds[c]
            assert (r0, ln) == (16 * c, 16) and plo <= r0 and r0 + 16 <= phi
        # the next level's plain rows are exactly the interface rows of this level's plain chunks
        assert (plan[l + 1][2], plan[l + 1][3]) == (2 * clo, 2 * chi)
        # "at most two special chunks per level" (what sizes the single-CTA kernels' staging area)
        assert clo + (chunks - chi) <= 2, (l, clo, chunks - chi)
        # and the planner does not give plain chunks away: a chunk outside [clo, chi) differs from the plain one
        if chi > clo:
            plain_key = tuple(labels[16 * clo:16 * clo + 16])
            for c in list(range(0, clo)) + list(range(chi, chunks)):
                r0, ln = bounds[c]
                assert tuple(labels[r0:r0 + ln]) != plain_key, (l, c)


def test_planner_rejects_bad_ranges():
    lib = _lib.load()
    out = np.zeros((20, 6), dtype=np.int64)
    p = out.ctypes.data_as(C.c_void_p)
    assert lib.fh_large_uniform_plan(100, 0, 2, p, 20) < 0       # a slab needs more than 2 rows
    assert lib.fh_large_uniform_plan(100, 50, 60, p, 20) < 0     # beyond the mesh
    assert lib.fh_large_uniform_plan(1 << 20, 0, 1 << 20, p, 2) < 0  # output too short


@pytest.mark.parametrize("N,k", [(40, 7.0), (1040, 30.0), (4111, 100.0), (16384, 250.0), (131088, 1000.0), (300_017, 37.25)])
def test_periodic_recursion_model_solves_the_reference_system(N, k):
    """The algorithm itself, in numpy on the planner's geometry (tests/models/periodic_recursion_model.py): constants
    chain, <= 2 special chunks per level, symbolic Thomas tables -> the solution LAPACK gives for the oracle's bands."""
    from models import periodic_recursion_model as prm
    from oracle import banded1d
    for bc in (dict(), dict(bc_left="dirichlet", bc_right="dirichlet", g_left=1.0, g_right=0.5)):
        dl, d, du, b = banded1d.assemble_bands(N, k, 1.0 / N, **bc)
        plan = _plan(N, 0, N + 1)
        u = prm.solve(plan, (dl[1], d[1], du[1], b[1]), (dl[0], d[0], du[0], b[0]), (dl[N], d[N], du[N], b[N]))
        be, _ = banded1d.residual_metrics(dl, d, du, b, u)
        assert be < 1e-13, (N, k, bc, be)


@pytest.mark.parametrize("world", [2, 3, 5])
def test_periodic_recursion_model_partitioned(world):
    """The same per slab, glued by the 2P-row interface system (what distributed.solve_partitioned does over NCCL)."""
    from models import periodic_recursion_model as prm
    from models.recursive_partition_model import slab_ranges, spike_solve_interface
    from oracle import banded1d
    N, k = 100_003, 400.0
    dl, d, du, b = banded1d.assemble_bands(N, k, 1.0 / N)
    interior, row0, rowN = (dl[1], d[1], du[1], b[1]), (dl[0], d[0], du[0], b[0]), (dl[N], d[N], du[N], b[N])
    parts = []
    for begin, end in slab_ranges(N + 1, world):
        plan = _plan(N, begin, end - begin)
        parts.append(prm.reduce_slab(plan, interior, row0, rowN, has_row0=begin == 0, has_rowN=end == N + 1))
    x = spike_solve_interface([iface for iface, _ in parts])
    u = np.concatenate([backsub(x[2 * r:2 * r + 2]) for r, (_, backsub) in enumerate(parts)])
    assert banded1d.residual_metrics(dl, d, du, b, u)[0] < 1e-13
