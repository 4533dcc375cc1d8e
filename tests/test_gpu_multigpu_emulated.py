"""Row (e): the sharded-MSM path on ONE GPU. Each MSM handles the point range of `rank` and exchanges partial sums
through the all-gather callback; the ranks are run one after the other (never as concurrent waiting kernels) with a
callback that replays the other rank's recorded partials."""
import numpy as np
import pytest

import b200zk
import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("tables", [True, False])
def test_sharded_msm_equals_full_msm(ctx, world, tables):
    k = 12
    n = 1 << k
    ctx.set_msm_tables(tables)
    ctx.srs_setup(k)
    rng = np.random.default_rng(world)
    a = O.random_fr(rng, n)
    a[5] = 0
    full = ctx.msm(a, 1)
    g, gl = ctx.srs_download()
    assert np.array_equal(full, O.msm(a, gl))
    recorded = {}
    try:
        for rank in range(world):
            def fn(data, rank=rank):
                recorded[rank] = data
                parts = [recorded.get(r, bytes(len(data))) for r in range(world)]  # zeros = XYZZ identity
                return b"".join(parts)

            ctx.set_allgather(rank, world, fn)
            out = ctx.msm(a, 1)
        assert np.array_equal(out, full)  # the last rank saw every partial
        # partial of each rank alone = MSM over its point range
        per = n // world
        for rank in range(world):
            lo, hi = per * rank, (n if rank == world - 1 else per * (rank + 1))
            ctx.set_allgather(rank, world, lambda data, rank=rank: b"".join(data if r == rank else bytes(len(data)) for r in range(world)))
            assert np.array_equal(ctx.msm(a, 1), O.msm(a[lo:hi], gl[lo:hi])), rank
    finally:
        ctx.set_allgather(0, 1, None)
        ctx.set_msm_tables(True)


def test_tables_and_generic_paths_agree(ctx):
    k = 13
    rng = np.random.default_rng(1)
    a = O.random_fr(rng, 1 << k)
    small = O.fr_array([int(v) for v in rng.integers(0, 1 << 12, size=1 << k)])
    try:
        ctx.set_msm_tables(True)
        ctx.srs_setup(k)
        with_tab = [ctx.msm(a, 0), ctx.msm(a, 1), ctx.msm(small, 1), ctx.msm(a[:3000], 0)]
        ctx.set_msm_tables(False)
        without = [ctx.msm(a, 0), ctx.msm(a, 1), ctx.msm(small, 1), ctx.msm(a[:3000], 0)]
    finally:
        ctx.set_msm_tables(True)
    for x, y in zip(with_tab, without):
        assert np.array_equal(x, y)
    g, gl = ctx.srs_download()
    assert np.array_equal(with_tab[1], O.msm(a, gl))


def test_column_dealt_batch(ctx):
    """A batch of q·world + rem columns: the first q·world are dealt out by column (each rank commits its own over the full
    point range), the last rem are split by point range and every rank contributes a partial sum."""
    import ctypes

    k, world = 12, 2
    n = 1 << k
    ctx.srs_setup(k)
    rng = np.random.default_rng(4)
    cols = [O.random_fr(rng, n) for _ in range(5)]
    ptrs = []
    for c in cols:
        p = ctx.dev_alloc(32 * n)
        ctx.h2d(p, c)
        ptrs.append(p)
    want = ctx.msm_batch_dev(ptrs, n, 1)
    got = {}
    try:
        for rank in range(world):
            ctx.set_allgather(rank, world, lambda data, rank=rank: b"".join(data if r == rank else bytes(len(data)) for r in range(world)))
            got[rank] = ctx.msm_batch_dev(ptrs, n, 1)
    finally:
        ctx.set_allgather(0, 1, None)
    for j in range(4):
        assert np.array_equal(got[j % world][j], want[j]), j          # the owner produced the right commitment
        assert not got[1 - j % world][j].any()                        # the other rank left it to the exchange
    # the fifth column: each rank alone yields the sum over its half of the points; the halves add up to the commitment
    parts = np.ascontiguousarray(np.stack([got[0][4], got[1][4]]))
    assert parts[0].any() and parts[1].any() and not np.array_equal(parts[0], want[4])
    total = np.zeros(8, dtype=np.uint64)
    assert b200zk.lib().b200zk_g1_sum_host(parts.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(2), total.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.array_equal(total, want[4])
    for p in ptrs:
        ctx.dev_free(p)
