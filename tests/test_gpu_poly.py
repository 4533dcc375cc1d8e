"""GPU parity, rows E, F, I at primitive level: batch inversion, running product, Horner evaluation, synthetic division
and the lookup permutation through the C ABI vs the oracle (bit-exact), incl. zeros, ragged sizes and failure cases."""
import ctypes

import numpy as np
import pytest

import b200zk
import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 7, 64, 1000, 4097, 1 << 16, (1 << 18) + 3])
def test_batch_invert(ctx, n):
    rng = np.random.default_rng(n)
    a = O.random_fr(rng, n)
    a[::5] = 0  # zeros stay zero
    want = a.copy()
    O.lib().oracle_fr_batch_invert(O.ptr(want), ctypes.c_size_t(n))
    got = ctx.batch_invert(a)
    assert np.array_equal(got, want)
    assert not got[::5].any()


@pytest.mark.parametrize("n", [1, 2, 9, 2048, 2049, 5000, 1 << 16, (1 << 20) + 77])
def test_prefix_product(ctx, n):
    rng = np.random.default_rng(n)
    m = O.random_fr(rng, n)
    first = O.random_fr(rng, 1)[0]
    z = ctx.prefix_product(m, first)
    assert np.array_equal(z[0], first)
    # the recurrence everywhere via one more GPU multiply, and the head against python big-ints
    if n > 1:
        assert np.array_equal(ctx.field_vec_op(0, 2, z[:-1], m[:-1]), z[1:])
    acc = O.from_mont(first)
    for i in range(min(n, 50)):
        assert O.from_mont(z[i]) == acc
        acc = acc * O.from_mont(m[i]) % O.R_MOD


@pytest.mark.parametrize("n", [1, 5, 16, 17, 4096, 4100, 1 << 15, (1 << 18) + 1])
def test_eval_polynomial(ctx, n):
    rng = np.random.default_rng(n)
    poly = O.random_fr(rng, n)
    for point in (O.random_fr(rng, 1)[0], O.to_mont(0), O.to_mont(1)):
        assert np.array_equal(ctx.eval_polynomial(poly, point), O.eval_polynomial(poly, point))


@pytest.mark.parametrize("n", [2, 3, 8, 9, 2048, 2055, 1 << 14, (1 << 17) + 5])
def test_kate_division(ctx, n):
    rng = np.random.default_rng(n)
    a = O.random_fr(rng, n)
    for b in (O.random_fr(rng, 1)[0], O.to_mont(0), O.to_mont(1)):
        want = np.zeros((n - 1, 4), dtype=np.uint64)
        O.lib().oracle_kate_division(O.ptr(a), ctypes.c_size_t(n), O.ptr(np.ascontiguousarray(b)), O.ptr(want))
        assert np.array_equal(ctx.kate_division(a, b), want)


@pytest.mark.parametrize("k,kind", [(6, "uniform"), (9, "hot"), (12, "uniform"), (12, "single"), (14, "sparse")])
def test_permute_expression_pair(ctx, k, kind):
    rng = np.random.default_rng(k)
    n = 1 << k
    table = O.fr_array([i if i < n // 2 else 0 for i in range(n)])
    if kind == "uniform":
        vals = rng.integers(0, n // 2, size=n)
    elif kind == "hot":
        vals = np.where(rng.integers(0, 4, size=n) == 0, rng.integers(0, n // 2, size=n), 3)
    elif kind == "single":
        vals = np.full(n, 0)
    else:
        vals = rng.integers(0, 8, size=n) * (n // 16)
    inp = O.fr_array([int(v) for v in vals])
    a_want = np.zeros((n, 4), dtype=np.uint64)
    s_want = np.zeros((n, 4), dtype=np.uint64)
    assert O.lib().oracle_permute_expression_pair(k, O.ptr(inp), O.ptr(table), O.ptr(a_want), O.ptr(s_want)) == 1
    a, s = ctx.permute_expression_pair(k, inp, table)
    u = n - 7
    assert np.array_equal(a, a_want[:u]) and np.array_equal(s, s_want[:u])
    bad = inp.copy()
    bad[n // 3] = O.to_mont(n // 2 + 1)  # not in the table
    with pytest.raises(b200zk.B200zkError) as e:
        ctx.permute_expression_pair(k, bad, table)
    assert e.value.code == b200zk.ESYNTH
    # a TABLE value outside [0, n) is an unsupported table for this counting sort (EINVAL), not a constraint failure;
    # an input value outside [0, n) is missing from every supported table (ESYNTH)
    big_table = table.copy()
    big_table[n // 5] = O.to_mont(n + 3)
    with pytest.raises(b200zk.B200zkError) as e:
        ctx.permute_expression_pair(k, inp, big_table)
    assert e.value.code == b200zk.EINVAL
    big = inp.copy()
    big[n // 5] = O.to_mont(O.R_MOD - 2)
    with pytest.raises(b200zk.B200zkError) as e:
        ctx.permute_expression_pair(k, big, table)
    assert e.value.code == b200zk.ESYNTH
