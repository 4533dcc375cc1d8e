"""GPU parity, row B of SURVEY.md §8: b200zk_msm* vs the oracle's best_multiexp restatement (bit-exact affine)."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

K = 14


@pytest.fixture(scope="module")
def srs(ctx):
    params = O.Params.setup(K)  # ChaCha20Rng::from_seed([0;32]) like halo2-base gen_srs
    s, g, gl = params.get()
    ctx.srs_load(K, g, gl)
    return params, g, gl


def scalar_sets(rng, n, k):
    full = O.random_fr(rng, n)
    small = O.fr_array([int(v) for v in rng.integers(0, 1 << (k - 1), size=n)])
    advice = full.copy()
    kinds = rng.integers(0, 100, size=n)
    for i in range(n):
        if kinds[i] < 40:
            advice[i] = O.to_mont(int(rng.integers(0, 1 << 21)))
        elif kinds[i] < 75:
            advice[i] = O.to_mont(int.from_bytes(rng.bytes(10), "little"))
        elif kinds[i] < 80:
            advice[i] = O.to_mont(int(kinds[i]) & 1)
    sorted_small = O.fr_array(sorted(int(v) for v in rng.integers(0, 1 << (k - 1), size=n)))
    return {"uniform": full, "small": small, "advice_like": advice, "sorted_lookup": sorted_small}


@pytest.mark.parametrize("n", [1, 2, 3, 31, 100, 1 << 10, (1 << 12) + 17, 1 << K])
def test_msm_matches_oracle_sizes(ctx, srs, n):
    params, g, gl = srs
    rng = np.random.default_rng(n)
    scalars = O.random_fr(rng, n)
    for basis, bases in ((0, g), (1, gl)):
        got = ctx.msm(scalars, basis)
        want = O.msm(scalars, bases[:n])
        assert np.array_equal(got, want), (n, basis)


def test_msm_scalar_distributions(ctx, srs):
    params, g, gl = srs
    rng = np.random.default_rng(99)
    n = 1 << K
    for name, scalars in scalar_sets(rng, n, K).items():
        got = ctx.msm(scalars, 1)
        want = O.msm(scalars, gl)
        assert np.array_equal(got, want), name


def test_msm_edge_cases(ctx, srs):
    params, g, gl = srs
    n = 1 << 11
    zeros = np.zeros((n, 4), dtype=np.uint64)
    assert not ctx.msm(zeros, 0).any()  # identity = (0,0)
    assert not ctx.msm(np.zeros((0, 4), dtype=np.uint64), 0).any()  # empty input
    ones = np.tile(O.to_mont(1), (n, 1))
    assert np.array_equal(ctx.msm(ones, 0), O.msm(ones, g[:n]))
    minus1 = np.tile(O.to_mont(O.R_MOD - 1), (n, 1))
    assert np.array_equal(ctx.msm(minus1, 0), O.msm(minus1, g[:n]))
    # all scalars equal and large: one hot bucket per window
    hot = np.tile(O.to_mont(0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % O.R_MOD), (n, 1))
    assert np.array_equal(ctx.msm(hot, 0), O.msm(hot, g[:n]))


def test_msm_repeated_and_identity_bases(ctx):
    """best_multiexp with caller bases: equal points (forces the doubling branch), opposite points, identity points."""
    rng = np.random.default_rng(5)
    n = 600
    G = O.g1_generator()
    P = O.g1_mul(G, O.to_mont(12345))
    negP = P.copy()
    negP[4:] = O.field_op(1, 4, P[4:])
    bases = np.tile(P, (n, 1))
    bases[::3] = negP
    bases[5::7] = 0  # identity
    scalars = O.fr_array([int(v) for v in rng.integers(1, 40, size=n)])  # tiny scalars: many collisions per bucket
    got = ctx.msm_bases(scalars, bases)
    want = O.msm(scalars, bases, naive=True)
    assert np.array_equal(got, want)
    assert np.array_equal(O.msm(scalars, bases), want)  # oracle Pippenger vs oracle double-and-add
    scalars2 = O.random_fr(rng, n)
    assert np.array_equal(ctx.msm_bases(scalars2, bases), O.msm(scalars2, bases, naive=True))


def test_msm_linearity_full_size(ctx):
    """Size-independent property at a BASELINE size (2^20): MSM(a) + MSM(b) == MSM(a + b), and a known-trapdoor check
    commit(p) == p(s)·G, with bases s^i·G produced by the oracle."""
    k = 20
    rng = np.random.default_rng(1)
    params = O.Params.setup(k)
    s, g, gl = params.get()
    ctx.srs_load(k, g, gl)
    n = 1 << k
    a = O.random_fr(rng, n)
    b = O.random_fr(rng, n)
    ab = ctx.field_vec_op(0, 0, a, b)
    ca, cb, cab = ctx.msm(a, 0), ctx.msm(b, 0), ctx.msm(ab, 0)
    assert np.array_equal(O.g1_add(ca, cb), cab)
    ps = O.eval_polynomial(a, s)
    assert np.array_equal(ca, O.g1_mul(O.g1_generator(), ps))
    # Lagrange basis ties to the monomial basis: commit_lagrange(evals) == commit(coeffs)
    coeffs = ctx.lagrange_to_coeff(k, a)
    assert np.array_equal(ctx.msm(a, 1), ctx.msm(coeffs, 0))


def test_msm_batch_matches_single(ctx):
    k = 13
    n = 1 << k
    ctx.srs_setup(k)
    rng = np.random.default_rng(8)
    cols = [O.random_fr(rng, n), O.fr_array([int(v) for v in rng.integers(0, 1 << 12, size=n)]), np.zeros((n, 4), dtype=np.uint64), O.random_fr(rng, n)]
    ptrs = []
    for c in cols:
        p = ctx.dev_alloc(32 * n)
        ctx.h2d(p, c)
        ptrs.append(p)
    for basis in (0, 1):
        got = ctx.msm_batch_dev(ptrs, n, basis)
        for i, c in enumerate(cols):
            assert np.array_equal(got[i], ctx.msm(c, basis)), (basis, i)
    g, gl = ctx.srs_download()
    assert np.array_equal(ctx.msm_batch_dev(ptrs, n, 1)[0], O.msm(cols[0], gl))
    for p in ptrs:
        ctx.dev_free(p)


@pytest.mark.parametrize("rounds", [1, 2, 3, 5])
def test_batched_affine_rounds_are_bit_exact(ctx, srs, rounds):
    """The batched-affine pre-reduction (b200zk_set_msm_affine_rounds; measured slower than XYZZ alone, hence off by default)
    must give the same commitments as the default path for every scalar distribution, alone and in a batch."""
    params, g, gl = srs
    ctx.srs_load(K, g, gl)  # earlier tests of this module may have loaded other bases
    rng = np.random.default_rng(200 + rounds)
    n = 1 << K
    sets = scalar_sets(rng, n, K)
    want = {name: ctx.msm(sc, 1) for name, sc in sets.items()}
    try:
        ctx.set_msm_affine_rounds(rounds)
        for name, sc in sets.items():
            assert np.array_equal(ctx.msm(sc, 1), want[name]), name
        assert np.array_equal(ctx.msm(sets["uniform"], 0), O.msm(sets["uniform"], g))
        got = ctx.msm_batch(list(sets.values()), 1)
        for i, name in enumerate(sets):
            assert np.array_equal(got[i], want[name]), name
    finally:
        ctx.set_msm_affine_rounds(0)


def test_batched_affine_falls_back_on_exceptional_pairs(ctx):
    """Affine addition has no doubling and no identity: bases with repeated points (P + P inside a bucket), opposite points
    and identities must make the rounds hand the column back to the XYZZ path — same result as without the rounds."""
    k = 10
    n = 1 << k
    G = O.g1_generator()
    P5 = O.g1_mul(G, O.to_mont(5))
    neg = P5.copy()
    neg[4:] = O.field_op(1, 4, P5[4:])  # −y
    pts = np.stack([P5 if i % 3 == 0 else (neg if i % 3 == 1 else np.zeros(8, dtype=np.uint64)) for i in range(n)])
    rng = np.random.default_rng(8)
    scalars = O.random_fr(rng, n)
    try:
        ctx.srs_load(k, pts, pts)
        want = ctx.msm(scalars, 0)
        assert np.array_equal(want, O.msm(scalars, pts))
        ctx.set_msm_affine_rounds(3)
        assert np.array_equal(ctx.msm(scalars, 0), want)
        assert np.array_equal(ctx.msm(scalars, 1), want)
    finally:
        ctx.set_msm_affine_rounds(0)
