"""GPU parity (rows A, C, D of SURVEY.md §8): libb200zk through its C ABI vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def edge_values(mod):
    vals = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, 1 << 64, (1 << 128) - 1, (1 << 253) + 12345,
            0xFFFFFFFF00000001, (1 << 254) % mod]
    return [v % mod for v in vals]


@pytest.mark.parametrize("field,mod", [(0, O.R_MOD), (1, O.Q_MOD)])
def test_field_ops_bit_exact(ctx, field, mod):
    rng = np.random.default_rng(7 + field)
    ev = edge_values(mod)
    pairs = [(a, b) for a in ev for b in ev]
    a_int = [p[0] for p in pairs] + [int.from_bytes(rng.bytes(40), "little") % mod for _ in range(20000)]
    b_int = [p[1] for p in pairs] + [int.from_bytes(rng.bytes(40), "little") % mod for _ in range(20000)]
    # raw limbs are arbitrary canonical values: the Montgomery map is a bijection, so test on raw limbs directly
    a = np.array([[(v >> (64 * i)) & (2**64 - 1) for i in range(4)] for v in a_int], dtype=np.uint64)
    b = np.array([[(v >> (64 * i)) & (2**64 - 1) for i in range(4)] for v in b_int], dtype=np.uint64)
    rinv = pow(1 << 256, -1, mod)
    got_add = ctx.field_vec_op(field, 0, a, b)
    got_sub = ctx.field_vec_op(field, 1, a, b)
    got_mul = ctx.field_vec_op(field, 2, a, b)
    got_neg = ctx.field_vec_op(field, 4, a)
    got_from = ctx.field_vec_op(field, 5, a)
    got_to = ctx.field_vec_op(field, 6, a)
    for i in range(len(a_int)):
        x, y = a_int[i], b_int[i]
        assert O.limbs_to_int(got_add[i]) == (x + y) % mod
        assert O.limbs_to_int(got_sub[i]) == (x - y) % mod
        assert O.limbs_to_int(got_mul[i]) == x * y * rinv % mod, (hex(x), hex(y))
        assert O.limbs_to_int(got_neg[i]) == (-x) % mod
        assert O.limbs_to_int(got_from[i]) == x * rinv % mod
        assert O.limbs_to_int(got_to[i]) == (x << 256) % mod
    inv = ctx.field_vec_op(field, 3, a[:300])
    for i in range(300):
        x = a_int[i] * rinv % mod  # value represented
        want = 0 if x == 0 else pow(x, -1, mod) * (1 << 256) % mod
        assert O.limbs_to_int(inv[i]) == want


def test_field_mul_matches_oracle_bulk(ctx):
    rng = np.random.default_rng(3)
    a = O.random_fr(rng, 1 << 16)
    b = O.random_fr(rng, 1 << 16)
    got = ctx.field_vec_op(0, 2, a, b)
    for i in range(0, 1 << 16, 97):
        assert np.array_equal(got[i], O.field_op(0, 2, a[i], b[i]))


def test_g1_ops_match_oracle(ctx):
    rng = np.random.default_rng(11)
    G = O.g1_generator()
    n = 64
    scalars = O.random_fr(rng, n)
    scalars[0] = O.to_mont(0)
    scalars[1] = O.to_mont(1)
    scalars[2] = O.to_mont(O.R_MOD - 1)
    pts = ctx.g1_vec_op(1, np.tile(G, (n, 1)), scalars)
    for i in range(n):
        assert np.array_equal(pts[i], O.g1_mul(G, scalars[i])), i
    # additions incl. P + P, P + (-P), P + 0, 0 + P
    a = pts.copy()
    b = np.roll(pts, 1, axis=0).copy()
    b[5] = a[5]  # doubling through the add path
    neg = a[6].copy()
    neg[4:] = O.field_op(1, 4, a[6][4:])
    b[6] = neg  # P + (-P) = identity
    b[7] = 0  # P + 0
    a[8] = 0  # 0 + Q
    got = ctx.g1_vec_op(0, a, b)
    for i in range(n):
        assert np.array_equal(got[i], O.g1_add(a[i], b[i])), i
    assert not got[6].any()
    dbl = ctx.g1_vec_op(2, pts)
    for i in range(n):
        assert np.array_equal(dbl[i], O.g1_add(pts[i], pts[i])), i
    # EIP-196 known answer: 2·(1,2)
    two = ctx.g1_vec_op(2, G.reshape(1, 8))[0]
    assert O.g1_affine_ints(two) == (
        1368015179489954701390400359078579693043519447331113978918064868415326638035,
        9918110051302171585080402603319702774565515993150576347155970296011118125764,
    )


@pytest.mark.parametrize("log_n", [1, 2, 3, 5, 8, 9, 11, 13, 16, 17, 18])
def test_ntt_matches_best_fft(ctx, log_n):
    rng = np.random.default_rng(log_n)
    a = O.random_fr(rng, 1 << log_n)
    for which in (0, 1):  # the domain generator and its inverse
        omega = O.domain_constant(log_n, which)
        got = ctx.ntt(a, log_n, omega)
        want = O.best_fft(a, log_n, omega)
        assert np.array_equal(got, want), (log_n, which)


def test_ntt_arbitrary_root_and_roundtrip(ctx):
    rng = np.random.default_rng(5)
    log_n = 12
    a = O.random_fr(rng, 1 << log_n)
    omega = O.domain_constant(log_n, 0)
    w5 = omega
    for _ in range(4):  # omega^5 is another primitive root (not the domain's generator)
        w5 = O.field_op(0, 2, w5, omega)
    got = ctx.ntt(a, log_n, w5)
    assert np.array_equal(got, O.best_fft(a, log_n, w5))
    # size-independent property: NTT followed by lagrange_to_coeff is the identity
    back = ctx.lagrange_to_coeff(log_n, ctx.ntt(a, log_n, omega))
    assert np.array_equal(back, a)


@pytest.mark.parametrize("k", [3, 6, 10, 14, 16])
def test_domain_transforms_match_oracle(ctx, k):
    rng = np.random.default_rng(100 + k)
    a = O.random_fr(rng, 1 << k)
    assert np.array_equal(ctx.lagrange_to_coeff(k, a), O.lagrange_to_coeff(k, a))
    ext = ctx.coeff_to_extended(k, a)
    assert np.array_equal(ext, O.coeff_to_extended(k, a))
    e = O.random_fr(rng, 4 << k)
    assert np.array_equal(ctx.extended_to_coeff(k, e), O.extended_to_coeff(k, e))
    # round trip: extended_to_coeff(coeff_to_extended(p)) == p padded with zeros
    back = ctx.extended_to_coeff(k, ext)
    assert np.array_equal(back[: 1 << k], a)
    assert not back[1 << k :].any()


def test_ntt_full_size_properties(ctx):
    """BASELINE sizes (2^20 / 2^22) through size-independent properties: linearity, inverse round trip, and a
    spot check of single outputs against a direct evaluation by the oracle."""
    rng = np.random.default_rng(42)
    for log_n in (20, 22):
        n = 1 << log_n
        a = O.random_fr(rng, n)
        omega = O.domain_constant(log_n, 0)
        fa = ctx.ntt(a, log_n, omega)
        # out[j] = sum a[i] w^(ij) = a(w^j): check three outputs with the oracle's Horner evaluation
        for j in (0, 1, n - 3):
            wj = O.to_mont(pow(O.from_mont(omega), j, O.R_MOD))
            assert np.array_equal(fa[j], O.eval_polynomial(a, wj)), (log_n, j)
        back = ctx.lagrange_to_coeff(log_n, fa)
        assert np.array_equal(back, a)
