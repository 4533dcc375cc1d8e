"""Pins the CPU oracle (oracle/*.hpp) with known-answer tests that do NOT come from the oracle itself:
Python big-int arithmetic, hashlib BLAKE2b, RFC 7539 ChaCha20 vectors, the EIP-196 doubling vector, naive DFT /
double-and-add, and algebraic identities. The reference repo holds no golden vector for this
path, so these are what stands between the oracle and "parity unpinned" (DESIGN.md)."""
import ctypes
import hashlib
import random

import numpy as np
import pytest

import oracle_lib as O

R, Q = O.R_MOD, O.Q_MOD


def test_field_constants_rederived():
    p = np.empty(4, dtype=np.uint64); inv = np.empty(1, dtype=np.uint64)
    r1 = np.empty(4, dtype=np.uint64); r2 = np.empty(4, dtype=np.uint64); r3 = np.empty(4, dtype=np.uint64)
    for which, mod, inv_expect in ((0, R, 0xC2E1F593EFFFFFFF), (1, Q, 0x87D20782E4866389)):
        O.lib().oracle_field_params(which, O.ptr(p), O.ptr(inv), O.ptr(r1), O.ptr(r2), O.ptr(r3))
        assert O.limbs_to_int(p) == mod
        assert int(inv[0]) == inv_expect == (-pow(mod, -1, 1 << 64)) % (1 << 64)
        assert O.limbs_to_int(r1) == (1 << 256) % mod
        assert O.limbs_to_int(r2) == (1 << 512) % mod
        assert O.limbs_to_int(r3) == (1 << 768) % mod
    # SURVEY.md Appendix A.6 values
    assert (1 << 512) % R == 0x0216D0B17F4E44A58C49833D53BB808553FE3AB1E35C59E31BB8E645AE216DA7
    root = np.empty(4, dtype=np.uint64); zeta = np.empty(4, dtype=np.uint64); delta = np.empty(4, dtype=np.uint64)
    O.lib().oracle_fr_constants(O.ptr(root), O.ptr(zeta), O.ptr(delta))
    w, z, d = O.from_mont(root), O.from_mont(zeta), O.from_mont(delta)
    assert w == pow(7, (R - 1) >> 28, R) and pow(w, 1 << 28, R) == 1 and pow(w, 1 << 27, R) != 1
    assert pow(z, 3, R) == 1 and z != 1
    assert d == pow(7, 1 << 28, R)


@pytest.mark.parametrize("which,mod", [(0, R), (1, Q)])
def test_field_ops_vs_python(which, mod):
    rnd = random.Random(which)
    cases = [(0, 0), (1, mod - 1), (mod - 1, mod - 1), (2, (mod + 1) // 2)] + [(rnd.randrange(mod), rnd.randrange(mod)) for _ in range(300)]
    for a, b in cases:
        A, B = O.to_mont(a, mod), O.to_mont(b, mod)
        assert O.from_mont(O.field_op(which, 0, A, B), mod) == (a + b) % mod
        assert O.from_mont(O.field_op(which, 1, A, B), mod) == (a - b) % mod
        assert O.from_mont(O.field_op(which, 2, A, B), mod) == a * b % mod
        assert O.from_mont(O.field_op(which, 4, A), mod) == (-a) % mod
        if a:
            assert O.from_mont(O.field_op(which, 3, A), mod) == pow(a, -1, mod)
        assert O.limbs_to_int(O.field_op(which, 5, A)) == a
    for _ in range(50):
        w = rnd.randrange(1 << 512)
        words = np.array([(w >> (64 * i)) & (2**64 - 1) for i in range(8)], dtype=np.uint64)
        out = np.empty(4, dtype=np.uint64)
        O.lib().oracle_fr_from_u512(O.ptr(words), O.ptr(out))
        assert O.from_mont(out) == w % R


def test_blake2b_matches_hashlib():
    out = np.empty(64, dtype=np.uint8)
    for n in (0, 1, 64, 127, 128, 129, 1000):
        data = bytes((i * 7 + 3) & 0xFF for i in range(n))
        for person in (b"Halo2-Transcript", b"Halo2-Verify-Key"):
            O.lib().oracle_blake2b(person, data, len(data), O.ptr(out))
            assert out.tobytes() == hashlib.blake2b(data, digest_size=64, person=person).digest()


def test_transcript_matches_hashlib_model():
    """Blake2bWrite: prefix bytes 0/1/2, 64-byte digest of a state clone reduced mod r, compressed points in the proof."""
    G = O.g1_generator()
    P2 = O.g1_add(G, G)
    s = O.to_mont(0x1234567890ABCDEF << 100)
    kinds = np.array([2, 1, 0, 1, 0, 0], dtype=np.uint8)  # scalar, point, squeeze, point, squeeze, squeeze
    inp = np.concatenate([s, G, P2]).astype(np.uint64)
    out = np.empty(12, dtype=np.uint64)
    proof = np.empty(96, dtype=np.uint8)
    n = O.lib().oracle_transcript_script(O.ptr(kinds), len(kinds), O.ptr(inp), O.ptr(out), O.ptr(proof))
    h = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
    le = lambda v: v.to_bytes(32, "little")
    gx, gy = O.g1_affine_ints(G)
    px, py = O.g1_affine_ints(P2)
    expect = []
    h.update(b"\x02" + le(O.from_mont(s)))
    h.update(b"\x01" + le(gx) + le(gy))
    h.update(b"\x00"); expect.append(int.from_bytes(h.copy().digest(), "little") % R)
    h.update(b"\x01" + le(px) + le(py))
    h.update(b"\x00"); expect.append(int.from_bytes(h.copy().digest(), "little") % R)
    h.update(b"\x00"); expect.append(int.from_bytes(h.copy().digest(), "little") % R)
    assert [O.from_mont(out[4 * i : 4 * i + 4]) for i in range(3)] == expect
    comp = lambda x, y: (x | ((y & 1) << 254)).to_bytes(32, "little")
    assert n == 96 and proof.tobytes() == le(O.from_mont(s)) + comp(gx, gy) + comp(px, py)


def test_chacha_vectors_and_seed_expansion():
    w = np.empty(16, dtype=np.uint32)
    O.lib().oracle_chacha_words(bytes(32), 20, 16, O.ptr(w))
    assert w.tobytes().hex() == ("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                                 "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")
    # RFC 7539 §2.3.2 block function vector (key 00..1f, counter 1, nonce 00:00:00:09:00:00:00:4a:00:00:00:00)
    st = np.array([0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + [int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)]
                  + [1, 0x09000000, 0x4A000000, 0], dtype=np.uint32)
    out = np.empty(16, dtype=np.uint32)
    O.lib().oracle_chacha_block(O.ptr(st), O.ptr(out), 20)
    assert out.tobytes().hex().startswith("10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e")
    # second buffer block uses counter+1: compare with the block function directly
    w2 = np.empty(32, dtype=np.uint32)
    O.lib().oracle_chacha_words(bytes(range(32)), 12, 32, O.ptr(w2))
    st2 = st.copy(); st2[12:] = [1, 0, 0, 0]
    O.lib().oracle_chacha_block(O.ptr(st2), O.ptr(out), 12)
    assert np.array_equal(w2[16:], out)
    # rand_core seed_from_u64 = PCG32 output expansion
    seed = np.empty(32, dtype=np.uint8)
    for s0 in (0, 1, 0xDEADBEEF):
        O.lib().oracle_std_rng_seed(ctypes.c_uint64(s0), O.ptr(seed))
        state, exp = s0, b""
        for _ in range(8):
            state = (state * 6364136223846793005 + 11634580027462260723) % (1 << 64)
            xs = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
            rot = state >> 59
            exp += (((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF).to_bytes(4, "little")
        assert seed.tobytes() == exp


def test_g1_known_answers():
    G = O.g1_generator()
    assert O.lib().oracle_g1_on_curve(O.ptr(G))
    two = O.g1_add(G, G)
    assert O.g1_affine_ints(two) == (1368015179489954701390400359078579693043519447331113978918064868415326638035,
                                     9918110051302171585080402603319702774565515993150576347155970296011118125764)
    assert not O.g1_mul(G, O.to_mont(0)).any()
    assert np.array_equal(O.g1_mul(G, O.to_mont(R - 1))[:4], G[:4])  # (r-1)G = -G: same x
    a, b = 123456789123456789, R - 987654321
    assert np.array_equal(O.g1_add(O.g1_mul(G, O.to_mont(a)), O.g1_mul(G, O.to_mont(b))), O.g1_mul(G, O.to_mont((a + b) % R)))
    # compressed encoding round trip (bit 6 of byte 31 = y odd)
    enc = np.empty(32, dtype=np.uint8); dec = np.empty(8, dtype=np.uint64)
    P = O.g1_mul(G, O.to_mont(77))
    O.lib().oracle_g1_to_bytes(O.ptr(P), O.ptr(enc))
    x, y = O.g1_affine_ints(P)
    assert int.from_bytes(enc.tobytes(), "little") == x | ((y & 1) << 254)
    assert O.lib().oracle_g1_from_bytes(O.ptr(enc), O.ptr(dec)) == 1 and np.array_equal(dec, P)


def test_msm_and_fft_vs_naive():
    rng = np.random.default_rng(0)
    G = O.g1_generator()
    n = 300
    bases = np.stack([O.g1_mul(G, O.to_mont(int(v))) for v in rng.integers(1, 1 << 62, size=n)])
    scalars = O.random_fr(rng, n)
    scalars[3] = 0
    assert np.array_equal(O.msm(scalars, bases), O.msm(scalars, bases, naive=True))
    for log_n in (1, 4, 7):
        a = O.random_fr(rng, 1 << log_n)
        omega = O.domain_constant(log_n, 0)
        out = np.empty_like(a)
        O.lib().oracle_naive_dft(O.ptr(a), O.ptr(out), 1 << log_n, O.ptr(omega))
        assert np.array_equal(O.best_fft(a, log_n, omega), out)


def test_domain_and_kzg_identities():
    rng = np.random.default_rng(1)
    k = 6
    a = O.random_fr(rng, 1 << k)
    coeffs = O.lagrange_to_coeff(k, a)
    omega = O.domain_constant(k, 0)
    assert np.array_equal(O.best_fft(coeffs, k, omega), a)
    ext = O.coeff_to_extended(k, coeffs)
    back = O.extended_to_coeff(k, ext)
    assert np.array_equal(back[: 1 << k], coeffs) and not back[1 << k :].any()
    # extended evaluations are p(zeta * w_ext^i)
    zeta = 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23
    wext = O.from_mont(O.domain_constant(k, 2))
    for i in (0, 1, 5, (4 << k) - 1):
        pt = O.to_mont(zeta * pow(wext, i, R) % R)
        assert np.array_equal(ext[i], O.eval_polynomial(coeffs, pt))
    params = O.Params.setup(k)
    s, g, gl = params.get()
    assert np.array_equal(params.lagrange_via_group_fft(), gl)        # upstream's derivation of g_lagrange
    assert np.array_equal(params.commit(a, lagrange=True), params.commit(coeffs))
    assert np.array_equal(params.commit(coeffs), O.g1_mul(O.g1_generator(), O.eval_polynomial(coeffs, s)))
    # kate division: a(X) = q(X)(X - b) + a(b)
    b = O.to_mont(12345)
    q = np.zeros(((1 << k) - 1, 4), dtype=np.uint64)
    O.lib().oracle_kate_division(O.ptr(coeffs), 1 << k, O.ptr(b), O.ptr(q))
    x = O.to_mont(987654321)
    lhs = O.from_mont(O.eval_polynomial(coeffs, x))
    rhs = (O.from_mont(O.eval_polynomial(q, x)) * (987654321 - 12345) + O.from_mont(O.eval_polynomial(coeffs, b))) % R
    assert lhs == rhs


def test_permute_expression_pair_vs_python_model():
    rng = np.random.default_rng(3)
    k = 7
    n, u = 1 << k, (1 << k) - 7
    table_vals = [i if i < n // 2 else 0 for i in range(n)]
    inp_vals = [int(v) for v in rng.integers(0, 20, size=n)]
    a_out = np.zeros((n, 4), dtype=np.uint64); s_out = np.zeros((n, 4), dtype=np.uint64)
    ok = O.lib().oracle_permute_expression_pair(k, O.ptr(O.fr_array(inp_vals)), O.ptr(O.fr_array(table_vals)), O.ptr(a_out), O.ptr(s_out))
    assert ok == 1
    # python model of halo2's rule (sort, first occurrence keeps the value, leftovers ascending popped onto repeated rows from the end)
    a = sorted(inp_vals[:u]); left = {}
    for t in table_vals[:u]:
        left[t] = left.get(t, 0) + 1
    s = [0] * u; rep = []
    for row in range(u):
        if row == 0 or a[row] != a[row - 1]:
            s[row] = a[row]; left[a[row]] -= 1
        else:
            rep.append(row)
    for v in sorted(left):
        for _ in range(left[v]):
            s[rep.pop()] = v
    assert not rep
    assert O.fr_ints(a_out[:u]) == a and O.fr_ints(s_out[:u]) == s
    # an input outside the table is a synthesis failure
    bad = list(inp_vals); bad[5] = n // 2 + 3
    assert O.lib().oracle_permute_expression_pair(k, O.ptr(O.fr_array(bad)), O.ptr(O.fr_array(table_vals)), O.ptr(a_out), O.ptr(s_out)) == 0


def test_pairing_identities():
    """oracle/pairing.hpp: bilinearity in both arguments, non-degeneracy, e(P,Q)^r == 1, s·g2 stays on the twist."""
    rnd = random.Random(5)
    for _ in range(2):
        a, b = rnd.randrange(1, O.R_MOD), rnd.randrange(1, O.R_MOD)
        mask = O.lib().oracle_pairing_selfcheck(O.ptr(O.to_mont(a)), O.ptr(O.to_mont(b)))
        assert mask == 63, bin(mask)


def test_pairing_checks_the_kzg_relation_of_the_srs():
    """e(g[i+1], g2) == e(g[i], s·g2) for the oracle's ParamsKZG::setup, with s·g2 from independent Python Fq2 arithmetic."""
    from test_gpu_srs import Q, _g2_scalar_mul

    gen = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
            11559732032986387107991004021392285783925812861821192530917403151452391805634),
           (8495653923123431417604973247489272438418190587263600148770280649306958101930,
            4082367875863433681332203403145435568316851327593401208105741076214120093531))
    s, g, _ = O.Params.setup(5).get()
    sg2 = _g2_scalar_mul(gen, O.from_mont(s))
    raw = lambda pt: np.concatenate([O.to_mont(c, Q) for c in (pt[0][0], pt[0][1], pt[1][0], pt[1][1])]).astype(np.uint64)
    g2r, sg2r = raw(gen), raw(sg2)
    check = O.lib().oracle_pairing_product_is_one
    for i in (0, 30):
        neg = g[i].copy()
        neg[4:] = O.to_mont((Q - O.from_mont(g[i][4:], Q)) % Q, Q)
        assert check(O.ptr(np.ascontiguousarray(g[i + 1])), O.ptr(g2r), O.ptr(neg), O.ptr(sg2r)) == 1
    assert check(O.ptr(np.ascontiguousarray(g[2])), O.ptr(g2r), O.ptr(neg), O.ptr(sg2r)) == 0
    off = g2r.copy()
    off[0] ^= np.uint64(1)
    assert check(O.ptr(np.ascontiguousarray(g[1])), O.ptr(off), O.ptr(neg), O.ptr(sg2r)) == -1  # not on the twist
