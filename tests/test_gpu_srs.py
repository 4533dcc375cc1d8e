"""GPU parity, row K: ParamsKZG::setup on the device vs the oracle (same ChaCha20 seed -> same trapdoor -> same bases)."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", [4, 9, 13])
def test_srs_setup_matches_oracle(ctx, k):
    s = ctx.srs_setup(k)
    params = O.Params.setup(k)
    so, g, gl = params.get()
    assert np.array_equal(s, so)
    dg, dgl = ctx.srs_download()
    assert np.array_equal(dg, g)
    assert np.array_equal(dgl, gl)


def test_srs_lagrange_is_group_ifft_of_monomial(ctx):
    """Upstream derives g_lagrange by a group inverse FFT of g: check the device bases against that derivation."""
    k = 6
    ctx.srs_setup(k, seed=bytes(range(32)))
    dg, dgl = ctx.srs_download()
    params = O.Params.load(k, O.to_mont(1), dg, dgl)
    assert np.array_equal(params.lagrange_via_group_fft(), dgl)
    # generator first: g[0] = G
    assert np.array_equal(dg[0], O.g1_generator())


def test_commit_with_device_srs_is_p_of_s_times_g(ctx):
    k = 12
    s = ctx.srs_setup(k)
    rng = np.random.default_rng(2)
    a = O.random_fr(rng, 1 << k)
    assert np.array_equal(ctx.msm(a, 0), O.g1_mul(O.g1_generator(), O.eval_polynomial(a, s)))
    assert np.array_equal(ctx.msm(a, 1), ctx.msm(ctx.lagrange_to_coeff(k, a), 0))


Q = O.Q_MOD


def _fq2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def _g2_scalar_mul(pt, k):
    """affine double-and-add over Fq2 in Python big-ints (independent check of the host G2 code)"""
    def inv(a):
        d = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
        return (a[0] * d % Q, -a[1] * d % Q)

    def add(p, q):
        if p is None:
            return q
        if q is None:
            return p
        if p[0] == q[0]:
            if p[1] != q[1]:
                return None
            lam = _fq2_mul(_fq2_mul((3, 0), _fq2_mul(p[0], p[0])), inv(((2 * p[1][0]) % Q, (2 * p[1][1]) % Q)))
        else:
            lam = _fq2_mul(((q[1][0] - p[1][0]) % Q, (q[1][1] - p[1][1]) % Q), inv(((q[0][0] - p[0][0]) % Q, (q[0][1] - p[0][1]) % Q)))
        l2 = _fq2_mul(lam, lam)
        x3 = ((l2[0] - p[0][0] - q[0][0]) % Q, (l2[1] - p[0][1] - q[0][1]) % Q)
        t = _fq2_mul(lam, ((p[0][0] - x3[0]) % Q, (p[0][1] - x3[1]) % Q))
        return (x3, ((t[0] - p[1][0]) % Q, (t[1] - p[1][1]) % Q))

    acc = None
    for bit in bin(k)[2:]:
        acc = add(acc, acc)
        if bit == "1":
            acc = add(acc, pt)
    return acc


def test_srs_file_round_trip_and_g2(ctx):
    k = 7
    n = 1 << k
    s = ctx.srs_setup(k)
    g, gl = ctx.srs_download()
    raw = ctx.srs_write(0)
    assert len(raw) == 4 + 2 * n * 64 + 256 and raw[:4] == k.to_bytes(4, "little")
    assert raw[4 : 4 + 64 * n] == g.tobytes() and raw[4 + 64 * n : 4 + 128 * n] == gl.tobytes()
    # G2 section: EIP-197 generator and s·g2, raw Montgomery limbs (x.c0, x.c1, y.c0, y.c1)
    g2 = np.frombuffer(raw[-256:], dtype=np.uint64).reshape(2, 4, 4)
    gen = tuple(O.from_mont(g2[0][i], Q) for i in range(4))
    assert gen == (10857046999023057135944570762232829481370756359578518086990519993285655852781,
                   11559732032986387107991004021392285783925812861821192530917403151452391805634,
                   8495653923123431417604973247489272438418190587263600148770280649306958101930,
                   4082367875863433681332203403145435568316851327593401208105741076214120093531)
    sg2 = tuple(O.from_mont(g2[1][i], Q) for i in range(4))
    want = _g2_scalar_mul(((gen[0], gen[1]), (gen[2], gen[3])), O.from_mont(s))
    assert sg2 == (want[0][0], want[0][1], want[1][0], want[1][1])
    # the KZG relation of the file, trapdoor-free, with the oracle's pairing: e(g[i+1], g2) == e(g[i], s·g2)
    g2_raw = np.frombuffer(raw[-256:], dtype=np.uint64).copy()
    for i in (0, n - 2):
        neg = g[i].copy()
        neg[4:] = O.to_mont((Q - O.from_mont(g[i][4:], Q)) % Q, Q)
        assert O.lib().oracle_pairing_product_is_one(O.ptr(np.ascontiguousarray(g[i + 1])), O.ptr(g2_raw[:16]), O.ptr(neg), O.ptr(g2_raw[16:])) == 1
    assert O.lib().oracle_pairing_product_is_one(O.ptr(np.ascontiguousarray(g[2])), O.ptr(g2_raw[:16]), O.ptr(neg), O.ptr(g2_raw[16:])) == 0
    # read back (RawBytes) into the context: same bases, same commitments
    rng = np.random.default_rng(1)
    a = O.random_fr(rng, n)
    c0 = ctx.msm(a, 1)
    ctx.srs_read(raw, 0)
    g_b, gl_b = ctx.srs_download()
    assert np.array_equal(g_b, g) and np.array_equal(gl_b, gl) and np.array_equal(ctx.msm(a, 1), c0)
    assert ctx.srs_write(0) == raw
    # Processed (compressed G1): build the file from the oracle's encoder, read it, and write it back identically
    enc = np.empty(32, dtype=np.uint8)
    body = b""
    for pts in (g, gl):
        for p in pts:
            O.lib().oracle_g1_to_bytes(O.ptr(np.ascontiguousarray(p)), O.ptr(enc))
            body += enc.tobytes()
    processed = k.to_bytes(4, "little") + body + bytes(range(128))  # opaque G2 bytes are preserved
    ctx.srs_read(processed, 1)
    g_c, gl_c = ctx.srs_download()
    assert np.array_equal(g_c, g) and np.array_equal(gl_c, gl)
    assert ctx.srs_write(1) == processed
    # corrupt files are rejected
    bad = bytearray(raw); bad[4 + 70] ^= 1
    with pytest.raises(Exception):
        ctx.srs_read(bytes(bad), 0)
    with pytest.raises(Exception):
        ctx.srs_read(raw[:-1], 0)
    badp = bytearray(processed); badp[4 + 31] |= 0x3F
    with pytest.raises(Exception):
        ctx.srs_read(bytes(badp), 1)


def test_gen_srs_caches_params_on_disk(ctx, tmp_path):
    """halo2-base gen_srs: the first call generates and writes params/kzg_bn254_<k>.srs, the second reads it back; both
    leave the same bases on the device, and commitments agree."""
    k = 10
    d = str(tmp_path / "params")
    assert ctx.gen_srs(k, d) is False
    g1, gl1 = ctx.srs_download()
    assert os.path.exists(os.path.join(d, f"kzg_bn254_{k}.srs"))
    ctx.srs_setup(k, seed=bytes([1] * 32))  # something else in between
    assert ctx.gen_srs(k, d) is True
    g2, gl2 = ctx.srs_download()
    assert np.array_equal(g1, g2) and np.array_equal(gl1, gl2)
    a = O.random_fr(np.random.default_rng(3), 1 << k)
    assert np.array_equal(ctx.msm(a, 1), O.msm(a, gl1))
