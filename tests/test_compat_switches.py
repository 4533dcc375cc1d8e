"""SURVEY.md §8c items 1–4: the upstream details that change proof BYTES and could not be checked against halo2-axiom's
source are each ONE switch, present in the oracle and in the product with the same meaning. Both settings of every
switch must give (CPU) a proof the oracle verifier accepts under the same setting and different bytes from the default,
and (GPU) the oracle's bytes."""
import numpy as np
import pytest

import b200zk
import oracle_lib as O
from synth_small import make_circuit

SETTINGS = [(0, 0), (1, 0), (2, 0), (4, 0), (0, 3), (0, 4), (7, 5)]


@pytest.fixture(autouse=True)
def _reset():
    yield
    O.set_compat(0, 0)


def test_oracle_proofs_verify_under_every_setting():
    k, A, L, F = 7, 2, 1, 1
    fixed, advice, copies = make_circuit(k, A, L, F, seed=11)
    params = O.Params.setup(k)
    proofs = {}
    for flags, chunks in SETTINGS:
        O.set_compat(flags, chunks)
        pk = O.ProvingKey(params, k, A, L, F, fixed, copies)
        proof = pk.create_proof(advice, 3)
        ok, err = pk.verify(proof, pairing=True)
        assert ok, (flags, chunks, err)
        proofs[(flags, chunks)] = proof
    base = proofs[(0, 0)]
    for key, p in proofs.items():
        if key != (0, 0):
            assert p != base and len(p) == len(base), key  # every switch changes the bytes, never the size
    # the point-encoding switch alone changes only flag bits of compressed points
    a, b = np.frombuffer(base, dtype=np.uint8), np.frombuffer(proofs[(4, 0)], dtype=np.uint8)
    diff = np.nonzero(a != b)[0]
    assert len(diff) > 0 and all(i % 32 == 31 for i in diff) and all((int(a[i]) ^ int(b[i])) & 0x3F == 0 for i in diff)
    # a verifier with another setting of the point encoding must not accept
    O.set_compat(0, 0)
    pk = O.ProvingKey(params, k, A, L, F, fixed, copies)
    assert not pk.verify(proofs[(4, 0)])[0] or proofs[(4, 0)] == base


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(8, 2, 1, 1), (10, 4, 2, 1)])
def test_gpu_matches_oracle_under_every_setting(ctx, shape):
    k, A, L, F = shape
    fixed, advice, copies = b200zk.synth_circuit(k, A, L, F, seed=k + 1)
    ctx.srs_setup(k)
    params = O.Params.setup(k)
    try:
        for flags, chunks in SETTINGS:
            O.set_compat(flags, chunks)
            ctx.set_compat(flags, chunks)
            opk = O.ProvingKey(params, k, A, L, F, fixed, copies)
            gpk = ctx.keygen(k, A, L, F, fixed, copies)
            want = opk.create_proof(advice, 5)
            got = gpk.create_proof(advice, 5)
            assert got == want, (flags, chunks)
            assert opk.verify(got, pairing=True)[0]
            gpk.close()
    finally:
        ctx.set_compat(0, 0)


@pytest.mark.gpu
def test_lookup_fill_rules_on_the_primitive(ctx):
    """b200zk_permute_expression_pair under both fill rules == the oracle's permute_expression_pair under the same rule."""
    k = 9
    n = 1 << k
    rng = np.random.default_rng(2)
    table = O.fr_array([i if i < n // 2 else 0 for i in range(n)])
    inp = O.fr_array([int(v) for v in np.where(rng.integers(0, 3, size=n) == 0, rng.integers(0, n // 2, size=n), 7)])
    outs = []
    try:
        for flags in (0, 2):
            O.set_compat(flags, 0)
            ctx.set_compat(flags, 0)
            a_want = np.zeros((n, 4), dtype=np.uint64)
            s_want = np.zeros((n, 4), dtype=np.uint64)
            assert O.lib().oracle_permute_expression_pair(k, O.ptr(inp), O.ptr(table), O.ptr(a_want), O.ptr(s_want)) == 1
            a, s = ctx.permute_expression_pair(k, inp, table)
            assert np.array_equal(a, a_want[: n - 7]) and np.array_equal(s, s_want[: n - 7])
            outs.append(s)
    finally:
        ctx.set_compat(0, 0)
    assert not np.array_equal(outs[0], outs[1])  # the two rules really differ on this input
