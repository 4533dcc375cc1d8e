"""The oracle prover end to end on CPU: proofs of satisfying witnesses verify (vanishing identity at x + SHPLONK opening
checked in G1 with the trapdoor AND with the real BN254 pairing), tampering and unsatisfied witnesses are rejected, outputs are deterministic, and the
committed golden vectors (tests/golden/, made by tests/golden/make_golden.py) still reproduce."""
import hashlib
import json
import os

import numpy as np
import pytest

import b200zk
import oracle_lib as O
from synth_small import make_circuit

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors.json")


@pytest.mark.parametrize("shape", [(5, 1, 1, 1), (8, 3, 2, 1), (9, 2, 0, 1), (10, 4, 1, 2)])
def test_oracle_proof_verifies_and_rejects_tampering(shape):
    k, A, L, F = shape
    fixed, advice, copies = make_circuit(k, A, L, F, seed=k)
    ok, err = O.mock_check(k, A, L, F, fixed, advice, copies)
    assert ok, err
    params = O.Params.setup(k)
    pk = O.ProvingKey(params, k, A, L, F, fixed, copies)
    proof = pk.create_proof(advice, 0)
    assert len(proof) == O.lib().oracle_proof_size(k, A, L, F)
    assert pk.verify(proof) == (True, "")
    assert pk.verify(proof, pairing=True) == (True, "")  # e(L, g2)·e(−H', s·g2) == 1, as halo2's verifier checks it
    assert proof == pk.create_proof(advice, 0)           # deterministic
    assert proof != pk.create_proof(advice, 1)           # rng seed matters (blinding)
    npoints = (A + L) + 2 * L + (A + L + F + 1) // 2 + L + 1 + 3
    for pos in (1, 32 * npoints + 5, len(proof) - 70, len(proof) - 3):
        bad = bytearray(proof); bad[pos] ^= 1
        assert not pk.verify(bytes(bad))[0]
        assert not pk.verify(bytes(bad), pairing=True)[0]
    assert not pk.verify(proof[:-32])[0]
    bad_w = advice.copy(); bad_w[0, 3] = bad_w[0, 2]
    assert not O.mock_check(k, A, L, F, fixed, bad_w, copies)[0]
    assert not pk.verify(pk.create_proof(bad_w, 0))[0]


def test_synth_circuits_satisfy_the_constraint_system():
    """csrc/synth.cu (host-only part of libb200zk) generates satisfying circuits of the halo2-base shape."""
    for (k, A, L, F) in [(6, 1, 1, 1), (10, 4, 1, 2), (12, 14, 3, 1), (11, 3, 0, 1)]:
        fixed, advice, copies = b200zk.synth_circuit(k, A, L, F, seed=k)
        ok, err = O.mock_check(k, A, L, F, fixed, advice, copies)
        assert ok, err
        f2, a2, c2 = b200zk.synth_circuit(k, A, L, F, seed=k)
        assert np.array_equal(advice, a2) and np.array_equal(copies, c2)  # deterministic in the seed
        # distribution: most cells are range-check cells (< 2^84), as in the reference's bn254_rev.svg profile
        sample = advice[0, : min(4000, 1 << k)]
        small = sum(1 for v in O.fr_ints(sample) if v < (1 << 84)) / len(sample)
        assert small > 0.5


def test_golden_vectors_reproduce():
    g = json.load(open(GOLDEN))
    for case in g["proofs"]:
        k, A, L, F = case["shape"]
        fixed, advice, copies = make_circuit(k, A, L, F, seed=case["circuit_seed"])
        params = O.Params.setup(k)
        s, gg, gl = params.get()
        assert hashlib.sha256(gg.tobytes() + gl.tobytes()).hexdigest() == case["srs_sha256"]
        pk = O.ProvingKey(params, k, A, L, F, fixed, copies)
        proof = pk.create_proof(advice, case["rng_seed"])
        assert proof.hex() == case["proof_hex"]
    assert O.fr_ints(np.array(g["std_rng_seed0_first_fr_mont"], dtype=np.uint64))[0] == int(g["std_rng_seed0_first_fr"])
    out = np.empty(4, dtype=np.uint64)
    O.lib().oracle_std_rng_random_fr(0, 1, O.ptr(out))
    assert O.from_mont(out) == int(g["std_rng_seed0_first_fr"])
