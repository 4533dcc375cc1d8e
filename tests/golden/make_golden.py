"""Generates tests/golden/oracle_vectors.json from the CPU oracle. The reference repo has no vectors for this path
 and cannot be built here, so these fixtures pin the ORACLE's byte output (regressions in either the
oracle or the GPU path show up against them); run: python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np

import oracle_lib as O
from synth_small import make_circuit

out = {"proofs": []}
for shape, cseed, rseed in (((5, 1, 1, 1), 5, 0), ((6, 2, 1, 1), 9, 3), ((7, 2, 0, 1), 2, 1)):
    k, A, L, F = shape
    fixed, advice, copies = make_circuit(k, A, L, F, seed=cseed)
    params = O.Params.setup(k)
    s, g, gl = params.get()
    pk = O.ProvingKey(params, k, A, L, F, fixed, copies)
    proof = pk.create_proof(advice, rseed)
    assert pk.verify(proof)[0]
    out["proofs"].append({"shape": list(shape), "circuit_seed": cseed, "rng_seed": rseed, "srs_sha256": hashlib.sha256(g.tobytes() + gl.tobytes()).hexdigest(),
                          "proof_hex": proof.hex()})
v = np.empty(4, dtype=np.uint64)
O.lib().oracle_std_rng_random_fr(0, 1, O.ptr(v))
out["std_rng_seed0_first_fr_mont"] = [[int(x) for x in v]]
out["std_rng_seed0_first_fr"] = str(O.from_mont(v))
json.dump(out, open(os.path.join(HERE, "oracle_vectors.json"), "w"), indent=1)
print("wrote", len(out["proofs"]), "proofs")
