"""One full create_proof on the GPU for a named shape, with the per-stage split and (optionally) the oracle check."""
import sys, os, time, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, b200zk, oracle_lib as O

k, A, L, F = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (20, 14, 3, 1))]
check = "--check" in sys.argv
t = time.time(); fixed, advice, copies = b200zk.synth_circuit(k, A, L, F, seed=0); print("synth s", round(time.time() - t, 2), "copies", len(copies), flush=True)
t = time.time(); params = O.Params.setup(k); s, g, gl = params.get(); print("oracle srs s", round(time.time() - t, 2), flush=True)
ctx = b200zk.Context(0)
ctx.srs_load(k, g, gl)
t = time.time(); pk = ctx.keygen(k, A, L, F, fixed, copies); print("gpu keygen s", round(time.time() - t, 2), flush=True)
for it in range(3):
    t = time.time(); proof, tm = pk.create_proof(advice, 0, timings=True); dt = time.time() - t
    print("gpu create_proof s", round(dt, 4), {k_: round(v * 1e3, 1) for k_, v in tm.items()}, flush=True)
t = time.time(); proof2 = pk.create_proof(advice, 0); print("gpu create_proof (no lap syncs) s", round(time.time() - t, 4), flush=True)
assert proof2 == proof
if check:
    t = time.time(); opk = O.ProvingKey(params, k, A, L, F, fixed, copies); print("oracle keygen s", round(time.time() - t, 2), "threads", O.lib().oracle_get_threads(), flush=True)
    want = opk.create_proof(advice, 0); print("oracle create_proof s", round(opk.last_seconds, 2), flush=True)
    print("bytes equal:", want == proof, "verify:", opk.verify(proof))
pk.close(); ctx.close()
