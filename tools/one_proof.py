"""keygen + N create_proof calls for a shape (used under ncu for launch lists)."""
import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, b200zk
k, A, L, F = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (20, 14, 3, 1))]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
ctx = b200zk.Context(0)
ctx.srs_setup(k)
fixed, advice, copies = b200zk.synth_circuit(k, A, L, F, seed=0)
pk = ctx.keygen(k, A, L, F, fixed, copies)
for _ in range(reps):
    t = time.time(); proof, tm = pk.create_proof(advice, 0, timings=True); print("proof s", round(time.time() - t, 4), {a: round(b * 1e3, 1) for a, b in tm.items()}, flush=True)
ctx.close()
