"""One full create_proof on the GPU for a named shape, with the per-stage split and (with --check) the byte comparison
against the CPU oracle's proof of the same circuit, SRS and rng seed.

  python tools/prove.py [k A L F] [--check] [--gpus N]

--gpus N > 1 proves through b200zk_create_multi (one process, N devices). The SRS is generated on the device
(ParamsKZG::setup(k, ChaCha20Rng::from_seed([0;32]))) and handed to the oracle, so the CPU side only pays keygen + proof."""
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np  # noqa: E402

import b200zk  # noqa: E402
import oracle_lib as O  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
k, A, L, F = [int(v) for v in (args[:4] if len(args) >= 4 else (20, 14, 3, 1))]
check = "--check" in sys.argv
gpus = int(sys.argv[sys.argv.index("--gpus") + 1]) if "--gpus" in sys.argv else 1
print(f"shape k={k} A={A} L={L} F={F}: {A + L + 2 * L + (A + L + F + 1) // 2 + L + 6} MSMs of 2^{k}; gpus={gpus}", flush=True)
t = time.time(); fixed, advice, copies = b200zk.synth_circuit(k, A, L, F, seed=0); print("synth s", round(time.time() - t, 2), "copies", len(copies), flush=True)
if gpus > 1:
    import torch  # noqa: F401  (brings libnccl.so.2 into the process)

    ctx = b200zk.Context.multi(list(range(gpus)))
else:
    ctx = b200zk.Context(0)
t = time.time(); trapdoor = ctx.srs_setup(k); print("device srs s", round(time.time() - t, 2), flush=True)
t = time.time(); pk = ctx.keygen(k, A, L, F, fixed, copies); print("gpu keygen s", round(time.time() - t, 2), flush=True)
for it in range(3):
    t = time.time(); proof, tm = pk.create_proof(advice, 0, timings=True); dt = time.time() - t
    print("gpu create_proof s", round(dt, 4), {k_: round(v * 1e3, 1) for k_, v in tm.items()}, flush=True)
t = time.time(); proof2 = pk.create_proof(advice, 0); print("gpu create_proof (no lap syncs, pageable host witness) s", round(time.time() - t, 4), flush=True)
assert proof2 == proof
print("proof bytes", len(proof), flush=True)
if check:
    g, gl = ctx.srs_download()
    params = O.Params.load(k, trapdoor, g, gl)
    del g, gl
    t = time.time(); opk = O.ProvingKey(params, k, A, L, F, fixed, copies); print("oracle keygen s", round(time.time() - t, 2), "threads", O.lib().oracle_get_threads(), flush=True)
    fc, pc = pk.commitments()
    print("vk commitments equal:", bool(np.array_equal(fc, opk.get(0)) and np.array_equal(pc, opk.get(1))), flush=True)
    want = opk.create_proof(advice, 0); print("oracle create_proof s", round(opk.last_seconds, 2), flush=True)
    print("bytes equal:", want == proof, "verify (pairing):", opk.verify(proof, pairing=True), flush=True)
pk.close()
ctx.close()
