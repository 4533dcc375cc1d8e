"""Device-side timing of b200zk_msm_dev on the KZG bases (CUDA events on the library's stream)."""
import sys, os, json, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch, b200zk, oracle_lib as O

ctx = b200zk.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
rng = np.random.default_rng(0)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
t = time.time(); params = O.Params.setup(k); s, g, gl = params.get(); print("oracle setup s", time.time() - t, flush=True)
ctx.srs_load(k, g, gl)
res = {}
for log_n in range(16, k + 1, 2):
    n = 1 << log_n
    for name in ("uniform", "small21"):
        host = O.random_fr(rng, n) if name == "uniform" else O.fr_array([int(v) for v in rng.integers(0, 1 << 21, size=n)])
        buf = torch.empty(n * 4, dtype=torch.int64, device="cuda")
        ctx.h2d(buf.data_ptr(), host)
        for _ in range(2):
            ctx.msm_dev(buf.data_ptr(), n, 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        t0 = time.time()
        e0.record(stream)
        for _ in range(iters):
            ctx.msm_dev(buf.data_ptr(), n, 0)
        e1.record(stream)
        ctx.sync()
        wall = (time.time() - t0) / iters * 1e3
        ms = e0.elapsed_time(e1) / iters
        res[f"{log_n}/{name}"] = dict(ms=ms, wall_ms=wall, Mpts_s=n / ms / 1e3)
        print(log_n, name, res[f"{log_n}/{name}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/time_msm.json", "w"), indent=1)
