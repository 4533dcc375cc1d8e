"""Quick device-side timing of the NTT entry points (CUDA events on the library's stream)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import b200zk
import oracle_lib as O

ctx = b200zk.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
rng = np.random.default_rng(0)
res = {}
for log_n in (16, 18, 20, 22, 24):
    n = 1 << log_n
    batch = 4 if log_n <= 22 else 1
    host = O.random_fr(rng, n)
    buf = torch.empty(batch * n * 4, dtype=torch.int64, device="cuda")
    for b in range(batch):
        ctx.h2d(buf.data_ptr() + 32 * n * b, host)
    omega = O.domain_constant(log_n, 0)
    for _ in range(3):
        ctx.ntt_dev(buf.data_ptr(), log_n, omega, batch, n)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(iters):
            ctx.ntt_dev(buf.data_ptr(), log_n, omega, batch, n)
        e1.record(stream)
    ctx.sync()
    ms = e0.elapsed_time(e1) / iters
    gbs = 64.0 * n * batch / (ms * 1e-3) / 1e9
    butterflies = batch * (n // 2) * log_n
    res[log_n] = dict(ms=ms, batch=batch, GBps=gbs, Gbutterfly_s=butterflies / (ms * 1e-3) / 1e9)
    print(log_n, res[log_n], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/time_ntt.json", "w"), indent=1)
