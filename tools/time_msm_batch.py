"""Per-column cost of batched commits (b200zk_msm_batch_dev) on the KZG bases at 2^k: batches of 1, 4, 16 columns of
uniform / 21-bit scalars, CUDA events on the library's stream.   python tools/time_msm_batch.py [k] [batches...]"""
import sys, os, json
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch, b200zk

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
batches = [int(v) for v in sys.argv[2:]] or [1, 4, 16]
ctx = b200zk.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
ctx.srs_setup(k)
n = 1 << k
rng = np.random.default_rng(0)
res = {}
for name in ("uniform", "small21"):
    cols = []
    for j in range(max(batches)):
        a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 60) - 1)
        if name == "small21":
            a[:, 0] &= np.uint64((1 << 21) - 1)
            a[:, 1:] = 0
            a = ctx.field_vec_op(0, 6, a)
        t = torch.empty(n * 4, dtype=torch.int64, device="cuda")
        ctx.h2d(t.data_ptr(), a)
        cols.append(t)
    for B in batches:
        ptrs = [c.data_ptr() for c in cols[:B]]
        for _ in range(2):
            ctx.msm_batch_dev(ptrs, n, 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        e0.record(stream)
        for _ in range(iters):
            ctx.msm_batch_dev(ptrs, n, 0)
        e1.record(stream)
        ctx.sync()
        ms = e0.elapsed_time(e1) / iters
        res[f"{name}/B{B}"] = dict(ms=ms, ms_per_col=ms / B, Mpts_s=B * n / ms / 1e3)
        print(name, "batch", B, res[f"{name}/B{B}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/time_msm_batch.json", "w"), indent=1)
ctx.close()
