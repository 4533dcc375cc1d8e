"""Per-rank MSM cost under point-range sharding, emulated on one GPU (the other ranks' partials are zeros)."""
import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch, b200zk
k = 20
n = 1 << k
ctx = b200zk.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
ctx.srs_setup(k)
rng = np.random.default_rng(0)
a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 60) - 1)
buf = torch.empty(n * 4, dtype=torch.int64, device="cuda"); ctx.h2d(buf.data_ptr(), a)
for world in (1, 2, 4, 8):
    if world > 1:
        ctx.set_allgather(0, world, lambda data, w=world: data + bytes(len(data) * (w - 1)))
    else:
        ctx.set_allgather(0, 1, None)
    for _ in range(3): ctx.msm_dev(buf.data_ptr(), n, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record(stream)
    for _ in range(5): ctx.msm_dev(buf.data_ptr(), n, 0)
    e1.record(stream); ctx.sync()
    print("world", world, "per-rank msm ms", round(e0.elapsed_time(e1) / 5, 3), "wall", round((time.time() - t0) / 5 * 1e3, 3), flush=True)
ctx.close()
