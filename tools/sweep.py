"""BASELINE.json configs [2] and [3]: G1 MSM sweep (uniform / advice-like / sorted-lookup scalars) and Fr NTT sweep
(forward, lagrange_to_coeff, coeff_to_extended, extended_to_coeff; batches 17 and 29), device-resident inputs, CUDA events on the
library's stream. Under torchrun the MSM sweep is sharded by point range across the ranks (partial sums over NCCL inside
the library) and the NTT sweep deals the columns of a batch round-robin to the ranks (no exchange: aggregate throughput,
time = the slowest rank).

    python tools/sweep.py [--max-k 26] [--out gpurun_out/sweep.json]
"""
import argparse
import json
import os
import sys
import time

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
import torch

import b200zk

p = argparse.ArgumentParser()
p.add_argument("--max-k", type=int, default=26)
p.add_argument("--min-k", type=int, default=16)
p.add_argument("--out", default="gpurun_out/sweep.json")
p.add_argument("--no-ntt", action="store_true")
args = p.parse_args()

rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local_rank)
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
ctx = b200zk.Context(local_rank)
stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
if world > 1:
    ctx.set_allgather(rank, world, b200zk.torch_allgather(dist, torch.device("cuda", local_rank)))
    ctx.comm_init()
HBM = 6541.8
try:
    HBM = float(json.load(open(os.path.join(R, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, iters):
    for _ in range(2):
        fn()
    ctx.sync()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    ctx.sync()
    ms = e0.elapsed_time(e1) / iters
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return ms


def scalars(kind, n, k, rng):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    if kind == "uniform":
        return a  # any canonical limbs are SOME uniformly distributed field elements in Montgomery form
    canon = np.zeros((n, 4), dtype=np.uint64)
    if kind == "advice_like":  # 40 % 21-bit limbs, 35 % <= 84-bit accumulators, 5 % exact 0/1, 20 % full width
        sel = rng.integers(0, 100, size=n)
        canon[:, 0] = a[:, 0]
        canon[:, 1] = a[:, 1]
        canon[:, 2] = a[:, 2]
        canon[:, 3] = a[:, 3] >> np.uint64(8)
        m = sel < 40
        canon[m, 0] &= np.uint64((1 << 21) - 1)
        canon[m, 1:] = 0
        m = (sel >= 40) & (sel < 75)
        canon[m, 1] &= np.uint64((1 << 20) - 1)
        canon[m, 2:] = 0
        m = (sel >= 75) & (sel < 80)
        canon[m, 0] &= np.uint64(1)
        canon[m, 1:] = 0
    else:  # sorted_lookup: sorted values < 2^(k-1)
        canon[:, 0] = np.sort(a[:, 0] & np.uint64((1 << max(k - 1, 1)) - 1))
    return ctx.field_vec_op(0, 6, canon)  # to Montgomery form on the device


res = {"world": world, "msm": {}, "ntt": {}}
rng = np.random.default_rng(0)
kmax = args.max_k
for k in range(args.min_k, kmax + 1, 2):
    n = 1 << k
    t0 = time.time()
    ctx.srs_setup(k)  # bases s^i·G of exactly this size (+ their window tables), generated on the device
    if rank == 0:
        print(f"srs k={k} (incl. window tables) {time.time() - t0:.1f} s", flush=True)
    for kind in ("uniform", "advice_like", "sorted_lookup"):
        host = scalars(kind, n, k, rng)
        buf = torch.empty(n * 4, dtype=torch.int64, device="cuda")
        ctx.h2d(buf.data_ptr(), host)
        ms = timed(lambda: ctx.msm_dev(buf.data_ptr(), n, 0), 5 if k <= 22 else 3)
        res["msm"][f"2^{k}/{kind}"] = {"ms": ms, "Mpts_s": n / ms / 1e3}
        if rank == 0:
            print("msm", k, kind, round(ms, 3), "ms", round(n / ms / 1e3, 1), "Mpts/s", flush=True)
        del buf
if not args.no_ntt:
    root = np.array([[0xd34f1ed960c37c9c, 0x3215cf6dd39329c8, 0x98865ea93dd31f74, 0x03ddb9f5166d18b7]], dtype=np.uint64)
    cases = [(k, B) for B in (17, 29) for k in range(args.min_k, min(kmax, 22) + 1, 2)]  # S20-bn's and S22-gl's A+L
    if kmax >= 24:
        cases.append((24, 4))
    for k, batch in cases:
        n = 1 << k
        col = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
        total_batch = batch
        batch = len(range(rank, total_batch, world))  # this rank's columns of the batch (round-robin)
        if batch == 0:
            batch = 1  # keeps the collective timing calls in step; not counted (total_batch < world never happens here)
        buf = torch.empty(batch * n * 4, dtype=torch.int64, device="cuda")
        for b in range(batch):
            ctx.h2d(buf.data_ptr() + 32 * n * b, col)
        w = ctx.field_vec_op(0, 6, root)
        for _ in range(28 - k):
            w = ctx.field_vec_op(0, 2, w, w)
        omega = w[0].copy()
        ms_f = timed(lambda: ctx.ntt_dev(buf.data_ptr(), k, omega, batch, n), 5)
        row = {"batch": total_batch, "columns_on_slowest_rank": len(range(0, total_batch, world)), "forward_ms": ms_f,
               "forward_GBps_64nB": 64.0 * n * total_batch / ms_f / 1e6, "forward_frac_hbm": 64.0 * n * total_batch / ms_f / 1e6 / (HBM * world),
               "Gbutterfly_s": total_batch * (n // 2) * k / ms_f / 1e6}
        ms_l = timed(lambda: ctx.lagrange_to_coeff_dev(k, buf.data_ptr(), batch, n), 5)
        row["lagrange_to_coeff_ms"] = ms_l
        ext = torch.empty(4 * n * 4, dtype=torch.int64, device="cuda")
        ms_c = timed(lambda: ctx.coeff_to_extended_dev(k, buf.data_ptr(), ext.data_ptr()), 5)
        row["coeff_to_extended_ms"] = ms_c
        row["coeff_to_extended_GBps_160n"] = 160.0 * n / ms_c / 1e6
        out3 = torch.empty(3 * n * 4, dtype=torch.int64, device="cuda")
        ms_e = timed(lambda: ctx.extended_to_coeff_dev(k, ext.data_ptr(), out3.data_ptr()), 5)
        row["extended_to_coeff_ms"] = ms_e
        row["extended_to_coeff_GBps_224n"] = 224.0 * n / ms_e / 1e6
        del ext, out3
        res["ntt"][f"2^{k}/B{total_batch}"] = row
        if rank == 0:
            print("ntt", k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in row.items()}, flush=True)
        del buf
if rank == 0:
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
ctx.close()
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
