#!/usr/bin/env python
"""bench.py — the headline benchmark of BASELINE.json: `create_proof` seconds for the FRI-verifier-shaped circuit at
k=20 (shape S20-bn: 14 gate + 3 lookup + 1 constant columns, 41 MSMs of 2^20, 35 iNTTs, 35 coset NTTs),
KZG-BN254 / SHPLONK / Blake2b transcript, plus the MSM Mpts/s and NTT GB/s lines the metric names.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--k 20] [--A 14 --L 3 --F 1]

One JSON line on stdout (rank 0). A "step" = one create_proof over one synthetic witness (workload/synth.cpp).
  value  : seconds per proof with the witness already resident in HBM (b200zk_create_proof_dev), CUDA events on the
           library's stream, max over ranks.
  e2e    : the same through the reference-facing call with HOST buffers (b200zk_create_proof) from PAGEABLE host memory —
           what halo2's create_proof holds (a Rust Vec<Fr>) — host-to-device copies inside the timed region, proof bytes
           back on the host. `e2e.pinned` is the same call from page-locked memory.
  roofline : the dominant kernel (msm_accumulate_kernel) against the roof that binds it, the INT32 multiply-add pipe
           (north_star: "ncu integer-pipe utilisation against the B200's INT32 peak"); the HBM view the contract names
           (96 bytes per point) is carried under roofline.hbm.
  --impl reference : the CPU restatement of halo2's prover (oracle/, all host threads) on the SAME configuration: the same
           circuit at the same k, at most two measured steps (≈ 50 s each at k=20 on 16 threads) and no warm-up, so that the
           run ends within minutes; `steps` in its line is the number of steps actually timed. The real Rust prover cannot
           be built here (no Rust toolchain, un-vendored dependencies).
N > 1 (torchrun): ONE proof is computed by all ranks together (strong scaling): commit batches dealt by column with the
remainder split by point range, NTTs by column, h(X) by row slice; the partial sums travel over NCCL inside the library.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = dict(A=14, L=3, F=1)  # S20-bn / its k-scaled versions
emit = None
# multiply-add instructions on the FMA-heavy pipe per Montgomery product / square / fused dual product a·b ± c·d
# (cuobjdump of build/*.o)
IMAD_PER_MUL, IMAD_PER_SQR, IMAD_PER_DUAL = 138, 108, 200
MADD_MUL, MADD_SQR, MADD_DUAL = 6, 2, 1  # XYZZ mixed addition (madd-2008-s: 8M + 2S) with Y3 as one dual product


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200")
    p.add_argument("--k", type=int, default=20)
    p.add_argument("--A", type=int, default=SHAPE["A"], help="gate advice columns (default: S20-bn)")
    p.add_argument("--L", type=int, default=SHAPE["L"], help="lookup advice columns")
    p.add_argument("--F", type=int, default=SHAPE["F"], help="constant columns")
    p.add_argument("--sample-k", type=int, default=17, help="k of the bounded CPU sample in the GPU arm's cpu_baseline")
    p.add_argument("--ref-max-steps", type=int, default=2, help="reference arm: measured steps are capped at this")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extras", action="store_true", help="skip the MSM / NTT side lines")
    return p.parse_args()


def workload_name(k):
    A, L, F = SHAPE["A"], SHAPE["L"], SHAPE["F"]
    name = {(14, 3, 1): "bn", (4, 1, 1): "bn", (23, 6, 1): "gl"}.get((A, L, F), "custom")
    return (f"S{k}-{name}: halo2-base-shaped FRI-verifier stand-in, k={k}, {A} gate + {L} lookup + {F} constant columns, "
            f"create_proof KZG-BN254 SHPLONK Blake2b, {A + L + 2 * L + (A + L + F + 1) // 2 + L + 6} MSMs of 2^{k}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.thread.join(timeout=5)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def cpu_prover(k, threads=None):
    """Oracle (restated halo2 CPU prover, std::thread on all host cores) for the bench shape at k. The circuit comes from the
    workload library, so this path maps neither libb200zk nor torch. Returns (step() -> (seconds, proof), cores, pk)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import workload

    if threads:
        O.lib().oracle_set_threads(threads)
    fixed, advice, copies = workload.synth_circuit(k, SHAPE["A"], SHAPE["L"], SHAPE["F"], seed=0)
    params = O.Params.setup(k)
    pk = O.ProvingKey(params, k, SHAPE["A"], SHAPE["L"], SHAPE["F"], fixed, copies)

    def step():
        proof = pk.create_proof(advice, 0)
        return pk.last_seconds, proof

    return step, O.lib().oracle_get_threads(), pk


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the box's host cores, on the
    configuration the GPU arm runs (same circuit, same k). Rank 0 alone works."""
    if rank != 0:
        return
    t_setup = time.time()
    step, cores, pk = cpu_prover(args.k)
    t_setup = time.time() - t_setup
    steps = max(1, min(args.steps, args.ref_max_steps))
    t = []
    for _ in range(steps):
        s, proof = step()
        t.append(s)
    ok, err = pk.verify(proof)
    if not ok:
        raise SystemExit(f"bench.py: the oracle rejected its own proof: {err}")
    value = sum(t) / len(t)
    line = {
        "impl": "reference", "metric": "create_proof_s", "value": value, "unit": "s", "n_gpus": args.gpus, "steps": steps, "warmup": 0,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "u64 limbs (254-bit Montgomery)",
        "data": "synthetic", "config": {"workload": workload_name(args.k)},
        "cpu_baseline": {"value": value, "unit": "s", "cores": cores, "kind": "port",
                         "sample": f"the whole configuration: create_proof at k={args.k}, {steps} measured step(s) of {[round(x, 2) for x in t]} s, no warm-up "
                                   f"(SRS + keygen_pk on the CPU took {t_setup:.0f} s, untimed); restated halo2 CPU algorithms (C++ oracle, std::thread), "
                                   "not the rayon binary — Rust is not available here"},
        "e2e": {"value": value, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture (profiles/ncu_traffic.json), or None."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]
        return float(rec["dram_bytes_per_launch"]), rec.get("source", "profiles/ncu_traffic.json")
    except Exception:
        return None, "no committed ncu capture for this kernel"


def main():
    args = parse()
    SHAPE.update(A=args.A, L=args.L, F=args.F)
    # stdout carries exactly ONE JSON line: anything else that writes to fd 1 (NCCL's version banner, library chatter)
    # is sent to stderr; the JSON goes through the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global emit
    emit = lambda line: os.write(json_fd, (json.dumps(line) + "\n").encode())  # noqa: E731
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return 0
    import numpy as np
    import torch

    import b200zk
    import workload

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libb200zk has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    k, A, L, F = args.k, SHAPE["A"], SHAPE["L"], SHAPE["F"]
    n = 1 << k
    ctx = b200zk.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))
    if world > 1:  # the library builds its own NCCL communicator; the callback only carries the 128-byte bootstrap id
        ctx.set_allgather(rank, world, b200zk.torch_allgather(dist, torch.device("cuda", local_rank)))
    # ---- untimed setup: SRS on the device, synthetic circuit, keygen ----
    t0 = time.time()
    ctx.srs_setup(k)  # ParamsKZG::setup(k, ChaCha20Rng::from_seed([0;32])) like halo2-base gen_srs
    t_srs = time.time() - t0
    fixed, advice, copies = workload.synth_circuit(k, A, L, F, seed=0)
    t0 = time.time()
    pk = ctx.keygen(k, A, L, F, fixed, copies)
    t_keygen = time.time() - t0
    del fixed, copies
    pageable_advice = np.ascontiguousarray(advice).reshape(-1)  # what a Rust Vec<Fr> is: ordinary pageable memory
    pinned = torch.from_numpy(pageable_advice.view(np.int64)).pin_memory()
    pinned_advice = pinned.numpy().view(np.uint64)
    dev_advice = torch.empty(pinned.numel(), dtype=torch.int64, device="cuda")
    dev_advice.copy_(pinned)
    torch.cuda.synchronize()
    advice_bytes = pinned.numel() * 8

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.time()
        e0.record(stream)
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        ctx.sync()
        barrier()
        wall = time.time() - w0
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms, wall * 1e3], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = t[0].item(), t[1].item() / 1e3
        return ms / steps, wall / steps, out

    step_dev = lambda: pk.create_proof(None, 0, device_ptr=dev_advice.data_ptr())  # noqa: E731
    step_pageable = lambda: pk.create_proof(pageable_advice, 0)  # noqa: E731
    step_pinned = lambda: pk.create_proof(pinned_advice, 0)  # noqa: E731
    for _ in range(max(args.warmup, 0)):
        step_dev()
    l_before = b200zk.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, wall_dev, proof = timed(step_dev, args.steps)
    launches = (b200zk.launch_count() - l_before) // max(args.steps, 1)
    for _ in range(min(args.warmup, 2)):
        step_pageable()
    ms_host, wall_host, proof_h = timed(step_pageable, args.steps)
    step_pinned()
    ms_pin, wall_pin, proof_p = timed(step_pinned, max(1, min(args.steps, 3)))
    clocks = sampler.stop()
    ok = proof == proof_h == proof_p and len(proof) == pk.proof_size()
    # ---- one profiled step: per-stage split + dominant-kernel durations by CUDA events on the launching stream ----
    # (two untimed steps: the stage split with the normal stream overlap, then the event profiler, under which the MSM
    # columns run one at a time so that a kernel's bracketed duration is its own and comparable with the ncu launch list)
    _, stages = pk.create_proof(None, 0, timings=True, device_ptr=dev_advice.data_ptr())
    ctx.profile_enable(True)
    pk.create_proof(None, 0, device_ptr=dev_advice.data_ptr())
    acc_adds = ctx.profile_work("msm_accumulate")
    ntt_bfly = ctx.profile_work("ntt_pass")
    acc_ms, acc_n = ctx.profile_get("msm_accumulate")
    ntt_ms, ntt_n = ctx.profile_get("ntt_pass")
    q_ms, q_n = ctx.profile_get("quotient")
    ctx.profile_enable(False)
    if world > 1:
        # the sharded proof must be the single-GPU proof: every rank recomputes it alone (untimed, after the last sharded
        # call) and compares the bytes; the verdict is agreed on by all ranks before anyone leaves
        ctx.set_allgather(0, 1, None)
        alone = pk.create_proof(None, 0, device_ptr=dev_advice.data_ptr())
        ok = ok and alone == proof
        flag = torch.tensor([0 if ok else 1], device="cuda", dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        ok = flag.item() == 0
    if not ok:
        pk.close()
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        raise SystemExit(f"bench.py: rank {rank}: proofs differ (device-resident / host witness" + (f" / sharded over {world} GPUs vs one GPU)" if world > 1 else ")"))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    imad_peak, imad_src = 18.0, "fallback 18.0 T/s (round-1 measurement)"
    try:
        mb = json.load(open(os.path.join(ROOT, "profiles", "microbench_r02.json")))
        imad_peak, imad_src = float(mb["imad_Tops"]), "measured: dependent-free IMAD issue rate, tools/microbench.cu (profiles/microbench_r02.json)"
    except Exception:
        pass
    # dominant kernel = msm_accumulate_kernel. Integer roof: one XYZZ mixed addition = 6 Montgomery products + 2 squares + 1
    # dual product = 6·138 + 2·108 + 200 = 1244 multiply-add instructions on the FMA-heavy pipe; additions per launch are
    # counted by the library.
    acc_avg_ms = acc_ms / max(acc_n, 1)
    imad_per_add = MADD_MUL * IMAD_PER_MUL + MADD_SQR * IMAD_PER_SQR + MADD_DUAL * IMAD_PER_DUAL
    adds_per_launch = acc_adds / max(acc_n, 1)
    achieved_tops = adds_per_launch * imad_per_add / (acc_avg_ms * 1e-3) / 1e12 if acc_avg_ms > 0 else 0.0
    alg_bytes = 96.0 * n  # 32 B scalar + 64 B base per point
    achieved_gbs = alg_bytes / (acc_avg_ms * 1e-3) / 1e9 if acc_avg_ms > 0 else 0.0
    traffic, traffic_src = ncu_traffic("msm_accumulate_kernel")
    roofline = {"kernel": "msm_accumulate_kernel", "bound": "int32", "achieved": achieved_tops, "peak": imad_peak, "unit": "T multiply-add/s",
                "frac": achieved_tops / imad_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": imad_src,
                "work": {"mixed_additions_per_launch": adds_per_launch, "multiply_adds_per_addition": imad_per_add,
                         "basis": "XYZZ madd-2008-s = 6 products (138 IMAD-class each) + 2 squares (108 each) + 1 dual product with one reduction "
                                  "(200); additions counted by the library"},
                "hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                        "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src},
                "launches": acc_n, "avg_launch_ms": acc_avg_ms, "share_of_step": acc_ms / (ms_dev if ms_dev else 1),
                "timing": "CUDA events on the launching stream around every launch of one untimed step in which the MSM columns run one at a time "
                          "(in the timed steps up to four columns overlap on separate streams)",
                "note": "integer-pipe bound by design (north_star: no tensor cores, IMAD carry chains); carry-flag forms of IMAD.WIDE issue at half "
                        "rate on sm_100 (profiles/microbench_r02.json), which caps a 32-bit-limb multiplier near 0.5 of this peak"}
    line = {
        "metric": "create_proof_s", "value": ms_dev / 1e3, "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
        "higher_is_better": False, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "u32 limbs (254-bit Montgomery integers)",
        "data": "synthetic",
        "config": {"workload": workload_name(k), "l2": f"inputs larger than L2 (witness {advice_bytes / 1e9:.2f} GB, every stage streams multi-GB device-resident columns)",
                   "parallelism": "1 GPU" if world == 1 else f"one proof on {world} GPUs: commit batches dealt by column (remainder by point range), NTTs by column, h(X) by row slice, grand products by set; byte-identical to the single-GPU proof (checked after the timed region)",
                   "rng": "StdRng::seed_from_u64(0)",
                   "srs": "ParamsKZG::setup(k, ChaCha20Rng::from_seed([0;32])) generated on device"},
        "clocks": clocks,
        "e2e": {"value": ms_host / 1e3, "unit": "s", "h2d_bytes_per_step": advice_bytes, "d2h_bytes_per_step": len(proof) + acc_n * 16 * 128,
                "wall_s": wall_host, "api": "b200zk_create_proof (host witness in PAGEABLE memory, as a Rust Vec<Fr> is -> proof bytes on host)",
                "pinned": {"value": ms_pin / 1e3, "wall_s": wall_pin, "note": "the same call from page-locked host memory"}},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "stages_ms": {k_: round(v * 1e3, 2) for k_, v in stages.items()},
        "setup_s": {"srs_device": round(t_srs, 2), "keygen_pk": round(t_keygen, 2)},
        "proof_bytes": len(proof),
        "kernels_in_profiled_step": {"msm_accumulate": {"ms": acc_ms, "launches": acc_n, "mixed_additions": acc_adds},
                                     "ntt_pass": {"ms": ntt_ms, "launches": ntt_n, "butterflies": ntt_bfly,
                                                  "Gbutterfly_s": ntt_bfly / (ntt_ms * 1e-3) / 1e9 if ntt_ms else None},
                                     "quotient": {"ms": q_ms, "launches": q_n}},
    }
    if dist is not None:
        # per-rank stage split of the one profiled step (stream synchronised at every stage boundary and around every
        # collective): `comm` = wall time inside collectives incl. waiting for the slowest peer, already contained in the stages
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {k_: round(v * 1e3, 2) for k_, v in stages.items()})
        line["per_rank_stages_ms"] = per_rank
    if rank == 0 and world == 1 and not args.no_extras:  # single-GPU side lines (a sharded context would wait for its peers)
        line.update(side_lines(ctx, stream, torch, np, k))
    pk.close()
    ctx.close()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # the contract asks for it at N=1 only
        step, cores, opk = cpu_prover(args.sample_k)
        step()  # warm-up (page faults, thread pool)
        secs, oproof = step()
        scale = float(1 << (k - args.sample_k))
        line["cpu_baseline"] = {"value": secs * scale, "unit": "s", "cores": cores, "kind": "port",
                                "sample": f"bounded sample: oracle create_proof on the same shape at k={args.sample_k} ({secs:.3f} s after one warm-up run), scaled "
                                          f"x{int(scale)} (rows) to k={k}; the reference arm (bench.py --impl reference) times the whole k={k} configuration; "
                                          "restated halo2 CPU algorithms (C++ oracle, std::thread), not the rayon binary"}
    if rank == 0:
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def side_lines(ctx, stream, torch, np, k):
    """MSM Mpts/s and NTT GB/s (the other two parts of BASELINE.json's metric) at the workload's size, device-resident."""
    n = 1 << k
    rng = np.random.default_rng(0)
    host = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    host[:, 3] &= np.uint64((1 << 60) - 1)  # 252-bit limbs: canonical (< r) whatever the other limbs are
    buf = torch.empty(4 * n * 4, dtype=torch.int64, device="cuda")
    for b in range(4):
        ctx.h2d(buf.data_ptr() + 32 * n * b, host)

    def ev(fn, iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            fn()
        ctx.sync()
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        ctx.sync()
        return e0.elapsed_time(e1) / iters

    msm_ms = ev(lambda: ctx.msm_dev(buf.data_ptr(), n, 0), 5)
    # NTT: 4 columns of 2^k (forward, standard root) and one coset NTT n -> 4n
    omega = np.empty(4, dtype=np.uint64)
    # the domain generator = ROOT_OF_UNITY^(2^(28-k)); computed with the library's own field kernels
    root = np.array([[0xd34f1ed960c37c9c, 0x3215cf6dd39329c8, 0x98865ea93dd31f74, 0x03ddb9f5166d18b7]], dtype=np.uint64)
    w = ctx.field_vec_op(0, 6, root)  # to Montgomery form
    for _ in range(28 - k):
        w = ctx.field_vec_op(0, 2, w, w)
    omega[:] = w[0]
    ntt_batch_ms = ev(lambda: ctx.ntt_dev(buf.data_ptr(), k, omega, 4, n), 5)
    ext = torch.empty(4 * n * 4, dtype=torch.int64, device="cuda")
    coset_ms = ev(lambda: ctx.coeff_to_extended_dev(k, buf.data_ptr(), ext.data_ptr()), 5)
    out = {
        "msm": {"n": n, "scalars": "uniform Fr", "ms": msm_ms, "Mpts_s": n / msm_ms / 1e3},
        "ntt": {"n": n, "batch": 4, "ms": ntt_batch_ms, "GBps_64nB": 64.0 * n * 4 / (ntt_batch_ms * 1e-3) / 1e9,
                "Gbutterfly_s": 4 * (n // 2) * k / (ntt_batch_ms * 1e-3) / 1e9},
        "coset_ntt": {"n": n, "ms": coset_ms, "GBps_160n": 160.0 * n / (coset_ms * 1e-3) / 1e9},
    }
    del buf, ext
    return out


if __name__ == "__main__":
    sys.exit(main())
