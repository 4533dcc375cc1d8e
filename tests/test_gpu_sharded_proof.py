"""Row (e) on real hardware: one proof computed by 2 (and, where the box has them, 4) ranks — commit batches dealt by
column with the remainder split by point range, per-column NTTs, h(X) row slices, grand products by set — must be
byte-identical to the single-GPU proof and to the oracle's. Needs >= 2 GPUs (skipped otherwise); launched through torchrun
like bench.py. The 4-rank shapes put permutation-product columns into the point-range remainder of their commit batch
(6 sets + the random polynomial on 4 ranks), which 2 ranks can never do."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    import numpy as np, torch, torch.distributed as dist
    import b200zk, oracle_lib as O
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    for (k, A, L, F) in [(10, 4, 1, 2), (12, 5, 3, 1), (11, 3, 0, 1), (11, 10, 0, 1), (10, 9, 1, 2)]:
        fixed, advice, copies = b200zk.synth_circuit(k, A, L, F, seed=k)
        ctx = b200zk.Context(rank)
        ctx.srs_setup(k)
        single = ctx.keygen(k, A, L, F, fixed, copies).create_proof(advice, 5)        # un-sharded on this GPU
        ctx.set_allgather(rank, world, b200zk.torch_allgather(dist, torch.device("cuda", rank)))
        pk = ctx.keygen(k, A, L, F, fixed, copies)                                     # sharded commits in keygen too
        proof = pk.create_proof(advice, 5)
        assert proof == single, (rank, k)
        if rank == 0:
            params = O.Params.setup(k)
            opk = O.ProvingKey(params, k, A, L, F, fixed, copies)
            assert proof == opk.create_proof(advice, 5)
            assert opk.verify(proof)[0]
        ctx.set_allgather(0, 1, None)
        dist.barrier()
    print("rank", rank, "sharded ok", flush=True)
    dist.destroy_process_group()
""") % (ROOT, ROOT)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_proof_is_identical(tmp_path, world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("sharded ok") == world, r.stdout  # (lines of the ranks may interleave)
