"""N > 1 host logic on CPU: the torch.distributed all-gather used to exchange partial MSM sums (world_size 2, gloo) and
the host-side combine of per-rank partial points."""
import ctypes
import os
import subprocess
import sys
import textwrap

import numpy as np

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, ctypes
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    import numpy as np, torch.distributed as dist
    import b200zk, oracle_lib as O
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    fn = b200zk.torch_allgather(dist)
    G = O.g1_generator()
    mine = O.g1_mul(G, O.to_mont(1000 + rank))          # this rank's "partial MSM result"
    out = fn(mine.tobytes())
    pts = np.frombuffer(out, dtype=np.uint64).reshape(world, 8).copy()
    assert np.array_equal(pts[rank], mine)
    total = np.empty(8, dtype=np.uint64)
    assert b200zk.lib().b200zk_g1_sum_host(pts.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(world), total.ctypes.data_as(ctypes.c_void_p)) == 0
    want = O.g1_mul(G, O.to_mont(sum(1000 + r for r in range(world))))
    assert np.array_equal(total, want)
    dist.barrier()
    print("rank-" + str(rank) + "-ok", flush=True)
""") % (ROOT, ROOT)


def test_allgather_and_combine_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    import socket

    with socket.socket() as sock:  # a free rendezvous port
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("-ok") == 2, r.stdout  # the two ranks' lines may interleave
