"""Row (b)/(e): ONE process, several GPUs — b200zk_create_multi. The group context must return exactly what a single-GPU
context returns: SRS, keygen commitments, MSM results and proof bytes (and the oracle's proof bytes)."""
import numpy as np
import pytest

import b200zk
import oracle_lib as O

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch

    return torch.cuda.device_count()


def test_single_device_group_is_a_plain_context():
    c = b200zk.Context.multi([0])
    try:
        assert c.group_size() == 1
        c.srs_setup(6)
        a = O.random_fr(np.random.default_rng(0), 64)
        g, gl = c.srs_download()
        assert np.array_equal(c.msm(a, 1), O.msm(a, gl))
    finally:
        c.close()


def test_duplicate_devices_are_rejected():
    with pytest.raises(b200zk.B200zkError) as e:
        b200zk.Context.multi([0, 0])
    assert e.value.code == b200zk.EINVAL


@pytest.mark.parametrize("ndev", [2, 4])
@pytest.mark.parametrize("shape", [(9, 3, 1, 1), (11, 5, 2, 1), (12, 9, 3, 2)])
def test_group_context_matches_single_gpu_and_oracle(ndev, shape):
    if _ngpu() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    import torch  # noqa: F401  (loads libnccl.so.2 into the process, as a host application would)

    k, A, L, F = shape
    fixed, advice, copies = b200zk.synth_circuit(k, A, L, F, seed=k)
    one = b200zk.Context(0)
    grp = b200zk.Context.multi(list(range(ndev)))
    try:
        assert grp.group_size() == ndev
        one.srs_setup(k)
        grp.srs_setup(k)
        g1, gl1 = one.srs_download()
        g2, gl2 = grp.srs_download()
        assert np.array_equal(g1, g2) and np.array_equal(gl1, gl2)
        rng = np.random.default_rng(k)
        a = O.random_fr(rng, 1 << k)
        assert np.array_equal(grp.msm(a, 1), one.msm(a, 1)) and np.array_equal(grp.msm(a, 0), one.msm(a, 0))
        pk1 = one.keygen(k, A, L, F, fixed, copies)
        pk2 = grp.keygen(k, A, L, F, fixed, copies)
        f1, p1 = pk1.commitments()
        f2, p2 = pk2.commitments()
        assert np.array_equal(f1, f2) and np.array_equal(p1, p2)
        want = pk1.create_proof(advice, 7)
        got = pk2.create_proof(advice, 7)
        assert got == want
        assert pk2.create_proof(advice, 7) == want  # and again: arenas, communicators and staging buffers are reusable
        timed, stages = pk2.create_proof(advice, 7, timings=True)  # the timed mode (collectives bracketed) on every rank alike
        assert timed == want and stages["msm"] > 0
        params = O.Params.setup(k)
        opk = O.ProvingKey(params, k, A, L, F, fixed, copies)
        assert got == opk.create_proof(advice, 7)
        pk2.close()
        pk1.close()
    finally:
        grp.close()
        one.close()
