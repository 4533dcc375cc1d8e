"""The CPU arm of bench.py (`--impl reference`: the oracle port timed on the host cores) prints exactly one JSON line
with the keys the driver reads. The GPU arm needs a device and is covered by the round-end bench itself."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "5", "--warmup", "3", "--k", "9", "--ref-max-steps", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "create_proof_s" and d["unit"] == "s" and d["higher_is_better"] is False
    # the arm times the configuration itself (here k=9), never a scaled sample; `steps` = the steps actually timed
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["steps_requested"] == 5 and d["warmup"] == 0 and d["value"] > 0 and abs(d["ms_per_step"] - 1e3 * d["value"]) < 1e-6 * d["ms_per_step"] + 1e-9
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "k=9" in d["config"]["workload"] and "k=9" in d["cpu_baseline"]["sample"]


def test_reference_arm_maps_no_product_library():
    """The CPU arm must not load libb200zk (its input comes from workload/, its prover from oracle/)."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--k', '7', '--steps', '1']\n"
            "try:\n    runpy.run_path(%r, run_name='__main__')\nexcept SystemExit:\n    pass\n"
            "maps = open('/proc/self/maps').read()\n"
            "sys.stderr.write('MAPPED_B200ZK=%%d MAPPED_ORACLE=%%d\\n' %% ('libb200zk' in maps, 'liboracle' in maps))\n") % os.path.join(ROOT, "bench.py")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "MAPPED_B200ZK=0 MAPPED_ORACLE=1" in r.stderr, r.stderr[-2000:]


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch

    if torch.cuda.is_available():
        import pytest

        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
