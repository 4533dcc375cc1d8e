"""CPU-side checks of the product library: it loads, exports every symbol include/b200zk.h declares, refuses to run
without a GPU (no CPU fallback), and its host-only code paths (shared field/curve code, G1 sum, synthetic circuits)
are correct. No compute call needs a GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest

import b200zk
import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "b200zk.h")).read()
    declared = set(re.findall(r"B200ZK_API[^;(]*?\b(b200zk_\w+)\s*\(", header))
    assert len(declared) >= 40
    lib = b200zk.lib()
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++ or torch types), and a C translation unit that
    references every declared entry point must link against the shared library."""
    import shutil
    import subprocess

    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    header = os.path.join(ROOT, "include", "b200zk.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", header])
    names = sorted(set(re.findall(r"B200ZK_API[^;(]*?\b(b200zk_\w+)\s*\(", open(header).read())))
    src = tmp_path / "link_all.c"
    src.write_text('#include "b200zk.h"\n#include <stdio.h>\nint main(void) {\n  void* p[] = {%s};\n  printf("%%d\\n", (int)(sizeof p / sizeof p[0]));\n  return 0;\n}\n'
                   % ", ".join("(void*)" + n for n in names))
    libdir = os.path.join(ROOT, "halo2-plonky2-verifier_b200")
    exe = tmp_path / "link_all"
    subprocess.check_call(["gcc", "-std=gnu99", "-I", os.path.join(ROOT, "include"), str(src), "-L", libdir, "-lb200zk", "-Wl,-rpath," + libdir, "-o", str(exe)])


def test_plain_c_host_builds_and_fails_loudly_without_a_gpu(tmp_path):
    """tools/prove_c.c — a C99 host of the library (SRS setup, keygen, create_proof through include/b200zk.h, no Python in
    the process) — must compile and link; without a device it must fail at b200zk_create, not fall back to anything."""
    import shutil
    import subprocess

    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    b200zk.synth_circuit(5, 1, 1, 1)  # makes sure workload/libfriworkload.so is built
    exe = tmp_path / "prove_c"
    libdir, wdir = os.path.join(ROOT, "halo2-plonky2-verifier_b200"), os.path.join(ROOT, "workload")
    subprocess.check_call(["gcc", "-O1", "-std=gnu99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "prove_c.c"),
                           "-L", libdir, "-lb200zk", "-L", wdir, "-lfriworkload", "-Wl,-rpath," + libdir, "-Wl,-rpath," + wdir, "-o", str(exe)])
    import torch

    if torch.cuda.is_available():
        out = subprocess.run([str(exe), "8", "2", "1", "1", "1", "2"], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and '"proof_bytes"' in out.stdout, out.stderr
    else:
        out = subprocess.run([str(exe), "6", "2", "1", "1", "1", "1"], capture_output=True, text=True, timeout=60)
        assert out.returncode != 0 and "b200zk_create" in out.stderr


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(b200zk.B200zkError) as e:
        b200zk.Context(0)
    assert e.value.code == b200zk.ENODEV


def test_host_selftest_and_g1_sum():
    assert b200zk.lib().b200zk_host_selftest(ctypes.c_uint64(7), ctypes.c_size_t(3000)) == 0
    G = O.g1_generator()
    pts = np.stack([O.g1_mul(G, O.to_mont(v)) for v in (5, 11, O.R_MOD - 16, 0, 9)])
    out = np.empty(8, dtype=np.uint64)
    assert b200zk.lib().b200zk_g1_sum_host(pts.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(pts)), out.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.array_equal(out, O.g1_mul(G, O.to_mont(9)))
    assert b200zk.lib().b200zk_proof_size(20, 14, 3, 1) == O.lib().oracle_proof_size(20, 14, 3, 1) == 5632


def test_synth_rejects_bad_arguments():
    with pytest.raises(b200zk.B200zkError):
        b200zk.synth_circuit(3, 1, 1, 1)


def test_workload_generator_is_not_in_the_product_library():
    """The synthetic-circuit generator is its own host library (workload/): the product library exports no test-input code,
    so bench.py's reference arm can build its input without mapping libb200zk."""
    import workload

    assert not hasattr(b200zk.lib(), "b200zk_synth_circuit")
    f, a, c = workload.synth_circuit(6, 2, 1, 1, seed=5)
    f2, a2, c2 = b200zk.synth_circuit(6, 2, 1, 1, seed=5)
    assert np.array_equal(f, f2) and np.array_equal(a, a2) and np.array_equal(c, c2)
