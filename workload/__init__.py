"""Synthetic workloads for bench.py and the tests: satisfying circuits of the halo2-base shape that replay the cell mix of the
reference's FRI-verifier circuit (see synth.cpp). Host only; a separate small library so that neither bench arm has to map
the other arm's code to obtain its input."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfriworkload.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.friworkload_max_copies.restype = ctypes.c_size_t
    return _lib


def synth_circuit(k, A, L, F, seed=0):
    """Returns (fixed[(F+1+A), n, 4], advice[(A+L), n, 4], copies[m, 4]) with every gate, lookup and copy constraint of
    the halo2-base shape satisfied; field elements are uint64 Montgomery limbs as halo2curves stores them."""
    n = 1 << k
    fixed = np.zeros((F + 1 + A, n, 4), dtype=np.uint64)
    advice = np.zeros((A + L, n, 4), dtype=np.uint64)
    maxc = int(lib().friworkload_max_copies(k, A, L, F))
    copies = np.zeros((maxc, 4), dtype=np.uint32)
    nc = ctypes.c_size_t(0)
    rc = lib().friworkload_synth_circuit(k, A, L, F, ctypes.c_uint64(seed), fixed.ctypes.data_as(ctypes.c_void_p), advice.ctypes.data_as(ctypes.c_void_p),
                                         copies.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nc))
    if rc != 0:
        raise ValueError(f"synth_circuit({k}, {A}, {L}, {F}) failed with code {rc}")
    return fixed, advice, copies[: nc.value].copy()
