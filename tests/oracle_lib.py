"""ctypes binding of oracle/liboracle.so — the CPU restatement used ONLY as the checker
(tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs).

Field elements travel as numpy uint64 arrays of shape (..., 4): little-endian limbs in Montgomery form,
exactly the halo2curves in-memory layout. Points are (..., 8): x limbs then y limbs.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
MONT_R = 1 << 256


def build_oracle(force=False):
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.oracle_last_error.restype = ctypes.c_char_p
        for name in ("oracle_params_setup", "oracle_params_from_trapdoor", "oracle_params_load", "oracle_keygen", "oracle_verifier_new"):
            getattr(_lib, name).restype = ctypes.c_void_p
        for name in ("oracle_create_proof", "oracle_proof_size", "oracle_transcript_script"):
            getattr(_lib, name).restype = ctypes.c_size_t
    return _lib


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def last_error():
    return lib().oracle_last_error().decode()


# ---- python big-int <-> limb helpers (independent of the oracle: used to pin it) --------------------------
def int_to_limbs(v):
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def limbs_to_int(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1)
    return sum(int(a[i]) << (64 * i) for i in range(len(a)))


def to_mont(v, mod=R_MOD):
    return int_to_limbs(v * MONT_R % mod)


def from_mont(a, mod=R_MOD):
    return limbs_to_int(a) * pow(MONT_R, -1, mod) % mod


def fr_array(values):
    """list of python ints (canonical) -> (n,4) Montgomery limbs"""
    out = np.empty((len(values), 4), dtype=np.uint64)
    for i, v in enumerate(values):
        out[i] = to_mont(v % R_MOD)
    return out


def fr_ints(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [from_mont(x) for x in a]


def random_fr(rng, n):
    """n uniformly random canonical Fr in Montgomery form, from a numpy Generator (fast path: any value < r
    is a valid Montgomery representative of SOME element, and the map is a bijection, so uniform limbs
    rejected to < r are uniform field elements)."""
    out = np.empty((n, 4), dtype=np.uint64)
    filled = 0
    top = R_MOD >> 192
    while filled < n:
        m = n - filled
        cand = rng.integers(0, 1 << 64, size=(m + 16, 4), dtype=np.uint64)
        cand[:, 3] &= np.uint64((1 << 62) - 1)
        ok = cand[:, 3] < np.uint64(top)  # strictly below the top limb of r: certainly < r
        cand = cand[ok][:m]
        out[filled : filled + len(cand)] = cand
        filled += len(cand)
    return out


def field_op(which, op, a, b=None):
    out = np.empty(4, dtype=np.uint64)
    a = np.ascontiguousarray(a, dtype=np.uint64)
    bb = np.ascontiguousarray(b, dtype=np.uint64) if b is not None else None
    lib().oracle_field_op(which, op, ptr(a), ptr(bb) if bb is not None else None, ptr(out))
    return out


def g1_generator():
    return np.concatenate([to_mont(1, Q_MOD), to_mont(2, Q_MOD)])


def g1_mul(p, s):
    out = np.empty(8, dtype=np.uint64)
    p = np.ascontiguousarray(p, dtype=np.uint64)
    s = np.ascontiguousarray(s, dtype=np.uint64)
    lib().oracle_g1_mul(ptr(p), ptr(s), ptr(out))
    return out


def g1_add(p, q):
    out = np.empty(8, dtype=np.uint64)
    p = np.ascontiguousarray(p, dtype=np.uint64)
    q = np.ascontiguousarray(q, dtype=np.uint64)
    lib().oracle_g1_add(ptr(p), ptr(q), ptr(out))
    return out


def g1_affine_ints(p):
    p = np.asarray(p, dtype=np.uint64).reshape(8)
    return from_mont(p[:4], Q_MOD), from_mont(p[4:], Q_MOD)


def msm(scalars, bases, naive=False):
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    bases = np.ascontiguousarray(bases, dtype=np.uint64)
    out = np.empty(8, dtype=np.uint64)
    fn = lib().oracle_naive_msm if naive else lib().oracle_msm
    fn(ptr(scalars), ptr(bases), ctypes.c_size_t(len(scalars)), ptr(out))
    return out


def best_fft(a, log_n, omega):
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    omega = np.ascontiguousarray(omega, dtype=np.uint64)
    lib().oracle_best_fft(ptr(a), ctypes.c_uint32(log_n), ptr(omega))
    return a


def domain_constant(k, which):
    out = np.empty(4, dtype=np.uint64)
    lib().oracle_domain_constant(ctypes.c_uint32(k), ctypes.c_int(which), ptr(out))
    return out


def lagrange_to_coeff(k, a):
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    lib().oracle_lagrange_to_coeff(ctypes.c_uint32(k), ptr(a))
    return a


def coeff_to_extended(k, a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty((4 << k, 4), dtype=np.uint64)
    lib().oracle_coeff_to_extended(ctypes.c_uint32(k), ptr(a), ptr(out))
    return out


def extended_to_coeff(k, a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty((3 << k, 4), dtype=np.uint64)
    lib().oracle_extended_to_coeff(ctypes.c_uint32(k), ptr(a), ptr(out))
    return out


def eval_polynomial(poly, point):
    poly = np.ascontiguousarray(poly, dtype=np.uint64)
    point = np.ascontiguousarray(point, dtype=np.uint64)
    out = np.empty(4, dtype=np.uint64)
    lib().oracle_eval_polynomial(ptr(poly), ctypes.c_size_t(len(poly)), ptr(point), ptr(out))
    return out


def set_compat(flags=0, random_poly_chunks=0):
    """The [UNVERIFIED-1..4] switches of SURVEY.md §8c on the oracle side (oracle/curve.hpp `Compat`): same bits as
    b200zk_set_compat. Global to the oracle library; tests reset it to (0, 0)."""
    lib().oracle_set_compat(ctypes.c_uint32(flags), ctypes.c_uint32(random_poly_chunks))


class Params:
    """ParamsKZG<Bn256> (oracle side). `setup(k)` mirrors halo2-base gen_srs: ChaCha20Rng::from_seed([0;32])."""

    def __init__(self, handle, k):
        self.h, self.k, self.n = handle, k, 1 << k

    @classmethod
    def setup(cls, k, seed=bytes(32)):
        return cls(lib().oracle_params_setup(ctypes.c_uint32(k), seed), k)

    @classmethod
    def from_trapdoor(cls, k, s):
        s = np.ascontiguousarray(s, dtype=np.uint64)
        return cls(lib().oracle_params_from_trapdoor(ctypes.c_uint32(k), ptr(s)), k)

    @classmethod
    def load(cls, k, s, g, g_lagrange):
        s = np.ascontiguousarray(s, dtype=np.uint64)
        g = np.ascontiguousarray(g, dtype=np.uint64)
        gl = np.ascontiguousarray(g_lagrange, dtype=np.uint64)
        return cls(lib().oracle_params_load(ctypes.c_uint32(k), ptr(s), ptr(g), ptr(gl)), k)

    def get(self):
        s = np.empty(4, dtype=np.uint64)
        g = np.empty((self.n, 8), dtype=np.uint64)
        gl = np.empty((self.n, 8), dtype=np.uint64)
        lib().oracle_params_get(ctypes.c_void_p(self.h), ptr(s), ptr(g), ptr(gl))
        return s, g, gl

    def commit(self, poly, lagrange=False):
        poly = np.ascontiguousarray(poly, dtype=np.uint64)
        out = np.empty(8, dtype=np.uint64)
        lib().oracle_commit(ctypes.c_void_p(self.h), ctypes.c_int(int(lagrange)), ptr(poly), ptr(out))
        return out

    def lagrange_via_group_fft(self):
        out = np.empty((self.n, 8), dtype=np.uint64)
        lib().oracle_lagrange_via_group_fft(ctypes.c_void_p(self.h), ptr(out))
        return out

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().oracle_params_free(ctypes.c_void_p(self.h))
                self.h = None
        except Exception:  # interpreter shutdown
            pass


class ProvingKey:
    def __init__(self, params, k, A, L, F, fixed, copies):
        self.params, self.shape = params, (k, A, L, F)
        fixed = np.ascontiguousarray(fixed, dtype=np.uint64)
        copies = np.ascontiguousarray(copies, dtype=np.uint32).reshape(-1, 4)
        self.h = lib().oracle_keygen(ctypes.c_void_p(params.h), k, A, L, F, ptr(fixed), ptr(copies), ctypes.c_size_t(len(copies)))
        if not self.h:
            raise RuntimeError(last_error())

    def get(self, which, idx=0):
        k, A, L, F = self.shape
        n = 1 << k
        shape = {0: (F + 1 + A, 8), 1: (F + A + L, 8), 2: (n, 4)}.get(which, (4 * n, 4))
        out = np.empty(shape, dtype=np.uint64)
        lib().oracle_pk_get(ctypes.c_void_p(self.h), which, idx, ptr(out))
        return out

    def transcript_repr(self):
        out = np.empty(4, dtype=np.uint64)
        lib().oracle_pk_transcript_repr(ctypes.c_void_p(self.h), ptr(out))
        return out

    def set_gates(self, calcs, constants=(), results=()):
        """Custom gates (same encoding as b200zk's ProvingKey.set_gates); used by evaluate_h, create_proof and verify."""
        arr = np.zeros((len(calcs), 7), dtype=np.uint32)
        for j, (op, a, b) in enumerate(calcs):
            b = b if b is not None else (0, 0, 0)
            arr[j] = [op, a[0], a[1], np.int32(a[2]).astype(np.uint32), b[0], b[1], np.int32(b[2]).astype(np.uint32)]
        cs = np.ascontiguousarray(np.asarray(constants, dtype=np.uint64).reshape(-1, 4)) if len(constants) else np.zeros((0, 4), dtype=np.uint64)
        rs = np.ascontiguousarray(results, dtype=np.uint32)
        lib().oracle_pk_set_gates(ctypes.c_void_p(self.h), ptr(arr), ctypes.c_size_t(len(calcs)), ptr(cs), ctypes.c_size_t(len(cs)), ptr(rs),
                                  ctypes.c_size_t(len(rs)))

    def create_proof(self, advice, rng_seed=0):
        advice = np.ascontiguousarray(advice, dtype=np.uint64)
        size = lib().oracle_proof_size(*self.shape)
        buf = np.empty(size, dtype=np.uint8)
        secs = ctypes.c_double(0)
        n = lib().oracle_create_proof(ctypes.c_void_p(self.params.h), ctypes.c_void_p(self.h), ptr(advice), ctypes.c_uint64(rng_seed), ptr(buf),
                                      ctypes.byref(secs))
        if n == 0:
            raise RuntimeError(last_error())
        assert n == size
        self.last_seconds = secs.value
        return buf.tobytes()

    def evaluate_h(self, advice_coeff, z_coeff, lookup_coeff, y, beta, gamma):
        """evaluate_h + divide_by_vanishing_poly on coefficient-form inputs -> [4n, 4]"""
        k, _, L, _ = self.shape
        n = 1 << k
        out = np.empty((4 * n, 4), dtype=np.uint64)
        lk = np.ascontiguousarray(lookup_coeff) if L else None
        ok = lib().oracle_evaluate_h(ctypes.c_void_p(self.h), ptr(np.ascontiguousarray(advice_coeff)), ptr(np.ascontiguousarray(z_coeff)),
                                     ptr(lk) if L else None, ptr(y), ptr(beta), ptr(gamma), ptr(out))
        assert ok, last_error()
        return out

    def verify(self, proof, pairing=False):
        """verify_proof; `pairing=True` checks the opening with the real pairing equation instead of the trapdoor."""
        buf = np.frombuffer(proof, dtype=np.uint8)
        fn = lib().oracle_verify_proof_pairing if pairing else lib().oracle_verify_proof
        ok = fn(ctypes.c_void_p(self.params.h), ctypes.c_void_p(self.h), ptr(buf), ctypes.c_size_t(len(buf)))
        return bool(ok), last_error()

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().oracle_pk_free(ctypes.c_void_p(self.h))
                self.h = None
        except Exception:
            pass


def mock_check(k, A, L, F, fixed, advice, copies):
    fixed = np.ascontiguousarray(fixed, dtype=np.uint64)
    advice = np.ascontiguousarray(advice, dtype=np.uint64)
    copies = np.ascontiguousarray(copies, dtype=np.uint32).reshape(-1, 4)
    ok = lib().oracle_mock_check(k, A, L, F, ptr(fixed), ptr(advice), ptr(copies), ctypes.c_size_t(len(copies)))
    return bool(ok), last_error()


class Verifier:
    """Oracle verifier over externally produced vk commitments (e.g. the GPU keygen's): verifies proofs at sizes where the
    oracle's own keygen would take minutes. Needs the SRS trapdoor (openings are checked in G1 instead of by a pairing)."""

    def __init__(self, k, A, L, F, trapdoor, fixed_commitments, perm_commitments, transcript_repr):
        t = np.ascontiguousarray(trapdoor, dtype=np.uint64)
        fc = np.ascontiguousarray(fixed_commitments, dtype=np.uint64)
        pc = np.ascontiguousarray(perm_commitments, dtype=np.uint64)
        tr = np.ascontiguousarray(transcript_repr, dtype=np.uint64)
        assert fc.shape == (F + 1 + A, 8) and pc.shape == (F + A + L, 8)
        self.h = lib().oracle_verifier_new(k, A, L, F, ptr(t), ptr(fc), ptr(pc), ptr(tr))

    def verify(self, proof, pairing=False):
        buf = np.frombuffer(proof, dtype=np.uint8)
        ok = lib().oracle_verifier_verify(ctypes.c_void_p(self.h), ptr(buf), ctypes.c_size_t(len(buf)), ctypes.c_int(int(pairing)))
        return bool(ok), last_error()

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().oracle_verifier_free(ctypes.c_void_p(self.h))
                self.h = None
        except Exception:
            pass
