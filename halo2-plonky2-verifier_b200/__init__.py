"""b200zk — B200-native KZG-BN254 proving backend (host-side Python mirror of the C ABI in include/b200zk.h).

The product path is libb200zk.so (hand-written sm_100a CUDA behind an extern "C" boundary). This module only
binds it with ctypes; there is NO CPU fallback — without the built library or without a CUDA device every entry
point raises.

Field elements are numpy uint64 arrays of shape (..., 4) (halo2curves Montgomery limbs); G1 affine points are
(..., 8).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__)) if "__file__" in globals() else os.getcwd()
if os.path.basename(_HERE) == "b200zk":  # executed through the alias package
    _HERE = os.path.join(os.path.dirname(_HERE), "halo2-plonky2-verifier_b200")
LIB_PATH = os.environ.get("B200ZK_LIB") or os.path.join(_HERE, "libb200zk.so")  # B200ZK_LIB: an experimental build for A/B runs
CSRC = os.path.join(_HERE, "csrc")

OK, ENODEV, EINVAL, ECUDA, ESTATE, ESYNTH = 0, -1, -2, -3, -4, -5


class B200zkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200zk error {code}: {msg}")
        self.code = code


def build(verbose=False):
    """Compile libb200zk.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", CSRC, "-j8"] + ([] if verbose else ["-s"]))
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200zkError(ENODEV, f"{LIB_PATH} is not built (run __graft_entry__.build()); there is no CPU fallback")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.b200zk_last_error.restype = ctypes.c_char_p
        _lib.b200zk_stream.restype = ctypes.c_void_p
        _lib.b200zk_launch_count.restype = ctypes.c_ulonglong
        _lib.b200zk_proof_size.restype = ctypes.c_size_t
        _lib.b200zk_num_sets.restype = ctypes.c_uint32
        _lib.b200zk_srs_file_size.restype = ctypes.c_size_t
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _c(a, dtype=np.uint64):
    return np.ascontiguousarray(a, dtype=dtype)


def launch_count():
    return int(lib().b200zk_launch_count())


ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)


def torch_allgather(dist, device=None):
    """All-gather of a small byte string over torch.distributed (NCCL when `device` is a CUDA device, gloo otherwise)."""
    import torch

    world = dist.get_world_size()

    def fn(data):
        t = torch.frombuffer(bytearray(data), dtype=torch.uint8)
        if device is not None:
            t = t.to(device)
        out = torch.empty(world * len(data), dtype=torch.uint8, device=t.device)
        dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy().tobytes()

    return fn


class Context:
    """One CUDA device + stream (b200zk_create). Mirrors the reference-side objects that own prover state."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        rc = lib().b200zk_create(int(device), ctypes.byref(self._h))
        if rc != OK:
            raise B200zkError(rc, "no usable CUDA device (libb200zk has no CPU fallback)" if rc == ENODEV else "b200zk_create failed")
        self.device = device

    @classmethod
    def multi(cls, devices):
        """b200zk_create_multi: ONE process driving several GPUs (one worker thread per device inside the library, NCCL among
        them). srs_setup / srs_load / keygen / create_proof / msm on the returned context run sharded over all devices and
        return the single-GPU results byte for byte."""
        devices = [int(d) for d in devices]
        self = cls.__new__(cls)
        self._h = ctypes.c_void_p()
        arr = (ctypes.c_int * len(devices))(*devices)
        rc = lib().b200zk_create_multi(arr, len(devices), ctypes.byref(self._h))
        if rc != OK:
            raise B200zkError(rc, "b200zk_create_multi failed (no usable CUDA devices, duplicate ordinals, or NCCL not loadable)")
        self.device = devices[0]
        self.devices = devices
        return self

    def group_size(self):
        return int(lib().b200zk_group_size(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().b200zk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise B200zkError(rc, lib().b200zk_last_error(self._h).decode())

    @property
    def stream(self):
        return lib().b200zk_stream(self._h)

    def sync(self):
        self._check(lib().b200zk_sync(self._h))

    # ---- multi-GPU: MSM point-range shards + all-gather of the partial sums ----
    def set_allgather(self, rank, world, fn):
        """`fn(send: bytes) -> bytes` must return the concatenation of every rank's `send` in rank order (see
        `torch_allgather`). All ranks must then call the same MSM / create_proof sequence."""
        if world <= 1:
            self._check(lib().b200zk_set_allgather(self._h, 0, 1, None, None))
            self._ag = None
            return

        def trampoline(user, send, nbytes, recv):
            try:
                data = ctypes.string_at(send, nbytes)
                out = fn(data)
                if len(out) != nbytes * world:
                    return -1
                ctypes.memmove(recv, out, len(out))
                return 0
            except Exception:  # never unwind into C
                return -2

        self._ag = ALLGATHER_FN(trampoline)
        self._check(lib().b200zk_set_allgather(self._h, int(rank), int(world), self._ag, None))

    def comm_init(self):
        """Bring up the library's NCCL communicator (collective over all ranks) so that standalone MSM calls exchange their
        partial sums inside the library."""
        self._check(lib().b200zk_comm_init(self._h))

    COMPAT_NO_UNUSED_BLIND_DRAWS, COMPAT_LOOKUP_FILL_ASCENDING, COMPAT_POINT_SIGN_BIT7 = 1, 2, 4

    def set_compat(self, flags=0, random_poly_chunks=0):
        """The [UNVERIFIED-1..4] switches of SURVEY.md §8c (include/b200zk.h); (0, 0) = defaults."""
        self._check(lib().b200zk_set_compat(self._h, ctypes.c_uint32(flags), ctypes.c_uint32(random_poly_chunks)))

    def set_msm_affine_rounds(self, rounds=0):
        """Batched-affine pre-reduction rounds for dense MSM columns (0 = off, the default: measured slower on B200)."""
        self._check(lib().b200zk_set_msm_affine_rounds(self._h, int(rounds)))

    def set_msm_tables(self, on=True):
        self._check(lib().b200zk_set_msm_tables(self._h, int(bool(on))))

    # ---- per-kernel-family event timing (bench.py roofline section) ----
    PROF_IDS = {"msm_accumulate": 0, "ntt_pass": 2, "quotient": 3}

    def profile_enable(self, on=True):
        self._check(lib().b200zk_profile_enable(self._h, int(bool(on))))

    def profile_work(self, name):
        """Algorithmic work of the recorded spans (mixed additions / butterflies / rows); call before profile_get."""
        units = ctypes.c_double(0)
        self._check(lib().b200zk_profile_work(self._h, self.PROF_IDS[name], ctypes.byref(units)))
        return units.value

    def profile_get(self, name):
        ms = ctypes.c_double(0)
        cnt = ctypes.c_ulonglong(0)
        self._check(lib().b200zk_profile_get(self._h, self.PROF_IDS[name], ctypes.byref(ms), ctypes.byref(cnt)))
        return ms.value, int(cnt.value)

    # ---- raw device memory ----
    def dev_alloc(self, nbytes):
        p = ctypes.c_void_p()
        self._check(lib().b200zk_dev_alloc(self._h, ctypes.c_size_t(nbytes), ctypes.byref(p)))
        return p.value

    def dev_free(self, ptr):
        self._check(lib().b200zk_dev_free(self._h, ctypes.c_void_p(ptr)))

    def h2d(self, dst, src):
        src = np.ascontiguousarray(src)
        self._check(lib().b200zk_h2d(self._h, ctypes.c_void_p(dst), _p(src), ctypes.c_size_t(src.nbytes)))

    def d2h(self, src, shape, dtype=np.uint64):
        out = np.empty(shape, dtype=dtype)
        self._check(lib().b200zk_d2h(self._h, _p(out), ctypes.c_void_p(src), ctypes.c_size_t(out.nbytes)))
        return out

    # ---- row A ----
    def field_vec_op(self, field, op, a, b=None):
        a = _c(a).reshape(-1, 4)
        bb = _c(b).reshape(-1, 4) if b is not None else None
        out = np.empty_like(a)
        self._check(lib().b200zk_field_vec_op(self._h, field, op, _p(a), _p(bb), _p(out), ctypes.c_size_t(len(a))))
        return out

    def g1_vec_op(self, op, a, b=None):
        a = _c(a).reshape(-1, 8)
        bb = _c(b) if b is not None else None
        out = np.empty_like(a)
        self._check(lib().b200zk_g1_vec_op(self._h, op, _p(a), _p(bb), _p(out), ctypes.c_size_t(len(a))))
        return out

    # ---- row C / D ----
    def ntt(self, a, log_n, omega):
        """halo2_proofs::arithmetic::best_fft(a, omega, log_n); returns the transformed copy."""
        a = _c(a).reshape(-1, 4).copy()
        assert len(a) == 1 << log_n
        omega = _c(omega)
        self._check(lib().b200zk_ntt(self._h, _p(a), ctypes.c_uint32(log_n), _p(omega)))
        return a

    def ntt_dev(self, ptr, log_n, omega, batch=1, stride=0):
        omega = _c(omega)
        self._check(lib().b200zk_ntt_batch_dev(self._h, ctypes.c_void_p(ptr), ctypes.c_uint32(log_n), _p(omega), ctypes.c_uint32(batch),
                                               ctypes.c_size_t(stride)))

    def lagrange_to_coeff(self, k, a):
        a = _c(a).reshape(-1, 4).copy()
        assert len(a) == 1 << k
        self._check(lib().b200zk_lagrange_to_coeff(self._h, ctypes.c_uint32(k), _p(a)))
        return a

    def coeff_to_extended(self, k, a):
        a = _c(a).reshape(-1, 4)
        assert len(a) == 1 << k
        out = np.empty((4 << k, 4), dtype=np.uint64)
        self._check(lib().b200zk_coeff_to_extended(self._h, ctypes.c_uint32(k), _p(a), _p(out)))
        return out

    def extended_to_coeff(self, k, a):
        a = _c(a).reshape(-1, 4)
        assert len(a) == 4 << k
        out = np.empty((3 << k, 4), dtype=np.uint64)
        self._check(lib().b200zk_extended_to_coeff(self._h, ctypes.c_uint32(k), _p(a), _p(out)))
        return out

    def lagrange_to_coeff_dev(self, k, ptr, batch=1, stride=0):
        self._check(lib().b200zk_lagrange_to_coeff_dev(self._h, ctypes.c_uint32(k), ctypes.c_void_p(ptr), ctypes.c_uint32(batch), ctypes.c_size_t(stride)))

    def coeff_to_extended_dev(self, k, src, dst, batch=1, stride_in=0, stride_out=0):
        self._check(lib().b200zk_coeff_to_extended_dev(self._h, ctypes.c_uint32(k), ctypes.c_void_p(src), ctypes.c_void_p(dst), ctypes.c_uint32(batch),
                                                       ctypes.c_size_t(stride_in), ctypes.c_size_t(stride_out)))

    def extended_to_coeff_dev(self, k, src, dst):
        self._check(lib().b200zk_extended_to_coeff_dev(self._h, ctypes.c_uint32(k), ctypes.c_void_p(src), ctypes.c_void_p(dst)))

    # ---- rows B / K ----
    def srs_load(self, k, g, g_lagrange):
        """Copy ParamsKZG's `g` and `g_lagrange` (2^k affine points each) to the device."""
        g = _c(g).reshape(-1, 8)
        gl = _c(g_lagrange).reshape(-1, 8)
        assert len(g) == 1 << k and len(gl) == 1 << k
        self._check(lib().b200zk_srs_load(self._h, ctypes.c_uint32(k), _p(g), _p(gl)))
        self.srs_k = k

    def srs_setup(self, k, seed=bytes(32), trapdoor=None):
        """ParamsKZG::<Bn256>::setup(k, ChaCha20Rng::from_seed(seed)) generated on the device (halo2-base gen_srs uses
        seed [0;32]); returns the trapdoor s (tests use it to check openings without a pairing)."""
        out = np.empty(4, dtype=np.uint64)
        if trapdoor is not None:
            t = _c(trapdoor)
            self._check(lib().b200zk_srs_setup_trapdoor(self._h, ctypes.c_uint32(k), _p(t)))
            out[:] = t
        else:
            self._check(lib().b200zk_srs_setup(self._h, ctypes.c_uint32(k), bytes(seed), _p(out)))
        self.srs_k = k
        return out

    def srs_write(self, fmt=0):
        """ParamsKZG::write bytes (fmt 0 = RawBytes, 1 = Processed)."""
        size = int(lib().b200zk_srs_file_size(ctypes.c_uint32(self.srs_k), int(fmt)))
        buf = np.empty(size, dtype=np.uint8)
        n = ctypes.c_size_t(0)
        self._check(lib().b200zk_srs_write(self._h, int(fmt), _p(buf), ctypes.c_size_t(size), ctypes.byref(n)))
        return buf[: n.value].tobytes()

    def srs_read(self, data, fmt=0):
        """ParamsKZG::read: loads both G1 bases (validated on the device) and keeps the G2 bytes."""
        buf = np.frombuffer(data, dtype=np.uint8)
        self._check(lib().b200zk_srs_read(self._h, _p(buf), ctypes.c_size_t(len(buf)), int(fmt)))
        self.srs_k = int.from_bytes(data[:4], "little")

    def gen_srs(self, k, params_dir="params"):
        """halo2-base `utils::fs::gen_srs(k)`: read `<params_dir>/kzg_bn254_<k>.srs` if it exists, else run
        ParamsKZG::setup(k, ChaCha20Rng::from_seed([0;32])) and write the file (the reference keeps that directory out of
        git: /root/reference/.gitignore:5 `params/`). The file is ParamsKZG::write in RawBytes form. Returns True when the
        cached file was used. (The MSM window tables are rebuilt on the device either way: 0.11 s at k=20, less than
        reading their 1.7 GB back from disk would take.)"""
        path = os.path.join(params_dir, f"kzg_bn254_{k}.srs")
        if os.path.exists(path):
            with open(path, "rb") as f:
                self.srs_read(f.read(), fmt=0)
            return True
        self.srs_setup(k)
        os.makedirs(params_dir, exist_ok=True)
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(self.srs_write(fmt=0))
        os.replace(tmp, path)
        return False

    def srs_download(self):
        n = 1 << self.srs_k
        g = np.empty((n, 8), dtype=np.uint64)
        gl = np.empty((n, 8), dtype=np.uint64)
        self._check(lib().b200zk_srs_download(self._h, _p(g), _p(gl)))
        return g, gl

    def msm(self, scalars, basis=0):
        """ParamsKZG::commit (basis=0) / commit_lagrange (basis=1) == best_multiexp(scalars, bases); affine result."""
        scalars = _c(scalars).reshape(-1, 4)
        out = np.empty(8, dtype=np.uint64)
        self._check(lib().b200zk_msm(self._h, int(basis), _p(scalars), ctypes.c_size_t(len(scalars)), _p(out)))
        return out

    def msm_dev(self, scalars_ptr, n, basis=0):
        out = np.empty(8, dtype=np.uint64)
        self._check(lib().b200zk_msm_dev(self._h, int(basis), ctypes.c_void_p(scalars_ptr), ctypes.c_size_t(n), _p(out)))
        return out

    def msm_batch_dev(self, col_ptrs, n, basis=0):
        """Commit several device-resident columns over one basis (one bucket reduction for the batch)."""
        arr = (ctypes.c_void_p * len(col_ptrs))(*col_ptrs)
        out = np.empty((len(col_ptrs), 8), dtype=np.uint64)
        self._check(lib().b200zk_msm_batch_dev(self._h, int(basis), arr, ctypes.c_size_t(len(col_ptrs)), ctypes.c_size_t(n), _p(out)))
        return out

    def msm_batch(self, cols, basis=0):
        """ParamsKZG::commit / commit_lagrange over several HOST columns ([ncols, n, 4] uint64 or a list of [n, 4] arrays)."""
        cols = [_c(c).reshape(-1, 4) for c in cols]
        n = len(cols[0]) if cols else 0
        assert all(len(c) == n for c in cols)
        arr = (ctypes.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        out = np.empty((len(cols), 8), dtype=np.uint64)
        self._check(lib().b200zk_msm_batch(self._h, int(basis), arr, ctypes.c_size_t(len(cols)), ctypes.c_size_t(n), _p(out)))
        return out

    def msm_bases(self, scalars, bases):
        """halo2curves::msm::best_multiexp(coeffs, bases) with caller-supplied bases."""
        scalars = _c(scalars).reshape(-1, 4)
        bases = _c(bases).reshape(-1, 8)
        assert len(scalars) == len(bases)
        out = np.empty(8, dtype=np.uint64)
        self._check(lib().b200zk_msm_bases(self._h, _p(bases), _p(scalars), ctypes.c_size_t(len(scalars)), _p(out)))
        return out

    def msm_bases_dev(self, bases_ptr, scalars_ptr, n):
        out = np.empty(8, dtype=np.uint64)
        self._check(lib().b200zk_msm_bases_dev(self._h, ctypes.c_void_p(bases_ptr), ctypes.c_void_p(scalars_ptr), ctypes.c_size_t(n), _p(out)))
        return out

    # ---- rows E, F, I: column primitives ----
    def batch_invert(self, a):
        a = _c(a).reshape(-1, 4).copy()
        self._check(lib().b200zk_batch_invert(self._h, _p(a), ctypes.c_size_t(len(a))))
        return a

    def prefix_product(self, m, first):
        m = _c(m).reshape(-1, 4)
        first = _c(first)
        z = np.empty_like(m)
        self._check(lib().b200zk_prefix_product(self._h, _p(m), _p(first), _p(z), ctypes.c_size_t(len(m))))
        return z

    def eval_polynomial(self, poly, point):
        poly = _c(poly).reshape(-1, 4)
        point = _c(point)
        out = np.empty(4, dtype=np.uint64)
        self._check(lib().b200zk_eval_polynomial(self._h, _p(poly), ctypes.c_size_t(len(poly)), _p(point), _p(out)))
        return out

    def kate_division(self, a, b):
        a = _c(a).reshape(-1, 4)
        b = _c(b)
        q = np.empty((len(a) - 1, 4), dtype=np.uint64)
        self._check(lib().b200zk_kate_division(self._h, _p(a), ctypes.c_size_t(len(a)), _p(b), _p(q)))
        return q

    def permute_expression_pair(self, k, inp, table):
        inp = _c(inp).reshape(-1, 4)
        table = _c(table).reshape(-1, 4)
        u = (1 << k) - 7
        a = np.empty((u, 4), dtype=np.uint64)
        s = np.empty((u, 4), dtype=np.uint64)
        self._check(lib().b200zk_permute_expression_pair(self._h, ctypes.c_uint32(k), _p(inp), _p(table), _p(a), _p(s)))
        return a, s

    # ---- rows J, E-I: keygen + create_proof ----
    def keygen(self, k, A, L, F, fixed, copies):
        """plonk::keygen_vk + keygen_pk for the halo2-base shape; returns a device-resident ProvingKey."""
        return ProvingKey(self, k, A, L, F, fixed, copies)


TIMING_KEYS = ("upload", "msm", "ntt", "lookup", "products", "quotient", "evals", "shplonk", "other", "comm")


class ProvingKey:
    """halo2_proofs::plonk::ProvingKey (device-resident columns, host-side vk commitments)."""

    def __init__(self, ctx, k, A, L, F, fixed, copies):
        self.ctx, self.shape = ctx, (k, A, L, F)
        fixed = _c(fixed)
        copies = np.ascontiguousarray(copies, dtype=np.uint32).reshape(-1, 4)
        assert fixed.size == (F + 1 + A) * (1 << k) * 4
        self._h = ctypes.c_void_p()
        ctx._check(lib().b200zk_keygen(ctx._h, k, A, L, F, _p(fixed), _p(copies), ctypes.c_size_t(len(copies)), ctypes.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self.ctx._h:
            lib().b200zk_pk_free(self.ctx._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_gates(self, calcs, constants=(), results=()):
        """b200zk_pk_set_gates. `calcs`: list of (op, a, b) with a, b = (kind, index, rotation) (b may be None for unary
        calculations); `constants`: field elements (uint64[4] each); `results`: indices of the gate intermediates. An empty
        `calcs` restores the built-in halo2-base gates."""
        arr = np.zeros((len(calcs), 7), dtype=np.uint32)
        for j, (op, a, b) in enumerate(calcs):
            b = b if b is not None else (0, 0, 0)
            arr[j] = [op, a[0], a[1], np.int32(a[2]).astype(np.uint32), b[0], b[1], np.int32(b[2]).astype(np.uint32)]
        cs = _c(np.asarray(constants, dtype=np.uint64).reshape(-1, 4)) if len(constants) else None
        rs = np.ascontiguousarray(results, dtype=np.uint32)
        self.ctx._check(lib().b200zk_pk_set_gates(self.ctx._h, self._h, _p(arr) if len(calcs) else None, ctypes.c_size_t(len(calcs)), _p(cs),
                                                  ctypes.c_size_t(0 if cs is None else len(cs)), _p(rs) if len(rs) else None, ctypes.c_size_t(len(rs))))

    def commitments(self):
        k, A, L, F = self.shape
        fc = np.empty((F + 1 + A, 8), dtype=np.uint64)
        pc = np.empty((F + A + L, 8), dtype=np.uint64)
        self.ctx._check(lib().b200zk_pk_commitments(self.ctx._h, self._h, _p(fc), _p(pc)))
        return fc, pc

    def transcript_repr(self, set_to=None):
        out = np.empty(4, dtype=np.uint64)
        st = _c(set_to) if set_to is not None else None
        self.ctx._check(lib().b200zk_pk_transcript_repr(self.ctx._h, self._h, _p(out), _p(st)))
        return out

    def get_column(self, which, idx):
        n = 1 << self.shape[0]
        out = np.empty((n if which == 0 else 4 * n, 4), dtype=np.uint64)
        self.ctx._check(lib().b200zk_pk_get_column(self.ctx._h, self._h, which, idx, _p(out)))
        return out

    def proof_size(self):
        return int(lib().b200zk_proof_size(*self.shape))

    def num_sets(self):
        _, A, L, F = self.shape
        return int(lib().b200zk_num_sets(A, L, F))

    def evaluate_h(self, advice_coeff, perm_z_coeff, lookup_coeff, y, beta, gamma):
        """evaluation::Evaluator::evaluate_h + divide_by_vanishing_poly on coefficient-form inputs: advice_coeff
        [(A+L), n, 4], perm_z_coeff [num_sets, n, 4], lookup_coeff [L, 3, n, 4] (Z, a', s' per lookup) or None;
        returns h on the extended domain [4n, 4]."""
        k, A, L, F = self.shape
        n = 1 << k
        advice_coeff, perm_z_coeff = _c(advice_coeff), _c(perm_z_coeff)
        assert advice_coeff.size == (A + L) * n * 4 and perm_z_coeff.size == self.num_sets() * n * 4
        if L:
            lookup_coeff = _c(lookup_coeff)
            assert lookup_coeff.size == 3 * L * n * 4
        out = np.empty((4 * n, 4), dtype=np.uint64)
        self.ctx._check(lib().b200zk_evaluate_h(self.ctx._h, self._h, _p(advice_coeff), _p(perm_z_coeff), _p(lookup_coeff) if L else None,
                                                _p(_c(y)), _p(_c(beta)), _p(_c(gamma)), _p(out)))
        return out

    def create_proof(self, advice, rng_seed=0, timings=False, device_ptr=None):
        """plonk::create_proof with StdRng::seed_from_u64(rng_seed) and a Blake2b transcript; returns the proof bytes.
        `advice`: host array ((A+L) × 2^k × 4 uint64), or pass `device_ptr` for columns already resident in HBM."""
        k, A, L, F = self.shape
        buf = np.empty(self.proof_size(), dtype=np.uint8)
        plen = ctypes.c_size_t(0)
        tm = np.zeros(len(TIMING_KEYS), dtype=np.float64)
        if device_ptr is not None:
            self.ctx._check(lib().b200zk_create_proof_dev(self.ctx._h, self._h, ctypes.c_void_p(device_ptr), ctypes.c_uint64(rng_seed), _p(buf),
                                                          ctypes.byref(plen), _p(tm) if timings else None))
        else:
            if not (isinstance(advice, np.ndarray) and advice.dtype == np.uint64 and advice.flags["C_CONTIGUOUS"]):
                advice = _c(advice)
            assert advice.size == (A + L) * (1 << k) * 4
            self.ctx._check(lib().b200zk_create_proof(self.ctx._h, self._h, _p(advice), ctypes.c_uint64(rng_seed), _p(buf), ctypes.byref(plen),
                                                      _p(tm) if timings else None))
        proof = buf[: plen.value].tobytes()
        if timings:
            return proof, dict(zip(TIMING_KEYS, tm.tolist()))
        return proof


RNG_FILL_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint8), ctypes.c_size_t)


def create_proof_rng(pk, advice, fill):
    """b200zk_create_proof_rng: `fill(nbytes) -> bytes` plays the host generator's RngCore::fill_bytes."""
    k, A, L, F = pk.shape
    advice = _c(advice)
    assert advice.size == (A + L) * (1 << k) * 4

    def trampoline(user, out, nbytes):
        try:
            data = fill(int(nbytes))
            if len(data) != nbytes:
                return -1
            ctypes.memmove(out, data, nbytes)
            return 0
        except Exception:  # never unwind into C
            return -2

    cb = RNG_FILL_FN(trampoline)
    buf = np.empty(pk.proof_size(), dtype=np.uint8)
    plen = ctypes.c_size_t(0)
    pk.ctx._check(lib().b200zk_create_proof_rng(pk.ctx._h, pk._h, _p(advice), cb, None, _p(buf), ctypes.byref(plen), None))
    return buf[: plen.value].tobytes()


def synth_circuit(k, A, L, F, seed=0):
    """Synthetic circuit of the halo2-base shape: forwards to the workload generator (workload/synth.cpp), which is its own
    small host library — the product library carries no test-input code."""
    import sys

    root = os.path.dirname(_HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    import workload

    try:
        return workload.synth_circuit(k, A, L, F, seed)
    except ValueError as e:
        raise B200zkError(EINVAL, str(e))
