"""Pure-Python (big-int) generator of a tiny satisfying circuit of the halo2-base shape:
A vertical-gate advice columns q·(a + b·c − d), L lookup-advice columns copied from limb cells, one
[0, 2^(k-1)) table column, F constant columns, copy constraints between them. Small k only — it exists to
pin the C++ generator (`b200zk_synth_circuit`) and to feed the oracle prover in CPU tests independently of
any product code.
"""
import numpy as np

from oracle_lib import R_MOD, fr_array


def make_circuit(k, A, L, F, seed=0):
    rng = np.random.default_rng(seed)
    n = 1 << k
    usable = n - 9  # halo2-base leaves 9 unusable rows
    lookup_bits = k - 1
    NF = F + 1 + A
    fixed = [[0] * n for _ in range(NF)]
    advice = [[0] * n for _ in range(A + L)]
    copies = []
    # table
    for i in range(1 << lookup_bits):
        fixed[F][i] = i
    # constants: a handful per constant column
    consts = {}
    for f in range(F):
        for r in range(min(8, usable)):
            v = [0, 1, 2, (1 << lookup_bits), R_MOD - 1, 7, 1 << 64, 12345][r] % R_MOD
            fixed[f][r] = v
            consts.setdefault(v, (f, r))
    lookup_fill = [0] * L
    for c in range(A):
        row = 0
        prev_d = None
        while row + 4 <= usable:
            kind = int(rng.integers(0, 4))
            if prev_d is not None and kind == 0 and row >= 1:
                # chained gate: a of this gate is d of the previous one (overlap by one cell)
                row -= 1
                a = prev_d
            elif kind == 1:
                a = int(rng.integers(0, 1 << lookup_bits))  # limb
            elif kind == 2:
                a = int.from_bytes(rng.bytes(40), "little") % R_MOD  # full width
            else:
                a = int(rng.integers(0, 1 << 62))
            b = int(rng.integers(0, 1 << lookup_bits)) if kind != 2 else int.from_bytes(rng.bytes(40), "little") % R_MOD
            cc = [0, 1, int(rng.integers(0, 1 << 32)), int.from_bytes(rng.bytes(40), "little") % R_MOD][int(rng.integers(0, 4))]
            d = (a + b * cc) % R_MOD
            advice[c][row : row + 4] = [a, b, cc, d]
            fixed[F + 1 + c][row] = 1
            # copy limbs into a lookup column
            if L and b < (1 << lookup_bits):
                l = int(rng.integers(0, L))
                if lookup_fill[l] < usable:
                    advice[A + l][lookup_fill[l]] = b
                    copies.append((F + A + l, lookup_fill[l], F + c, row + 1))
                    lookup_fill[l] += 1
            # tie constants
            if cc in consts:
                f, r = consts[cc]
                copies.append((f, r, F + c, row + 2))
            prev_d = d
            row += 4
            if rng.integers(0, 8) == 0:
                row += int(rng.integers(0, 3))  # leave a gap (unconstrained zero cells)
                prev_d = None
    # a few advice-advice copies of equal values (zeros in gaps are equal)
    flat_fixed = np.concatenate([fr_array(col) for col in fixed]).reshape(NF, n, 4)
    flat_advice = np.concatenate([fr_array(col) for col in advice]).reshape(A + L, n, 4)
    return flat_fixed, flat_advice, np.array(copies, dtype=np.uint32).reshape(-1, 4)
