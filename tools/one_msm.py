import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, b200zk
k = int(sys.argv[1]) if len(sys.argv) > 1 else 18
kind = sys.argv[2] if len(sys.argv) > 2 else "uniform"
ctx = b200zk.Context(0)
ctx.srs_setup(k)
rng = np.random.default_rng(0)
n = 1 << k
a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
a[:, 3] &= np.uint64((1 << 60) - 1)
if kind == "small":
    a[:, 1:] = 0
    a[:, 0] &= np.uint64((1 << 19) - 1)
    a = ctx.field_vec_op(0, 6, a)
for _ in range(2):
    r = ctx.msm(a, 0)
print("ok", r[:2])
ctx.close()
