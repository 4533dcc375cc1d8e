import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, b200zk
k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = b200zk.Context(0)
n = 1 << k
rng = np.random.default_rng(0)
a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
a[:, 3] &= np.uint64((1 << 60) - 1)
src = ctx.dev_alloc(32 * n); dst = ctx.dev_alloc(128 * n)
ctx.h2d(src, a)
for _ in range(3):
    ctx.coeff_to_extended_dev(k, src, dst)
ctx.sync()
print("ok")
ctx.close()
