import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, b200zk, oracle_lib as O
k = int(sys.argv[1]) if len(sys.argv) > 1 else 18
ctx = b200zk.Context(0)
params = O.Params.setup(k); s, g, gl = params.get(); ctx.srs_load(k, g, gl)
rng = np.random.default_rng(0)
a = O.random_fr(rng, 1 << k)
for _ in range(2):
    r = ctx.msm(a, 0)
print("ok", r[:2])
os._exit(0)
