import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch, b200zk
ctx=b200zk.Context(0); stream=torch.cuda.ExternalStream(ctx.stream)
k=int(sys.argv[1]) if len(sys.argv)>1 else 20; n=1<<k
src=ctx.dev_alloc(32*n); dst=ctx.dev_alloc(128*n)
a=np.random.default_rng(0).integers(0,1<<62,size=(n,4),dtype=np.uint64); ctx.h2d(src,a)
for name,fn in (("coset n->4n", lambda: ctx.coeff_to_extended_dev(k,src,dst)), ("l2c n", lambda: ctx.lagrange_to_coeff_dev(k,src))):
    for _ in range(3): fn()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): fn()
    e1.record(stream); ctx.sync(); print(name, "ms", round(e0.elapsed_time(e1)/10,4), flush=True)
ctx.close()
