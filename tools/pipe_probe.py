"""DRAM traffic / time of the pipelined smoother for option combinations (run under ncu for bytes)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gsb200 as gsb
from gsb200 import synth
ctx = gsb.Context()
c = int(sys.argv[1]) if len(sys.argv) > 1 else 128
lp = synth.make_level_part((c,) * 3, (1, 1, 1), 0)
rp, col, val, b = synth.poisson_rows(lp)
A = gsb.SparseMatrix(ctx, lp.n_own, lp.n_own, 0, rp, col, val)
nnz = int(rp[-1])
s = gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), 10, 2.0 / 3.0)
ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
x, r = gsb.allocate_in_domain(A), gsb.allocate_in_domain(A)
r.set(np.sin(np.arange(lp.n_own, dtype=np.float64)))
for cfg in sys.argv[2:]:
    for kv in cfg.split(","):
        k, v = kv.split("=")
        ctx.set_option(k, v)
    for _ in range(2):
        gsb.solve_(x, ns, r)
    ctx.timer_start()
    for _ in range(5):
        gsb.solve_(x, ns, r)
    ms = ctx.timer_stop() / 5
    print(cfg, round(ms, 4), "ms", round((10 * (12 * nnz + 44 * lp.n_own)) / ms / 1e6), "alg GB/s", flush=True)
