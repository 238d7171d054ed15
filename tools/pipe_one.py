import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import gsb200 as gsb
from gsb200 import synth
ctx = gsb.Context()
c = 128
lp = synth.make_level_part((c,) * 3, (1, 1, 1), 0)
rp, col, val, b = synth.poisson_rows(lp)
A = gsb.SparseMatrix(ctx, lp.n_own, lp.n_own, 0, rp, col, val)
s = gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), 4, 2.0 / 3.0)
ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
x, r = gsb.allocate_in_domain(A), gsb.allocate_in_domain(A)
r.set(np.sin(np.arange(lp.n_own, dtype=np.float64)))
ctx.set_option("pipe_stages", int(sys.argv[1]) if len(sys.argv) > 1 else 2)
for _ in range(2):
    gsb.solve_(x, ns, r)
