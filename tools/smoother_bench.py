"""Time one RichardsonSmoother(Jacobi,10,2/3) application (prologue + 10 sweeps) on the fine-level Poisson
matrix for several pipeline depths.  usage: python tools/smoother_bench.py [cells ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gsb200 as gsb
from gsb200 import synth

ctx = gsb.Context()
for c in [int(a) for a in sys.argv[1:] if a.isdigit()] or [128]:
    lp = synth.make_level_part((c,) * 3, (1, 1, 1), 0)
    rp, col, val, b = synth.poisson_rows(lp)
    A = gsb.SparseMatrix(ctx, lp.n_own, lp.n_own, 0, rp, col, val)
    nnz = int(rp[-1])
    s = gsb.RichardsonSmoother(gsb.JacobiLinearSolver(), 10, 2.0 / 3.0)
    ns = gsb.numerical_setup(gsb.symbolic_setup(s, A), A)
    x, r = gsb.allocate_in_domain(A), gsb.allocate_in_domain(A)
    r.set(np.sin(np.arange(lp.n_own, dtype=np.float64)))
    out = {"cells": c, "rows": lp.n_own}
    for stages in (1, 2, 3, 4, 5, 10):
        ctx.set_option("pipe_stages", stages)
        for _ in range(3):
            gsb.solve_(x, ns, r)
        ctx.timer_start()
        reps = 10
        for _ in range(reps):
            gsb.solve_(x, ns, r)
        ms = ctx.timer_stop() / reps
        bytes_alg = 10 * (12 * nnz + 44 * lp.n_own) + 40 * lp.n_own
        out["S%d" % stages] = {"ms": round(ms, 4), "alg_GBps": round(bytes_alg / ms / 1e6, 0)}
    print(json.dumps(out), flush=True)
