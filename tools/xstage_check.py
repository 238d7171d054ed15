"""Correctness (bit-exact vs oracle) and speed of the opt-in staged-x-window kernel."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import gsb200 as gsb
from gsb200 import synth
from oracle import linalg as ola
from util import host_to_scipy

ctx = gsb.Context()
ctx.set_option("xstage", "1")
hh = synth.poisson_hierarchy_host((48, 40, 36), 1)
n = hh.levels[0].n_own
As = host_to_scipy(hh.A[0], n)
A = gsb.SparseMatrix.from_scipy(As, ctx)
x = np.sin(np.arange(n, dtype=np.float64))
xd, yd = gsb.allocate_in_domain(A), gsb.allocate_in_range(A)
xd.set(x)
gsb.mul_(yd, A, xd)
yo = np.zeros(n)
ola.mul(yo, ola.CSR(As), x)
print("xstage spmv bit-exact:", bool(np.array_equal(yd.get(), yo)), flush=True)
c = 128
lp = synth.make_level_part((c,) * 3, (1, 1, 1), 0)
rp, col, val, b = synth.poisson_rows(lp)
A2 = gsb.SparseMatrix(ctx, lp.n_own, lp.n_own, 0, rp, col, val)
nnz = int(rp[-1])
out = {}
for mode, pr in (("spmv", 20), ("sweep", 44)):
    ms = A2.bench_rows(mode, 20)
    out[mode] = {"us": round(ms * 1e3, 1), "GBps": round((12 * nnz + pr * lp.n_own) / ms / 1e6)}
print(json.dumps(out), flush=True)
