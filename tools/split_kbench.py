"""Single-GPU diagnosis of the own/ghost split kernels: rank-0 part of a 2x1x1 partition."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsb200 as gsb
from gsb200 import synth

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = gsb.Context()
lp = synth.make_level_part((2 * cells, cells, cells), (2, 1, 1), 0, (2.0, 1.0, 1.0))
rp, col, val, b = synth.poisson_rows(lp)
A = gsb.SparseMatrix(ctx, lp.n_own, lp.n_own, lp.n_ghost, rp, col, val)
out = {}
for name, opts in [("plain", {"force_split": "0"}), ("split", {"force_split": "1", "overlap": "1"})]:
    for k, v in opts.items():
        ctx.set_option(k, v)
    out[name] = {m: round(A.bench_rows(m, 20) * 1e3, 1) for m in ("sweep", "residual", "spmv")}
print(json.dumps(out))
