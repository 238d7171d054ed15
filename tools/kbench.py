"""Row-kernel micro-benchmark: achieved algorithmic GB/s of every row-kernel mode on the fine-level
Poisson matrices (C2: 128^3, C3 unit: 256^3).  usage: python tools/kbench.py [cells ...] [--opt k=v ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gsb200 as gsb
from gsb200 import synth

PER_ROW = {"spmv": 20, "spmv_dot": 28, "residual": 28, "sweep": 44, "spmv_add": 36}


def main():
    cells = [int(a) for a in sys.argv[1:] if a.isdigit()] or [128]
    opts = [a.split("=", 1) for a in sys.argv[1:] if "=" in a and not a.startswith("--")]
    ctx = gsb.Context()
    for k, v in opts:
        ctx.set_option(k, v)
    peak = 6545.3
    for c in cells:
        lp = synth.make_level_part((c,) * 3, (1, 1, 1), 0)
        rp, col, val, b = synth.poisson_rows(lp)
        A = gsb.SparseMatrix(ctx, lp.n_own, lp.n_own, 0, rp, col, val)
        nnz = int(rp[-1])
        out = {"cells": c, "rows": lp.n_own, "nnz": nnz, "opts": dict(opts)}
        for mode in ("spmv", "residual", "sweep", "spmv_dot"):
            ms = A.bench_rows(mode, 30)
            gbs = (12 * nnz + PER_ROW[mode] * lp.n_own) / (ms * 1e-3) / 1e9
            out[mode] = {"us": round(ms * 1e3, 1), "GBps": round(gbs, 0), "frac": round(gbs / peak, 3)}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
