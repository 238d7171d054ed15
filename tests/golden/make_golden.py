"""Generates tests/golden/oracle_histories.json from the CPU ORACLE (not from the reference: the
reference is Julia and cannot run in this container -- see oracle/__init__.py "PARITY STATUS").
The fixture pins the oracle itself against regressions and gives the GPU tests a fixed target.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fem  # noqa: E402
from oracle import linalg as ola  # noqa: E402
from oracle import solvers as S  # noqa: E402

P = S.JacobiLinearSolver
CASES = {
    "gmres40_PrPl": lambda: S.GMRESSolver(40, Pr=P(), Pl=P(), rtol=1e-8),
    "gmres10": lambda: S.GMRESSolver(10, rtol=1e-8),
    "gmres10_restart": lambda: S.GMRESSolver(10, restart=True, rtol=1e-8),
    "fgmres10": lambda: S.FGMRESSolver(10, P(), rtol=1e-8),
    "fgmres10_restart": lambda: S.FGMRESSolver(10, P(), restart=True, rtol=1e-8),
    "cg": lambda: S.CGSolver(rtol=1e-8),
    "pcg": lambda: S.CGSolver(P(), rtol=1e-8),
    "fpcg": lambda: S.CGSolver(P(), flexible=True, rtol=1e-8),
    "minres": lambda: S.MINRESSolver(Pl=P(), rtol=1e-8),
    "cg_richardson": lambda: S.CGSolver(S.LinearSolverFromSmoother(S.RichardsonSmoother(P(), 5, 2.0 / 3.0)), rtol=1e-8),
}


def main():
    out = {"generator": "tests/golden/make_golden.py (oracle, numpy %s)" % np.__version__, "krylov": {}, "gmg": {}}
    for nc in [(8, 8), (8, 8, 8)]:
        sysm = fem.poisson(nc)
        A = ola.CSR(sysm.A)
        for name, mk in CASES.items():
            s = mk()
            ns = S.numerical_setup(S.symbolic_setup(s, A), A)
            x = S.allocate_in_domain(A)
            S.solve_(x, ns, sysm.b)
            out["krylov"]["%s/%s" % ("x".join(map(str, nc)), name)] = {
                "num_iters": s.log.num_iters, "flag": s.log.flag, "residuals": s.log.history().tolist(),
                "l2_error_sq": fem.l2_error_sq(sysm, x)}
    for nc, nlev, cyc in [((16, 16), 3, "v_cycle"), ((16, 16, 16), 3, "v_cycle"), ((16, 16, 16), 3, "w_cycle"),
                          ((16, 16, 16), 3, "f_cycle"), ((32, 32, 32), 4, "v_cycle")]:
        H = fem.poisson_hierarchy(nc, nlev)
        mats = [ola.CSR(m) for m in H.mats]
        sm = [S.RichardsonSmoother(P(), 10, 2.0 / 3.0)] * (nlev - 1)
        gmg = S.GMGLinearSolver(mats, [ola.CSR(p) for p in H.P], [ola.CSR(r) for r in H.R], pre_smoothers=sm,
                                post_smoothers=sm, maxiter=1, cycle_type=cyc)
        s = S.CGSolver(gmg, maxiter=20, atol=1e-14, rtol=1e-8) if cyc == "v_cycle" else S.FGMRESSolver(5, gmg, maxiter=20, atol=1e-14, rtol=1e-8)
        ns = S.numerical_setup(S.symbolic_setup(s, mats[0]), mats[0])
        x = S.allocate_in_domain(mats[0])
        S.solve_(x, ns, H.systems[0].b)
        out["gmg"]["%s/%d/%s" % ("x".join(map(str, nc)), nlev, cyc)] = {
            "num_iters": s.log.num_iters, "residuals": s.log.history().tolist(),
            "l2_error_sq": fem.l2_error_sq(H.systems[0], x)}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_histories.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["krylov"]), "krylov and", len(out["gmg"]), "gmg cases")


if __name__ == "__main__":
    main()
