"""gsb200 -- B200-native solve phase for GridapSolvers.jl (host-side mirror of the reference API).

The directory is named after the reference (`gridapsolvers.jl_b200`); because of the dot it is
imported through the root-level shim `gsb200.py` (`import gsb200`).
"""
from . import _lib
from ._lib import GSBError, build
from .api import *  # noqa: F401,F403
from .api import (Context, ExchangePlan, SparseMatrix, BlockSparseMatrix, Vector, allocate_in_domain, allocate_in_range,
                  mul_, dot, norm, axpby_, copy_, consistent_, assemble_, SolverTolerances, ConvergenceLog, symbolic_setup,
                  numerical_setup, numerical_setup_, solve_, ldiv_, IdentitySolver, JacobiLinearSolver, LUSolver,
                  RichardsonSmoother, LinearSolverFromSmoother, Fill, GMGLinearSolver, CGSolver, GMRESSolver,
                  FGMRESSolver, MINRESSolver, BlockTriangularSolver, BlockDiagonalSolver, LanczosDiagnostic,
                  RichardsonLinearSolver, SchurComplementSolver, HierarchicalArray, num_levels, with_level,
                  get_solver_tolerances, set_solver_tolerances_)
