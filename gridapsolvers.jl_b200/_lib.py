"""ctypes binding of libgsb200.so (the C ABI declared in include/gsb200.h).

There is no CPU fallback: if the library is missing this module raises, and every numerical
entry point fails with GSB_ECUDA when no CUDA device is present.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libgsb200.so")
SYNTH_PATH = os.path.join(_PKG, "lib", "libgsb200_synth.so")
CSRC = os.path.join(_PKG, "csrc")

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_i64 = ctypes.c_int64
c_d = ctypes.c_double
PP = ctypes.POINTER

# name -> argtypes ; every function returns int unless listed in _RESTYPES
SIGNATURES = {
    "gsb_version": [],
    "gsb_last_error": [c_p],
    "gsb_nccl_unique_id": [c_p],
    "gsb_init": [c_i, c_i, c_i, c_p, PP(c_p)],
    "gsb_finalize": [c_p],
    "gsb_synchronize": [c_p],
    "gsb_timer_start": [c_p],
    "gsb_timer_stop": [c_p, PP(ctypes.c_float)],
    "gsb_launch_count": [c_p, PP(c_i64)],
    "gsb_set_option": [c_p, ctypes.c_char_p, ctypes.c_char_p],
    "gsb_diag_sell_plan": [c_i64, c_i64, c_i64, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "gsb_bench_rows": [c_p, c_i, c_i, PP(ctypes.c_float)],
    "gsb_profile_start": [c_p],
    "gsb_profile_stop": [c_p, c_i, PP(c_i), c_p, c_p, c_p, c_p, c_p, c_p],
    "gsb_plan_create": [c_p, c_i64, c_i64, c_i, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_i, PP(c_p)],
    "gsb_plan_destroy": [c_p],
    "gsb_redist_create": [c_p, c_i64, c_i64, c_i, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_i, PP(c_p)],
    "gsb_vec_redistribute": [c_p, c_p, c_p],
    "gsb_mat_create": [c_p, c_i64, c_i64, c_i64, c_i, c_i, c_i, c_p, c_p, c_p, c_p, PP(c_p)],
    "gsb_mat_update_values": [c_p, c_p],
    "gsb_mat_info": [c_p, PP(c_i64), PP(c_i64), PP(c_i64), PP(c_i64)],
    "gsb_mat_format": [c_p, PP(c_i), PP(c_i), PP(c_i), PP(c_i64), PP(c_i64)],
    "gsb_mat_destroy": [c_p],
    "gsb_block_mat_create": [c_p, c_i, PP(c_p), PP(c_p)],
    "gsb_vec_create": [c_p, c_i64, c_i64, PP(c_p)],
    "gsb_vec_create_domain": [c_p, PP(c_p)],
    "gsb_vec_create_range": [c_p, PP(c_p)],
    "gsb_vec_destroy": [c_p],
    "gsb_vec_size": [c_p, PP(c_i64), PP(c_i64)],
    "gsb_vec_set": [c_p, c_p, c_i64],
    "gsb_vec_get": [c_p, c_p, c_i64],
    "gsb_vec_get_local": [c_p, c_p, c_i64],
    "gsb_vec_fill": [c_p, c_d],
    "gsb_vec_copy": [c_p, c_p],
    "gsb_vec_consistent": [c_p, c_p],
    "gsb_vec_assemble": [c_p, c_p],
    "gsb_host_register": [c_p, c_p, c_i64],
    "gsb_host_unregister": [c_p, c_p],
    "gsb_spmv": [c_p, c_p, c_p, c_d, c_d],
    "gsb_dot": [c_p, c_p, PP(c_d)],
    "gsb_norm2": [c_p, PP(c_d)],
    "gsb_axpby": [c_p, c_d, c_p, c_d, c_p],
    "gsb_identity_create": [c_p, PP(c_p)],
    "gsb_jacobi_create": [c_p, PP(c_p)],
    "gsb_richardson_create": [c_p, c_p, c_i, c_d, PP(c_p)],
    "gsb_from_smoother_create": [c_p, c_p, PP(c_p)],
    "gsb_dense_lu_create": [c_p, PP(c_p)],
    "gsb_gmg_create": [c_p, c_i, PP(c_p), PP(c_p), PP(c_p), PP(c_p), PP(c_p), c_p, c_i, c_i, c_i, c_d, c_d, PP(c_p)],
    "gsb_gmg_create_redist": [c_p, c_i, PP(c_p), PP(c_p), PP(c_p), PP(c_p), PP(c_p), c_p, c_i, c_i, c_i, c_d, c_d, PP(c_p), PP(c_p), PP(c_p)],
    "gsb_cg_create": [c_p, c_p, c_i, c_i, c_d, c_d, PP(c_p)],
    "gsb_gmres_create": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_d, c_d, PP(c_p)],
    "gsb_fgmres_create": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_d, c_d, PP(c_p)],
    "gsb_minres_create": [c_p, c_p, c_i, c_d, c_d, PP(c_p)],
    "gsb_block_solver_create": [c_p, c_i, PP(c_p), PP(c_p), c_p, c_i, c_i, PP(c_p)],
    "gsb_richardson_linear_create": [c_p, c_p, c_d, c_i, c_d, c_d, PP(c_p)],
    "gsb_schur_complement_create": [c_p, c_p, c_p, c_p, c_p, PP(c_p)],
    "gsb_cg_record_coefficients": [c_p, c_i],
    "gsb_cg_coefficients": [c_p, c_p, c_p, c_i64, PP(c_i64)],
    "gsb_solver_update": [c_p, c_p],
    "gsb_solve": [c_p, c_p, c_p],
    "gsb_solve_host": [c_p, c_p, c_p, c_i64],
    "gsb_solve_host_zero_guess": [c_p, c_p, c_p, c_i64],
    "gsb_solver_log": [c_p, PP(c_i), c_p, c_i64, PP(c_i)],
    "gsb_solver_destroy": [c_p],
}
_RESTYPES = {"gsb_last_error": ctypes.c_char_p}

_lib = None
_synth = None


class GSBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libgsb200 error {code}: {msg}")
        self.code = code


def build(verbose: bool = False) -> None:
    """Compile libgsb200.so and libgsb200_synth.so in-tree for sm_100a (nvcc cross-compiles)."""
    out = subprocess.run(["make", "-C", CSRC, "all"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
        print(out.stderr)
    if out.returncode != 0:
        raise RuntimeError("building libgsb200 failed")


def _preload_nccl():
    """Make sure libnccl.so.2 is resolvable: prefer the copy torch ships (already loaded when
    torch.distributed is in use), else the system one."""
    try:
        import torch  # noqa: F401  (loads its bundled NCCL into the process)
    except Exception:
        pass
    for cand in ("libnccl.so.2",):
        try:
            ctypes.CDLL(cand, mode=ctypes.RTLD_GLOBAL)
            return
        except OSError:
            continue
    try:
        import nvidia.nccl  # type: ignore

        p = os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so.2")
        ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
    except Exception:
        pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        _preload_nccl()
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_i)
        _lib = L
    return _lib


def check(code: int):
    if code != 0:
        msg = lib().gsb_last_error(None)
        raise GSBError(code, msg.decode() if msg else "?")


def synth():
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_PATH):
            raise RuntimeError(f"{SYNTH_PATH} is missing: run build()")
        S = ctypes.CDLL(SYNTH_PATH)
        rows = [c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p]
        S.synth_poisson_rows.argtypes = rows + [c_p]
        S.synth_mass_rows.argtypes = rows
        tr = [c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p]
        S.synth_prolong_rows.argtypes = tr
        S.synth_restrict_rows.argtypes = tr
        S.synth_fe_rows.argtypes = [c_i, c_i, c_i, c_p, c_p, c_p, c_p, c_i64, c_p, c_p, c_i, c_p, c_p, c_p, c_p]
        for f in (S.synth_poisson_rows, S.synth_mass_rows, S.synth_prolong_rows, S.synth_restrict_rows, S.synth_fe_rows):
            f.restype = None
        _synth = S
    return _synth
