// ops.h -- launchers of the device primitives (defined in core.cu), used by the solver layer.
#pragma once
#include "gsb_internal.h"

namespace gsb {

// fill!(v,val) on the whole local vector (own + ghost)
void vec_fill(gsb_vec_s &v, double val);
// copy!(dst,src): own values
void vec_copy(gsb_vec_s &dst, const gsb_vec_s &src);
// z = ((a*x + b*y) + c*w) / d over own values; y/w/d optional (nullptr / has_d=false)
void ew_axpby(gsb_vec_s &z, ScalarRef a, const gsb_vec_s &x, ScalarRef b, const gsb_vec_s *y,
              ScalarRef c = imm(0.0), const gsb_vec_s *w = nullptr, bool has_d = false, ScalarRef d = imm(1.0));
// z = x ./ d
void ew_div(gsb_vec_s &z, const gsb_vec_s &x, ScalarRef d);
// z = dvec .* x   (dvec a raw device array of n_own doubles)
void ew_mul_raw(gsb_vec_s &z, const double *dvec, const gsb_vec_s &x);
// scal[slot] = dot(own(a), own(b)), allreduced over ranks
void dot(const gsb_vec_s &a, const gsb_vec_s &b, int slot);
// consistent!(v): owner -> ghost update through the plan (no-op for nranks == 1 / no ghosts)
void consistent(gsb_vec_s &v, gsb_plan_t plan);
// same exchange on the communication stream; returns immediately, `consistent_end` makes the compute
// stream wait for the ghosts
void consistent_begin(gsb_vec_s &v, gsb_plan_t plan);
void consistent_end(gsb_vec_s &v, gsb_plan_t plan);

// row kernels (halo exchange of the gathered vector included)
void spmv(gsb_mat_t A, gsb_vec_s &x, gsb_vec_s &y, double alpha, double beta);          // mul!(y,A,x,alpha,beta)
void resid(gsb_mat_t A, gsb_vec_s &x, const gsb_vec_s &b, gsb_vec_s &out);             // out = b - A x
void sweep(gsb_mat_t A, gsb_vec_s &dx_in, gsb_vec_s &r, const double *invd, double omega, gsb_vec_s &dx_out,
           gsb_vec_s &xacc);                                                            // fused Jacobi-Richardson sweep
void spmv_dot(gsb_mat_t A, gsb_vec_s &x, gsb_vec_s &y, const gsb_vec_s &dotv, int slot); // y = A x ; scal[slot] = dotv.y
void spmv_add(gsb_mat_t A, gsb_vec_s &x, gsb_vec_s &y, gsb_vec_s &xacc);                // y = A x ; xacc += y

// dx = omega*(invd.*r) ; x += dx  (x_is_zero: x = dx)
void jacobi_step(const double *invd, const gsb_vec_s &r, double omega, gsb_vec_s &dx, gsb_vec_s &x, bool x_is_zero);
// z = invd.*r ; scal[slot] = z.r
void jacobi_dot(const double *invd, const gsb_vec_s &r, gsb_vec_s &z, int slot);
// x += alpha p ; r -= alpha w ; scal[slot] = r.r
void cg_update(ScalarRef alpha, const gsb_vec_s &p, const gsb_vec_s &w, gsb_vec_s &x, gsb_vec_s &r, int slot);
// invd = 1 ./ diag(A_own_own)
void inv_diag(gsb_mat_t A, double *invd);
// matrix set-up helpers (device side of matrix.cu): diagonal + its CSR positions from the CSR arrays;
// refresh of the diagonal from CSR-ordered values; CSR -> block-SELL conversion
void csr_diag(gsb_mat_t A, const double *dval);
void refresh_diag(gsb_mat_t A, const double *dval);
void sell_fill(gsb_mat_t A, bool values_only, const double *dval);
// modified Gram-Schmidt step: w -= scal[slot_prev]*vprev (vprev may be null) ; scal[slot] = w.v (v null: w.w)
void mgs_step(gsb_vec_s &w, const gsb_vec_s *vprev, int slot_prev, const gsb_vec_s *v, int slot);
// x += sum_i g[i]*z[i] (applied vector after vector per element)
void multi_axpy(gsb_vec_s &x, const std::vector<const gsb_vec_s *> &z, const double *g);
// assemble!(v): ghost -> owner accumulation through the reversed plan, then ghosts zeroed
void assemble(gsb_vec_s &v, gsb_plan_t plan);
// redistribution of own values between two row partitions (plan created by gsb_redist_create)
void redistribute(gsb_plan_t plan, gsb_vec_s &src, gsb_vec_s &dst);

// dense coarse solver pieces
void dense_inverse_rows(gsb_mat_t A, DevBuf<double> &inv_rows, int64_t &n_global, int64_t &row_off);
void dense_apply(gsb_ctx_t ctx, const DevBuf<double> &inv_rows, int64_t n_global, int64_t row_off, int64_t n_own,
                 const gsb_vec_s &b, gsb_vec_s &x, DevBuf<double> &bfull);

// sub-vector view (block solvers)
gsb_vec_s view(gsb_vec_s &v, int64_t off, int64_t n);

}  // namespace gsb
