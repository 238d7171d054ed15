// solvers.cu -- NumericalSetup mirrors of the reference's LinearSolvers, running on the device
// primitives of core.cu.  Control flow follows the reference statement by statement (file:line
// cited at each step); Krylov coefficients stay in device scalars and the host reads back one
// number per iteration (the residual the ConvergenceLog needs for its stopping test).
#include <cmath>
#include <cstring>

#include "ops.h"

#include "api_macros.h"
using namespace gsb;

namespace {

using Vec = gsb_vec_s;
using VecP = std::unique_ptr<Vec>;

VecP make_vec(gsb_ctx_t ctx, int64_t n_own, int64_t n_ghost) {
  VecP v(new Vec());
  v->ctx = ctx; v->n_own = n_own; v->n_ghost = n_ghost;
  GSB_CUDA(cudaMalloc(&v->d, sizeof(double) * std::max<int64_t>(1, n_own + n_ghost)));
  GSB_CUDA(cudaMemsetAsync(v->d, 0, sizeof(double) * std::max<int64_t>(1, n_own + n_ghost), ctx->stream));
  return v;
}
// device scalar slots owned by a solver: returned to the context's free list when the solver is destroyed
struct Slots {
  gsb_ctx_t ctx = nullptr;
  int start = -1, n = 0;
  void take(gsb_ctx_t c, int count) { ctx = c; n = count; start = c->alloc_slots(count); }
  ~Slots() {
    if (ctx && n > 0 && gsb::ctx_alive(ctx)) ctx->free_slots(start, n);
  }
};

VecP domain_vec(gsb_mat_t A) { return make_vec(A->ctx, A->n_own_cols, A->n_ghost_cols); }  // allocate_in_domain
VecP range_vec(gsb_mat_t A) { return make_vec(A->ctx, A->n_rows, 0); }                     // allocate_in_range

void set_slot(gsb_ctx_t ctx, int slot, double v) {
  double *stage = ctx->h_scal + gsb_ctx_s::H_SCAL_READ;  // pinned staging word behind the read-back window
  *stage = v;
  GSB_CUDA(cudaMemcpyAsync(ctx->scal.p + slot, stage, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  GSB_CUDA(cudaStreamSynchronize(ctx->stream));
}

// LinearAlgebra.givensAlgorithm(f,g) for reals (Julia stdlib, port of LAPACK 3.x dlartg), normal range
void givens(double f, double g, double &cs, double &sn, double &r) {
  if (g == 0.0) { cs = 1.0; sn = 0.0; r = f; return; }
  if (f == 0.0) { cs = 0.0; sn = 1.0; r = g; return; }
  r = std::sqrt(f * f + g * g);
  cs = f / r; sn = g / r;
  if (std::fabs(f) > std::fabs(g) && cs < 0) { cs = -cs; sn = -sn; r = -r; }
}

// ---------------------------------------------------------------- IdentitySolver
struct IdentityNS : gsb_solver_s {
  const char *name() const override { return "Identity"; }
  bool capturable() const override { return true; }
  void solve(Vec &x, Vec &b) override { vec_copy(x, b); }  // IdentityLinearSolvers.jl:23-26
};

// ---------------------------------------------------------------- JacobiLinearSolver
struct JacobiNS : gsb_solver_s {
  gsb_mat_t A;
  DevBuf<double> invd;
  explicit JacobiNS(gsb_mat_t A_) : A(A_) {
    ctx = A->ctx;
    GSB_CHECK(A->nb == 0, "Jacobi: block matrices not supported (use a block-diagonal solver)");
    invd.alloc((size_t)std::max<int64_t>(1, A->n_rows));
    inv_diag(A, invd.p);  // JacobiLinearSolvers.jl:20-23,29-34
  }
  const char *name() const override { return "Jacobi"; }
  bool capturable() const override { return true; }
  void update(gsb_mat_t A_) override { A = A_; inv_diag(A, invd.p); }    // :25-27,36-41
  void solve(Vec &x, Vec &b) override { ew_mul_raw(x, invd.p, b); }      // :43-56 own values only
};

// ---------------------------------------------------------------- RichardsonSmoother
struct RichardsonNS : gsb_solver_s {
  gsb_mat_t A;
  gsb_solver_t M;
  int niter;
  double omega;
  VecP dx, Adx;  // RichardsonSmoothers.jl:58-63 ; the fused path uses Adx as the second dx buffer
  RichardsonNS(gsb_mat_t A_, gsb_solver_t M_, int niter_, double omega_) : A(A_), M(M_), niter(niter_), omega(omega_) {
    ctx = A->ctx;
    dx = domain_vec(A);
    Adx = domain_vec(A);
  }
  const char *name() const override { return "Richardson"; }
  bool capturable() const override { return M->capturable(); }
  void update(gsb_mat_t A_) override { M->update(A_); A = A_; }  // :72-76
  // solve!(x,ns,r): updates x AND r in place (:84-98)
  void solve(Vec &x, Vec &r) override { apply(x, r, false); }
  void apply(Vec &x, Vec &r, bool x_is_zero) {
    JacobiNS *J = dynamic_cast<JacobiNS *>(M);
    const bool fused = J && J->A == A && ctx->opt("fuse_smoother", "1") == "1";
    if (x_is_zero && !(fused && niter > 0)) { vec_fill(x, 0.0); x_is_zero = false; }
    if (fused) {
      if (niter <= 0) return;
      // iteration 1 prologue: dx = w*(invD*r) ; x += dx          (:91-93)
      jacobi_step(J->invd.p, r, omega, *dx, x, x_is_zero);
      for (int it = 1; it <= niter; ++it) {
        if (it < niter) {
          // r -= A dx  (:94-95) fused with the next iteration's dx = w*(invD*r) ; x += dx
          sweep(A, *dx, r, J->invd.p, omega, *Adx, x);
          std::swap(dx, Adx);
        } else {
          resid(A, *dx, r, r);
        }
      }
      return;
    }
    // literal sequence for any other inner solver
    vec_fill(*dx, 0.0);                                      // :89
    for (int it = 1; it <= niter; ++it) {
      M->solve(*dx, r);                                      // :91
      ew_axpby(*dx, imm(omega), *dx, imm(0.0), nullptr);     // :92  dx .= w .* dx
      ew_axpby(x, imm(1.0), x, imm(1.0), dx.get());          // :93  x .= x .+ dx
      spmv(A, *dx, *Adx, 1.0, 0.0);                          // :94
      ew_axpby(r, imm(1.0), r, imm(-1.0), Adx.get());        // :95  r .= r .- Adx
    }
  }
};

// ---------------------------------------------------------------- LinearSolverFromSmoother
struct FromSmootherNS : gsb_solver_s {
  gsb_solver_t smoother;
  VecP r;
  FromSmootherNS(gsb_mat_t A, gsb_solver_t s) : smoother(s) { ctx = A->ctx; r = domain_vec(A); }
  const char *name() const override { return "LinearSolverFromSmoother"; }
  bool capturable() const override { return smoother->capturable(); }
  void update(gsb_mat_t A) override { smoother->update(A); }
  void solve(Vec &x, Vec &b) override {  // LinearSolverFromSmoothers.jl:44-50
    vec_copy(*r, b);
    if (RichardsonNS *R = dynamic_cast<RichardsonNS *>(smoother)) {
      R->apply(x, *r, true);  // fill!(x,0) folded into the first x update
    } else {
      vec_fill(x, 0.0);
      smoother->solve(x, *r);
    }
  }
};

// ---------------------------------------------------------------- dense coarse solver (LUSolver stand-in)
struct DenseLUNS : gsb_solver_s {
  gsb_mat_t A;
  DevBuf<double> inv_rows, bfull;
  int64_t n_global = 0, row_off = 0;
  explicit DenseLUNS(gsb_mat_t A_) : A(A_) { ctx = A->ctx; update(A_); }
  const char *name() const override { return "DenseLU"; }
  bool capturable() const override { return true; }
  void update(gsb_mat_t A_) override {
    A = A_;
    dense_inverse_rows(A, inv_rows, n_global, row_off);
    bfull.alloc((size_t)std::max<int64_t>(1, n_global));
  }
  void solve(Vec &x, Vec &b) override { dense_apply(ctx, inv_rows, n_global, row_off, A->n_rows, b, x, bfull); }
};

// ---------------------------------------------------------------- GMG
struct GMGNS : gsb_solver_s {
  int nlev, mode, cycle_type;
  std::vector<gsb_mat_t> mats, interp, restrict_;
  std::vector<gsb_solver_t> pre, post;
  gsb_solver_t coarse;
  VecP rh;  // finest level cache, GMGLinearSolvers.jl:391-396
  struct Work { VecP dxh, Adxh, dxH, rH, tP, tR, dxH_red, rH_red; };
  std::vector<gsb_plan_t> to_coarse, to_fine;  // per level boundary: redistribution plans (nullptr: same parts)
  std::vector<Work> work;  // :451-466
  Slots slots;
  int slot_rr;
  bool children_capturable = true, graph_failed = false;
  const char *name() const override { return "GMG"; }
  gsb_mat_t matrix() override { return mats[0]; }
  GMGNS(gsb_ctx_t c, int nlev_, const gsb_mat_t *m, const gsb_mat_t *ip, const gsb_mat_t *rs, const gsb_solver_t *pr,
        const gsb_solver_t *po, gsb_solver_t cs, int mode_, int cyc, int maxiter, double atol, double rtol,
        const gsb_plan_t *to_coarse_ = nullptr, const gsb_plan_t *to_fine_ = nullptr)
      : nlev(nlev_), mode(mode_), cycle_type(cyc), coarse(cs) {
    ctx = c;
    GSB_CHECK(nlev >= 1, "GMG: need at least one level");
    GSB_CHECK(mode == GSB_GMG_PRECONDITIONER || mode == GSB_GMG_SOLVER, "GMG: bad mode");
    GSB_CHECK(cyc == GSB_V_CYCLE || cyc == GSB_W_CYCLE || cyc == GSB_F_CYCLE, "GMG: bad cycle type");
    mats.assign(m, m + nlev);
    interp.assign(ip, ip + nlev - 1);
    restrict_.assign(rs, rs + nlev - 1);
    pre.assign(pr, pr + nlev - 1);
    post.assign(po, po + nlev - 1);
    to_coarse.assign((size_t)std::max(nlev - 1, 0), nullptr);
    to_fine.assign((size_t)std::max(nlev - 1, 0), nullptr);
    for (int l = 0; l < nlev - 1; ++l) {
      if (to_coarse_) to_coarse[(size_t)l] = to_coarse_[l];
      if (to_fine_) to_fine[(size_t)l] = to_fine_[l];
      GSB_CHECK((to_coarse[(size_t)l] == nullptr) == (to_fine[(size_t)l] == nullptr), "GMG: a redistributed level needs both plans");
    }
    log.configure(maxiter, atol, rtol);
    has_log = true;
    slots.take(ctx, 2);
    slot_rr = slots.start;
    // the preconditioner application may be replayed from a CUDA graph only if no child reads scalars back
    // to the host or takes data-dependent host decisions (any LinearSolver is legal as smoother / coarsest
    // solver, GMGLinearSolvers.jl:48-58; an inner Krylov solver is not capturable)
    children_capturable = coarse->capturable();
    for (int l = 0; l < nlev - 1; ++l) children_capturable = children_capturable && pre[(size_t)l]->capturable() && post[(size_t)l]->capturable();
    rh = domain_vec(mats[0]);
    work.resize((size_t)nlev - 1);
    for (int l = 0; l < nlev - 1; ++l) {
      Work &w = work[(size_t)l];
      const bool red = to_coarse[(size_t)l] != nullptr;
      GSB_CHECK(interp[(size_t)l]->n_rows == mats[(size_t)l]->n_rows && restrict_[(size_t)l]->n_own_cols == mats[(size_t)l]->n_rows,
                "GMG: transfer operator shape mismatch (fine side) at level " + std::to_string(l + 1));
      if (!red) {
        GSB_CHECK(interp[(size_t)l]->n_own_cols == mats[(size_t)l + 1]->n_rows,
                  "GMG: prolongation shape mismatch at level " + std::to_string(l + 1));
        GSB_CHECK(restrict_[(size_t)l]->n_rows == mats[(size_t)l + 1]->n_rows,
                  "GMG: restriction shape mismatch at level " + std::to_string(l + 1));
      } else {
        // level l+2 lives on another (smaller) set of parts: P / R act on the coarse space in the partition of
        // level l+1's parts, the plans move own values between the two layouts (GridTransferOperators.jl:391-401,
        // 536-561 with Val{true}: redistribute_free_values! before the interpolation / after the restriction)
        gsb_plan_t tc = to_coarse[(size_t)l], tf = to_fine[(size_t)l];
        GSB_CHECK(tc->redist && tf->redist, "GMG: redistribution plans expected");
        GSB_CHECK(tc->n_own == restrict_[(size_t)l]->n_rows && tc->n_ghost == mats[(size_t)l + 1]->n_rows,
                  "GMG: to-coarse redistribution plan does not match level " + std::to_string(l + 2));
        GSB_CHECK(tf->n_own == mats[(size_t)l + 1]->n_rows && tf->n_ghost == interp[(size_t)l]->n_own_cols,
                  "GMG: to-fine redistribution plan does not match level " + std::to_string(l + 2));
        w.rH_red = range_vec(restrict_[(size_t)l]);
        w.dxH_red = domain_vec(interp[(size_t)l]);
      }
      w.dxh = domain_vec(mats[(size_t)l]);
      w.Adxh = range_vec(mats[(size_t)l]);
      w.dxH = domain_vec(mats[(size_t)l + 1]);
      w.rH = domain_vec(mats[(size_t)l + 1]);
      // transfer operators may use a different ghost layout than the level matrices ("FE layout"
      // vs "matrix layout", GridTransferOperators.jl:395-398): stage through a copy of own values
      if (!red && (interp[(size_t)l]->plan != mats[(size_t)l + 1]->plan || interp[(size_t)l]->n_ghost_cols != mats[(size_t)l + 1]->n_ghost_cols))
        w.tP = domain_vec(interp[(size_t)l]);
      if (restrict_[(size_t)l]->plan != mats[(size_t)l]->plan || restrict_[(size_t)l]->n_ghost_cols != mats[(size_t)l]->n_ghost_cols)
        w.tR = domain_vec(restrict_[(size_t)l]);
    }
  }
  void update(gsb_mat_t) override {  // GMGLinearSolvers.jl:249-258 (@error, not a throw)
    fail(GSB_EUNSUPPORTED, "GMGLinearSolverFromMatrices does not support updates");
  }
  void correct(int lev, Vec &xh, Vec &rh_) {
    Work &w = work[(size_t)lev];
    gsb_mat_t P = interp[(size_t)lev], Ah = mats[(size_t)lev];
    Vec *src = w.dxH.get();
    if (to_fine[(size_t)lev]) { redistribute(to_fine[(size_t)lev], *w.dxH, *w.dxH_red); src = w.dxH_red.get(); }
    if (w.tP) { vec_copy(*w.tP, *w.dxH); src = w.tP.get(); }
    spmv_add(P, *src, *w.dxh, xh);     // :491,494  dxh = P dxH ; xh .= xh .+ dxh
    resid(Ah, *w.dxh, rh_, rh_);       // :495-496  rh .= rh .- Ah dxh
  }
  void restrict_residual(int lev, Vec &rh_) {
    Work &w = work[(size_t)lev];
    Vec *src = &rh_;
    if (w.tR) { vec_copy(*w.tR, rh_); src = w.tR.get(); }
    if (to_coarse[(size_t)lev]) {
      spmv(restrict_[(size_t)lev], *src, *w.rH_red, 1.0, 0.0);  // :484 then redistribute_free_values!
      redistribute(to_coarse[(size_t)lev], *w.rH_red, *w.rH);
    } else {
      spmv(restrict_[(size_t)lev], *src, *w.rH, 1.0, 0.0);  // :484
    }
    vec_fill(*w.dxH, 0.0);                                // :487
  }
  void cycle(int kind, int lev, Vec &xh, Vec &rh_) {  // gmg_v_cycle! :468-502, w :504-556, f :558-610
    if (lev == nlev - 1) { coarse->solve(xh, rh_); return; }  // :472-474
    Work &w = work[(size_t)lev];
    pre[(size_t)lev]->solve(xh, rh_);             // :481
    restrict_residual(lev, rh_);
    coarser(kind, lev);                           // :488 / :524 / :578
    correct(lev, xh, rh_);
    if (kind != GSB_V_CYCLE) {
      post[(size_t)lev]->solve(xh, rh_);          // re-smooth :533 / :587
      restrict_residual(lev, rh_);
      coarser(kind == GSB_W_CYCLE ? GSB_W_CYCLE : GSB_V_CYCLE, lev);  // :540 / :594
      correct(lev, xh, rh_);
    }
    post[(size_t)lev]->solve(xh, rh_);            // :499
  }
  // the recursive call of a cycle on level lev+1.  Below the finest level a cycle always works on the SAME vectors
  // (the level's dxH / rH work vectors), whatever pair (x, b) the caller of the solver passed: the sub-cycle from
  // the second level down -- a few hundred tiny launches -- is captured once per cycle kind into a CUDA graph and
  // replayed, in every mode (mode=:solver iterations, FGMRES's changing vector pairs, ...).
  cudaGraphExec_t sub_exec[3] = {nullptr, nullptr, nullptr};
  int64_t sub_launches[3] = {0, 0, 0};
  bool sub_failed = false, in_capture = false;
  bool graphs_allowed() const {
    return !ctx->profiling && ctx->opt("graph", "1") == "1" && children_capturable &&
           (ctx->nranks == 1 || (!ctx->nccl_halo_in_use && ctx->opt("overlap", "0") != "1"));
  }
  void coarser(int kind, int lev) {
    Work &w = work[(size_t)lev];
    const int gi = kind == GSB_V_CYCLE ? 0 : (kind == GSB_W_CYCLE ? 1 : 2);
    if (lev != 0 || nlev <= 2 || in_capture || sub_failed || !graphs_allowed() || ctx->opt("gmg_subgraph", "1") != "1") {
      cycle(kind, lev + 1, *w.dxH, *w.rH);
      return;
    }
    if (!sub_exec[gi]) {
      const int64_t l0 = ctx->launches;
      cudaGraph_t g = nullptr;
      bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
      if (ok) {
        in_capture = true;
        try {
          cycle(kind, 1, *w.dxH, *w.rH);
        } catch (...) {
          ok = false;
        }
        in_capture = false;
        if (cudaStreamEndCapture(ctx->stream, &g) != cudaSuccess || g == nullptr) ok = false;
        if (ok && cudaGraphInstantiate(&sub_exec[gi], g, 0) != cudaSuccess) { ok = false; sub_exec[gi] = nullptr; }
        if (g) cudaGraphDestroy(g);
      }
      sub_launches[gi] = ctx->launches - l0;
      ctx->launches = l0;
      if (!ok) {
        (void)cudaGetLastError();
        sub_failed = true;
        cycle(kind, 1, *w.dxH, *w.rH);
        return;
      }
    }
    GSB_CUDA(cudaGraphLaunch(sub_exec[gi], ctx->stream));
    ctx->launches += sub_launches[gi];
  }
  double norm_rh() {
    dot(*rh, *rh, slot_rr);
    return std::sqrt(ctx->read_scalar(slot_rr));
  }
  void solve(Vec &x, Vec &b) override {  // :612-645
    if (mode == GSB_GMG_PRECONDITIONER) {
      vec_fill(x, 0.0);
      vec_copy(*rh, b);
    } else {
      resid(mats[0], x, b, *rh);
    }
    pending = false;
    // Preconditioner fast path (maxiter == 1): init!(log,res0) can only stop the solve when
    // res0 < atol (or 1.0 < rtol); when the caller has just told us ||b|| on the host we know the
    // decision without a synchronisation, run the single cycle, and leave both norms of the
    // reference's log (GMGLinearSolvers.jl:627,639) in device slots until somebody asks for them.
    const bool have_hint = ctx->hint_vec == b.d && ctx->hint_vec != nullptr;
    ctx->hint_vec = nullptr;
    if (log.maxiter == 1 && have_hint && !(1.0 < log.rtol) && ctx->hint_norm > 4.0 * log.atol &&
        std::isfinite(ctx->hint_norm) && ctx->opt("gmg_defer_log", "1") == "1") {
      // single rank: the whole application (norm, V-cycle, norm: ~100 launches, most of them tiny
      // coarse-level kernels) is captured once into a CUDA graph and replayed
      // (multi-rank: NCCL collectives are capturable and the peer-memory halo exchange keeps its
      //  sequence number on the device, so the replay is valid there too; the two-stream overlap is not)
      const bool use_graph = graphs_allowed() && !graph_failed;
      if (use_graph && graph_exec && graph_x == x.d && graph_b == b.d) {
        GSB_CUDA(cudaGraphLaunch(graph_exec, ctx->stream));
        ctx->launches += graph_launches;
      } else if (use_graph && last_x == x.d && last_b == b.d) {
        // second consecutive application on the same vector pair (the PCG pattern): worth capturing.
        // Callers that pass a different pair every time (FGMRES: Z[j], V[j]) never pay for a capture.
        if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
        const int64_t l0 = ctx->launches;
        cudaGraph_t g = nullptr;
        bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
          in_capture = true;  // the sub-cycle graph is not nested into this capture: everything is recorded flat
          try {
            dot(*rh, *rh, slot_rr);
            cycle(cycle_type, 0, x, *rh);
            dot(*rh, *rh, slot_rr + 1);
          } catch (...) {
            ok = false;
          }
          in_capture = false;
          if (cudaStreamEndCapture(ctx->stream, &g) != cudaSuccess || g == nullptr) ok = false;
          if (ok && cudaGraphInstantiate(&graph_exec, g, 0) != cudaSuccess) { ok = false; graph_exec = nullptr; }
          if (g) cudaGraphDestroy(g);
        }
        if (ok) {
          graph_launches = ctx->launches - l0;
          graph_x = x.d; graph_b = b.d;
          GSB_CUDA(cudaGraphLaunch(graph_exec, ctx->stream));
        } else {
          // something in the cycle could not be captured: clear the error state, never try again, run eagerly
          (void)cudaGetLastError();
          ctx->launches = l0;
          graph_failed = true;
          dot(*rh, *rh, slot_rr);
          cycle(cycle_type, 0, x, *rh);
          dot(*rh, *rh, slot_rr + 1);
        }
      } else {
        dot(*rh, *rh, slot_rr);
        cycle(cycle_type, 0, x, *rh);
        dot(*rh, *rh, slot_rr + 1);
      }
      last_x = x.d; last_b = b.d;
      pending = true;
      return;
    }
    double res = norm_rh();
    bool done = log.init(res);
    while (!done) {
      cycle(cycle_type, 0, x, *rh);
      res = norm_rh();
      done = log.update(res);
    }
    log.finalize(res);
  }
  bool pending = false;
  cudaGraphExec_t graph_exec = nullptr;
  const double *graph_x = nullptr, *graph_b = nullptr, *last_x = nullptr, *last_b = nullptr;
  int64_t graph_launches = 0;
  ~GMGNS() override {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    for (cudaGraphExec_t e : sub_exec)
      if (e) cudaGraphExecDestroy(e);
  }
  void finish() override {
    if (!pending) return;
    double v[2];
    ctx->read_scalars(slot_rr, 2, v);
    log.init(std::sqrt(v[0]));
    const double res = std::sqrt(v[1]);
    log.update(res);
    log.finalize(res);
    pending = false;
  }
};

// ---------------------------------------------------------------- CG
struct CGNS : gsb_solver_s {
  gsb_mat_t A;
  gsb_solver_t Pl;
  bool flexible;
  VecP w, p, z, r;  // CGSolvers.jl:42-48
  Slots slots;
  int s_g0, s_pw, s_rr, s_delta;  // five consecutive slots: gamma[2], p.w, r.r, delta
  bool record = false;            // LanczosDiagnostic support: keep alpha_k, beta_k
  std::vector<double> alphas, betas;
  const char *name() const override { return "CG"; }
  gsb_mat_t matrix() override { return A; }
  CGNS(gsb_mat_t A_, gsb_solver_t Pl_, bool flex, int maxiter, double atol, double rtol) : A(A_), Pl(Pl_), flexible(flex) {
    ctx = A->ctx;
    log.configure(maxiter, atol, rtol);
    has_log = true;
    w = domain_vec(A); p = domain_vec(A); z = domain_vec(A); r = domain_vec(A);
    slots.take(ctx, 5);
    s_g0 = slots.start; s_pw = s_g0 + 2; s_rr = s_g0 + 3; s_delta = s_g0 + 4;
  }
  void update(gsb_mat_t A_) override {  // :57-63
    if (Pl) Pl->update(A_);
    A = A_;
  }
  void solve(Vec &x, Vec &b) override {  // :73-120
    resid(A, x, b, *r);                  // :79  mul!(w,A,x); r .= b .- w
    vec_fill(*p, 0.0);                   // :80
    vec_fill(*z, 0.0);                   // :81
    int cur = 0;
    alphas.clear(); betas.clear();
    set_slot(ctx, s_g0 + 1, 1.0);        // gamma = 1   :82
    dot(*r, *r, s_rr);
    double res = std::sqrt(ctx->read_scalar(s_rr));  // :85
    bool done = log.init(res);                       // :86
    JacobiNS *J = dynamic_cast<JacobiNS *>(Pl);
    while (!done) {
      const int gcur = s_g0 + cur, gprev = s_g0 + (1 - cur);
      ScalarRef beta;
      Vec *zz = z.get();
      if (!Pl) {  // :90-92   z .= r ; gamma = dot(r,r)
        zz = r.get();
        dot(*r, *r, gcur);
        beta = slot_ratio(gcur, gprev);
      } else if (!flexible) {  // :93-95
        if (J) jacobi_dot(J->invd.p, *r, *z, gcur);
        else { ctx->hint_vec = r->d; ctx->hint_norm = res; Pl->solve(*z, *r); ctx->hint_vec = nullptr; dot(*z, *r, gcur); }
        beta = slot_ratio(gcur, gprev);
      } else {  // :96-99
        dot(*z, *r, s_delta);
        ctx->hint_vec = r->d; ctx->hint_norm = res;
        Pl->solve(*z, *r);
        ctx->hint_vec = nullptr;
        dot(*z, *r, gcur);
        beta = slot_ratio(gcur, gprev, 0, s_delta);
      }
      ew_axpby(*p, imm(1.0), *zz, beta, p.get());                 // :101  p .= z .+ beta .* p
      spmv_dot(A, *p, *w, *p, s_pw);                              // :104-105
      cg_update(slot_ratio(gcur, s_pw), *p, *w, x, *r, s_rr);     // :105-109 alpha = gamma/dot(p,w)
      if (record) {  // same IEEE operations the kernels perform on the device slots
        double v[5];
        ctx->read_scalars(s_g0, 5, v);
        const double gc = v[cur], gp = v[1 - cur];
        alphas.push_back(gc / v[2]);
        betas.push_back((flexible && Pl) ? (gc - v[4]) / gp : gc / gp);
        res = std::sqrt(v[3]);
      } else {
        res = std::sqrt(ctx->read_scalar(s_rr));                  // :111
      }
      done = log.update(res);                                     // :112
      cur ^= 1;
    }
    log.finalize(res);  // :118
  }
};

// ---------------------------------------------------------------- GMRES / FGMRES
struct GMRESNS : gsb_solver_s {
  gsb_mat_t A;
  gsb_solver_t Pr, Pl;
  int m0, m_add;
  bool restart, flex;
  std::vector<VecP> V, Z;  // GMRESSolvers.jl:57-69 ; FGMRESSolvers.jl:58-70
  VecP zr, zl;
  std::vector<double> H, g, c, s;  // H column-major (ld = m+1)
  int ldH = 0;
  Slots slots;
  int s_h, s_h_cap;
  const char *name() const override { return flex ? "FGMRES" : "GMRES"; }
  gsb_mat_t matrix() override { return A; }
  GMRESNS(gsb_mat_t A_, gsb_solver_t Pr_, gsb_solver_t Pl_, int m, bool restart_, int m_add_, bool flex_, int maxiter,
          double atol, double rtol)
      : A(A_), Pr(Pr_), Pl(Pl_), m0(m), m_add(m_add_), restart(restart_), flex(flex_) {
    ctx = A->ctx;
    GSB_CHECK(m >= 1 && m_add >= 1, "GMRES: m and m_add must be >= 1");
    GSB_CHECK(!flex || Pr, "FGMRES needs a right preconditioner");
    log.configure(maxiter, atol, rtol);
    has_log = true;
    for (int i = 0; i < m + 1; ++i) V.push_back(domain_vec(A));
    if (flex) for (int i = 0; i < m; ++i) Z.push_back(domain_vec(A));
    if (Pr && !flex) zr = domain_vec(A);
    zl = domain_vec(A);
    resize_host(m);
    s_h_cap = std::max(m, maxiter) + 4;
    slots.take(ctx, s_h_cap);
    s_h = slots.start;
  }
  int mcur() const { return (int)V.size() - 1; }
  void resize_host(int m) {
    std::vector<double> Hn((size_t)(m + 1) * m, 0.0);
    const int mo = ldH ? ldH - 1 : 0;
    for (int j = 0; j < mo; ++j)
      for (int i = 0; i < mo + 1; ++i) Hn[(size_t)j * (m + 1) + i] = H[(size_t)j * ldH + i];
    H.swap(Hn); ldH = m + 1;
    g.resize((size_t)m + 1, 0.0); c.resize((size_t)m, 0.0); s.resize((size_t)m, 0.0);
  }
  double &Hij(int i, int j) { return H[(size_t)j * ldH + i]; }
  void expand() {  // GMRESSolvers.jl:76-92
    for (int k = 0; k < m_add; ++k) {
      V.push_back(domain_vec(A));
      if (flex) Z.push_back(domain_vec(A));
    }
    resize_host(mcur());
    GSB_CHECK(mcur() + 2 <= s_h_cap, "GMRES: Krylov basis outgrew the scalar slots");
  }
  void update(gsb_mat_t A_) override {
    if (Pr) Pr->update(A_);
    if (Pl) Pl->update(A_);
    A = A_;
  }
  void apply_Pr(Vec &out, Vec &x) {
    ctx->hint_vec = x.d; ctx->hint_norm = 1.0;  // x = V[j], normalised just before
    Pr->solve(out, x);
    ctx->hint_vec = nullptr;  // only the solve it was meant for may consume the hint
  }
  void krylov_mul(Vec &y, Vec &x, Vec *wr) {  // KrylovUtils.jl:17-32
    if (Pr && Pl) { apply_Pr(*wr, x); spmv(A, *wr, *zl, 1.0, 0.0); Pl->solve(y, *zl); }
    else if (Pr) { apply_Pr(*wr, x); spmv(A, *wr, y, 1.0, 0.0); }
    else if (Pl) { spmv(A, x, *zl, 1.0, 0.0); Pl->solve(y, *zl); }
    else spmv(A, x, y, 1.0, 0.0);
  }
  void krylov_residual(Vec &r, Vec &x, Vec &b) {  // KrylovUtils.jl:46-54
    if (Pl) { resid(A, x, b, *zl); Pl->solve(r, *zl); }
    else resid(A, x, b, r);
  }
  double norm_of(Vec &v, int slot) {
    dot(v, v, slot);
    return std::sqrt(ctx->read_scalar(slot));
  }
  void solve(Vec &x, Vec &b) override {  // GMRESSolvers.jl:132-210 ; FGMRESSolvers.jl:130-199
    vec_fill(*V[0], 0.0);
    if (zr) vec_fill(*zr, 0.0);
    vec_fill(*zl, 0.0);
    krylov_residual(*V[0], x, b);
    double beta = norm_of(*V[0], s_h);
    bool done = log.init(beta);
    while (!done) {
      int j = 1;
      ew_div(*V[0], *V[0], imm(beta));  // V[1] ./= beta
      std::fill(H.begin(), H.end(), 0.0);
      std::fill(g.begin(), g.end(), 0.0);
      g[0] = beta;
      while (!done && !(restart && j > m0)) {
        if (j > mcur()) expand();
        Vec &Vn = *V[(size_t)j];
        vec_fill(Vn, 0.0);
        if (flex) {
          vec_fill(*Z[(size_t)j - 1], 0.0);  // FGMRESSolvers.jl:158
          krylov_mul(Vn, *V[(size_t)j - 1], Z[(size_t)j - 1].get());
        } else {
          krylov_mul(Vn, *V[(size_t)j - 1], zr.get());  // zr not re-zeroed, GMRESSolvers.jl:161
        }
        // modified Gram-Schmidt (:162-165), one launch per step: the axpy of step i-1 is fused with the dot
        // of step i (same order, same roundings as the separate broadcasts); the coefficients stay on the
        // device until the column is complete, the closing step also yields the norm of :166
        if (ctx->opt("fuse_mgs", "1") == "1") {
          for (int i = 0; i <= j; ++i)
            mgs_step(Vn, i > 0 ? V[(size_t)i - 1].get() : nullptr, s_h + i - 1, i < j ? V[(size_t)i].get() : nullptr, s_h + i);
        } else {
          for (int i = 0; i < j; ++i) {
            dot(Vn, *V[(size_t)i], s_h + i);
            ew_axpby(Vn, imm(1.0), Vn, slot_ratio(s_h + i, -1, 1), V[(size_t)i].get());
          }
          dot(Vn, Vn, s_h + j);
        }
        std::vector<double> hcol((size_t)j + 1);
        ctx->read_scalars(s_h, j + 1, hcol.data());
        for (int i = 0; i < j; ++i) Hij(i, j - 1) = hcol[(size_t)i];
        Hij(j, j - 1) = std::sqrt(hcol[(size_t)j]);    // :166
        ew_div(Vn, Vn, imm(Hij(j, j - 1)));            // :167
        for (int i = 0; i < j - 1; ++i) {              // :170-174
          const double gam = c[(size_t)i] * Hij(i, j - 1) + s[(size_t)i] * Hij(i + 1, j - 1);
          Hij(i + 1, j - 1) = -s[(size_t)i] * Hij(i, j - 1) + c[(size_t)i] * Hij(i + 1, j - 1);
          Hij(i, j - 1) = gam;
        }
        double rr;
        givens(Hij(j - 1, j - 1), Hij(j, j - 1), c[(size_t)j - 1], s[(size_t)j - 1], rr);  // :177
        Hij(j - 1, j - 1) = c[(size_t)j - 1] * Hij(j - 1, j - 1) + s[(size_t)j - 1] * Hij(j, j - 1);
        Hij(j, j - 1) = 0.0;
        g[(size_t)j] = -s[(size_t)j - 1] * g[(size_t)j - 1];
        g[(size_t)j - 1] = c[(size_t)j - 1] * g[(size_t)j - 1];
        beta = std::fabs(g[(size_t)j]);
        j += 1;
        done = log.update(beta);
      }
      j -= 1;
      for (int i = j - 1; i >= 0; --i) {  // :188-190 back substitution
        double acc = 0.0;
        for (int k = i + 1; k < j; ++k) acc += Hij(i, k) * g[(size_t)k];
        g[(size_t)i] = (g[(size_t)i] - acc) / Hij(i, i);
      }
      // solution update: x .+= g[i] .* Z[i] (FGMRES :191-193) / V[i] (:193-196), one pass over x for all i
      auto basis = [&](std::vector<VecP> &B) {
        std::vector<const Vec *> z;
        for (int i = 0; i < j; ++i) z.push_back(B[(size_t)i].get());
        return z;
      };
      if (flex) {
        multi_axpy(x, basis(Z), g.data());
      } else if (!Pr) {
        multi_axpy(x, basis(V), g.data());
      } else {
        vec_fill(*zl, 0.0);
        multi_axpy(*zl, basis(V), g.data());
        Pr->solve(*zr, *zl);
        ew_axpby(x, imm(1.0), x, imm(1.0), zr.get());
      }
      krylov_residual(*V[0], x, b);  // :205
    }
    log.finalize(beta);
  }
};

// ---------------------------------------------------------------- MINRES
struct MINRESNS : gsb_solver_s {
  gsb_mat_t A;
  gsb_solver_t Pl;
  VecP Vs[3], Ws[3], Zs[3];  // MINRESSolvers.jl:39-44
  Slots slots;
  int s0;
  const char *name() const override { return "MINRES"; }
  gsb_mat_t matrix() override { return A; }
  MINRESNS(gsb_mat_t A_, gsb_solver_t Pl_, int maxiter, double atol, double rtol) : A(A_), Pl(Pl_) {
    ctx = A->ctx;
    log.configure(maxiter, atol, rtol);
    has_log = true;
    for (int i = 0; i < 3; ++i) { Vs[i] = domain_vec(A); Ws[i] = domain_vec(A); Zs[i] = domain_vec(A); }
    slots.take(ctx, 4);
    s0 = slots.start;
  }
  void update(gsb_mat_t A_) override {
    if (Pl) Pl->update(A_);
    A = A_;
  }
  void solve(Vec &x, Vec &b) override {  // :75-149
    Vec *Vnew = Vs[0].get(), *V = Vs[1].get(), *Vold = Vs[2].get();
    Vec *Wnew = Ws[0].get(), *W = Ws[1].get(), *Wold = Ws[2].get();
    Vec *Znew = Zs[0].get(), *Z = Zs[1].get(), *Zold = Zs[2].get();
    vec_fill(*W, 0.0); vec_fill(*Wold, 0.0); vec_fill(*Vold, 0.0); vec_fill(*Zold, 0.0);
    resid(A, x, b, *Vnew);  // :90-91
    vec_fill(*Znew, 0.0);
    if (Pl) Pl->solve(*Znew, *Vnew); else vec_copy(*Znew, *Vnew);
    dot(*Znew, *Znew, s0);
    dot(*Znew, *Vnew, s0 + 1);
    double two[2];
    ctx->read_scalars(s0, 2, two);
    double beta_r = std::sqrt(two[0]);
    double beta_p = two[1];
    GSB_CHECK(beta_p > 0.0, "MINRES: preconditioner is not positive definite (beta_p <= 0)");  // :97
    double gnew = 0.0, gam = std::sqrt(beta_p), gold = 1.0;
    double cnew = 0.0, c = 1.0, cold = 1.0;
    double snew = 0.0, s = 0.0, sold = 0.0;
    ew_div(*V, *Vnew, imm(gam));
    ew_div(*Z, *Znew, imm(gam));
    double eta = gam;
    bool done = log.init(beta_r);
    while (!done) {
      spmv(A, *Z, *Vnew, 1.0, 0.0);                                        // :110
      if (Pl) Pl->solve(*Znew, *Vnew); else vec_copy(*Znew, *Vnew);       // :111
      dot(*Vnew, *Z, s0);                                                  // :112 delta
      const ScalarRef mdelta = slot_ratio(s0, -1, 1);
      ew_axpby(*Vnew, imm(1.0), *Vnew, mdelta, V, imm(-gam), Vold);        // :113
      ew_axpby(*Znew, imm(1.0), *Znew, mdelta, Z, imm(-gam), Zold);        // :114
      dot(*Znew, *Vnew, s0 + 1);                                           // :115
      ctx->read_scalars(s0, 2, two);
      const double delta = two[0];
      beta_p = two[1];
      gnew = std::sqrt(beta_p);                                            // :116
      ew_div(*Vnew, *Vnew, imm(gnew));                                     // :118
      ew_div(*Znew, *Znew, imm(gnew));                                     // :119
      const double a0 = c * delta - cold * s * gam;                        // :122
      double a1;
      givens(a0, gnew, cnew, snew, a1);                                    // :123
      const double a2 = s * delta + cold * c * gam;
      const double a3 = sold * gam;
      ew_axpby(*Wnew, imm(1.0), *Z, imm(-a2), W, imm(-a3), Wold, true, imm(a1));  // :128
      ew_axpby(x, imm(1.0), x, imm(cnew * eta), Wnew);                     // :129
      eta = -snew * eta;
      beta_r = std::fabs(snew) * beta_r;                                   // :133
      // swap3(xnew,x,xold) = xold, xnew, x   :136-142
      { Vec *t = Vnew; Vnew = Vold; Vold = V; V = t; }
      { Vec *t = Wnew; Wnew = Wold; Wold = W; W = t; }
      { Vec *t = Znew; Znew = Zold; Zold = Z; Z = t; }
      { double t = gnew; gnew = gold; gold = gam; gam = t; }
      { double t = cnew; cnew = cold; cold = c; c = t; }
      { double t = snew; snew = sold; sold = s; s = t; }
      done = log.update(beta_r);
    }
    log.finalize(beta_r);
  }
};

// ---------------------------------------------------------------- block triangular / diagonal
struct BlockNS : gsb_solver_s {
  int nb, half;
  bool diagonal;
  std::vector<gsb_mat_t> blocks;
  std::vector<gsb_solver_t> solvers;
  std::vector<double> coeffs;
  std::vector<int64_t> off;
  std::vector<VecP> w, y;  // BlockTriangularSolvers.jl:135-143 (y zeroed at set-up only)
  const char *name() const override { return diagonal ? "BlockDiagonal" : "BlockTriangular"; }
  bool capturable() const override {
    for (gsb_solver_t c : solvers)
      if (!c->capturable()) return false;
    return true;
  }
  BlockNS(gsb_ctx_t c, int nb_, const gsb_mat_t *b, const gsb_solver_t *s, const double *co, int half_, bool diag)
      : nb(nb_), half(half_), diagonal(diag) {
    ctx = c;
    blocks.assign(b, b + (size_t)nb * nb);
    solvers.assign(s, s + nb);
    coeffs.assign((size_t)nb * nb, 1.0);
    if (co) coeffs.assign(co, co + (size_t)nb * nb);
    off.assign((size_t)nb + 1, 0);
    for (int i = 0; i < nb; ++i) {
      int64_t n = -1;
      for (int j = 0; j < nb; ++j)
        if (blocks[(size_t)i * nb + j]) n = blocks[(size_t)i * nb + j]->n_rows;
      for (int j = 0; j < nb && n < 0; ++j)
        if (blocks[(size_t)j * nb + i]) n = blocks[(size_t)j * nb + i]->n_own_cols;
      GSB_CHECK(n >= 0, "block solver: cannot infer the size of block " + std::to_string(i));
      off[(size_t)i + 1] = off[(size_t)i] + n;
      w.push_back(make_vec(ctx, n, 0));
      y.push_back(make_vec(ctx, n, 0));
    }
  }
  void solve(Vec &x, Vec &b) override {
    GSB_CHECK(x.n_own == off.back() && b.n_own == off.back(), "block solver: vector size mismatch");
    for (int t = 0; t < nb; ++t) {
      const int iB = (diagonal || half == GSB_LOWER) ? t : nb - 1 - t;
      Vec bi = view(b, off[(size_t)iB], off[(size_t)iB + 1] - off[(size_t)iB]);
      Vec xi = view(x, off[(size_t)iB], off[(size_t)iB + 1] - off[(size_t)iB]);
      Vec *rhs = &bi;
      if (!diagonal) {  // BlockTriangularSolvers.jl:195-205 / :223-233
        vec_copy(*w[(size_t)iB], bi);
        const int j0 = (half == GSB_LOWER) ? 0 : iB + 1, j1 = (half == GSB_LOWER) ? iB : nb;
        for (int jB = j0; jB < j1; ++jB) {
          const double cij = coeffs[(size_t)iB * nb + jB];
          gsb_mat_t Bij = blocks[(size_t)iB * nb + jB];
          if (Bij && std::fabs(cij) > std::nextafter(std::fabs(cij), INFINITY) - std::fabs(cij)) {
            Vec xj = view(x, off[(size_t)jB], off[(size_t)jB + 1] - off[(size_t)jB]);
            spmv(Bij, xj, *w[(size_t)iB], -cij, 1.0);  // mul!(wi,A_ij,xj,-cij,1.0)
          }
        }
        rhs = w[(size_t)iB].get();
      }
      solvers[(size_t)iB]->solve(*y[(size_t)iB], *rhs);  // :208-212 / BlockDiagonalSolvers.jl:169-174
      vec_copy(xi, *y[(size_t)iB]);
    }
  }
};

// ---------------------------------------------------------------- RichardsonLinearSolver
struct RichardsonLinearNS : gsb_solver_s {
  gsb_mat_t A;
  gsb_solver_t Pl;
  double omega;
  VecP z, r;  // RichardsonLinearSolvers.jl:33-40
  Slots slots;
  int s_rr;
  const char *name() const override { return "RichardsonLinearSolver"; }
  gsb_mat_t matrix() override { return A; }
  RichardsonLinearNS(gsb_mat_t A_, gsb_solver_t Pl_, double w, int maxiter, double atol, double rtol) : A(A_), Pl(Pl_), omega(w) {
    ctx = A->ctx;
    log.configure(maxiter, atol, rtol);
    has_log = true;
    z = domain_vec(A); r = domain_vec(A);
    slots.take(ctx, 1);
    s_rr = slots.start;
  }
  void update(gsb_mat_t A_) override {
    if (Pl) Pl->update(A_);
    A = A_;
  }
  double residual(Vec &x, Vec &b) {  // r .= b ; mul!(r, A, x, -1, 1) ; norm(r)
    vec_copy(*r, b);
    spmv(A, x, *r, -1.0, 1.0);
    dot(*r, *r, s_rr);
    return std::sqrt(ctx->read_scalar(s_rr));
  }
  void solve(Vec &x, Vec &b) override {  // :79-106
    double res = residual(x, b);
    bool done = log.init(res);
    while (!done) {
      if (Pl) {
        Pl->solve(*z, *r);
        ew_axpby(x, imm(1.0), x, imm(omega), z.get());  // x .+= w .* z
      } else {
        ew_axpby(x, imm(1.0), x, imm(omega), r.get());
      }
      res = residual(x, b);
      done = log.update(res);
    }
    log.finalize(res);
  }
};

// ---------------------------------------------------------------- SchurComplementSolver
struct SchurNS : gsb_solver_s {
  gsb_solver_t A, S;
  gsb_mat_t B, C;
  VecP du, bu, bp;  // SchurComplementSolvers.jl:40-45
  int64_t nu, np;
  const char *name() const override { return "SchurComplement"; }
  bool capturable() const override { return A->capturable() && S->capturable(); }
  SchurNS(gsb_ctx_t c, gsb_solver_t A_, gsb_mat_t B_, gsb_mat_t C_, gsb_solver_t S_) : A(A_), S(S_), B(B_), C(C_) {
    ctx = c;
    GSB_CHECK(B->n_rows == C->n_own_cols && C->n_rows == B->n_own_cols, "Schur complement: B and C shapes do not match");
    nu = B->n_rows; np = C->n_rows;
    du = domain_vec(C); bu = domain_vec(C); bp = domain_vec(B);
  }
  void solve(Vec &x, Vec &y) override {  // :55-74
    GSB_CHECK(x.n_own == nu + np && y.n_own == nu + np, "Schur complement: vector size mismatch");
    Vec xu = view(x, 0, nu), xp = view(x, nu, np), yu = view(y, 0, nu), yp = view(y, nu, np);
    A->solve(xu, yu);                                   // x_u = A^-1 y_u
    vec_copy(*bp, yp);
    spmv(C, xu, *bp, -1.0, 1.0);                        // bp = y_p - C x_u
    S->solve(xp, *bp);                                  // x_p = S^-1 bp
    spmv(B, xp, *bu, 1.0, 0.0);                         // bu = B x_p
    A->solve(*du, *bu);                                 // du = A^-1 bu
    ew_axpby(xu, imm(1.0), xu, imm(-1.0), du.get());    // x_u .-= du
  }
};

}  // namespace

// ------------------------------------------------------------------------------------------ C ABI
extern "C" {

int gsb_identity_create(gsb_ctx_t ctx, gsb_solver_t *out) {
  GSB_NULLCHK(ctx)
  API_BEGIN
  auto *s = new IdentityNS();
  s->ctx = ctx;
  *out = s;
  API_END(ctx)
}
int gsb_jacobi_create(gsb_mat_t A, gsb_solver_t *out) {
  GSB_NULLCHK(A)
  API_BEGIN
  *out = new JacobiNS(A);
  API_END(A->ctx)
}
int gsb_richardson_create(gsb_mat_t A, gsb_solver_t M, int niter, double omega, gsb_solver_t *out) {
  GSB_NULLCHK(A)
  API_BEGIN
  GSB_CHECK(M != nullptr, "Richardson: inner solver is NULL");
  *out = new RichardsonNS(A, M, niter, omega);
  API_END(A->ctx)
}
int gsb_from_smoother_create(gsb_mat_t A, gsb_solver_t smoother, gsb_solver_t *out) {
  GSB_NULLCHK(A)
  API_BEGIN
  *out = new FromSmootherNS(A, smoother);
  API_END(A->ctx)
}
int gsb_dense_lu_create(gsb_mat_t A, gsb_solver_t *out) {
  GSB_NULLCHK(A)
  API_BEGIN
  *out = new DenseLUNS(A);
  API_END(A->ctx)
}
int gsb_gmg_create(gsb_ctx_t ctx, int nlev, const gsb_mat_t *mats, const gsb_mat_t *interp, const gsb_mat_t *restrict_,
                   const gsb_solver_t *pre, const gsb_solver_t *post, gsb_solver_t coarsest, int mode, int cycle_type,
                   int maxiter, double atol, double rtol, gsb_solver_t *out) {
  GSB_NULLCHK(ctx)
  API_BEGIN
  *out = new GMGNS(ctx, nlev, mats, interp, restrict_, pre, post, coarsest, mode, cycle_type, maxiter, atol, rtol);
  API_END(ctx)
}
int gsb_gmg_create_redist(gsb_ctx_t ctx, int nlev, const gsb_mat_t *mats, const gsb_mat_t *interp, const gsb_mat_t *restrict_,
                          const gsb_solver_t *pre, const gsb_solver_t *post, gsb_solver_t coarsest, int mode, int cycle_type,
                          int maxiter, double atol, double rtol, const gsb_plan_t *to_coarse, const gsb_plan_t *to_fine,
                          gsb_solver_t *out) {
  GSB_NULLCHK(ctx)
  API_BEGIN
  *out = new GMGNS(ctx, nlev, mats, interp, restrict_, pre, post, coarsest, mode, cycle_type, maxiter, atol, rtol, to_coarse, to_fine);
  API_END(ctx)
}
int gsb_cg_create(gsb_mat_t A, gsb_solver_t Pl, int flexible, int maxiter, double atol, double rtol, gsb_solver_t *out) {
  GSB_NULLCHK(A)
  API_BEGIN
  *out = new CGNS(A, Pl, flexible != 0, maxiter, atol, rtol);
  API_END(A->ctx)
}
int gsb_gmres_create(gsb_mat_t A, gsb_solver_t Pr, gsb_solver_t Pl, int m, int restart, int m_add, int maxiter,
                     double atol, double rtol, gsb_solver_t *out) {
  GSB_NULLCHK(A)
  API_BEGIN
  *out = new GMRESNS(A, Pr, Pl, m, restart != 0, m_add, false, maxiter, atol, rtol);
  API_END(A->ctx)
}
int gsb_fgmres_create(gsb_mat_t A, gsb_solver_t Pr, gsb_solver_t Pl, int m, int restart, int m_add, int maxiter,
                      double atol, double rtol, gsb_solver_t *out) {
  GSB_NULLCHK(A)
  API_BEGIN
  *out = new GMRESNS(A, Pr, Pl, m, restart != 0, m_add, true, maxiter, atol, rtol);
  API_END(A->ctx)
}
int gsb_minres_create(gsb_mat_t A, gsb_solver_t Pl, int maxiter, double atol, double rtol, gsb_solver_t *out) {
  GSB_NULLCHK(A)
  API_BEGIN
  *out = new MINRESNS(A, Pl, maxiter, atol, rtol);
  API_END(A->ctx)
}
int gsb_block_solver_create(gsb_ctx_t ctx, int nb, const gsb_mat_t *blocks, const gsb_solver_t *solvers,
                            const double *coeffs, int half, int diagonal, gsb_solver_t *out) {
  GSB_NULLCHK(ctx)
  API_BEGIN
  *out = new BlockNS(ctx, nb, blocks, solvers, coeffs, half, diagonal != 0);
  API_END(ctx)
}

int gsb_richardson_linear_create(gsb_mat_t A, gsb_solver_t Pl, double omega, int maxiter, double atol, double rtol,
                                 gsb_solver_t *out) {
  GSB_NULLCHK(A)
  API_BEGIN
  *out = new RichardsonLinearNS(A, Pl, omega, maxiter, atol, rtol);
  API_END(A->ctx)
}
int gsb_schur_complement_create(gsb_ctx_t ctx, gsb_solver_t A_ns, gsb_mat_t B, gsb_mat_t C, gsb_solver_t S_ns,
                                gsb_solver_t *out) {
  GSB_NULLCHK(ctx)
  API_BEGIN
  GSB_CHECK(A_ns && S_ns && B && C, "Schur complement: NULL argument");
  *out = new SchurNS(ctx, A_ns, B, C, S_ns);
  API_END(ctx)
}
int gsb_cg_record_coefficients(gsb_solver_t ns, int enable) {
  GSB_NULLCHK(ns)
  API_BEGIN
  CGNS *cg = dynamic_cast<CGNS *>(ns);
  GSB_CHECK(cg != nullptr, "not a CG numerical setup");
  cg->record = enable != 0;
  API_END(ns->ctx)
}
int gsb_cg_coefficients(gsb_solver_t ns, double *alpha, double *beta, int64_t cap, int64_t *n) {
  GSB_NULLCHK(ns)
  API_BEGIN
  CGNS *cg = dynamic_cast<CGNS *>(ns);
  GSB_CHECK(cg != nullptr, "not a CG numerical setup");
  *n = (int64_t)cg->alphas.size();
  for (int64_t i = 0; i < std::min<int64_t>(cap, *n); ++i) { alpha[i] = cg->alphas[(size_t)i]; beta[i] = cg->betas[(size_t)i]; }
  API_END(ns->ctx)
}

int gsb_solver_update(gsb_solver_t ns, gsb_mat_t A) {
  GSB_NULLCHK(ns)
  API_BEGIN
  ns->update(A);
  API_END(ns->ctx)
}
int gsb_solve(gsb_solver_t ns, gsb_vec_t x, gsb_vec_t b) {
  GSB_NULLCHK(ns)
  API_BEGIN
  ns->solve(*x, *b);
  GSB_CUDA(cudaStreamSynchronize(ns->ctx->stream));
  API_END(ns->ctx)
}
static int solve_host_impl(gsb_solver_t ns, double *x_host, const double *b_host, int64_t n, bool x0_is_zero) {
  API_BEGIN
  gsb_ctx_t ctx = ns->ctx;
  // staging vectors sized like the caller's own values; ghost room is taken from the solver's
  // matrix when it has one (domain layout), see gsb_solver_s::host_x
  struct Stage { VecP &x, &b; } st{ns->host_x, ns->host_b};
  if (!st.x || st.x->n_own != n) {
    gsb_mat_t M = ns->matrix();
    const int64_t ng = M ? M->n_ghost_cols : 0;
    st.x = make_vec(ctx, n, ng);
    st.b = make_vec(ctx, n, ng);
  }
  GSB_CUDA(cudaMemcpyAsync(st.b->d, b_host, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  if (x0_is_zero) vec_fill(*st.x, 0.0);  // the caller vouches for x0 = 0: no upload of the initial guess
  else GSB_CUDA(cudaMemcpyAsync(st.x->d, x_host, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  ns->solve(*st.x, *st.b);
  GSB_CUDA(cudaMemcpyAsync(x_host, st.x->d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  GSB_CUDA(cudaStreamSynchronize(ctx->stream));
  API_END(ns->ctx)
}
int gsb_solve_host(gsb_solver_t ns, double *x_host, const double *b_host, int64_t n) {
  GSB_NULLCHK(ns)
  return solve_host_impl(ns, x_host, b_host, n, false);
}
int gsb_solve_host_zero_guess(gsb_solver_t ns, double *x_host, const double *b_host, int64_t n) {
  GSB_NULLCHK(ns)
  return solve_host_impl(ns, x_host, b_host, n, true);
}
int gsb_solver_log(gsb_solver_t ns, int *num_iters, double *residuals, int64_t cap, int *flag) {
  GSB_NULLCHK(ns)
  API_BEGIN
  GSB_CHECK(ns->has_log, std::string(ns->name()) + " has no convergence log");
  ns->finish();
  if (num_iters) *num_iters = ns->log.num_iters;
  if (flag) *flag = ns->log.flag;
  if (residuals) {
    const int64_t n = std::min<int64_t>(cap, (int64_t)ns->log.num_iters + 1);
    for (int64_t i = 0; i < n; ++i) residuals[i] = ns->log.residuals[(size_t)i];
  }
  API_END(ns->ctx)
}
int gsb_solver_destroy(gsb_solver_t ns) {
  delete ns;
  return GSB_OK;
}

}  // extern "C"
