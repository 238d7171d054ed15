// gsb_internal.h -- host-side objects behind the opaque handles of include/gsb200.h.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/gsb200.h"
#include "types.h"

namespace gsb {

struct Error {
  int code;
  std::string msg;
};
[[noreturn]] void fail(int code, const std::string &msg);

#define GSB_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t e_ = (expr);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      ::gsb::fail(GSB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                 std::to_string(__LINE__) + ")");                                        \
  } while (0)
#define GSB_NCCL(expr)                                                                                    \
  do {                                                                                                    \
    ncclResult_t e_ = (expr);                                                                             \
    if (e_ != ncclSuccess)                                                                                \
      ::gsb::fail(GSB_ENCCL, std::string(#expr) + ": " + ncclGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                 std::to_string(__LINE__) + ")");                                         \
  } while (0)
#define GSB_CHECK(cond, msg)                          \
  do {                                                \
    if (!(cond)) ::gsb::fail(GSB_EINVAL, (msg));      \
  } while (0)

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t n_) { alloc(n_); }
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t n_) {
    release();
    n = n_;
    if (n) GSB_CUDA(cudaMalloc(&p, n * sizeof(T)));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

}  // namespace gsb

struct gsb_ctx_s {
  int device = 0, nranks = 1, rank = 0, num_sms = 148;
  cudaStream_t stream = nullptr;       // compute stream
  cudaStream_t comm_stream = nullptr;  // halo stream
  ncclComm_t comm = nullptr;
  cudaEvent_t t0 = nullptr, t1 = nullptr, ev_a = nullptr, ev_b = nullptr;
  int64_t launches = 0;
  // device scalars + reduction scratch
  gsb::DevBuf<double> scal;
  int next_slot = 8;  // slots 0..7: scratch for the C-ABI dot/norm calls
  std::vector<std::pair<int, int>> free_ranges;  // (start, count) of slot ranges returned by destroyed solvers
  gsb::DevBuf<double> partials;
  gsb::DevBuf<unsigned int> ticket;
  double *h_scal = nullptr;  // pinned host mirror for read-backs: H_SCAL_READ slots + one staging word
  static constexpr int H_SCAL_READ = 4096;
  std::map<std::string, std::string> opts;
  std::string err;
  // a caller that already knows ||b|| on the host (CG knows ||r||, GMRES knows ||V_j|| = 1) tells the
  // next inner solve, so that a maxiter=1 GMG preconditioner need not synchronise to decide `done`
  const double *hint_vec = nullptr;
  double hint_norm = 0.0;
  bool nccl_halo_in_use = false;  // some plan fell back to (or was asked to use) ncclSend/ncclRecv
  // optional per-launch profiling of the row kernels (bench.py's roofline numbers)
  struct ProfRec { int mode; int stream; int64_t nrows, nnz; cudaEvent_t e0, e1; };
  bool profiling = false;
  std::vector<ProfRec> prof;
  int alloc_slots(int n);
  void free_slots(int start, int n);
  double read_scalar(int slot);                 // blocking
  void read_scalars(int slot, int n, double *out);
  void allreduce_slot(int slot, int n = 1);     // in-place NCCL allreduce when nranks > 1
  gsb::ReduceOut reduce_out(int slot);
  std::string opt(const std::string &k, const std::string &dflt) const {
    auto it = opts.find(k);
    return it == opts.end() ? dflt : it->second;
  }
};

struct gsb_plan_s {
  gsb_ctx_t ctx;
  int64_t n_own = 0, n_ghost = 0;
  std::vector<int> nbr_snd, nbr_rcv;
  std::vector<int64_t> snd_ptrs, rcv_ptrs;  // per-neighbour offsets into the id lists
  gsb::DevBuf<int> snd_ids, rcv_ids;        // 0-based local ids
  gsb::DevBuf<double> snd_buf, rcv_buf;
  bool redist = false;          // redistribution plan: n_own = own size of the source layout, n_ghost = own size of the
                                // destination layout, rcv ids index OWN entries of the destination vector
  bool rcv_contiguous = false;  // ghosts of each neighbour form one ascending run -> no unpack
  // NVLink peer-memory exchange (CUDA IPC); see kernels.cuh p2p_push_kernel
  bool p2p = false;
  gsb::DevBuf<unsigned long long> seq_dev;  // exchange counter lives on the device (graph replay)
  void *block = nullptr;                 // [flags: nranks u64 | buf parity 0 | buf parity 1]
  size_t flag_bytes = 0;
  std::vector<void *> peer_base;         // opened IPC mappings of the send neighbours' blocks
  gsb::DevBuf<int> snd_nbr, nbr_rcv_dev;
  gsb::DevBuf<int64_t> snd_ptrs_dev;
  gsb::DevBuf<double *> peer_buf[2];
  gsb::DevBuf<unsigned long long *> peer_flag;
  gsb::DevBuf<unsigned int> ticket;
  ~gsb_plan_s();
};

struct gsb_vec_s {
  gsb_ctx_t ctx;
  int64_t n_own = 0, n_ghost = 0;
  double *d = nullptr;
  bool owns = true;
  int64_t n_local() const { return n_own + n_ghost; }
  ~gsb_vec_s() {
    if (owns && d) cudaFree(d);
  }
};

struct gsb_mat_s {
  gsb_ctx_t ctx;
  int64_t n_rows = 0, n_own_cols = 0, n_ghost_cols = 0, nnz = 0;
  gsb_plan_t plan = nullptr;
  // CSR mirror: row pointers always; column ids / values only for small matrices and for matrices that have
  // no block-SELL form (the row kernels of large matrices stream the block-SELL arrays only)
  gsb::DevBuf<int> rowptr, col;
  gsb::DevBuf<double> val;
  bool csr_kept = false;
  gsb::DevBuf<double> diag;  // diag(A_own_own), 0 for a missing entry (JacobiLinearSolvers.jl:20-23)
  gsb::DevBuf<int> diag_pos; // CSR position of each row's diagonal entry (-1: none), for value refreshes
  // upload permutation: position in the caller's value array of each CSR entry (for update_values)
  std::vector<int64_t> perm;  // empty == identity
  int max_row_nnz = 0;
  int G = 1;  // lanes per row of the CSR fallback kernel
  // block-SELL-32 (kernels.cuh sell_kernel): BS x BS blocks, one lane per block row, optional sorting of
  // the block rows by length inside windows of 256
  bool sell_ok = false;
  int bs = 1;
  bool sorted = false;
  int64_t n_brows = 0, n_slices = 0;
  int64_t sell_blocks = 0;  // stored blocks incl. padding
  int64_t sell_explicit = 0;  // (slice, k) pairs whose 32 block-column ids are stored explicitly (the others are affine)
  int64_t sell_aligned = 0;   // diagonal-aligned slices (all their column words are affine)
  gsb::DevBuf<int> sell_perm, sell_lmask, sell_off, sell_kbase, sell_kind, sell_bcol;
  gsb::DevBuf<double> sell_val;
  // halo overlap: slices whose rows touch no ghost column ("interior") run while the exchange is in
  // flight, the remaining ("boundary") slices after it
  bool split_ok = false;
  int64_t n_int_slices = 0, n_bnd_slices = 0;
  gsb::DevBuf<int> int_slices, bnd_slices;
  // block matrix (acts on concatenated vectors)
  int nb = 0;
  std::vector<gsb_mat_t> blocks;  // row-major nb*nb, may contain nullptr
  std::vector<int64_t> row_off, col_off;
  // bytes the row kernels stream per pass over the matrix (values + ids + per-row metadata)
  int64_t format_bytes() const;
};

namespace gsb {

struct Log {
  int maxiter = 1000;
  double atol = 1e-12, rtol = 1e-6;
  int num_iters = 0;
  int flag = 0;
  std::vector<double> residuals;
  void configure(int maxiter_, double atol_, double rtol_) {
    maxiter = maxiter_; atol = atol_; rtol = rtol_;
    residuals.assign((size_t)maxiter + 1, 0.0);
  }
  bool finished(int niter, double e_a, double e_r) const {  // SolverTolerances.jl:117-128
    return (niter >= maxiter) || (e_r < rtol) || (e_a < atol);
  }
  bool init(double r0) {  // ConvergenceLogs.jl:101-112
    num_iters = 0;
    std::fill(residuals.begin(), residuals.end(), 0.0);
    residuals[0] = r0;
    return finished(0, r0, 1.0);
  }
  bool update(double r) {  // ConvergenceLogs.jl:119-129
    num_iters += 1;
    residuals[(size_t)num_iters] = r;
    return finished(num_iters, r, r / residuals[0]);
  }
  int finalize(double r) {  // ConvergenceLogs.jl:136-150 ; SolverTolerances.jl:97-110
    const double e_r = r / residuals[0];
    if (e_r < rtol) flag = GSB_CONVERGED_RTOL;
    else if (r < atol) flag = GSB_CONVERGED_ATOL;
    else if (num_iters >= maxiter) flag = GSB_DIVERGED_MAXITER;
    else flag = GSB_DIVERGED_BREAKDOWN;
    return flag;
  }
};

}  // namespace gsb

struct gsb_solver_s {
  gsb_ctx_t ctx = nullptr;
  gsb::Log log;
  bool has_log = false;
  virtual ~gsb_solver_s() {}
  virtual void solve(gsb_vec_s &x, gsb_vec_s &b) = 0;  // solve!(x,ns,b)
  virtual void update(gsb_mat_t A) { (void)A; }       // numerical_setup!(ns,A)
  virtual const char *name() const = 0;
  virtual gsb_mat_t matrix() { return nullptr; }       // the system matrix, when the solver has one
  virtual void finish() {}                             // resolve a deferred (device-resident) log
  // true when solve() only enqueues device work (no host read-back, no data-dependent host control flow),
  // i.e. it may run inside a CUDA-graph capture
  virtual bool capturable() const { return false; }
  std::unique_ptr<gsb_vec_s> host_x, host_b;           // staging for gsb_solve_host
};
