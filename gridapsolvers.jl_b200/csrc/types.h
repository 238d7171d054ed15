// types.h -- plain structs shared by the kernels (kernels.cuh) and the host code.
#pragma once
#include <stdint.h>
#ifndef __CUDACC__
#define GSB_HD
#else
#define GSB_HD __host__ __device__
#endif

namespace gsb {

struct ScalarRef {
  double v;   // immediate value, used when num < 0
  int num;    // slot of numerator (or -1)
  int sub;    // slot subtracted from the numerator (or -1)   s = (scal[num]-scal[sub])/scal[den]
  int den;    // slot of the denominator (or -1)
  int neg;    // negate the result
};
GSB_HD inline ScalarRef imm(double v) { return ScalarRef{v, -1, -1, -1, 0}; }
GSB_HD inline ScalarRef slot_ratio(int num, int den, int neg = 0, int sub = -1) {
  return ScalarRef{0.0, num, sub, den, neg};
}

struct ReduceOut {
  double *partials;      // >= gridDim.x * nred doubles
  unsigned int *ticket;  // zero-initialised, self-resetting
  double *scal;          // device scalar array
  int slot[2];           // where the totals go
};


enum RowMode {
  ROW_SPMV = 0,      // y = beta*y + A*(alpha*x)                         mul!(y,A,x,alpha,beta)
  ROW_RESID = 1,     // out = b - A*x                                    r .= b .- w / r .= r .- Adx
  ROW_SWEEP = 2,     // out = b - A*x ; d = omega*(invd*out) ; dxout = d ; xacc += d   (fused Jacobi-Richardson)
  ROW_SPMV_DOT = 3,  // y = A*x ; acc += dotv[row]*y[row]                 w = A p ; p.w
  ROW_SPMV_ADD = 4,  // y = A*x ; xacc += y                               dxh = P dxH ; xh .= xh .+ dxh
};

struct RowArgs {
  const double *x;  // gathered vector (own + ghost entries)
  double *y;        // SPMV / SPMV_DOT / SPMV_ADD output
  const double *b;  // RESID / SWEEP minuend (may alias out)
  double *out;      // RESID / SWEEP output
  const double *invd;
  double *dxout;
  double *xacc;
  const double *dotv;
  double alpha, beta, omega;
  ReduceOut red;
};





struct EwArgs {
  double *z;
  const double *x, *y, *w;
  ScalarRef a, b, c, d;
  int has_y, has_w, has_d, mul_xy;  // mul_xy: z = x .* y (Jacobi apply)
  const double *scal;
  int64_t n;
};


}  // namespace gsb
