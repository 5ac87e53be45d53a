// hostmath.cpp -- TEST SHIM: compiles the scalar device math of gp_math.h for the host so that the CPU
// test-suite can check the exact formulas the kernels evaluate against the oracle without a GPU.
// Not part of the product path (libb200bo.so never links it).
#include "gp_math.h"
#include "host_qr.h"

extern "C" {
double b2h_corr(int corr, const double* theta, const double* x, const double* y, int D) {
  double acc = b2::corr_init(corr);
  for (int d = 0; d < D; ++d) acc = b2::corr_accum(corr, acc, theta[d], x[d] - y[d]);
  return b2::corr_finish(corr, acc);
}
double b2h_acq(int acq, double yhat, double mse, double sigma2, double plugin, double par, int minimize) {
  if (acq == b2::ACQ_MGFI && par > 22.36) par = 22.36;
  return b2::acq_value(acq, yhat, mse, sigma2, plugin, par, minimize);
}
double b2h_kv(double nu, double x) { return b2::bessel_kv(nu, x); }
double b2h_matern(double h, double nu) { return b2::matern_general(h, nu); }
void b2h_thin_qr(const double* Ft, const double* yt, int N, int p, double* G, double* beta, double* rho) {
  b2::thin_qr_beta_rho(Ft, yt, N, p, G, beta, rho);
}
int b2h_arg_better(double av, long long ai, double bv, long long bi) { return b2::arg_better(av, ai, bv, bi); }
}
