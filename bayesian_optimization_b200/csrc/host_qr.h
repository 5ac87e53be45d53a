// host_qr.h -- host-side thin QR of the (N, p) panel Ft = L^-1 F for p > 1 regression trends (p <= 64, O(N p^2)).
//
// What it stands for upstream: Q, G = scipy.linalg.qr(Ft, mode="economic") (gpr.py:805), beta = G^-1 Q^T Yt (gpr.py:787)
// and rho = Yt - Q Q^T Yt (gpr.py:806).  Unblocked Householder reflections with LAPACK's conventions (dgeqr2 / dlarfg:
// R_jj = -sign(alpha) ||x||, v_0 = 1), so that G -- an attribute callers can read -- matches scipy sign for sign.
// Shared by libb200bo.so (factor path) and the host test shim (tests/test_hostmath.py checks it against scipy).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace b2 {

// Ft: (N, p) row-major; yt: (N,).  Outputs: G (p, p) row-major upper-triangular, beta (p,), rho (N,).
inline void thin_qr_beta_rho(const double* Ft, const double* yt, int N, int p, double* G, double* beta, double* rho) {
  std::vector<double> a((size_t)N * p), qty(yt, yt + N), tau(p, 0.0);
  for (int c = 0; c < p; ++c)
    for (int i = 0; i < N; ++i) a[(size_t)c * N + i] = Ft[(size_t)i * p + c];  // column-major working copy
  for (int j = 0; j < p && j < N; ++j) {
    double* x = &a[(size_t)j * N];
    double xn2 = 0.0;
    for (int i = j + 1; i < N; ++i) xn2 += x[i] * x[i];
    const double alpha = x[j];
    if (xn2 == 0.0) {
      tau[j] = 0.0;  // H = I
    } else {
      const double bt = -std::copysign(std::sqrt(alpha * alpha + xn2), alpha);
      tau[j] = (bt - alpha) / bt;
      const double sc = 1.0 / (alpha - bt);
      for (int i = j + 1; i < N; ++i) x[i] *= sc;
      x[j] = bt;
    }
    auto apply = [&](double* col) {  // col <- (I - tau v v^T) col, v = [1, x[j+1:]]
      double w = col[j];
      for (int i = j + 1; i < N; ++i) w += x[i] * col[i];
      w *= tau[j];
      col[j] -= w;
      for (int i = j + 1; i < N; ++i) col[i] -= w * x[i];
    };
    if (tau[j] != 0.0) {
      for (int c = j + 1; c < p; ++c) apply(&a[(size_t)c * N]);
      apply(qty.data());  // the reflectors go onto Yt as they are formed: qty = Q^T Yt (full)
    }
  }
  for (int i = 0; i < p * p; ++i) G[i] = 0.0;
  for (int i = 0; i < p; ++i)
    for (int c = i; c < p; ++c) G[(size_t)i * p + c] = a[(size_t)c * N + i];
  for (int i = p - 1; i >= 0; --i) {  // beta = G^-1 (Q^T Yt)[:p]
    double v = qty[i];
    for (int c = i + 1; c < p; ++c) v -= G[(size_t)i * p + c] * beta[c];
    beta[i] = v / G[(size_t)i * p + i];
  }
  std::vector<double> r(qty);  // rho = Q [0; (Q^T Yt)[p:]]
  for (int i = 0; i < p && i < N; ++i) r[i] = 0.0;
  for (int j = std::min(p, N) - 1; j >= 0; --j) {
    if (tau[j] == 0.0) continue;
    const double* x = &a[(size_t)j * N];
    double w = r[j];
    for (int i = j + 1; i < N; ++i) w += x[i] * r[i];
    w *= tau[j];
    r[j] -= w;
    for (int i = j + 1; i < N; ++i) r[i] -= w * x[i];
  }
  for (int i = 0; i < N; ++i) rho[i] = r[i];
}

}  // namespace b2
