// gp_math.h -- scalar float64 math of the hot path, shared by device kernels and a host test shim.
// Every function cites the reference statement it evaluates (paths relative to
// /root/reference/bayes_optim/).  Written from the formulas, not translated from the numpy code.
#pragma once
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b2 {

enum Corr { RBF = 0, MATERN12 = 1, MATERN32 = 2, MATERN52 = 3, ABSEXP = 4, CUBIC = 5, GENEXP = 6 };
enum Acq { ACQ_EI = 0, ACQ_PI = 1, ACQ_UCB = 2, ACQ_MGFI = 3 };

// ---- correlation: accumulate over features, then finish -------------------------------------------
// acc starts at corr_init(); one corr_accum per feature with diff = x_d - x'_d; corr_finish -> r.
B2_HD double corr_init(int corr) { return corr == CUBIC ? 1.0 : 0.0; }

B2_HD double corr_accum(int corr, double acc, double theta, double diff) {
  if (corr == ABSEXP) return acc + theta * fabs(diff);  // surrogate/gaussian_process/kernel.py:283-286
  if (corr == CUBIC) {                                   // kernel.py:455-464
    double td = fabs(diff) * theta;
    td = td > 1.0 ? 1.0 : td;
    return acc * (1.0 - td * td * (3.0 - 2.0 * td));
  }
  return acc + theta * (diff * diff);  // kernel.py:326-329 (RBF), :184-187 (Matern): sum theta_j d_j^2
}

// generalized_exponential carries its exponent next to theta: exp(-sum theta_j |d_j|^pw)   kernel.py:372-373
B2_HD double corr_accum_p(int corr, double acc, double theta, double diff, double pw) {
  if (corr == GENEXP) return acc + theta * pow(fabs(diff), pw);
  return corr_accum(corr, acc, theta, diff);
}

B2_HD double corr_finish(int corr, double acc) {
  switch (corr) {
    case RBF:
    case ABSEXP:
    case GENEXP:
      return exp(-acc);
    case MATERN12:
      return exp(-sqrt(acc));  // kernel.py:189-190
    case MATERN32: {           // kernel.py:192-195
      double k = sqrt(acc) * 1.7320508075688772;
      return (1.0 + k) * exp(-k);
    }
    case MATERN52: {  // kernel.py:197-200
      double k = sqrt(acc) * 2.23606797749979;
      return (1.0 + k + k * k / 3.0) * exp(-k);
    }
    default:
      return acc;  // CUBIC: the running product is the value (kernel.py:464)
  }
}

// ---- d r / d theta_j = corr_dtheta_factor(acc) * corr_dtheta_weight(diff_j) -------------------------------
// The reference's corr_grad_theta (surrogate/gaussian_process/gpr.py:736-770) implements RBF (:748:
// -diff^2 R), Matern-3/2 (:757: -3 exp(-sqrt3 h) diff^2 / 2) and absolute_exponential (:761: -|diff| R) and
// leaves the rest unbound (fit() crashes upstream).  Matern-5/2 is provided here from its
// derivative  d/ds[(1+k+k^2/3)e^-k] = -(5/6)(1+k)e^-k,  k = sqrt(5 s),  s = sum theta_j d_j^2.
B2_HD bool corr_has_dtheta(int corr) { return corr == RBF || corr == MATERN32 || corr == MATERN52 || corr == ABSEXP; }
B2_HD double corr_dtheta_factor(int corr, double acc) {
  switch (corr) {
    case RBF:
    case ABSEXP:
      return -exp(-acc);
    case MATERN32:
      return -1.5 * exp(-sqrt(acc) * 1.7320508075688772);
    case MATERN52: {
      double k = sqrt(acc) * 2.23606797749979;
      return -(5.0 / 6.0) * (1.0 + k) * exp(-k);
    }
    default:
      return 0.0;
  }
}
B2_HD double corr_dtheta_weight(int corr, double diff) { return corr == ABSEXP ? fabs(diff) : diff * diff; }

// ---- normal distribution ----------------------------------------------------------------------------
// scipy.stats.norm.cdf -> special.ndtr (erfc based); norm.pdf = exp(-x^2/2)/sqrt(2 pi)
B2_HD double norm_cdf(double z) { return 0.5 * erfc(-z * 0.7071067811865476); }
B2_HD double norm_pdf(double z) { return exp(-(z * z) / 2.0) / 2.5066282746310002; }

// ---- acquisition values from (yhat, mse) of ONE candidate -------------------------------------------
// yhat is negated when maximising and sd = sqrt(mse): acquisition/acquisition_fun.py:61-64.
// `plugin` arrives sign-adjusted (acquisition_fun.py:100-104).

// EI: acquisition_fun.py:162-164 (0 when sd/sqrt(sigma2) < 1e-6), :170-174.
B2_HD double acq_ei(double y, double sd, double sigma2, double plugin) {
  if (sd / sqrt(sigma2) < 1e-6) return 0.0;
  double d = plugin - y;
  double z = d / sd;
  return d * norm_cdf(z) + sd * norm_pdf(z);
}

// epsilon-PI (epsilon = 0: PI): acquisition_fun.py:212-216.
B2_HD double acq_pi(double y, double sd, double plugin, double eps) {
  double coef = y > 0 ? 1.0 - eps : 1.0 + eps;
  return norm_cdf((plugin - coef * y) / sd);
}

// UCB: acquisition_fun.py:133.
B2_HD double acq_ucb(double y, double sd, double alpha) { return y + alpha * sd; }

// MGFI: acquisition_fun.py:262 (t <= 22.36 -- applied by the caller once), :274 (isclose(sd,0): |sd| <=
// 1e-8), :280-283, :284-290 (exp overflow raises -> 0; inf -> 0).
B2_HD double acq_mgfi(double y, double sd, double plugin, double t) {
  if (fabs(sd) <= 1e-8) return 0.0;
  double sd2 = sd * sd;
  double y_p = y - t * sd2;
  double beta_p = (plugin - y_p) / sd;
  double term = t * (plugin - y - 1.0);
  double e = term + t * t * sd2 / 2.0;
  if (e > 709.782712893384) return 0.0;  // log(DBL_MAX): numpy's exp would overflow -> warning -> 0
  double v = norm_cdf(beta_p) * exp(e);
  if (!(fabs(v) <= DBL_MAX)) return 0.0;  // inf or nan
  return v;
}

B2_HD double acq_value(int acq, double yhat, double mse, double sigma2, double plugin, double par,
                       int minimize) {
  double y = minimize ? yhat : -yhat;
  double sd = sqrt(mse);
  switch (acq) {
    case ACQ_EI:
      return acq_ei(y, sd, sigma2, plugin);
    case ACQ_PI:
      return acq_pi(y, sd, plugin, par);
    case ACQ_UCB:
      return acq_ucb(y, sd, par);
    default:
      return acq_mgfi(y, sd, plugin, par);
  }
}

// numpy.argmax ordering: a beats b if a is NaN and b is not, or a > b, or equal value and lower index.
B2_HD bool arg_better(double av, long long ai, double bv, long long bi) {
  bool an = av != av, bn = bv != bv;
  if (an || bn) return an && (!bn || ai < bi);
  return av > bv || (av == bv && ai < bi);
}

}  // namespace b2
