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

enum Corr { RBF = 0, MATERN12 = 1, MATERN32 = 2, MATERN52 = 3, ABSEXP = 4, CUBIC = 5, GENEXP = 6, MATERN_NU = 7 };
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
  return corr_accum(corr, acc, theta, diff);  // MATERN_NU accumulates sum theta_j d_j^2 like the other Matern kernels
}
B2_HD bool corr_has_extra_param(int corr) { return corr == GENEXP || corr == MATERN_NU; }

// ---- modified Bessel function of the second kind K_nu(x), real nu >= 0, x > 0 -----------------------------------
// What the reference's general-nu Matern calls scipy.special.kv for (kernel.py:201-207).  Temme's method: with
// nu = n + mu, |mu| <= 1/2, K_mu and K_mu+1 come from a power series (x <= 2) or from Steed's continued fraction
// (x > 2), then the upward recurrence K_{k+1} = K_{k-1} + (2 k / x) K_k, which is stable for K.
B2_HD double rgamma1p_odd(double mu) {
  // Gamma1(mu) = (1/Gamma(1 - mu) - 1/Gamma(1 + mu)) / (2 mu) from the Taylor series of 1/Gamma(1 + z) (odd part)
  const double m2 = mu * mu;
  return -(0.5772156649015329 + m2 * (-0.0420026350340952 + m2 * (-0.0421977345555443 + m2 * (0.0072189432466630 +
           m2 * (-0.0002152416741149 + m2 * (-0.0000201348547807 + m2 * 0.0000011330272320))))));
}
B2_HD void bessel_k_pair(double mu, double x, double* kmu, double* kmu1) {
  const double PI = 3.141592653589793, EPS = 1e-17;
  const double gampl = 1.0 / tgamma(1.0 + mu), gammi = 1.0 / tgamma(1.0 - mu);
  if (x <= 2.0) {
    const double b = 0.5 * x, d0 = -log(b), e0 = mu * d0;
    const double pimu = PI * mu;
    const double fact = fabs(pimu) < 1e-9 ? 1.0 : pimu / sin(pimu);
    const double fact2 = fabs(e0) < 1e-9 ? 1.0 : sinh(e0) / e0;
    const double gam1 = fabs(mu) < 0.1 ? rgamma1p_odd(mu) : (gammi - gampl) / (2.0 * mu);
    const double gam2 = 0.5 * (gammi + gampl);
    double ff = fact * (gam1 * cosh(e0) + gam2 * fact2 * d0);
    double sum = ff;
    const double ee = exp(e0);
    double p = 0.5 * ee / gampl, q = 0.5 / (ee * gammi), c = 1.0, sum1 = p;
    const double d = b * b;
    for (int i = 1; i <= 500; ++i) {
      ff = (i * ff + p + q) / ((double)i * i - mu * mu);
      c *= d / i;
      p /= (i - mu);
      q /= (i + mu);
      const double del = c * ff;
      sum += del;
      sum1 += c * (p - i * ff);
      if (fabs(del) < fabs(sum) * EPS) break;
    }
    *kmu = sum;
    *kmu1 = sum1 * 2.0 / x;
  } else {
    double b = 2.0 * (1.0 + x), d = 1.0 / b, h = d, delh = d, q1 = 0.0, q2 = 1.0;
    const double a1 = 0.25 - mu * mu;
    double q = a1, c = a1, a = -a1, s = 1.0 + q * delh;
    for (int i = 2; i <= 500; ++i) {
      a -= 2 * (i - 1);
      c = -a * c / i;
      const double qnew = (q1 - b * q2) / a;
      q1 = q2;
      q2 = qnew;
      q += c * qnew;
      b += 2.0;
      d = 1.0 / (b + a * d);
      delh = (b * d - 1.0) * delh;
      h += delh;
      const double dels = q * delh;
      s += dels;
      if (fabs(dels / s) < EPS) break;
    }
    h = a1 * h;
    *kmu = sqrt(PI / (2.0 * x)) * exp(-x) / s;
    *kmu1 = *kmu * (mu + x + 0.5 - h) / x;
  }
}
B2_HD double bessel_kv(double nu, double x) {
  const int nl = (int)(nu + 0.5);
  const double mu = nu - nl;
  double k0, k1;
  bessel_k_pair(mu, x, &k0, &k1);
  for (int i = 1; i <= nl; ++i) {  // K_{mu+i+1} = K_{mu+i-1} + 2 (mu + i) / x K_{mu+i}
    const double kn = k0 + 2.0 * (mu + i) / x * k1;
    k0 = k1;
    k1 = kn;
  }
  return k0;
}
// general-nu Matern: 2^(1-nu) / Gamma(nu) t^nu K_nu(t), t = sqrt(2 nu) h, h = 0 replaced by eps   kernel.py:201-207
B2_HD double matern_general(double h, double nu) {
  if (h == 0.0) h = DBL_EPSILON;
  const double t = sqrt(2.0 * nu) * h;
  if (t > 705.0) return 0.0;  // K_nu underflows
  return exp2(1.0 - nu) / tgamma(nu) * pow(t, nu) * bessel_kv(nu, t);
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

B2_HD double corr_finish_p(int corr, double acc, double pw) {
  if (corr == MATERN_NU) return matern_general(sqrt(acc), pw);  // pw = nu
  return corr_finish(corr, acc);
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
