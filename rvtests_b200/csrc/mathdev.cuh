// mathdev.cuh -- fp64 special functions used after the genotype sweep (O(M^3) per gene, no N).
// Every routine is __host__ __device__ so that tests/hostcheck can compile the very same source
// with g++ and compare it with the oracle on the CPU; the product only ever runs them on the GPU.
//
// Reference call sites these replace:
//   gsl_ran_beta_pdf           src/Model.h:2652 (SKAT weights), :2807 (SKAT-O weights)
//   gsl_cdf_chisq_Q(stat,1)    regression/LinearRegressionScoreTest.cpp:259-261
//   cdfchn -> cumchn -> cumchi -> cumgam   regression/cdflib.cpp:2634,5172,5141 (Liu tail)
//   MixtureChiSquare::getLiuPvalue         regression/MixtureChiSquare.cpp:44-83
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define RVT_HD __host__ __device__ __forceinline__
#define RVT_HDN __host__ __device__
#else
#define RVT_HD inline
#define RVT_HDN inline
#endif

namespace rvt {

// Regularised upper incomplete gamma Q(a,x), a>0.  Series for x<a+1, modified-Lentz continued
// fraction otherwise; fp64, relative accuracy ~1e-14 away from the far tails.
RVT_HDN double gamma_q(double a, double x) {
  if (!(x > 0.0)) return 1.0;
  const double lg = lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 20000; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (fabs(del) < fabs(sum) * 1e-17) break;
    }
    return 1.0 - sum * exp(-x + a * log(x) - lg);
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 20000; ++i) {
    double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < 1e-16) break;
  }
  return exp(-x + a * log(x) - lg) * h;
}

RVT_HDN double gamma_p(double a, double x) {
  if (!(x > 0.0)) return 0.0;
  if (x < a + 1.0) {
    const double lg = lgamma(a);
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 20000; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (fabs(del) < fabs(sum) * 1e-17) break;
    }
    return sum * exp(-x + a * log(x) - lg);
  }
  return 1.0 - gamma_q(a, x);
}

// P(chi2_df > x)
RVT_HDN double chisq_q(double x, double df) {
  if (!(x > 0.0)) return 1.0;
  if (df == 1.0) return erfc(sqrt(0.5 * x));
  return gamma_q(0.5 * df, 0.5 * x);
}
RVT_HDN double chisq_p(double x, double df) {
  if (!(x > 0.0)) return 0.0;
  if (df == 1.0) return erf(sqrt(0.5 * x));
  return gamma_p(0.5 * df, 0.5 * x);
}

// gsl_ran_chisq_pdf (gsl-1.16 randist/chisq.c:40-62)
RVT_HDN double chisq_pdf(double x, double nu) {
  if (x < 0) return 0.0;
  if (nu == 2.0) return exp(-x / 2.0) / 2.0;
  return exp((nu / 2 - 1) * log(x / 2) - x / 2 - lgamma(nu / 2)) / 2;
}

// gsl_ran_beta_pdf (gsl-1.16 randist/beta.c:43-74)
RVT_HDN double beta_pdf(double x, double a, double b) {
  if (x < 0 || x > 1) return 0.0;
  double gab = lgamma(a + b), ga = lgamma(a), gb = lgamma(b);
  if (x == 0.0 || x == 1.0) {
    if (a > 1.0 && b > 1.0) return 0.0;
    return exp(gab - ga - gb) * pow(x, a - 1) * pow(1 - x, b - 1);
  }
  return exp(gab - ga - gb + log(x) * (a - 1) + log1p(-x) * (b - 1));
}

// Beta(MAF) weight: src/Model.h:2644-2661 (squared, SKAT) / :2799-2813 (unsquared, SKAT-O)
RVT_HDN double beta_weight(double freq, double b1, double b2, bool squared) {
  if (freq > 0.5) freq = 1.0 - freq;
  if (freq > 1e-30) {
    double w = beta_pdf(freq, b1, b2);
    return squared ? w * w : w;
  }
  return 0.0;
}

// MixtureChiSquare::getLiuPvalue (regression/MixtureChiSquare.cpp:44-83).  lambda[0..n)
RVT_HDN double liu_pvalue(const double* lambda, int n, double Q) {
  double c1 = 0, c2 = 0, c3 = 0, c4 = 0;
  for (int i = 0; i < n; ++i) {
    double l = lambda[i];
    c1 += l;
    c2 += l * l;
    c3 += l * l * l;
    c4 += l * l * l * l;
  }
  double s1 = c3 / c2 / sqrt(c2);
  double s2 = c4 / c2 / c2;
  double sigmaQ = sqrt(2.0 * c2);
  double tstar = (Q - c1) / sigmaQ;
  double a, delta, l;
  if (s1 * s1 > s2) {
    a = 1 / (s1 - sqrt(s1 * s1 - s2));
    delta = (s1 * a - 1) * a * a;
    l = a * a - 2.0 * delta;
  } else {
    a = 1.0 / s1;
    delta = 0.0;
    l = c2 * c2 * c2 / c3 / c3;
  }
  double x = tstar * (sqrt(2.0) * a) + (l + delta);
  // cdfchn input checks (regression/cdflib.cpp:2768-2790): a non-zero status returns 1.
  if (!(x >= 0.0) || !(l > 0.0) || !(delta >= 0.0)) return 1.0;
  if (delta > 1.0e-10) {
    // non-central chi-square tail (Poisson mixture); only reachable through rounding of s1^2>s2
    double half = 0.5 * delta, wt = exp(-half), sum = 0.0;
    for (int i = 0; i < 2000; ++i) {
      sum += wt * gamma_p(0.5 * l + i, 0.5 * x);
      wt *= half / (i + 1);
      if (wt < 1e-18 && i > half) break;
    }
    return 1.0 - sum;
  }
  return gamma_q(0.5 * l, 0.5 * x);
}

}  // namespace rvt
