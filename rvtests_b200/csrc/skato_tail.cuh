// skato_tail.cuh -- SKAT-O after the genotype sweep (K4): everything SkatO::Fit does, restated on
// the M x M sufficient statistics so that no further access to the N x M genotypes is needed.
//
// Replaces regression/SkatO.cpp:101-281 (Fit), :60-99 (FitSKAT, M == 1), helpers :343-455,
// integrands :303-337, and Integration::integrateLU = gsl_integration_qags with limit 1000,
// epsabs 1e-25, epsrel 0.0001220703 (regression/GSLIntegration.cpp:37-49, SkatO.cpp:236-242).
//
// Algebra.  With G_w = G diag(w) (UN-squared Beta weights, src/Model.h:2799-2813)
//   Wm   := Z1'Z1 = W (G'G - G'X (X'X)^-1 X'G) W / 2                       (SkatO.cpp:150-160)
//   Q_rho = [(1-rho) sum_j v_j^2 + rho (sum_j v_j)^2] / (2 s2),  v = W G'r   (SkatO.cpp:141-147)
//   K_rho = L'Wm L, L = chol(R_rho)  has the eigenvalues of  R^1/2 Wm R^1/2, and
//           R^1/2 = a I + b 11' with a = sqrt(1-rho), b = (sqrt(1-rho+rho M) - a)/M, so
//           R^1/2 Wm R^1/2 = a^2 Wm + a b (1 c' + c 1') + b^2 (1'Wm 1) 11',  c = Wm 1
//           -- a rank-2 update instead of a Cholesky + two M x M products per rho (SkatO.cpp:163-175)
//   zbar'Z1 = c'/M,  ||zbar||^2 = 1'Wm1 / M^2,  ZMZ = (c c')/(M^2 ||zbar||^2),  ZIMZ = Wm - ZMZ  (:178-185)
// The quadrature is a QAGS state machine (QUADPACK dqagse as implemented by GSL 1.16
// integration/qags.c, qelg.c, qpsrt.c, qk.c with the 21-point Kronrod rule): one thread owns the
// interval list, the whole group evaluates the integrand (a Davies evaluation per node).
#pragma once
#include "eigen.cuh"

namespace rvt {

// ---- chi-square quantile: x with P(chi2_df > x) = q  (gsl_cdf_chisq_Qinv, SkatO.cpp:431) ----
RVT_HDN double chisq_qinv(double q, double df) {
  if (!(q < 1.0)) return 0.0;
  if (!(q > 0.0)) return INFINITY;
  // Wilson-Hilferty start
  const double a = 0.5 * df;
  double t;  // normal upper quantile of q (rational approx, refined by the Newton steps below)
  {
    const double pp = (q < 0.5) ? q : 1.0 - q;
    const double s = sqrt(-2.0 * log(pp));
    double z = s - (2.515517 + 0.802853 * s + 0.010328 * s * s) / (1.0 + 1.432788 * s + 0.189269 * s * s + 0.001308 * s * s * s);
    t = (q < 0.5) ? z : -z;
  }
  double x = df * pow(1.0 - 2.0 / (9.0 * df) + t * sqrt(2.0 / (9.0 * df)), 3.0);
  if (!(x > 0.0)) x = 1e-8 * df;
  double lo = 0.0, hi = INFINITY;
  const double lg = lgamma(a);
  for (int it = 0; it < 200; ++it) {
    const double f = gamma_q(a, 0.5 * x) - q;  // decreasing in x
    if (f > 0.0)
      lo = x;
    else
      hi = x;
    if (f == 0.0) break;
    // pdf of chi2_df at x
    const double lpdf = (a - 1.0) * log(0.5 * x) - 0.5 * x - lg;
    const double pdf = 0.5 * exp(lpdf);
    double xn = x + f / pdf;  // Newton on the survival function
    if (!(xn > lo) || !(xn < hi) || !(pdf > 0.0)) xn = (hi < INFINITY) ? 0.5 * (lo + hi) : 2.0 * x + 1.0;
    if (fabs(xn - x) <= 1e-15 * fabs(xn)) {
      x = xn;
      break;
    }
    x = xn;
  }
  return x;
}

// ---- QAGS ------------------------------------------------------------------------------------
struct QagsWork {   // `limit` entries each; owned by ONE thread
  double *a, *b, *r, *e;
  int *order, *level;
  int limit;
  long long deadline;       // device watchdog: clock64() value after which the quadrature gives up (0 = none)
};

constexpr int kQagsLimit = 1000;   // Integration::limit (regression/GSLIntegration.cpp:7-15)

// per-gene QAGS interval list in global memory (touched by one thread only)
struct QagsScratch {
  double a[kQagsLimit], b[kQagsLimit], r[kQagsLimit], e[kQagsLimit];
  int order[kQagsLimit], level[kQagsLimit];
};

constexpr double kDblEps = 2.2204460492503131e-16;
constexpr double kDblMin = 2.2250738585072014e-308;
constexpr double kDblMax = 1.7976931348623157e+308;

struct GkResult {
  double result, abserr, resabs, resasc;
};

// nodes of the 21-point Kronrod rule in the order the integrand is sampled (qk.c):
// centre; the 5 Gauss nodes (-,+); the 5 Kronrod-only nodes (-,+).  node in [-1,1].
RVT_HD double gk21_node(int i) {
  const double xgk[11] = {0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
                          0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
                          0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
                          0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
                          0.294392862701460198131126603103866, 0.148874338981631210884826001129720,
                          0.000000000000000000000000000000000};
  if (i == 0) return 0.0;
  const int k = (i - 1) >> 1;                 // 0..9
  const int idx = (k < 5) ? (2 * k + 1) : (2 * (k - 5));
  return ((i - 1) & 1) ? xgk[idx] : -xgk[idx];
}

// combine the 21 samples fv[i] = f(centre + half*gk21_node(i)) exactly as qk.c does
RVT_HDN GkResult gk21_combine(const double* fv, double a, double b) {
  const double wg[5] = {0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
                        0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
                        0.295524224714752870173892994651338};
  const double wgk[11] = {0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
                          0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
                          0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
                          0.123491976262065851077958109831074, 0.134709217311473325928054001771707,
                          0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
                          0.149445554002916905664936468389821};
  const double half = 0.5 * (b - a), ahalf = fabs(half);
  const double fc = fv[0];
  double fv1[10], fv2[10];
  double rg = 0.0, rk = fc * wgk[10], rabs = fabs(rk);
  for (int j = 0; j < 5; ++j) {
    const int jtw = 2 * j + 1;
    const double f1 = fv[1 + 2 * j], f2 = fv[2 + 2 * j];
    fv1[jtw] = f1;
    fv2[jtw] = f2;
    rg += wg[j] * (f1 + f2);
    rk += wgk[jtw] * (f1 + f2);
    rabs += wgk[jtw] * (fabs(f1) + fabs(f2));
  }
  for (int j = 0; j < 5; ++j) {
    const int jt = 2 * j;
    const double f1 = fv[11 + 2 * j], f2 = fv[12 + 2 * j];
    fv1[jt] = f1;
    fv2[jt] = f2;
    rk += wgk[jt] * (f1 + f2);
    rabs += wgk[jt] * (fabs(f1) + fabs(f2));
  }
  const double mean = rk * 0.5;
  double rasc = wgk[10] * fabs(fc - mean);
  for (int j = 0; j < 10; ++j) rasc += wgk[j] * (fabs(fv1[j] - mean) + fabs(fv2[j] - mean));
  double err = (rk - rg) * half;
  rk *= half;
  rabs *= ahalf;
  rasc *= ahalf;
  // rescale_error (err.c)
  err = fabs(err);
  if (rasc != 0 && err != 0) {
    const double scale = pow((200 * err / rasc), 1.5);
    err = (scale < 1) ? rasc * scale : rasc;
  }
  if (rabs > kDblMin / (50 * kDblEps)) {
    const double min_err = 50 * kDblEps * rabs;
    if (min_err > err) err = min_err;
  }
  GkResult o;
  o.result = rk;
  o.abserr = err;
  o.resabs = rabs;
  o.resasc = rasc;
  return o;
}

// The adaptive driver as a resumable machine: want() hands out the next interval to sample,
// give() consumes its Gauss-Kronrod result.  status: 0 success, else the QUADPACK error class
// (1 max iterations, 2 roundoff, 3 singularity, 4 extrapolation roundoff, 5 divergent, 6 failed,
//  7 bad tolerance) -- SkatO::Fit only tests for non-zero (SkatO.cpp:243-255).
struct QagsMachine {
  QagsWork w;
  double epsabs, epsrel;
  // workspace bookkeeping
  int size, nrmax, cur, maximum_level;
  // driver state
  int stage;  // 0 first interval, 1 left half, 2 right half, 3 finished
  double result, abserr;
  int status;
  double area, errsum, res_ext, err_ext, tolerance, ertest, error_over_large_intervals;
  double reseps, abseps, correc, resabs0;
  int ktmin, roundoff_type1, roundoff_type2, roundoff_type3, error_type, error_type2, iteration;
  int positive_integrand, extrapolate, disallow_extrapolation;
  // extrapolation table (qelg.c)
  int tab_n, tab_nres;
  double rlist2[52], res3la[3];
  // the interval being bisected
  double a1, b1, a2, b2, r_i, e_i;
  int current_level;
  GkResult g1;

  RVT_HDN void init(const QagsWork& work, double a, double b, double ea, double er) {
    w = work;
    epsabs = ea;
    epsrel = er;
    size = 0;
    nrmax = 0;
    cur = 0;
    maximum_level = 0;
    w.a[0] = a;
    w.b[0] = b;
    w.r[0] = 0.0;
    w.e[0] = 0.0;
    w.order[0] = 0;
    w.level[0] = 0;
    result = 0;
    abserr = 0;
    status = 0;
    stage = 0;
    ertest = 0;
    error_over_large_intervals = 0;
    reseps = abseps = correc = 0;
    ktmin = 0;
    roundoff_type1 = roundoff_type2 = roundoff_type3 = 0;
    error_type = error_type2 = 0;
    iteration = 0;
    positive_integrand = extrapolate = disallow_extrapolation = 0;
    tab_n = tab_nres = 0;
    a1 = a;
    b1 = b;
    if (epsabs <= 0 && (epsrel < 50 * kDblEps || epsrel < 0.5e-28)) {
      status = 7;
      stage = 3;
    }
  }
  RVT_HDN bool want(double* lo, double* hi) const {
    if (stage == 3) return false;
    if (stage == 2) {
      *lo = a2;
      *hi = b2;
    } else {
      *lo = a1;
      *hi = b1;
    }
    return true;
  }

  RVT_HDN void qpsrt() {
    const int last = size - 1, limit = w.limit;
    int i_nrmax = nrmax, i_maxerr = w.order[i_nrmax];
    if (last < 2) {
      w.order[0] = 0;
      w.order[1] = 1;
      cur = i_maxerr;
      return;
    }
    const double errmax = w.e[i_maxerr];
    while (i_nrmax > 0 && errmax > w.e[w.order[i_nrmax - 1]]) {
      w.order[i_nrmax] = w.order[i_nrmax - 1];
      i_nrmax--;
    }
    const int top = (last < (limit / 2 + 2)) ? last : limit - last + 1;
    int i = i_nrmax + 1;
    while (i < top && errmax < w.e[w.order[i]]) {
      w.order[i - 1] = w.order[i];
      i++;
    }
    w.order[i - 1] = i_maxerr;
    const double errmin = w.e[last];
    int k = top - 1;
    while (k > i - 2 && errmin >= w.e[w.order[k]]) {
      w.order[k + 1] = w.order[k];
      k--;
    }
    w.order[k + 1] = last;
    i_maxerr = w.order[i_nrmax];
    cur = i_maxerr;
    nrmax = i_nrmax;
  }
  RVT_HDN void update(double area1, double error1, double area2, double error2) {
    const int i_max = cur, i_new = size;
    const int new_level = w.level[i_max] + 1;
    if (error2 > error1) {
      w.a[i_max] = a2;
      w.r[i_max] = area2;
      w.e[i_max] = error2;
      w.level[i_max] = new_level;
      w.a[i_new] = a1;
      w.b[i_new] = b1;
      w.r[i_new] = area1;
      w.e[i_new] = error1;
      w.level[i_new] = new_level;
    } else {
      w.b[i_max] = b1;
      w.r[i_max] = area1;
      w.e[i_max] = error1;
      w.level[i_max] = new_level;
      w.a[i_new] = a2;
      w.b[i_new] = b2;
      w.r[i_new] = area2;
      w.e[i_new] = error2;
      w.level[i_new] = new_level;
    }
    size++;
    if (new_level > maximum_level) maximum_level = new_level;
    qpsrt();
  }
  RVT_HDN bool increase_nrmax() {
    const int id = nrmax, limit = w.limit, last = size - 1;
    const int jupbnd = (last > (1 + limit / 2)) ? limit + 1 - last : last;
    for (int k = id; k <= jupbnd; k++) {
      const int i_max = w.order[nrmax];
      cur = i_max;
      if (w.level[i_max] < maximum_level) return true;
      nrmax++;
    }
    return false;
  }
  RVT_HDN void qelg(double* res_out, double* err_out) {
    double* epstab = rlist2;
    const int n = tab_n - 1;
    const double current = epstab[n];
    double absolute = kDblMax, relative = 5 * kDblEps * fabs(current);
    const int newelm = n / 2, n_orig = n;
    int n_final = n;
    const int nres_orig = tab_nres;
    *res_out = current;
    *err_out = kDblMax;
    if (n < 2) {
      *res_out = current;
      *err_out = fmax(absolute, relative);
      return;
    }
    epstab[n + 2] = epstab[n];
    epstab[n] = kDblMax;
    for (int i = 0; i < newelm; i++) {
      double res = epstab[n - 2 * i + 2];
      const double e0 = epstab[n - 2 * i - 2], e1 = epstab[n - 2 * i - 1], e2 = res;
      const double e1abs = fabs(e1), delta2 = e2 - e1, err2 = fabs(delta2);
      const double tol2 = fmax(fabs(e2), e1abs) * kDblEps;
      const double delta3 = e1 - e0, err3 = fabs(delta3), tol3 = fmax(e1abs, fabs(e0)) * kDblEps;
      if (err2 <= tol2 && err3 <= tol3) {
        *res_out = res;
        absolute = err2 + err3;
        relative = 5 * kDblEps * fabs(res);
        *err_out = fmax(absolute, relative);
        return;
      }
      const double e3 = epstab[n - 2 * i];
      epstab[n - 2 * i] = e1;
      const double delta1 = e1 - e3, err1 = fabs(delta1), tol1 = fmax(e1abs, fabs(e3)) * kDblEps;
      if (err1 <= tol1 || err2 <= tol2 || err3 <= tol3) {
        n_final = 2 * i;
        break;
      }
      const double ss = (1 / delta1 + 1 / delta2) - 1 / delta3;
      if (fabs(ss * e1) <= 0.0001) {
        n_final = 2 * i;
        break;
      }
      res = e1 + 1 / ss;
      epstab[n - 2 * i] = res;
      const double error = err2 + fabs(res - e2) + err3;
      if (error <= *err_out) {
        *err_out = error;
        *res_out = res;
      }
    }
    const int limexp = 50 - 1;
    if (n_final == limexp) n_final = 2 * (limexp / 2);
    if (n_orig % 2 == 1) {
      for (int i = 0; i <= newelm; i++) epstab[1 + i * 2] = epstab[i * 2 + 3];
    } else {
      for (int i = 0; i <= newelm; i++) epstab[i * 2] = epstab[i * 2 + 2];
    }
    if (n_orig != n_final)
      for (int i = 0; i <= n_final; i++) epstab[i] = epstab[n_orig - n_final + i];
    tab_n = n_final + 1;
    if (nres_orig < 3) {
      res3la[nres_orig] = *res_out;
      *err_out = kDblMax;
    } else {
      *err_out = (fabs(*res_out - res3la[2]) + fabs(*res_out - res3la[1]) + fabs(*res_out - res3la[0]));
      res3la[0] = res3la[1];
      res3la[1] = res3la[2];
      res3la[2] = *res_out;
    }
    tab_nres = nres_orig + 1;
    *err_out = fmax(*err_out, 5 * kDblEps * fabs(*res_out));
  }
  RVT_HDN void next_bisection() {
    // retrieve the interval with the largest error estimate and split it
    const double a_i = w.a[cur], b_i = w.b[cur];
    r_i = w.r[cur];
    e_i = w.e[cur];
    current_level = w.level[cur] + 1;
    a1 = a_i;
    b1 = 0.5 * (a_i + b_i);
    a2 = b1;
    b2 = b_i;
    iteration++;
    stage = 1;
  }
  RVT_HDN void finish_sum() {  // compute_result
    double s = 0;
    for (int k = 0; k < size; k++) s += w.r[k];
    result = s;
    abserr = errsum;
    finish_error();
  }
  RVT_HDN void finish_error() {  // return_error
    if (error_type > 2) error_type--;
    status = error_type;  // 0 success
    stage = 3;
  }
  RVT_HDN void finish_ext() {  // after the main loop
    result = res_ext;
    abserr = err_ext;
    if (err_ext == kDblMax) return finish_sum();
    if (error_type || error_type2) {
      if (error_type2) err_ext += correc;
      if (error_type == 0) error_type = 3;
      if (res_ext != 0.0 && area != 0.0) {
        if (err_ext / fabs(res_ext) > errsum / fabs(area)) return finish_sum();
      } else if (err_ext > errsum) {
        return finish_sum();
      } else if (area == 0.0) {
        return finish_error();
      }
    }
    {
      const double max_area = fmax(fabs(res_ext), fabs(area));
      if (!positive_integrand && max_area < 0.01 * resabs0) return finish_error();
    }
    {
      const double ratio = res_ext / area;
      if (ratio < 0.01 || ratio > 100.0 || errsum > fabs(area)) error_type = 6;
    }
    finish_error();
  }

  RVT_HDN void give(const GkResult& g) {
    const int limit = w.limit;
    if (stage == 0) {
      // first integration over the whole range
      size = 1;
      w.r[0] = g.result;
      w.e[0] = g.abserr;
      resabs0 = g.resabs;
      tolerance = fmax(epsabs, epsrel * fabs(g.result));
      if (g.abserr <= 100 * kDblEps * g.resabs && g.abserr > tolerance) {
        result = g.result;
        abserr = g.abserr;
        status = 2;
        stage = 3;
        return;
      } else if ((g.abserr <= tolerance && g.abserr != g.resasc) || g.abserr == 0.0) {
        result = g.result;
        abserr = g.abserr;
        status = 0;
        stage = 3;
        return;
      } else if (limit == 1) {
        result = g.result;
        abserr = g.abserr;
        status = 1;
        stage = 3;
        return;
      }
      tab_n = 0;
      tab_nres = 0;
      rlist2[tab_n++] = g.result;
      area = g.result;
      errsum = g.abserr;
      res_ext = g.result;
      err_ext = kDblMax;
      positive_integrand = (fabs(g.result) >= (1 - 50 * kDblEps) * g.resabs);
      iteration = 1;
      next_bisection();
      return;
    }
    if (stage == 1) {
      g1 = g;
      stage = 2;
      return;
    }
    // stage 2: both halves are known -> one pass of the main loop body
    const double area1 = g1.result, error1 = g1.abserr, resasc1 = g1.resasc;
    const double area2 = g.result, error2 = g.abserr, resasc2 = g.resasc;
    const double area12 = area1 + area2, error12 = error1 + error2;
    const double last_e_i = e_i;
    errsum = errsum + error12 - e_i;
    area = area + area12 - r_i;
    tolerance = fmax(epsabs, epsrel * fabs(area));
    if (resasc1 != error1 && resasc2 != error2) {
      const double delta = r_i - area12;
      if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) {
        if (!extrapolate)
          roundoff_type1++;
        else
          roundoff_type2++;
      }
      if (iteration > 10 && error12 > e_i) roundoff_type3++;
    }
    if (roundoff_type1 + roundoff_type2 >= 10 || roundoff_type3 >= 20) error_type = 2;
    if (roundoff_type2 >= 5) error_type2 = 1;
    {
      const double tmp = (1 + 100 * kDblEps) * (fabs(a2) + 1000 * kDblMin);
      if (fabs(a1) <= tmp && fabs(b2) <= tmp) error_type = 4;
    }
    update(area1, error1, area2, error2);
    if (errsum <= tolerance) return finish_sum();
    if (error_type) return finish_ext();
    if (iteration >= limit - 1) {
      error_type = 1;
      return finish_ext();
    }
    bool go_on = false;  // `continue` of the reference loop
    if (iteration == 2) {
      error_over_large_intervals = errsum;
      ertest = tolerance;
      rlist2[tab_n++] = area;
      go_on = true;
    } else if (disallow_extrapolation) {
      go_on = true;
    } else {
      error_over_large_intervals += -last_e_i;
      if (current_level < maximum_level) error_over_large_intervals += error12;
      if (!extrapolate) {
        if (w.level[cur] < maximum_level) {
          go_on = true;
        } else {
          extrapolate = 1;
          nrmax = 1;
        }
      }
      if (!go_on && !error_type2 && error_over_large_intervals > ertest) {
        if (increase_nrmax()) go_on = true;
      }
      if (!go_on) {
        rlist2[tab_n++] = area;
        qelg(&reseps, &abseps);
        ktmin++;
        if (ktmin > 5 && err_ext < 0.001 * errsum) error_type = 5;
        if (abseps < err_ext) {
          ktmin = 0;
          err_ext = abseps;
          res_ext = reseps;
          correc = error_over_large_intervals;
          ertest = fmax(epsabs, epsrel * fabs(reseps));
          if (err_ext <= ertest) return finish_ext();
        }
        if (tab_n == 1) disallow_extrapolation = 1;
        if (error_type == 5) return finish_ext();
        nrmax = 0;
        cur = w.order[0];
        extrapolate = 0;
        error_over_large_intervals = errsum;
      }
    }
    if (iteration < limit)
      next_bisection();
    else
      finish_ext();
  }
};

// ---- SKAT-O proper -----------------------------------------------------------------------------
struct SkatoMoment {
  double muQ, varQ, df;
};

// SkatO.cpp:350-382: from eigenvalues sorted DESCENDING in ev[0..n): number kept (>= mean of the
// positive ones / 1e5, scanning from the small end), or -1 when none is positive.
RVT_HDN int skato_keep(const double* ev, int n) {
  int npos = 0;
  double spos = 0.0;
  for (int i = n - 1; i >= 0; --i)  // ascending order, like the reference's accumulation
    if (ev[i] > 0) {
      ++npos;
      spos += ev[i];
    }
  if (npos == 0) return -1;
  const double t = spos / npos / 100000;
  int keep = n;
  for (int i = n - 1; i >= 0; --i) {
    if (ev[i] < t)
      --keep;
    else
      break;
  }
  return keep;
}

// SkatO.cpp:383-416
RVT_HDN SkatoMoment skato_moment(const double* lam, int n) {
  double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  for (int i = 0; i < n; ++i) {
    const double l = lam[i], l2 = l * l;
    c0 += l;
    c1 += l2;
    c2 += l2 * l;
    c3 += l2 * l2;
  }
  const double sigmaQ = sqrt(2 * c1);
  const double s1 = c2 / c1 / sqrt(c1);
  const double s2 = c3 / (c1 * c1);
  double l;
  if (s1 * s1 > s2) {
    const double a = 1 / (s1 - sqrt(s1 * s1 - s2));
    const double d = (s1 * a - 1.0 * a * a);
    l = a * a - 2 * d;
  } else {
    l = 1. / s2;
  }
  SkatoMoment m;
  m.muQ = c0;
  m.varQ = sigmaQ * sigmaQ;
  m.df = l;
  return m;
}

struct SkatoParams {   // everything the integrands need (SkatO.cpp:303-337)
  double Qs_minP[11], taus[11], rhos[11];
  double MuQ, VarQ, VarZeta, Df, lam_sum;
  const double* lam;   // eigenvalues of Z(I-M)Z', kept, descending
  int n_lam;
};

template <class Par>
RVT_HDN double skato_integrand_davies(const SkatoParams& P, double x, int* th, const Par& par) {
  double kappa = 0.0;
  for (int i = 0; i < 11; ++i) {
    const double v = (P.Qs_minP[i] - P.taus[i] * x) / (1.0 - P.rhos[i]);
    if (i == 0 || v < kappa) kappa = v;
  }
  double temp;
  if (kappa > P.lam_sum * 10000) {
    temp = 0.0;
  } else {
    const double Q = (kappa - P.MuQ) * sqrt(P.VarQ - P.VarZeta) / sqrt(P.VarQ) + P.MuQ;
    int fault = 0;
    temp = mixchisq_pvalue(P.lam, P.n_lam, Q, th, &fault, par);
    if (temp <= 0.0 || temp == 1.0) temp = liu_pvalue(P.lam, P.n_lam, Q);
  }
  return (1.0 - temp) * chisq_pdf(x, 1.0);
}
RVT_HDN double skato_integrand_liu(const SkatoParams& P, double x) {
  double kappa = kDblMax;
  for (int i = 0; i < 11; ++i) {
    const double v = (P.Qs_minP[i] - P.taus[i] * x) / (1.0 - P.rhos[i]);
    if (v < kappa) kappa = v;
  }
  const double Q = (kappa - P.MuQ) / sqrt(P.VarQ) * sqrt(2.0 * P.Df) + P.Df;
  return chisq_p(Q, P.Df) * chisq_pdf(x, 1.0);
}

#if defined(__CUDACC__)
// One warp as a cooperating group: barriers are __syncwarp, reductions are shuffles.  Used to
// evaluate several quadrature nodes (one Davies evaluation each) concurrently inside a CTA.
struct WarpPar {
  __device__ __forceinline__ int tid() const { return threadIdx.x & 31; }
  __device__ __forceinline__ int nt() const { return 32; }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ void allreduce2(double& a, double& b) const {
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
  }
  __device__ __forceinline__ void allreduce4(double& a, double& b, double& c, double& d) const {
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
      d += __shfl_xor_sync(0xffffffffu, d, o);
    }
  }
};
#endif

// Cooperative QAGS of one of the two integrands over [0, 40].  `mach` and `fv` live in memory
// visible to the whole group; thread 0 drives the machine.  Returns the status to every thread.
template <class Par>
RVT_HDN int skato_integrate(const SkatoParams& P, bool use_davies, QagsMachine* mach, const QagsWork& work,
                            double* fv /*21*/, double* bcast /*3*/, int* th, int th_stride /* ints per warp, 0: one group */,
                            double* result, const Par& par) {
  if (par.tid() == 0) mach->init(work, 0.0, 40.0, 1e-25, 0.0001220703);
  par.sync();
  for (;;) {
    if (par.tid() == 0) {
#if defined(__CUDA_ARCH__)
      // watchdog (status 8): the machine is bounded by `limit` bisections of 42 budgeted Davies evaluations each, which
      // is finite but can be minutes; a gene that exceeds its cycle budget is reported, not waited for
      if (mach->stage != 3 && mach->w.deadline != 0 && clock64() > mach->w.deadline) {
        mach->status = 8;
        mach->stage = 3;
      }
#endif
      double lo = 0, hi = 0;
      const bool more = mach->want(&lo, &hi);
      bcast[0] = more ? 1.0 : 0.0;
      bcast[1] = lo;
      bcast[2] = hi;
    }
    par.sync();
    if (bcast[0] == 0.0) break;
    const double lo = bcast[1], hi = bcast[2];
    const double centre = 0.5 * (lo + hi), half = 0.5 * (hi - lo);
    if (use_davies) {
#if defined(__CUDA_ARCH__)
      if (th_stride > 0) {
        // the 21 Kronrod nodes are independent: each warp of the CTA takes nodes w, w+nw, ... and
        // runs its own Davies evaluation with warp-level reductions (no block barriers inside)
        const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        WarpPar wp;
        for (int i = w; i < 21; i += nw) {
          const double fx = skato_integrand_davies(P, centre + half * gk21_node(i), th + w * th_stride, wp);
          if (wp.tid() == 0) fv[i] = fx;
        }
      } else
#endif
      {
        for (int i = 0; i < 21; ++i) {   // every node is a group-wide Davies evaluation
          const double fx = skato_integrand_davies(P, centre + half * gk21_node(i), th, par);
          if (par.tid() == 0) fv[i] = fx;
        }
      }
    } else {
      for (int i = par.tid(); i < 21; i += par.nt()) fv[i] = skato_integrand_liu(P, centre + half * gk21_node(i));
    }
    par.sync();
    if (par.tid() == 0) mach->give(gk21_combine(fv, lo, hi));
    par.sync();
  }
  *result = mach->result;
  const int st = mach->status;
  par.sync();
  return st;
}

struct SkatoOut {
  double Q, rho, pvalue;
  int ok;
  int timed_out;       // the quadrature hit the device watchdog (QagsWork::deadline): ok = 0, record status RVT_GENE_TIMEOUT
};

// Wm: M x M (lda), symmetric, = Z1'Z1 (kept intact).  Km: M x M scratch (lda).  v[M] = w_j * (g_j'r).
// s2 = ||r||^2/(N-1).  ev/e/vv/pp: eigen scratch (>= M+2 each); lamz[M]: receives the ZIMZ spectrum.
// c[M]: scratch.  All scratch group-visible.  Every thread receives the same SkatoOut.
template <class Par>
RVT_HDN SkatoOut skato_tail(const double* Wm, double* Km, int M, int lda, const double* v, double s2, double* ev,
                            double* e, double* vv, double* pp, double* lamz, double* c, QagsMachine* mach,
                            const QagsWork& work, double* fv, double* bcast, int* th, int th_stride, const Par& par) {
  SkatoOut out;
  out.Q = 0;
  out.rho = 0;
  out.pvalue = -999.0;
  out.ok = 0;
  out.timed_out = 0;
  if (M == 1) {  // FitSKAT, SkatO.cpp:60-99 / :118-120
    const double Q = v[0] * v[0] / s2 / 2.0;
    const double lam1 = Wm[0];
    if (!(lam1 > 0.0)) return out;   // getEigen fails: numNonZero == 0
    int fault = 0;
    out.Q = Q;
    out.rho = 0.0;
    out.pvalue = mixchisq_pvalue(&lam1, 1, Q, th, &fault, par);
    out.ok = 1;
    return out;
  }
  // c = Wm 1, tot = 1'Wm 1
  for (int k = par.tid(); k < M; k += par.nt()) {
    double s = 0.0;
    for (int j = 0; j < M; ++j) s += Wm[j * lda + k];
    c[k] = s;
  }
  par.sync();
  double tot = 0.0, sv = 0.0, sv2 = 0.0, su2 = 0.0;
  for (int k = 0; k < M; ++k) {
    tot += c[k];
    sv += v[k];
    sv2 += v[k] * v[k];
    su2 += (c[k] / M) * (c[k] / M);
  }
  const double z_norm = tot / ((double)M * (double)M);
  SkatoParams P;
  SkatoMoment mom[11];
  double Qs[11], pvals[11];
  for (int i = 0; i < 11; ++i) {
    const double rho_o = (double)i / 10;
    const double rho = (rho_o > 0.999) ? 0.999 : rho_o;   // capRhos, SkatO.cpp:436-446
    P.rhos[i] = rho;
    Qs[i] = ((1.0 - rho) * sv2 + rho * sv * sv) / s2 / 2.0;
    const double a = sqrt(1.0 - rho), b = (sqrt(1.0 - rho + rho * M) - a) / M;
    for (int idx = par.tid(); idx < M * M; idx += par.nt()) {
      const int j = idx / M, k = idx - j * M;
      Km[j * lda + k] = a * a * Wm[j * lda + k] + a * b * (c[j] + c[k]) + b * b * tot;
    }
    par.sync();
    sym_eigenvalues_tridiag(Km, M, lda, ev, e, vv, pp, lamz, par);
    const int keep = skato_keep(lamz, M);
    if (keep < 0) return out;
    mom[i] = skato_moment(lamz, keep);
    P.taus[i] = (double)M * (double)M * rho * z_norm + (1.0 - rho) * su2 / z_norm;
    par.sync();
  }
  // Z(I-M)Z' = Wm - (c c')/(M^2 z_norm)
  double vz_part = 0.0, dummy = 0.0;
  for (int idx = par.tid(); idx < M * M; idx += par.nt()) {
    const int j = idx / M, k = idx - j * M;
    const double zmz = (c[j] / M) * (c[k] / M) / z_norm;
    const double zimz = Wm[j * lda + k] - zmz;
    Km[j * lda + k] = zimz;
    vz_part += zmz * zimz;
  }
  par.allreduce2(vz_part, dummy);
  sym_eigenvalues_tridiag(Km, M, lda, ev, e, vv, pp, lamz, par);
  const int nl = skato_keep(lamz, M);
  if (nl < 0) return out;
  P.lam = lamz;
  P.n_lam = nl;
  P.VarZeta = 4.0 * vz_part;
  double l1 = 0, l2 = 0, l4 = 0;
  for (int i = 0; i < nl; ++i) {
    const double l = lamz[i];
    l1 += l;
    l2 += l * l;
    l4 += (l * l) * (l * l);
  }
  P.MuQ = l1;
  P.lam_sum = l1;
  P.VarQ = 2.0 * l2 + P.VarZeta;
  const double KerQ = l4 / l2 / l2 * 12;
  P.Df = 12 / KerQ;
  // per-rho p-values by moment matching, the minimum, and its quantiles (SkatO.cpp:206-233)
  int minIndex = 0;
  for (int i = 0; i < 11; ++i) {
    const double qn = (Qs[i] - mom[i].muQ) / sqrt(mom[i].varQ) * sqrt(2. * mom[i].df) + mom[i].df;
    pvals[i] = chisq_q(qn, mom[i].df);
  }
  double minP = pvals[0];
  for (int i = 1; i < 11; ++i)
    if (pvals[i] < minP) {
      minP = pvals[i];
      minIndex = i;
    }
  for (int i = 0; i < 11; ++i) {
    const double q_org = chisq_qinv(minP, mom[i].df);
    P.Qs_minP[i] = (q_org - mom[i].df) / sqrt(2. * mom[i].df) * sqrt(mom[i].varQ) + mom[i].muQ;
  }
  double integral = 0.0;
  int st = skato_integrate(P, true, mach, work, fv, bcast, th, th_stride, &integral, par);
  if (st) st = skato_integrate(P, false, mach, work, fv, bcast, th, th_stride, &integral, par);
  if (st == 8) {
    out.timed_out = 1;
    return out;
  }
  double pvalue = 1.0 - integral;
  // sanity rules, SkatO.cpp:262-277 (nRho = 11 -> multi = 3)
  if (pvalue <= 0) {
    const double p3 = minP * 3;
    if (pvalue < p3) pvalue = p3;
  }
  if (pvalue == 0.0) {
    pvalue = pvals[0];
    for (int i = 1; i < 11; ++i)
      if (pvals[i] > 0 && pvals[i] < pvalue) pvalue = pvals[i];
  }
  out.Q = Qs[minIndex];
  out.rho = (P.rhos[minIndex] >= 0.999) ? 1.0 : P.rhos[minIndex];   // uncapRhos, :447-455
  out.pvalue = pvalue;
  out.ok = 1;
  return out;
}

}  // namespace rvt
