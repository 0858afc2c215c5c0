// davies.cuh -- Davies' algorithm (AS 155) for P(sum_j lambda_j chi2_1 > Q), cooperative version.
//
// Replaces regression/qfc.c:304-452 (qf) and its helpers (:77-301) as called by
// MixtureChiSquare::getPvalue (regression/MixtureChiSquare.cpp:7-29) with
//   noncen = 0, df = 1 for every term, sigma = 0, lim = 10000, acc = 1e-6
// (regression/MixtureChiSquare.h:7,33-35).
//
// Design (not a transcription): the reference keeps its state in file-scope statics and unwinds
// with longjmp when the evaluation budget `lim` is exhausted (qfc.c:26-30,77-83); that cannot run
// many genes at once.  Here the state is a per-call struct, budget exhaustion is a sticky flag
// that every loop checks, and the only O(terms x r) part -- the trapezoid sum `integrate`
// (qfc.c:237-268) -- is spread over the cooperating threads of a `Par` group (a CTA on the GPU,
// a single thread in the CPU host-check build) and combined with one all-reduce.  All scalar
// control flow is evaluated redundantly and identically by every thread of the group, so branches
// stay uniform and the fault codes match the serial algorithm.  Constants are the reference's:
// pi = 3.14159265358979, log28 = .0866 (qfc.c:23-24), exp cut-off -50 (:35-36).
#pragma once
#include "mathdev.cuh"

namespace rvt {

// Single-thread "group": used by the host-check build and by per-thread device calls.
struct SerialPar {
  RVT_HD int tid() const { return 0; }
  RVT_HD int nt() const { return 1; }
  RVT_HD void sync() const {}
  RVT_HD void allreduce2(double&, double&) const {}
  RVT_HD void allreduce4(double&, double&, double&, double&) const {}
};

struct QfState {
  double sigsq, lmax, lmin, mean, c;
  double intl, ersm;
  int count, r, lim;
  bool ndtsrt, fail, over;
  const double* lb;  // r coefficients
  int* th;           // r ints of scratch (order of |lb|, descending)
};

namespace qfd {
constexpr double kPi = 3.14159265358979;
constexpr double kLog28 = .0866;

RVT_HD double exp1(double x) { return x < -50.0 ? 0.0 : exp(x); }
RVT_HD double sq(double x) { return x * x; }

// qfc.c:89-105.  first ? log(1+x) : log(1+x)-x
RVT_HD double log1(double x, bool first) {
  if (fabs(x) > 0.1) return first ? log(1.0 + x) : (log(1.0 + x) - x);
  double y = x / (2.0 + x);
  double term = 2.0 * y * y * y;
  double k = 3.0;
  double s = (first ? 2.0 : -x) * y;
  y = y * y;
  // |y^2| <= (0.1/1.9)^2: the series is down to one ulp after < 12 terms.  The cap makes a NaN argument (s1 != s
  // for ever) return NaN instead of spinning -- on the device that spin was a hang.
  int guard = 0;
  for (double s1 = s + term / k; s1 != s && guard < 64; s1 = s + term / k, ++guard) {
    k = k + 2.0;
    term = term * y;
    s = s1;
  }
  return s;
}

// (int) of a double the way the reference's x86 build converts it (cvttsd2si): NaN and values outside int range give
// INT_MIN, where the GPU's cvt.rzi would give 0 / saturate.  qf() reaches this with xnt = NaN when its argument c is NaN:
// INT_MIN terms = an empty trapezoid sum on the CPU; 0 on the device meant ONE term evaluated at u = NaN.
RVT_HD int to_int_x86(double x) {
  if (!(x > -2147483649.0 && x < 2147483648.0)) return (-2147483647 - 1);
  return (int)x;
}

// budget tick (qfc.c:77-83); returns true when the budget is exhausted
RVT_HD bool tick(QfState& s) {
  s.count = s.count + 1;
  if (s.count > s.lim) s.over = true;
  return s.over;
}

// qfc.c:128-147 with n_j = 1, nc_j = 0.  The r terms are dealt to the threads of the group
// (serial group: j = r-1..0, the reference's order) and combined with one all-reduce.
template <class Par>
RVT_HD double errbd(QfState& s, double u, double* cx, const Par& par) {
  if (tick(s)) return 0.0;
  const double ncj = 0.0;
  const double xc0 = u * s.sigsq;
  const double s10 = u * xc0;
  u = 2.0 * u;
  double xconst = (par.tid() == 0) ? xc0 : 0.0;
  double sum1 = (par.tid() == 0) ? s10 : 0.0;
  for (int j = s.r - 1 - par.tid(); j >= 0; j -= par.nt()) {
    double lj = s.lb[j];
    double x = u * lj, y = 1.0 - x;
    xconst = xconst + lj * (ncj / y + 1) / y;
    sum1 = sum1 + ncj * sq(x / y) + (sq(x) / y + log1(-x, false));
  }
  par.allreduce2(xconst, sum1);
  *cx = xconst;
  return exp1(-0.5 * sum1);
}

// qfc.c:149-174
template <class Par>
RVT_HD double ctff(QfState& s, double accx, double* upn, const Par& par) {
  double u2 = *upn, u1 = 0.0, c1 = s.mean, c2 = 0.0, xconst = 0.0;
  double rb = 2.0 * ((u2 > 0.0) ? s.lmax : s.lmin);
  for (;;) {
    double u = u2 / (1.0 + u2 * rb);
    double e = errbd(s, u, &c2, par);
    if (s.over) return 0.0;
    if (!(e > accx)) break;
    u1 = u2;
    c1 = c2;
    u2 = 2.0 * u2;
  }
  for (double u = (c1 - s.mean) / (c2 - s.mean); u < 0.9; u = (c1 - s.mean) / (c2 - s.mean)) {
    u = (u1 + u2) / 2.0;
    double e = errbd(s, u / (1.0 + u * rb), &xconst, par);
    if (s.over) return 0.0;
    if (e > accx) {
      u1 = u;
      c1 = xconst;
    } else {
      u2 = u;
      c2 = xconst;
    }
  }
  *upn = u2;
  return c2;
}

// qfc.c:176-213 with n_j = 1, nc_j = 0; terms dealt to the group like errbd (serial: j = 0..r-1)
template <class Par>
RVT_HD double truncation(QfState& s, double u, double tausq, const Par& par) {
  if (tick(s)) return 0.0;
  const double ncj = 0.0;
  double sum1 = 0.0, prod2 = 0.0, prod3 = 0.0;
  double ssd = 0.0;
  const double sum2 = (s.sigsq + tausq) * sq(u);
  double prod1 = (par.tid() == 0) ? 2.0 * sum2 : 0.0;
  u = 2.0 * u;
  for (int j = par.tid(); j < s.r; j += par.nt()) {
    double x = sq(u * s.lb[j]);
    sum1 = sum1 + ncj * x / (1.0 + x);
    if (x > 1.0) {
      prod2 = prod2 + log(x);
      prod3 = prod3 + log1(x, true);
      ssd = ssd + 1.0;
    } else
      prod1 = prod1 + log1(x, true);
  }
  par.allreduce4(prod1, prod2, prod3, ssd);
  const int ss = (int)ssd;
  sum1 = 0.5 * sum1;  // identically 0 on this path (nc_j = 0)
  prod2 = prod1 + prod2;
  prod3 = prod1 + prod3;
  double x = exp1(-sum1 - 0.25 * prod2) / kPi;
  double y = exp1(-sum1 - 0.25 * prod3) / kPi;
  double err1 = (ss == 0) ? 1.0 : x * 2.0 / ss;
  double err2 = (prod3 > 1.0) ? 2.5 * y : 1.0;
  if (err2 < err1) err1 = err2;
  x = 0.5 * sum2;
  err2 = (x <= y) ? 1.0 : y / x;
  return (err1 < err2) ? err1 : err2;
}

// qfc.c:215-234
template <class Par>
RVT_HD void findu(QfState& s, double* utx, double accx, const Par& par) {
  const double divis[4] = {2.0, 1.4, 1.2, 1.1};
  double ut = *utx, u = ut / 4.0;
  double t = truncation(s, u, 0.0, par);
  if (s.over) return;
  if (t > accx) {
    for (;;) {
      u = ut;
      t = truncation(s, u, 0.0, par);
      if (s.over) return;
      if (!(t > accx)) break;
      ut = ut * 4.0;
    }
  } else {
    ut = u;
    for (;;) {
      u = u / 4.0;
      t = truncation(s, u, 0.0, par);
      if (s.over) return;
      if (!(t <= accx)) break;
      ut = u;
    }
  }
  for (int i = 0; i < 4; i++) {
    u = ut / divis[i];
    t = truncation(s, u, 0.0, par);
    if (s.over) return;
    if (t <= accx) ut = u;
  }
  *utx = ut;
}

// qfc.c:107-125 -- insertion order of |lb|, descending.  One thread writes, the group syncs.
template <class Par>
RVT_HD void order(QfState& s, const Par& par) {
  if (par.tid() == 0) {
    for (int j = 0; j < s.r; j++) {
      double lj = fabs(s.lb[j]);
      int k = j - 1;
      for (; k >= 0; k--) {
        if (lj > fabs(s.lb[s.th[k]]))
          s.th[k + 1] = s.th[k];
        else
          break;
      }
      s.th[k + 1] = j;
    }
  }
  par.sync();
  s.ndtsrt = false;
}

// qfc.c:270-301 with n_j = 1, nc_j = 0
template <class Par>
RVT_HD double cfe(QfState& s, double x, const Par& par) {
  if (tick(s)) return 1.0;
  if (s.ndtsrt) order(s, par);
  double axl = fabs(x), sxl = (x > 0.0) ? 1.0 : -1.0, sum1 = 0.0;
  for (int j = s.r - 1; j >= 0; j--) {
    int t = s.th[j];
    if (s.lb[t] * sxl > 0.0) {
      double lj = fabs(s.lb[t]);
      double axl1 = axl - lj * (1 + 0.0);
      double axl2 = lj / kLog28;
      if (axl1 > axl2)
        axl = axl1;
      else {
        if (axl > axl2) axl = axl2;
        sum1 = (axl - axl1) / lj;
        for (int k = j - 1; k >= 0; k--) sum1 = sum1 + (1 + 0.0);
        break;
      }
    }
  }
  if (sum1 > 100.0) {
    s.fail = true;
    return 1.0;
  }
  return pow(2.0, (sum1 / 4.0)) / (kPi * sq(axl));
}

// qfc.c:237-268, terms k = nterm..0 dealt round-robin to the threads of the group.
template <class Par>
RVT_HD void integrate(QfState& s, int nterm, double interv, double tausq, bool mainx,
                      const Par& par) {
  const double ncj = 0.0;
  double inpi = interv / kPi;
  double a_intl = 0.0, a_ersm = 0.0;
  if (nterm < 0) nterm = -1;   // (INT_MIN from to_int_x86: no term; and nterm - tid must not wrap)
  for (int k = nterm - par.tid(); k >= 0; k -= par.nt()) {
    double u = (k + 0.5) * interv;
    double sum1 = -2.0 * u * s.c;
    double sum2 = fabs(sum1);
    double sum3 = -0.5 * s.sigsq * sq(u);
    for (int j = s.r - 1; j >= 0; j--) {
      double x = 2.0 * s.lb[j] * u;
      double y = sq(x);
      sum3 = sum3 - 0.25 * log1(y, true);
      y = ncj * x / (1.0 + y);
      double z = atan(x) + y;
      sum1 = sum1 + z;
      sum2 = sum2 + fabs(z);
      sum3 = sum3 - 0.5 * x * y;
    }
    double x = inpi * exp1(sum3) / u;
    if (!mainx) x = x * (1.0 - exp1(-0.5 * tausq * sq(u)));
    a_intl += sin(0.5 * sum1) * x;
    a_ersm += 0.5 * sum2 * x;
  }
  par.allreduce2(a_intl, a_ersm);
  s.intl = s.intl + a_intl;
  s.ersm = s.ersm + a_ersm;
}
}  // namespace qfd

// qf(): returns P(sum lambda_j chi2_1 < c); *ifault as qfc.c:304-325 (4 = budget exhausted).
// th: r ints of scratch visible to the whole group.
template <class Par>
RVT_HDN double davies_qf(const double* lb, int r, double c1, int lim1, double acc, int* th,
                         int* ifault, const Par& par) {
  using namespace qfd;
  QfState s;
  s.r = r;
  s.lim = lim1;
  s.c = c1;
  s.lb = lb;
  s.th = th;
  s.count = 0;
  s.intl = 0.0;
  s.ersm = 0.0;
  s.ndtsrt = true;
  s.fail = false;
  s.over = false;
  *ifault = 0;
  double qfval = -1.0, acc1 = acc;
  double xlim = (double)s.lim;
  const double sigma = 0.0;
  s.sigsq = sq(sigma);
  double sd = s.sigsq;
  s.lmax = 0.0;
  s.lmin = 0.0;
  s.mean = 0.0;
  for (int j = 0; j < r; j++) {
    double lj = lb[j];
    sd = sd + sq(lj) * (2 * 1 + 4.0 * 0.0);
    s.mean = s.mean + lj * (1 + 0.0);
    if (s.lmax < lj)
      s.lmax = lj;
    else if (s.lmin > lj)
      s.lmin = lj;
  }
  if (sd == 0.0) return (s.c > 0.0) ? 1.0 : 0.0;
  if (s.lmin == 0.0 && s.lmax == 0.0 && sigma == 0.0) {
    *ifault = 3;
    return qfval;
  }
  sd = sqrt(sd);
  double almx = (s.lmax < -s.lmin) ? -s.lmin : s.lmax;

  double utx = 16.0 / sd, up = 4.5 / sd, un = -up;
  double tausq, intv = 0.0, xnt = 0.0;
  findu(s, &utx, .5 * acc1, par);
  if (s.over) goto budget;
  if (s.c != 0.0 && (almx > 0.07 * sd)) {
    double cf = cfe(s, s.c, par);
    if (s.over) goto budget;
    tausq = .25 * acc1 / cf;
    if (s.fail)
      s.fail = false;
    else {
      double t = truncation(s, utx, tausq, par);
      if (s.over) goto budget;
      if (t < .2 * acc1) {
        s.sigsq = s.sigsq + tausq;
        findu(s, &utx, .25 * acc1, par);
        if (s.over) goto budget;
      }
    }
  }
  acc1 = 0.5 * acc1;

  for (;;) {  // label l1 of qfc.c
    double d1 = ctff(s, acc1, &up, par);
    if (s.over) goto budget;
    d1 = d1 - s.c;
    if (d1 < 0.0) return 1.0;
    double d2 = ctff(s, acc1, &un, par);
    if (s.over) goto budget;
    d2 = s.c - d2;
    if (d2 < 0.0) return 0.0;
    intv = 2.0 * kPi / ((d1 > d2) ? d1 : d2);
    xnt = utx / intv;
    double xntm = 3.0 / sqrt(acc1);
    if (!(xnt > xntm * 1.5)) break;
    // auxiliary integration
    if (xntm > xlim) {
      *ifault = 1;
      return qfval;
    }
    int ntm = to_int_x86(floor(xntm + 0.5));
    double intv1 = utx / ntm;
    double x = 2.0 * kPi / intv1;
    if (x <= fabs(s.c)) break;
    double cf1 = cfe(s, s.c - x, par);
    if (s.over) goto budget;
    double cf2 = cfe(s, s.c + x, par);
    if (s.over) goto budget;
    tausq = .33 * acc1 / (1.1 * (cf1 + cf2));
    if (s.fail) break;
    acc1 = .67 * acc1;
    integrate(s, ntm, intv1, tausq, false, par);
    xlim = xlim - xntm;
    s.sigsq = s.sigsq + tausq;
    findu(s, &utx, .25 * acc1, par);
    if (s.over) goto budget;
    acc1 = 0.75 * acc1;
  }

  // main integration (label l2)
  if (xnt > xlim) {
    *ifault = 1;
    return qfval;
  }
  {
    int nt = to_int_x86(floor(xnt + 0.5));
    integrate(s, nt, intv, 0.0, true, par);
    qfval = 0.5 - s.intl;
    // round-off test, radix 8/16 allowance (qfc.c:444-446)
    double upv = s.ersm, x = upv + acc / 10.0;
    const int rats[4] = {1, 2, 4, 8};
    for (int j = 0; j < 4; j++)
      if (rats[j] * x == rats[j] * upv) *ifault = 2;
  }
  return qfval;

budget:
  *ifault = 4;
  return qfval;
}

// MixtureChiSquare::getPvalue (regression/MixtureChiSquare.cpp:7-29): Davies with lim=10000,
// acc=1e-6; a single lambda goes to Liu; any fault returns -1.  *fault receives qf's ifault.
template <class Par>
RVT_HDN double mixchisq_pvalue(const double* lambda, int n, double Q, int* th, int* fault,
                               const Par& par) {
  *fault = 0;
  if (n == 1) return liu_pvalue(lambda, n, Q);
  double p = 1.0 - davies_qf(lambda, n, Q, 10000, 0.000001, th, fault, par);
  if (p > 1.0) p = 1.0;
  if (*fault) p = -1.0;
  return p;
}

}  // namespace rvt
