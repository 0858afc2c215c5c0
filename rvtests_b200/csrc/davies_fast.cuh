// davies_fast.cuh -- Davies' algorithm (AS 155) for ONE thread, restructured for the SKAT-O quadrature.
//
// Same function as davies.cuh (regression/qfc.c:304-452 with noncen = 0, df = 1, sigma = 0, lim = 10000, acc = 1e-6,
// as called by MixtureChiSquare::getPvalue, regression/MixtureChiSquare.cpp:7-29) -- same decisions, same fault codes,
// same evaluation budget -- but built for the place where it is the dominant cost: SKAT-O integrates
// (1 - Davies(kappa(x))) over x with an adaptive 21-point rule (regression/SkatO.cpp:303-337), i.e. ~1 100 Davies
// evaluations per gene that all share ONE spectrum `lb` and differ only in the point c.  The kernel gives every
// quadrature node its own thread (skato_fast.cuh), so this version is serial, and it removes what made the serial
// form expensive:
//   * every O(r) inner loop of qfc.c evaluates one or two transcendentals per term (log1, log, atan).  Sums of logs are
//     logs of products and sums of arctangents are the argument of a product of complex numbers, so the loops here
//     multiply (a handful of FMAs per term, with an explicit exponent against over/underflow and a quarter-turn counter
//     that keeps the argument unwrapped) and take ONE log / atan per call.  Values agree with the term-by-term sums to
//     ~r ulp (tests/test_davies_fast.py: |dqf| <= 1e-12 against the reference build, identical fault codes).
//   * what does not depend on c is computed once per spectrum (DaviesPre): sd, lmax, lmin, mean, the order of |lb| that
//     cfe needs (qfc.c:107-125), and the first findu (qfc.c:372: utx and the budget ticks it consumed).
#pragma once
#include "davies.cuh"

// inner-loop variants (A/B-timed on the B200, profiles/r02_davies_variants.txt): 1 = terms in blocks with independent chains
#ifndef RVT_DF_ERRBD4
#define RVT_DF_ERRBD4 0
#endif
#ifndef RVT_DF_TRUNC2
#define RVT_DF_TRUNC2 1   // measured: 57.1 -> 51.1 ms per 2 500 genes (ERRBD4: 61.7, INT2: 64.3 -- register pressure)
#endif
#ifndef RVT_DF_INT2
#define RVT_DF_INT2 0
#endif
#ifndef RVT_DF_RCP
#define RVT_DF_RCP 0      // 1: errbd's reciprocals as rcp.approx + two Newton steps instead of an IEEE division
#endif

namespace rvt {

struct DaviesPre {
  double sd, lmax, lmin, mean, almx;
  double utx0;     // utx after the first findu(.5 * acc)
  int count0;      // budget ticks that findu consumed
  int over0;       // the budget ran out inside it (cannot happen with lim = 10000; kept for exactness)
  int degenerate;  // 1: sd == 0 (qf returns c > 0), 2: lmin == lmax == 0 (fault 3)
  int r;
};

namespace qff {
using qfd::exp1;
using qfd::kLog28;
using qfd::kPi;
using qfd::sq;

constexpr double kBig = 1e150, kSmall = 1e-150;
constexpr double kLnBig = 345.38776394910684;   // ln(1e150)

// running product m * 1e150^e of positive factors
struct Prod {
  double m;
  int e;
  RVT_HD void init() {
    m = 1.0;
    e = 0;
  }
  RVT_HD void mul(double f) {
    m *= f;
    if (m > kBig) {
      m *= kSmall;
      ++e;
    } else if (m < kSmall) {
      m *= kBig;
      --e;
    }
  }
  RVT_HD double ln() const { return log(m) + (double)e * kLnBig; }
};

// 1 / y to ~1 ulp without the IEEE division's special-case path (device; the host build divides)
RVT_HD double rcp_fast(double y) {
#if defined(__CUDA_ARCH__) && RVT_DF_RCP
  if (!(fabs(y) > 1e-290 && fabs(y) < 1e290)) return 1.0 / y;
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
  double e = fma(-y, r, 1.0);
  r = fma(r, e, r);
  e = fma(-y, r, 1.0);
  r = fma(r, e, r);
  e = fma(-y, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / y;
#endif
}

struct St {
  double sigsq, lmax, lmin, mean, c;
  double intl, ersm;
  int count, r, lim;
  bool fail, over;
  const double* lb;
  const int* th;   // order of |lb|, descending (precomputed)
};

RVT_HD bool tick(St& s) {
  s.count = s.count + 1;
  if (s.count > s.lim) s.over = true;
  return s.over;
}

#if RVT_DF_ERRBD4
// qfc.c:128-147.  sum1 = u^2 sigsq + sum_j [x_j^2 / y_j + log(1 - x_j) + x_j],  x_j = 2 u lb_j, y_j = 1 - x_j.
// Terms go four at a time: the four reciprocals are independent instruction chains (the kernel is latency-bound: a few
// warps per scheduler, each thread a serial Davies evaluation), and the product takes one range check per block.
RVT_HD double errbd(St& s, double u, double* cx) {
  if (tick(s)) return 0.0;
  double xconst = u * s.sigsq;
  double sum1 = u * xconst;
  u = 2.0 * u;
  Prod p;
  p.init();
  double sx = 0.0;
  bool neg = false;
  int j = s.r - 1;
  for (; j >= 3; j -= 4) {
    const double l0 = s.lb[j], l1 = s.lb[j - 1], l2 = s.lb[j - 2], l3 = s.lb[j - 3];
    const double x0 = u * l0, x1 = u * l1, x2 = u * l2, x3 = u * l3;
    const double y0 = 1.0 - x0, y1 = 1.0 - x1, y2 = 1.0 - x2, y3 = 1.0 - x3;
    const double i0 = 1.0 / y0, i1 = 1.0 / y1, i2 = 1.0 / y2, i3 = 1.0 / y3;
    xconst = xconst + l0 * i0;
    xconst = xconst + l1 * i1;
    xconst = xconst + l2 * i2;
    xconst = xconst + l3 * i3;
    sum1 = sum1 + sq(x0) * i0;
    sum1 = sum1 + sq(x1) * i1;
    sum1 = sum1 + sq(x2) * i2;
    sum1 = sum1 + sq(x3) * i3;
    sx += (x0 + x1) + (x2 + x3);
    neg |= !(y0 > 0.0) | !(y1 > 0.0) | !(y2 > 0.0) | !(y3 > 0.0);
    const double a0 = fabs(y0), a1 = fabs(y1), a2 = fabs(y2), a3 = fabs(y3);
    if (fmax(fmax(a0, a1), fmax(a2, a3)) < 1e30 && fmin(fmin(a0, a1), fmin(a2, a3)) > 1e-30)
      p.mul((a0 * a1) * (a2 * a3));   // |block| in (1e-120, 1e120): one range check
    else {
      p.mul(a0);
      p.mul(a1);
      p.mul(a2);
      p.mul(a3);
    }
  }
  for (; j >= 0; j--) {
    const double lj = s.lb[j];
    const double x = u * lj, y = 1.0 - x;
    const double inv = 1.0 / y;
    xconst = xconst + lj * inv;
    sum1 = sum1 + sq(x) * inv;
    sx += x;
    neg |= !(y > 0.0);
    p.mul(fabs(y));
  }
  // (a non-positive y makes the reference's log NaN; keep that outcome)
  sum1 = sum1 + (sx + (neg ? nan("") : p.ln()));
  *cx = xconst;
  return exp1(-0.5 * sum1);
}

#else
// qfc.c:128-147.  sum1 = u^2 sigsq + sum_j [x_j^2 / y_j + log(1 - x_j) + x_j],  x_j = 2 u lb_j, y_j = 1 - x_j
RVT_HD double errbd(St& s, double u, double* cx) {
  if (tick(s)) return 0.0;
  double xconst = u * s.sigsq;
  double sum1 = u * xconst;
  u = 2.0 * u;
  Prod p;
  p.init();
  double sx = 0.0;
  bool neg = false;
  for (int j = s.r - 1; j >= 0; j--) {
    const double lj = s.lb[j];
    const double x = u * lj, y = 1.0 - x;
    const double inv = rcp_fast(y);
    xconst = xconst + lj * inv;
    sum1 = sum1 + sq(x) * inv;
    sx += x;
    neg |= !(y > 0.0);
    p.mul(fabs(y));
  }
  // (a non-positive y makes the reference's log NaN; keep that outcome)
  sum1 = sum1 + (sx + (neg ? nan("") : p.ln()));
  *cx = xconst;
  return exp1(-0.5 * sum1);
}

#endif
// qfc.c:149-174
RVT_HD double ctff(St& s, double accx, double* upn) {
  double u2 = *upn, u1 = 0.0, c1 = s.mean, c2 = 0.0, xconst = 0.0;
  const double rb = 2.0 * ((u2 > 0.0) ? s.lmax : s.lmin);
  for (;;) {
    const double u = u2 / (1.0 + u2 * rb);
    const double e = errbd(s, u, &c2);
    if (s.over) return 0.0;
    if (!(e > accx)) break;
    u1 = u2;
    c1 = c2;
    u2 = 2.0 * u2;
  }
  for (double u = (c1 - s.mean) / (c2 - s.mean); u < 0.9; u = (c1 - s.mean) / (c2 - s.mean)) {
    u = (u1 + u2) / 2.0;
    const double e = errbd(s, u / (1.0 + u * rb), &xconst);
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

// qfc.c:176-213: prod1 = 2 (sigsq + tausq) u^2 + sum_{x<=1} log(1+x), prod2 = prod1 + sum_{x>1} log x,
// prod3 = prod1 + sum_{x>1} log(1+x), x = (2 u lb_j)^2
RVT_HD double truncation(St& s, double u, double tausq) {
  if (tick(s)) return 0.0;
  const double sum2 = (s.sigsq + tausq) * sq(u);
  u = 2.0 * u;
#if RVT_DF_TRUNC2
  Prod p1, p2, p3;
  p1.init();
  p2.init();
  p3.init();
  int ss = 0;
  // branch-free: the threads of a warp share lb but not u, so (x > 1) differs from lane to lane
  int j = 0;
  for (; j + 1 < s.r; j += 2) {
    const double xa = sq(u * s.lb[j]), xb = sq(u * s.lb[j + 1]);
    const bool ba = xa > 1.0, bb = xb > 1.0;
    ss += (int)ba + (int)bb;
    if (xa < 1e60 && xb < 1e60) {
      p1.mul((ba ? 1.0 : 1.0 + xa) * (bb ? 1.0 : 1.0 + xb));
      p2.mul((ba ? xa : 1.0) * (bb ? xb : 1.0));
      p3.mul((ba ? 1.0 + xa : 1.0) * (bb ? 1.0 + xb : 1.0));
    } else {
      p1.mul(ba ? 1.0 : 1.0 + xa);
      p1.mul(bb ? 1.0 : 1.0 + xb);
      p2.mul(ba ? xa : 1.0);
      p2.mul(bb ? xb : 1.0);
      p3.mul(ba ? 1.0 + xa : 1.0);
      p3.mul(bb ? 1.0 + xb : 1.0);
    }
  }
  for (; j < s.r; j++) {
    const double x = sq(u * s.lb[j]);
    const bool b = x > 1.0;
    ss += (int)b;
    p1.mul(b ? 1.0 : 1.0 + x);
    p2.mul(b ? x : 1.0);
    p3.mul(b ? 1.0 + x : 1.0);
  }
#else
  Prod p1, p2, p3;
  p1.init();
  p2.init();
  p3.init();
  int ss = 0;
  for (int j = 0; j < s.r; j++) {
    const double x = sq(u * s.lb[j]);
    if (x > 1.0) {
      p2.mul(x);
      p3.mul(1.0 + x);
      ++ss;
    } else
      p1.mul(1.0 + x);
  }
#endif
  const double prod1 = 2.0 * sum2 + p1.ln();
  const double prod2 = prod1 + (ss ? p2.ln() : 0.0);
  const double prod3 = prod1 + (ss ? p3.ln() : 0.0);
  double x = exp1(-0.25 * prod2) / kPi;
  const double y = exp1(-0.25 * prod3) / kPi;
  double err1 = (ss == 0) ? 1.0 : x * 2.0 / ss;
  double err2 = (prod3 > 1.0) ? 2.5 * y : 1.0;
  if (err2 < err1) err1 = err2;
  x = 0.5 * sum2;
  err2 = (x <= y) ? 1.0 : y / x;
  return (err1 < err2) ? err1 : err2;
}

// qfc.c:215-234
RVT_HD void findu(St& s, double* utx, double accx) {
  const double divis[4] = {2.0, 1.4, 1.2, 1.1};
  double ut = *utx, u = ut / 4.0;
  double t = truncation(s, u, 0.0);
  if (s.over) return;
  if (t > accx) {
    for (;;) {
      u = ut;
      t = truncation(s, u, 0.0);
      if (s.over) return;
      if (!(t > accx)) break;
      ut = ut * 4.0;
    }
  } else {
    ut = u;
    for (;;) {
      u = u / 4.0;
      t = truncation(s, u, 0.0);
      if (s.over) return;
      if (!(t <= accx)) break;
      ut = u;
    }
  }
  for (int i = 0; i < 4; i++) {
    u = ut / divis[i];
    t = truncation(s, u, 0.0);
    if (s.over) return;
    if (t <= accx) ut = u;
  }
  *utx = ut;
}

// qfc.c:270-301 with the order of |lb| precomputed
RVT_HD double cfe(St& s, double x) {
  if (tick(s)) return 1.0;
  double axl = fabs(x), sum1 = 0.0;
  const double sxl = (x > 0.0) ? 1.0 : -1.0;
  for (int j = s.r - 1; j >= 0; j--) {
    const int t = s.th[j];
    if (s.lb[t] * sxl > 0.0) {
      const double lj = fabs(s.lb[t]);
      const double axl1 = axl - lj;
      const double axl2 = lj / kLog28;
      if (axl1 > axl2)
        axl = axl1;
      else {
        if (axl > axl2) axl = axl2;
        sum1 = (axl - axl1) / lj;
        for (int k = j - 1; k >= 0; k--) sum1 = sum1 + 1.0;
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

#if RVT_DF_INT2
// qfc.c:237-268.  Per term k: prod_j (1 + i x_j), x_j = 2 lb_j u, gives sum_j atan(x_j) as its (unwrapped) argument and
// sum_j log(1 + x_j^2) as the log of its squared modulus.  Positive and negative coefficients are kept in two products
// (the error sum needs sum_j |atan x_j|): `p` multiplies (1 + i x_j) for x_j >= 0, `n` multiplies (1 + i |x_j|) for
// x_j < 0 (its argument enters with a minus sign).  Each product is held in the first quadrant by a quarter turn back
// whenever it leaves it -- one factor turns it by less than pi/2 -- and the turns are counted.
struct Cprod {
  double r, i;
  int q, e;   // quarter turns taken back, exponent in units of 1e150
  RVT_HD void init() {
    r = 1.0;
    i = 0.0;
    q = 0;
    e = 0;
  }
  RVT_HD void mul(double x /* >= 0 */) {
    const double tr = r - i * x, ti = i + r * x;   // argument now in [0, pi)
    const bool turn = tr <= 0.0;                   // past pi/2: multiply by -i
    r = turn ? ti : tr;
    i = turn ? -tr : ti;
    q += (int)turn;
    if (r + i > kBig) {
      r *= kSmall;
      i *= kSmall;
      ++e;
    }
  }
  RVT_HD double arg() const { return (double)q * (0.5 * 3.14159265358979323846) + atan2(i, r); }
  RVT_HD double mod2() const { return r * r + i * i; }
};

RVT_HD void integrate_term(St& s, double u, const Cprod& p, const Cprod& n, double inpi, double tausq, bool mainx) {
  double sum1 = -2.0 * u * s.c;
  double sum2 = fabs(sum1);
  double sum3 = -0.5 * s.sigsq * sq(u);
  const double thp = p.arg(), thn = n.arg();   // both >= 0
  const double lmod2 = log(p.mod2() * n.mod2()) + 2.0 * (double)(p.e + n.e) * kLnBig;
  sum3 = sum3 - 0.25 * lmod2;
  sum1 = sum1 + (thp - thn);
  sum2 = sum2 + (thp + thn);
  double x = inpi * exp1(sum3) / u;
  if (!mainx) x = x * (1.0 - exp1(-0.5 * tausq * sq(u)));
  s.intl = s.intl + sin(0.5 * sum1) * x;
  s.ersm = s.ersm + 0.5 * sum2 * x;
}

// two trapezoid terms per pass over the coefficients: two independent product chains (and the sign branch on lb_j is
// the same for every thread of the warp); the sums are accumulated in the reference's order k = nterm .. 0
RVT_HD void integrate(St& s, int nterm, double interv, double tausq, bool mainx) {
  const double inpi = interv / kPi;
  int k = nterm;
  for (; k >= 1; k -= 2) {
    const double ua = (k + 0.5) * interv, ub = (k - 0.5) * interv;
    const double ua2 = 2.0 * ua, ub2 = 2.0 * ub;
    Cprod pa, na, pb, nb;
    pa.init();
    na.init();
    pb.init();
    nb.init();
    for (int j = s.r - 1; j >= 0; j--) {
      const double l = s.lb[j];
      if (l >= 0.0) {
        pa.mul(l * ua2);
        pb.mul(l * ub2);
      } else {
        na.mul(-l * ua2);
        nb.mul(-l * ub2);
      }
    }
    integrate_term(s, ua, pa, na, inpi, tausq, mainx);
    integrate_term(s, ub, pb, nb, inpi, tausq, mainx);
  }
  if (k == 0) {
    const double u = 0.5 * interv, u2 = 2.0 * u;
    Cprod p, n;
    p.init();
    n.init();
    for (int j = s.r - 1; j >= 0; j--) {
      const double l = s.lb[j];
      if (l >= 0.0)
        p.mul(l * u2);
      else
        n.mul(-l * u2);
    }
    integrate_term(s, u, p, n, inpi, tausq, mainx);
  }
}
#else
// qfc.c:237-268.  Per term k: prod_j (1 + i x_j), x_j = 2 lb_j u, gives sum_j atan(x_j) as its (unwrapped) argument and
// sum_j log(1 + x_j^2) as the log of its squared modulus.  Positive and negative coefficients are kept in two products
// (the error sum needs sum_j |atan x_j|); each is held in the first / fourth quadrant by a quarter turn whenever it
// leaves it -- one factor turns it by less than pi/2 -- and the turns are counted.
RVT_HD void integrate(St& s, int nterm, double interv, double tausq, bool mainx) {
  const double inpi = interv / kPi;
  for (int k = nterm; k >= 0; k--) {
    const double u = (k + 0.5) * interv;
    double sum1 = -2.0 * u * s.c;
    double sum2 = fabs(sum1);
    double sum3 = -0.5 * s.sigsq * sq(u);
    double pr = 1.0, pi = 0.0, nr = 1.0, ni = 0.0;   // positive- and negative-coefficient products
    int pq = 0, nq = 0, pe = 0, ne = 0;              // quarter turns, exponents (units of 1e150)
    const double u2 = 2.0 * u;
    for (int j = s.r - 1; j >= 0; j--) {
      const double x = s.lb[j] * u2;
      if (x >= 0.0) {
        const double tr = pr - pi * x, ti = pi + pr * x;   // angle in [0, pi)
        if (tr <= 0.0) {                                   // past pi/2: turn back by a quarter (multiply by -i)
          pr = ti;
          pi = -tr;
          ++pq;
        } else {
          pr = tr;
          pi = ti;
        }
        if (pr + pi > kBig) {
          pr *= kSmall;
          pi *= kSmall;
          ++pe;
        }
      } else {
        const double tr = nr - ni * x, ti = ni + nr * x;   // angle in (-pi, 0]
        if (tr <= 0.0) {                                   // multiply by +i
          nr = -ti;
          ni = tr;
          ++nq;
        } else {
          nr = tr;
          ni = ti;
        }
        if (nr - ni > kBig) {
          nr *= kSmall;
          ni *= kSmall;
          ++ne;
        }
      }
    }
    const double thp = (double)pq * (0.5 * 3.14159265358979323846) + atan2(pi, pr);    // >= 0
    const double thn = -(double)nq * (0.5 * 3.14159265358979323846) + atan2(ni, nr);   // <= 0
    const double lmod2 = log((pr * pr + pi * pi) * (nr * nr + ni * ni)) + 2.0 * (double)(pe + ne) * kLnBig;
    sum3 = sum3 - 0.25 * lmod2;
    sum1 = sum1 + (thp + thn);
    sum2 = sum2 + (thp - thn);
    double x = inpi * exp1(sum3) / u;
    if (!mainx) x = x * (1.0 - exp1(-0.5 * tausq * sq(u)));
    s.intl = s.intl + sin(0.5 * sum1) * x;
    s.ersm = s.ersm + 0.5 * sum2 * x;
  }
}
#endif
}  // namespace qff

// what qf() computes before it looks at c, once per spectrum.  th: r ints, receives the order of |lb| (descending)
RVT_HDN void davies_prepare(const double* lb, int r, int lim, double acc, int* th, DaviesPre* pre) {
  using namespace qff;
  pre->r = r;
  pre->degenerate = 0;
  pre->over0 = 0;
  pre->count0 = 0;
  pre->utx0 = 0.0;
  double sd = 0.0, lmax = 0.0, lmin = 0.0, mean = 0.0;
  for (int j = 0; j < r; j++) {
    const double lj = lb[j];
    sd = sd + sq(lj) * 2.0;
    mean = mean + lj;
    if (lmax < lj)
      lmax = lj;
    else if (lmin > lj)
      lmin = lj;
  }
  pre->lmax = lmax;
  pre->lmin = lmin;
  pre->mean = mean;
  if (sd == 0.0) {
    pre->sd = 0.0;
    pre->almx = 0.0;
    pre->degenerate = 1;
    return;
  }
  if (lmin == 0.0 && lmax == 0.0) {
    pre->degenerate = 2;
    return;
  }
  sd = sqrt(sd);
  pre->sd = sd;
  pre->almx = (lmax < -lmin) ? -lmin : lmax;
  // qfc.c:107-125 (insertion order of |lb|, descending, stable)
  for (int j = 0; j < r; j++) {
    const double lj = fabs(lb[j]);
    int k = j - 1;
    for (; k >= 0; k--) {
      if (lj > fabs(lb[th[k]]))
        th[k + 1] = th[k];
      else
        break;
    }
    th[k + 1] = j;
  }
  St s;
  s.r = r;
  s.lim = lim;
  s.c = 0.0;
  s.lb = lb;
  s.th = th;
  s.count = 0;
  s.intl = s.ersm = 0.0;
  s.fail = s.over = false;
  s.sigsq = 0.0;
  s.lmax = lmax;
  s.lmin = lmin;
  s.mean = mean;
  double utx = 16.0 / sd;
  findu(s, &utx, .5 * acc);
  pre->utx0 = utx;
  pre->count0 = s.count;
  pre->over0 = s.over ? 1 : 0;
}

// qf() from the prepared state: P(sum lb_j chi2_1 < c1); *ifault as qfc.c:304-325 (4 = budget exhausted)
RVT_HDN double davies_qf_fast(const double* lb, const DaviesPre& pre, const int* th, double c1, int lim1, double acc, int* ifault) {
  using namespace qff;
  *ifault = 0;
  double qfval = -1.0, acc1 = acc;
  if (pre.degenerate == 1) return (c1 > 0.0) ? 1.0 : 0.0;
  if (pre.degenerate == 2) {
    *ifault = 3;
    return qfval;
  }
  St s;
  s.r = pre.r;
  s.lim = lim1;
  s.c = c1;
  s.lb = lb;
  s.th = th;
  s.count = pre.count0;
  s.intl = 0.0;
  s.ersm = 0.0;
  s.fail = false;
  s.over = pre.over0 != 0;
  s.sigsq = 0.0;
  s.lmax = pre.lmax;
  s.lmin = pre.lmin;
  s.mean = pre.mean;
  double xlim = (double)s.lim;
  const double sd = pre.sd, almx = pre.almx;
  double utx = pre.utx0, up = 4.5 / sd, un = -up;
  double tausq, intv = 0.0, xnt = 0.0;
  if (s.over) goto budget;
  if (s.c != 0.0 && (almx > 0.07 * sd)) {
    const double cf = cfe(s, s.c);
    if (s.over) goto budget;
    tausq = .25 * acc1 / cf;
    if (s.fail)
      s.fail = false;
    else {
      const double t = truncation(s, utx, tausq);
      if (s.over) goto budget;
      if (t < .2 * acc1) {
        s.sigsq = s.sigsq + tausq;
        findu(s, &utx, .25 * acc1);
        if (s.over) goto budget;
      }
    }
  }
  acc1 = 0.5 * acc1;

  for (;;) {
    double d1 = ctff(s, acc1, &up);
    if (s.over) goto budget;
    d1 = d1 - s.c;
    if (d1 < 0.0) return 1.0;
    double d2 = ctff(s, acc1, &un);
    if (s.over) goto budget;
    d2 = s.c - d2;
    if (d2 < 0.0) return 0.0;
    intv = 2.0 * kPi / ((d1 > d2) ? d1 : d2);
    xnt = utx / intv;
    const double xntm = 3.0 / sqrt(acc1);
    if (!(xnt > xntm * 1.5)) break;
    if (xntm > xlim) {
      *ifault = 1;
      return qfval;
    }
    const int ntm = qfd::to_int_x86(floor(xntm + 0.5));
    const double intv1 = utx / ntm;
    const double x = 2.0 * kPi / intv1;
    if (x <= fabs(s.c)) break;
    const double cf1 = cfe(s, s.c - x);
    if (s.over) goto budget;
    const double cf2 = cfe(s, s.c + x);
    if (s.over) goto budget;
    tausq = .33 * acc1 / (1.1 * (cf1 + cf2));
    if (s.fail) break;
    acc1 = .67 * acc1;
    integrate(s, ntm, intv1, tausq, false);
    xlim = xlim - xntm;
    s.sigsq = s.sigsq + tausq;
    findu(s, &utx, .25 * acc1);
    if (s.over) goto budget;
    acc1 = 0.75 * acc1;
  }

  if (xnt > xlim) {
    *ifault = 1;
    return qfval;
  }
  {
    const int nt = qfd::to_int_x86(floor(xnt + 0.5));
    integrate(s, nt, intv, 0.0, true);
    qfval = 0.5 - s.intl;
    const double upv = s.ersm, x = upv + acc / 10.0;
    const int rats[4] = {1, 2, 4, 8};
    for (int j = 0; j < 4; j++)
      if (rats[j] * x == rats[j] * upv) *ifault = 2;
  }
  return qfval;

budget:
  *ifault = 4;
  return qfval;
}

// MixtureChiSquare::getPvalue on the prepared spectrum (n >= 2; a single lambda goes to Liu before this is reached)
RVT_HDN double mixchisq_pvalue_fast(const double* lambda, const DaviesPre& pre, const int* th, double Q, int* fault) {
  *fault = 0;
  double p = 1.0 - davies_qf_fast(lambda, pre, th, Q, 10000, 0.000001, fault);
  if (p > 1.0) p = 1.0;
  if (*fault) p = -1.0;
  return p;
}

}  // namespace rvt
