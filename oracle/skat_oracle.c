/*
 * oracle/skat_oracle.c  --  TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement, in plain C / fp64, of the rvtests per-gene association hot path
 * (SKAT, CMC, Zeggini, the linear null model and the linear score test).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.  The product (librvtests_b200.so) never does.
 *
 * Each function cites the reference file:line it restates (paths relative to the
 * upstream tree, zhanxw/rvtests @ 8defd6f).  The reference's own linear algebra is
 * Eigen 3.3.9, which is NOT vendored upstream (third/Makefile:82-84 downloads it), so
 * those call sites are restated with small hand-written Cholesky / Jacobi routines.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - Davies / Liu p-values: pinned against the reference's own MixtureChiSquare.cpp +
 *     qfc.c + cdflib.cpp compiled in place into oracle/_ref/ (tests/test_oracle_pin.py)
 *     and against the three known-answer vectors of regression/test/testMixtureChiSquare.cpp.
 *   - Q statistics / burden U, V, p / null model / permutation loop: the reference ships no golden
 *     vector, so its OWN sources (Skat.cpp, LinearRegression*.cpp, Permutation.h ...)
 *     are compiled unmodified against oracle/eigen_standin into oracle/_ref/libskat_ref.so and run
 *     beside this file (tests/test_oracle_pin_reference_skat.py, tests/golden/ref_skat_golden.npz);
 *     also (i) the literal float32 N x N restatement orc_skat_faithful32() of
 *     regression/Skat.cpp:29-105 agreeing with the reduced fp64 algebra, (ii) the C1 example anchor.
 *   - flip-to-minor / collapse (src/DataConsolidator.cpp, src/Model.cpp): cannot be compiled in
 *     isolation => "parity unpinned" for those integer steps.
 */
#include <math.h>
#include <setjmp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * small dense helpers (stand-ins for Eigen LLT / SelfAdjointEigenSolver)
 * ---------------------------------------------------------------------------------------- */

/* in-place Cholesky A = L L^T of an n x n row-major SPD matrix; returns 0 on success. */
static int chol_decomp(int n, double* a) {
  for (int j = 0; j < n; ++j) {
    double d = a[j * n + j];
    for (int k = 0; k < j; ++k) d -= a[j * n + k] * a[j * n + k];
    if (!(d > 0.0)) return -1;
    d = sqrt(d);
    a[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = a[i * n + j];
      for (int k = 0; k < j; ++k) s -= a[i * n + k] * a[j * n + k];
      a[i * n + j] = s / d;
    }
  }
  return 0;
}

/* solve L L^T x = b in place (L from chol_decomp, lower triangle of a). */
static void chol_solve(int n, const double* l, double* b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= l[i * n + k] * b[k];
    b[i] = s / l[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= l[k * n + i] * b[k];
    b[i] = s / l[i * n + i];
  }
}

/* inverse of SPD matrix (row-major n x n) via Cholesky: `.llt().solve(Identity)`
 * (regression/LinearRegression.cpp:33-35, LinearRegressionScoreTest.cpp:229). */
static int spd_inverse(int n, const double* a, double* inv) {
  double* l = (double*)malloc(sizeof(double) * n * n);
  double* col = (double*)malloc(sizeof(double) * n);
  memcpy(l, a, sizeof(double) * n * n);
  if (chol_decomp(n, l)) {
    free(l);
    free(col);
    return -1;
  }
  for (int j = 0; j < n; ++j) {
    for (int i = 0; i < n; ++i) col[i] = (i == j) ? 1.0 : 0.0;
    chol_solve(n, l, col);
    for (int i = 0; i < n; ++i) inv[i * n + j] = col[i];
  }
  free(l);
  free(col);
  return 0;
}

/* cyclic Jacobi eigenvalues of a symmetric n x n matrix (row-major, destroyed).
 * Stand-in for Eigen::SelfAdjointEigenSolver (regression/Skat.cpp:75-76,
 * regression/SkatO.cpp:350-352).  Output ascending, like Eigen. */
static int cmp_dbl(const void* a, const void* b) {
  double x = *(const double*)a, y = *(const double*)b;
  return (x > y) - (x < y);
}
ORC_API void orc_sym_eigenvalues(int n, double* a, double* ev) {
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) {
      diag += a[i * n + i] * a[i * n + i];
      for (int j = i + 1; j < n; ++j) off += a[i * n + j] * a[i * n + j];
    }
    if (off <= 1e-300 || off <= 1e-32 * diag) break;
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
        double apq = a[p * n + q];
        if (apq == 0.0) continue;
        double app = a[p * n + p], aqq = a[q * n + q];
        double theta = (aqq - app) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          double akp = a[k * n + p], akq = a[k * n + q];
          a[k * n + p] = c * akp - s * akq;
          a[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          double apk = a[p * n + k], aqk = a[q * n + k];
          a[p * n + k] = c * apk - s * aqk;
          a[q * n + k] = s * apk + c * aqk;
        }
      }
    }
  }
  for (int i = 0; i < n; ++i) ev[i] = a[i * n + i];
  qsort(ev, n, sizeof(double), cmp_dbl);
}

/* ------------------------------------------------------------------------------------------
 * A3: linear null model.  regression/LinearRegression.cpp:20-69
 *   XtXinv = (X'X)^-1 (LLT) ; B = XtXinv X'y ; predict = X B ; resid = y - predict ;
 *   sigma2 = ||resid||^2 / n  (MLE, :60)
 * X is N x C column-major (base/MathMatrix.h:33-41), first column the intercept
 * (src/ModelUtil.h:102-130).
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_fit_null_linear(int64_t N, int C, const double* X, const double* y, double* resid,
                                double* sigma2, double* xtx_inv /* C*C row-major */,
                                double* beta /* C */) {
  double* xtx = (double*)calloc((size_t)C * C, sizeof(double));
  double* xty = (double*)calloc((size_t)C, sizeof(double));
  for (int a = 0; a < C; ++a) {
    const double* xa = X + (size_t)a * N;
    for (int b = a; b < C; ++b) {
      const double* xb = X + (size_t)b * N;
      double s = 0;
      for (int64_t i = 0; i < N; ++i) s += xa[i] * xb[i];
      xtx[a * C + b] = xtx[b * C + a] = s;
    }
    double s = 0;
    for (int64_t i = 0; i < N; ++i) s += xa[i] * y[i];
    xty[a] = s;
  }
  if (spd_inverse(C, xtx, xtx_inv)) {
    free(xtx);
    free(xty);
    return -1;
  }
  for (int a = 0; a < C; ++a) {
    double s = 0;
    for (int b = 0; b < C; ++b) s += xtx_inv[a * C + b] * xty[b];
    beta[a] = s;
  }
  double rss = 0;
  for (int64_t i = 0; i < N; ++i) {
    double p = 0;
    for (int a = 0; a < C; ++a) p += X[(size_t)a * N + i] * beta[a];
    double r = y[i] - p;
    resid[i] = r;
    rss += r * r;
  }
  *sigma2 = rss / (double)N;
  free(xtx);
  free(xty);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * A1: getFlippedToMinorPolymorphicGenotype = convertToMinorAlleleCount + removeMonomorphicMarker
 *   src/DataConsolidator.h:128-132 ; src/DataConsolidator.cpp:46-69 (flip when colsum > rows),
 *   :94-116 (monomorphic = all non-missing values equal), :118-142 (compaction).
 * G col-major N x M doubles; out must hold N*M doubles; keep[] receives source column ids.
 * Returns M' (number of columns kept).
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_flip_minor_polymorphic(int64_t N, int M, const double* G, double* out, int* keep,
                                       int* flipped) {
  int mo = 0;
  for (int j = 0; j < M; ++j) {
    const double* g = G + (size_t)j * N;
    double s = 0;
    for (int64_t i = 0; i < N; ++i) s += g[i];
    int flip = !(s <= (double)N);
    double* o = out + (size_t)mo * N;
    if (!flip)
      for (int64_t i = 0; i < N; ++i) o[i] = g[i];
    else
      for (int64_t i = 0; i < N; ++i) o[i] = 2 - g[i];
    /* isMonomorphicMarker on the flipped column */
    int64_t first = N;
    for (int64_t i = 0; i < N; ++i)
      if (o[i] >= 0) {
        first = i;
        break;
      }
    int mono = 1;
    for (int64_t i = first + 1; i < N; ++i) {
      if (o[i] < 0) continue;
      if (o[i] != o[first]) {
        mono = 0;
        break;
      }
    }
    if (!mono) {
      keep[mo] = j;
      if (flipped) flipped[mo] = flip;
      ++mo;
    }
  }
  return mo;
}

/* ------------------------------------------------------------------------------------------
 * A2: Beta(MAF; b1, b2) density weight.  src/Model.h:2644-2661 (SKAT: squared),
 * :2799-2813 (SKAT-O: unsquared); density = gsl_ran_beta_pdf (gsl-1.16 randist/beta.c:43-74).
 * ---------------------------------------------------------------------------------------- */
ORC_API double orc_beta_pdf(double x, double a, double b) {
  if (x < 0 || x > 1) return 0;
  double gab = lgamma(a + b), ga = lgamma(a), gb = lgamma(b);
  if (x == 0.0 || x == 1.0) {
    if (a > 1.0 && b > 1.0) return 0.0;
    return exp(gab - ga - gb) * pow(x, a - 1) * pow(1 - x, b - 1);
  }
  return exp(gab - ga - gb + log(x) * (a - 1) + log1p(-x) * (b - 1));
}
ORC_API double orc_skat_weight(double freq, double b1, double b2, int squared) {
  if (freq > 0.5) freq = 1.0 - freq;
  if (freq > 1e-30) {
    double w = orc_beta_pdf(freq, b1, b2);
    return squared ? w * w : w;
  }
  return 0.0;
}

/* ------------------------------------------------------------------------------------------
 * A5: Davies' algorithm (AS 155).  Restatement of regression/qfc.c:26-452 with the file-scope
 * statics (qfc.c:26-30) moved into a struct so that it is re-entrant; constants identical
 * (pi, log28: qfc.c:23-24; exp1 cut-off -50: :35-36).  `real` is double there
 * (qfc.c:1 "#define UseDouble 0" + "#ifdef UseDouble").
 * ---------------------------------------------------------------------------------------- */
#define QF_PI 3.14159265358979
#define QF_LOG28 .0866

typedef struct {
  double sigsq, lmax, lmin, mean, c;
  double intl, ersm;
  int count, r, lim;
  int ndtsrt, fail;
  const int* n;
  int* th;
  const double* lb;
  const double* nc;
  jmp_buf env;
} qf_t;

static double qf_exp1(double x) { return x < -50.0 ? 0.0 : exp(x); }
static void qf_counter(qf_t* s) { /* qfc.c:77-83 */
  s->count = s->count + 1;
  if (s->count > s->lim) longjmp(s->env, 1);
}
static double qf_square(double x) { return x * x; }
static double qf_cube(double x) { return x * x * x; }

static double qf_log1(double x, int first) { /* qfc.c:89-105 */
  if (fabs(x) > 0.1) {
    return (first ? log(1.0 + x) : (log(1.0 + x) - x));
  } else {
    double s, s1, term, y, k;
    y = x / (2.0 + x);
    term = 2.0 * qf_cube(y);
    k = 3.0;
    s = (first ? 2.0 : -x) * y;
    y = qf_square(y);
    for (s1 = s + term / k; s1 != s; s1 = s + term / k) {
      k = k + 2.0;
      term = term * y;
      s = s1;
    }
    return s;
  }
}

static void qf_order(qf_t* s) { /* qfc.c:107-125 */
  int j, k;
  for (j = 0; j < s->r; j++) {
    double lj = fabs(s->lb[j]);
    for (k = j - 1; k >= 0; k--) {
      if (lj > fabs(s->lb[s->th[k]]))
        s->th[k + 1] = s->th[k];
      else
        goto l1;
    }
    k = -1;
  l1:
    s->th[k + 1] = j;
  }
  s->ndtsrt = 0;
}

static double qf_errbd(qf_t* s, double u, double* cx) { /* qfc.c:128-147 */
  double sum1, lj, ncj, x, y, xconst;
  int j, nj;
  qf_counter(s);
  xconst = u * s->sigsq;
  sum1 = u * xconst;
  u = 2.0 * u;
  for (j = s->r - 1; j >= 0; j--) {
    nj = s->n[j];
    lj = s->lb[j];
    ncj = s->nc[j];
    x = u * lj;
    y = 1.0 - x;
    xconst = xconst + lj * (ncj / y + nj) / y;
    sum1 = sum1 + ncj * qf_square(x / y) + nj * (qf_square(x) / y + qf_log1(-x, 0));
  }
  *cx = xconst;
  return qf_exp1(-0.5 * sum1);
}

static double qf_ctff(qf_t* s, double accx, double* upn) { /* qfc.c:149-174 */
  double u1, u2, u, rb, xconst, c1, c2;
  u2 = *upn;
  u1 = 0.0;
  c1 = s->mean;
  rb = 2.0 * ((u2 > 0.0) ? s->lmax : s->lmin);
  for (u = u2 / (1.0 + u2 * rb); qf_errbd(s, u, &c2) > accx; u = u2 / (1.0 + u2 * rb)) {
    u1 = u2;
    c1 = c2;
    u2 = 2.0 * u2;
  }
  for (u = (c1 - s->mean) / (c2 - s->mean); u < 0.9; u = (c1 - s->mean) / (c2 - s->mean)) {
    u = (u1 + u2) / 2.0;
    if (qf_errbd(s, u / (1.0 + u * rb), &xconst) > accx) {
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

static double qf_truncation(qf_t* s, double u, double tausq) { /* qfc.c:176-213 */
  double sum1, sum2, prod1, prod2, prod3, lj, ncj, x, y, err1, err2;
  int j, nj, ss;
  qf_counter(s);
  sum1 = 0.0;
  prod2 = 0.0;
  prod3 = 0.0;
  ss = 0;
  sum2 = (s->sigsq + tausq) * qf_square(u);
  prod1 = 2.0 * sum2;
  u = 2.0 * u;
  for (j = 0; j < s->r; j++) {
    lj = s->lb[j];
    ncj = s->nc[j];
    nj = s->n[j];
    x = qf_square(u * lj);
    sum1 = sum1 + ncj * x / (1.0 + x);
    if (x > 1.0) {
      prod2 = prod2 + nj * log(x);
      prod3 = prod3 + nj * qf_log1(x, 1);
      ss = ss + nj;
    } else
      prod1 = prod1 + nj * qf_log1(x, 1);
  }
  sum1 = 0.5 * sum1;
  prod2 = prod1 + prod2;
  prod3 = prod1 + prod3;
  x = qf_exp1(-sum1 - 0.25 * prod2) / QF_PI;
  y = qf_exp1(-sum1 - 0.25 * prod3) / QF_PI;
  err1 = (ss == 0) ? 1.0 : x * 2.0 / ss;
  err2 = (prod3 > 1.0) ? 2.5 * y : 1.0;
  if (err2 < err1) err1 = err2;
  x = 0.5 * sum2;
  err2 = (x <= y) ? 1.0 : y / x;
  return (err1 < err2) ? err1 : err2;
}

static void qf_findu(qf_t* s, double* utx, double accx) { /* qfc.c:215-234 */
  double u, ut;
  int i;
  static const double divis[] = {2.0, 1.4, 1.2, 1.1};
  ut = *utx;
  u = ut / 4.0;
  if (qf_truncation(s, u, 0.0) > accx) {
    for (u = ut; qf_truncation(s, u, 0.0) > accx; u = ut) ut = ut * 4.0;
  } else {
    ut = u;
    for (u = u / 4.0; qf_truncation(s, u, 0.0) <= accx; u = u / 4.0) ut = u;
  }
  for (i = 0; i < 4; i++) {
    u = ut / divis[i];
    if (qf_truncation(s, u, 0.0) <= accx) ut = u;
  }
  *utx = ut;
}

static void qf_integrate(qf_t* s, int nterm, double interv, double tausq, int mainx) {
  /* qfc.c:237-268 */
  double inpi, u, sum1, sum2, sum3, x, y, z;
  int k, j, nj;
  inpi = interv / QF_PI;
  for (k = nterm; k >= 0; k--) {
    u = (k + 0.5) * interv;
    sum1 = -2.0 * u * s->c;
    sum2 = fabs(sum1);
    sum3 = -0.5 * s->sigsq * qf_square(u);
    for (j = s->r - 1; j >= 0; j--) {
      nj = s->n[j];
      x = 2.0 * s->lb[j] * u;
      y = qf_square(x);
      sum3 = sum3 - 0.25 * nj * qf_log1(y, 1);
      y = s->nc[j] * x / (1.0 + y);
      z = nj * atan(x) + y;
      sum1 = sum1 + z;
      sum2 = sum2 + fabs(z);
      sum3 = sum3 - 0.5 * x * y;
    }
    x = inpi * qf_exp1(sum3) / u;
    if (!mainx) x = x * (1.0 - qf_exp1(-0.5 * tausq * qf_square(u)));
    sum1 = sin(0.5 * sum1) * x;
    sum2 = 0.5 * sum2 * x;
    s->intl = s->intl + sum1;
    s->ersm = s->ersm + sum2;
  }
}

static double qf_cfe(qf_t* s, double x) { /* qfc.c:270-301 */
  double axl, axl1, axl2, sxl, sum1, lj;
  int j, k, t;
  qf_counter(s);
  if (s->ndtsrt) qf_order(s);
  axl = fabs(x);
  sxl = (x > 0.0) ? 1.0 : -1.0;
  sum1 = 0.0;
  for (j = s->r - 1; j >= 0; j--) {
    t = s->th[j];
    if (s->lb[t] * sxl > 0.0) {
      lj = fabs(s->lb[t]);
      axl1 = axl - lj * (s->n[t] + s->nc[t]);
      axl2 = lj / QF_LOG28;
      if (axl1 > axl2)
        axl = axl1;
      else {
        if (axl > axl2) axl = axl2;
        sum1 = (axl - axl1) / lj;
        for (k = j - 1; k >= 0; k--) sum1 = sum1 + (s->n[s->th[k]] + s->nc[s->th[k]]);
        goto l;
      }
    }
  }
l:
  if (sum1 > 100.0) {
    s->fail = 1;
    return 1.0;
  } else
    return pow(2.0, (sum1 / 4.0)) / (QF_PI * qf_square(axl));
}

/* qfc.c:304-452 */
ORC_API double orc_qf(const double* lb1, const double* nc1, const int* n1, int r1, double sigma,
                      double c1, int lim1, double acc, double* trace, int* ifault) {
  qf_t S;
  qf_t* s = &S;
  int j, nj, nt, ntm;
  double acc1, almx, xlim, xnt, xntm;
  double utx, tausq, sd, intv, intv1, x, up, un, d1, d2, lj, ncj;
  volatile double qfval = -1.0;
  static const int rats[] = {1, 2, 4, 8};
  s->th = NULL;

  if (setjmp(s->env) != 0) {
    *ifault = 4;
    goto endofproc;
  }
  s->r = r1;
  s->lim = lim1;
  s->c = c1;
  s->n = n1;
  s->lb = lb1;
  s->nc = nc1;
  for (j = 0; j < 7; j++) trace[j] = 0.0;
  *ifault = 0;
  s->count = 0;
  s->intl = 0.0;
  s->ersm = 0.0;
  qfval = -1.0;
  acc1 = acc;
  s->ndtsrt = 1;
  s->fail = 0;
  xlim = (double)s->lim;
  s->th = (int*)malloc(s->r * (sizeof(int)));
  if (!s->th) {
    *ifault = 5;
    goto endofproc;
  }

  s->sigsq = qf_square(sigma);
  sd = s->sigsq;
  s->lmax = 0.0;
  s->lmin = 0.0;
  s->mean = 0.0;
  for (j = 0; j < s->r; j++) {
    nj = s->n[j];
    lj = s->lb[j];
    ncj = s->nc[j];
    if (nj < 0 || ncj < 0.0) {
      *ifault = 3;
      goto endofproc;
    }
    sd = sd + qf_square(lj) * (2 * nj + 4.0 * ncj);
    s->mean = s->mean + lj * (nj + ncj);
    if (s->lmax < lj)
      s->lmax = lj;
    else if (s->lmin > lj)
      s->lmin = lj;
  }
  if (sd == 0.0) {
    qfval = (s->c > 0.0) ? 1.0 : 0.0;
    goto endofproc;
  }
  if (s->lmin == 0.0 && s->lmax == 0.0 && sigma == 0.0) {
    *ifault = 3;
    goto endofproc;
  }
  sd = sqrt(sd);
  almx = (s->lmax < -s->lmin) ? -s->lmin : s->lmax;

  utx = 16.0 / sd;
  up = 4.5 / sd;
  un = -up;
  qf_findu(s, &utx, .5 * acc1);
  if (s->c != 0.0 && (almx > 0.07 * sd)) {
    tausq = .25 * acc1 / qf_cfe(s, s->c);
    if (s->fail)
      s->fail = 0;
    else if (qf_truncation(s, utx, tausq) < .2 * acc1) {
      s->sigsq = s->sigsq + tausq;
      qf_findu(s, &utx, .25 * acc1);
      trace[5] = sqrt(tausq);
    }
  }
  trace[4] = utx;
  acc1 = 0.5 * acc1;

l1:
  d1 = qf_ctff(s, acc1, &up) - s->c;
  if (d1 < 0.0) {
    qfval = 1.0;
    goto endofproc;
  }
  d2 = s->c - qf_ctff(s, acc1, &un);
  if (d2 < 0.0) {
    qfval = 0.0;
    goto endofproc;
  }
  intv = 2.0 * QF_PI / ((d1 > d2) ? d1 : d2);
  xnt = utx / intv;
  xntm = 3.0 / sqrt(acc1);
  if (xnt > xntm * 1.5) {
    if (xntm > xlim) {
      *ifault = 1;
      goto endofproc;
    }
    ntm = (int)floor(xntm + 0.5);
    intv1 = utx / ntm;
    x = 2.0 * QF_PI / intv1;
    if (x <= fabs(s->c)) goto l2;
    tausq = .33 * acc1 / (1.1 * (qf_cfe(s, s->c - x) + qf_cfe(s, s->c + x)));
    if (s->fail) goto l2;
    acc1 = .67 * acc1;
    qf_integrate(s, ntm, intv1, tausq, 0);
    xlim = xlim - xntm;
    s->sigsq = s->sigsq + tausq;
    trace[2] = trace[2] + 1;
    trace[1] = trace[1] + ntm + 1;
    qf_findu(s, &utx, .25 * acc1);
    acc1 = 0.75 * acc1;
    goto l1;
  }

l2:
  trace[3] = intv;
  if (xnt > xlim) {
    *ifault = 1;
    goto endofproc;
  }
  nt = (int)floor(xnt + 0.5);
  qf_integrate(s, nt, intv, 0.0, 1);
  trace[2] = trace[2] + 1;
  trace[1] = trace[1] + nt + 1;
  qfval = 0.5 - s->intl;
  trace[0] = s->ersm;

  up = s->ersm;
  x = up + acc / 10.0;
  for (j = 0; j < 4; j++) {
    if (rats[j] * x == rats[j] * up) *ifault = 2;
  }

endofproc:
  free(s->th);
  trace[6] = (double)s->count;
  return qfval;
}

/* ------------------------------------------------------------------------------------------
 * regularised upper incomplete gamma Q(a,x) (fp64), the quantity the reference reaches through
 * cdfchn -> cumchn (pnonc<=1e-10) -> cumchi -> cumgam -> gratio
 * (regression/cdflib.cpp:2634, 5172-5235, 5141-5170, 5581).  Written from the textbook series /
 * Lentz continued fraction, NOT a transcription of gratio; agreement with the reference's own
 * cdflib is asserted in tests/test_oracle_pin.py.
 * ---------------------------------------------------------------------------------------- */
ORC_API double orc_gamma_q(double a, double x) {
  if (x <= 0.0) return 1.0;
  if (x < a + 1.0) { /* series for P */
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 100000; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (fabs(del) < fabs(sum) * 1e-17) break;
    }
    double p = sum * exp(-x + a * log(x) - lgamma(a));
    return 1.0 - p;
  } else { /* continued fraction for Q (modified Lentz) */
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 100000; ++i) {
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
    return exp(-x + a * log(x) - lgamma(a)) * h;
  }
}

/* gsl_cdf_chisq_Q(x, 1.0) as used by regression/LinearRegressionScoreTest.cpp:259-261 */
ORC_API double orc_chisq_q(double x, double df) {
  if (x <= 0) return 1.0;
  if (df == 1.0) return erfc(sqrt(0.5 * x));
  return orc_gamma_q(0.5 * df, 0.5 * x);
}

/* MixtureChiSquare::getLiuPvalue, regression/MixtureChiSquare.cpp:44-83 */
ORC_API double orc_liu_pvalue(const double* lambda, int n, double Q) {
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
  double muQ = c1;
  double sigmaQ = sqrt(2.0 * c2);
  double tstar = (Q - muQ) / sigmaQ;
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
  double muX = l + delta;
  double sigmaX = sqrt(2) * a;
  double x = tstar * sigmaX + muX;
  /* cdfchn argument checks (cdflib.cpp:2768-2790): x<0 -> status -4, df<=0 -> -5, ncp<0 -> -6;
     any non-zero status makes getLiuPvalue return 1 (MixtureChiSquare.cpp:79-81). */
  if (x < 0.0 || !(l > 0.0) || delta < 0.0) return 1.0;
  if (x != x || l != l) return 1.0;
  if (delta > 1.0e-10) {
    /* non-central branch: unreachable up to rounding on this path;
       evaluated with the Poisson-mixture definition so that the oracle is still total. */
    double half = 0.5 * delta, wt = exp(-half), sum = 0.0;
    for (int i = 0; i < 2000; ++i) {
      sum += wt * (1.0 - orc_gamma_q(0.5 * l + i, 0.5 * x));
      wt *= half / (i + 1);
      if (wt < 1e-18 && i > half) break;
    }
    return 1.0 - sum;
  }
  return orc_gamma_q(0.5 * l, 0.5 * x); /* cumchn: x<=0 -> ccum=1 (cdflib.cpp:5222-5226) */
}

/* MixtureChiSquare::getPvalue, regression/MixtureChiSquare.cpp:7-29 (lim=10000, acc=1e-6,
 * sigma=0: MixtureChiSquare.h:7).  *fault receives qf's ifault (0 when the Liu shortcut is used). */
ORC_API double orc_mixchisq_pvalue(const double* lambda, int n, double Q, int* fault) {
  *fault = 0;
  if (n == 1) return orc_liu_pvalue(lambda, n, Q);
  double* nc = (double*)calloc(n, sizeof(double));
  int* df = (int*)malloc(sizeof(int) * n);
  for (int i = 0; i < n; ++i) df[i] = 1;
  double trace[7];
  double p = 1.0 - orc_qf(lambda, nc, df, n, 0.0, Q, 10000, 0.000001, trace, fault);
  if (p > 1.0) p = 1.0;
  if (*fault) p = -1.0;
  free(nc);
  free(df);
  return p;
}

/* ------------------------------------------------------------------------------------------
 * A4: Skat::Fit  (regression/Skat.cpp:29-105) in the algebraically reduced fp64 form
 *   K_sqrt = diag(sqrt w) G'        (:41-47)
 *   Q      = || K_sqrt res ||^2     (:50-52)
 *   P0     = V - V X (X'VX)^-1 X'V  (:55-70), V = diag(v)
 *   lambda = eig( K_sqrt P0 K_sqrt' ) (:75-76)
 *         == eig( W^1/2 [ G'VG - (G'VX)(X'VX)^-1(X'VG) ] W^1/2 )   <- O(N M^2), no N x N
 *   keep lambda > 1e-30 from the top, stop at the first failure, at most min(N,M) (:84-98)
 *   p = Davies; if p<=0 or p==1 -> Liu (:100-103)
 * G: N x M col-major (already flipped/polymorphic), X: N x C col-major, v: per-sample variance.
 * out_lambda needs M entries.  Returns 0.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  double Q;
  double pvalue;   /* after the Davies -> Liu rule */
  double p_davies; /* raw getPvalue() result (-1 when fault) */
  double p_liu;
  int fault;    /* qf ifault */
  int n_lambda; /* eigenvalues kept */
} orc_skat_out;

ORC_API int orc_skat_reduced64(int64_t N, int M, int C, const double* G, const double* X,
                               const double* res, const double* v, const double* w,
                               orc_skat_out* out, double* out_lambda) {
  double* s = (double*)calloc(M, sizeof(double));
  double* A = (double*)calloc((size_t)M * M, sizeof(double));
  double* B = (double*)calloc((size_t)M * C, sizeof(double));
  double* XVX = (double*)calloc((size_t)C * C, sizeof(double));
  double* XVXi = (double*)calloc((size_t)C * C, sizeof(double));
  /* row-major scratch copy of a block of rows for the rank-1 style update (cache friendly) */
  const int RB = 256;
  double* blk = (double*)malloc(sizeof(double) * RB * M);
  for (int64_t i0 = 0; i0 < N; i0 += RB) {
    int nb = (int)((N - i0 < RB) ? (N - i0) : RB);
    for (int j = 0; j < M; ++j) {
      const double* g = G + (size_t)j * N + i0;
      for (int i = 0; i < nb; ++i) blk[(size_t)i * M + j] = g[i];
    }
    for (int i = 0; i < nb; ++i) {
      const double* gi = blk + (size_t)i * M;
      double vi = v[i0 + i], ri = res[i0 + i];
      for (int j = 0; j < M; ++j) {
        double gj = gi[j];
        if (gj == 0.0) continue; /* exact: contributes nothing */
        s[j] += gj * ri;
        double vg = vi * gj;
        double* Aj = A + (size_t)j * M;
        for (int k = 0; k < M; ++k) Aj[k] += vg * gi[k];
        for (int c = 0; c < C; ++c) B[(size_t)j * C + c] += vg * X[(size_t)c * N + i0 + i];
      }
    }
  }
  for (int a = 0; a < C; ++a)
    for (int b = a; b < C; ++b) {
      double t = 0;
      for (int64_t i = 0; i < N; ++i) t += X[(size_t)a * N + i] * v[i] * X[(size_t)b * N + i];
      XVX[a * C + b] = XVX[b * C + a] = t;
    }
  int rc = spd_inverse(C, XVX, XVXi);
  double Q = 0;
  for (int j = 0; j < M; ++j) Q += w[j] * s[j] * s[j];
  double* K = (double*)calloc((size_t)M * M, sizeof(double));
  double* T = (double*)calloc((size_t)M * C, sizeof(double));
  for (int j = 0; j < M; ++j)
    for (int c = 0; c < C; ++c) {
      double t = 0;
      for (int d = 0; d < C; ++d) t += B[(size_t)j * C + d] * XVXi[d * C + c];
      T[(size_t)j * C + c] = t;
    }
  for (int j = 0; j < M; ++j)
    for (int k = 0; k < M; ++k) {
      double t = 0;
      for (int c = 0; c < C; ++c) t += T[(size_t)j * C + c] * B[(size_t)k * C + c];
      K[(size_t)j * M + k] = sqrt(w[j]) * (A[(size_t)j * M + k] - t) * sqrt(w[k]);
    }
  /* symmetrise against rounding before Jacobi */
  for (int j = 0; j < M; ++j)
    for (int k = j + 1; k < M; ++k) {
      double t = 0.5 * (K[(size_t)j * M + k] + K[(size_t)k * M + j]);
      K[(size_t)j * M + k] = K[(size_t)k * M + j] = t;
    }
  double* ev = (double*)malloc(sizeof(double) * M);
  orc_sym_eigenvalues(M, K, ev);
  int r_ub = (int)((N < M) ? N : M);
  int r = 0;
  for (int i = M - 1; i >= 0; --i) {
    if (ev[i] > 1e-30 && r < r_ub) {
      out_lambda[r++] = ev[i];
    } else
      break;
  }
  out->Q = Q;
  out->n_lambda = r;
  out->fault = 0;
  out->p_davies = orc_mixchisq_pvalue(out_lambda, r, Q, &out->fault);
  out->p_liu = orc_liu_pvalue(out_lambda, r, Q);
  out->pvalue = out->p_davies;
  if (out->pvalue <= 0.0 || out->pvalue == 1.0) out->pvalue = out->p_liu;
  free(s); free(A); free(B); free(XVX); free(XVXi); free(blk); free(K); free(T); free(ev);
  return rc;
}

/* Literal restatement of regression/Skat.cpp:29-105 INCLUDING the float32 casts
 * (regression/EigenMatrixInterface.cpp:10-27) and the explicit N x N P0 (:55-70).
 * Only usable for small N; exists to validate the reduced algebra above. */
ORC_API int orc_skat_faithful32(int N, int M, int C, const double* G, const double* X,
                                const double* res, const double* v, const double* w,
                                orc_skat_out* out, double* out_lambda) {
  float* Ks = (float*)malloc(sizeof(float) * (size_t)M * N); /* M x N row-major */
  for (int j = 0; j < M; ++j) {
    float ws = sqrtf((float)w[j]);
    for (int i = 0; i < N; ++i) Ks[(size_t)j * N + i] = ws * (float)G[(size_t)j * N + i];
  }
  float Qf = 0;
  for (int j = 0; j < M; ++j) {
    float t = 0;
    for (int i = 0; i < N; ++i) t += Ks[(size_t)j * N + i] * (float)res[i];
    Qf += t * t;
  }
  float* P0 = (float*)malloc(sizeof(float) * (size_t)N * N);
  if (C == 1) {
    float vs = 0;
    for (int i = 0; i < N; ++i) vs += (float)v[i];
    for (int i = 0; i < N; ++i)
      for (int k = 0; k < N; ++k) P0[(size_t)i * N + k] = -(float)v[i] * (float)v[k] / vs;
    for (int i = 0; i < N; ++i) P0[(size_t)i * N + i] += (float)v[i];
  } else {
    /* XtV (C x N), inv(XtV X) in float via double Cholesky of the float product */
    float* XtV = (float*)malloc(sizeof(float) * (size_t)C * N);
    for (int c = 0; c < C; ++c)
      for (int i = 0; i < N; ++i) XtV[(size_t)c * N + i] = (float)X[(size_t)c * N + i] * (float)v[i];
    double* m = (double*)malloc(sizeof(double) * C * C);
    double* mi = (double*)malloc(sizeof(double) * C * C);
    for (int a = 0; a < C; ++a)
      for (int b = 0; b < C; ++b) {
        float t = 0;
        for (int i = 0; i < N; ++i) t += XtV[(size_t)a * N + i] * (float)X[(size_t)b * N + i];
        m[a * C + b] = t;
      }
    for (int a = 0; a < C; ++a)
      for (int b = a + 1; b < C; ++b) m[a * C + b] = m[b * C + a] = 0.5 * (m[a * C + b] + m[b * C + a]);
    spd_inverse(C, m, mi);
    for (int i = 0; i < N; ++i)
      for (int k = 0; k < N; ++k) {
        float t = 0;
        for (int a = 0; a < C; ++a) {
          float u = 0;
          for (int b = 0; b < C; ++b) u += (float)mi[a * C + b] * XtV[(size_t)b * N + k];
          t += XtV[(size_t)a * N + i] * u;
        }
        P0[(size_t)i * N + k] = -t;
      }
    for (int i = 0; i < N; ++i) P0[(size_t)i * N + i] += (float)v[i];
    free(XtV); free(m); free(mi);
  }
  /* K = Ks P0 Ks' in float */
  float* KP = (float*)malloc(sizeof(float) * (size_t)M * N);
  for (int j = 0; j < M; ++j)
    for (int k = 0; k < N; ++k) {
      float t = 0;
      for (int i = 0; i < N; ++i) t += Ks[(size_t)j * N + i] * P0[(size_t)i * N + k];
      KP[(size_t)j * N + k] = t;
    }
  double* K = (double*)malloc(sizeof(double) * (size_t)M * M);
  for (int j = 0; j < M; ++j)
    for (int k = 0; k < M; ++k) {
      float t = 0;
      for (int i = 0; i < N; ++i) t += KP[(size_t)j * N + i] * Ks[(size_t)k * N + i];
      K[(size_t)j * M + k] = t;
    }
  for (int j = 0; j < M; ++j)
    for (int k = j + 1; k < M; ++k) {
      double t = 0.5 * (K[(size_t)j * M + k] + K[(size_t)k * M + j]);
      K[(size_t)j * M + k] = K[(size_t)k * M + j] = t;
    }
  double* ev = (double*)malloc(sizeof(double) * M);
  orc_sym_eigenvalues(M, K, ev);
  int r_ub = (N < M) ? N : M, r = 0;
  for (int i = M - 1; i >= 0; --i) {
    float evf = (float)ev[i];
    if (evf > 1e-30f && r < r_ub)
      out_lambda[r++] = (double)evf;
    else
      break;
  }
  out->Q = (double)Qf;
  out->n_lambda = r;
  out->fault = 0;
  out->p_davies = orc_mixchisq_pvalue(out_lambda, r, out->Q, &out->fault);
  out->p_liu = orc_liu_pvalue(out_lambda, r, out->Q);
  out->pvalue = out->p_davies;
  if (out->pvalue <= 0.0 || out->pvalue == 1.0) out->pvalue = out->p_liu;
  free(Ks); free(P0); free(KP); free(K); free(ev);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * A8: cmcCollapse / zegginiCollapse  (src/Model.cpp:73-89, 115-130): `(int)g > 0`
 * A10: totalNonRefSite (src/Model.h:894-900)
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_cmc_collapse(int64_t N, int M, const double* G, double* out) {
  for (int64_t p = 0; p < N; ++p) {
    out[p] = 0.0;
    for (int m = 0; m < M; ++m) {
      int g = (int)(G[(size_t)m * N + p]);
      if (g > 0) {
        out[p] = 1.0;
        break;
      }
    }
  }
}
ORC_API void orc_zeggini_collapse(int64_t N, int M, const double* G, double* out) {
  for (int64_t p = 0; p < N; ++p) {
    out[p] = 0.0;
    for (int m = 0; m < M; ++m) {
      int g = (int)(G[(size_t)m * N + p]);
      if (g > 0) out[p] += 1.0;
    }
  }
}
ORC_API int orc_nonref_sites(int64_t N, const double* collapsed) {
  int s = 0;
  for (int64_t i = 0; i < N; ++i) s += collapsed[i] == 0.0 ? 0 : 1;
  return s;
}

/* ------------------------------------------------------------------------------------------
 * A9: LinearRegressionScoreTest::TestCovariate(Matrix Xnull, Vector y, Matrix Xcol), m = 1
 *   regression/LinearRegressionScoreTest.cpp:173-263
 *   U = S'r ; SS = S'S ; SZ = S'Z ; ZZ = Z'Z ; SS -= SZ ZZ^-1 SZ' ; V = SS*sigma2 ;
 *   stat = U (SS^-1 / sigma2) U ; stat<0 -> fail ; p = gsl_cdf_chisq_Q(stat, 1.0)
 * Returns 0 when fitOK, -1 otherwise.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_score_test_1(int64_t N, int C, const double* Z, const double* resid, double sigma2,
                             const double* S, double* U_out, double* V_out, double* stat_out,
                             double* p_out) {
  double U = 0, SS = 0;
  double* SZ = (double*)calloc(C, sizeof(double));
  double* ZZ = (double*)calloc((size_t)C * C, sizeof(double));
  double* ZZi = (double*)calloc((size_t)C * C, sizeof(double));
  for (int64_t i = 0; i < N; ++i) {
    U += S[i] * resid[i];
    SS += S[i] * S[i];
  }
  for (int c = 0; c < C; ++c) {
    double t = 0;
    for (int64_t i = 0; i < N; ++i) t += S[i] * Z[(size_t)c * N + i];
    SZ[c] = t;
    for (int d = c; d < C; ++d) {
      double u = 0;
      for (int64_t i = 0; i < N; ++i) u += Z[(size_t)c * N + i] * Z[(size_t)d * N + i];
      ZZ[c * C + d] = ZZ[d * C + c] = u;
    }
  }
  int rc = spd_inverse(C, ZZ, ZZi);
  double q = 0;
  for (int c = 0; c < C; ++c)
    for (int d = 0; d < C; ++d) q += SZ[c] * ZZi[c * C + d] * SZ[d];
  SS -= q;
  double V = SS * sigma2;
  /* SS.llt().solve(I) of a 1x1: 1/SS (NaN/inf when SS<=0, as Eigen's LLT would produce) */
  double inv = 1.0 / SS;
  double stat = U * (inv / sigma2) * U;
  *U_out = U;
  *V_out = V;
  *stat_out = stat;
  free(SZ); free(ZZ); free(ZZi);
  if (rc) return -1;
  if (stat < 0 || stat != stat) {
    *p_out = NAN;
    return -1;
  }
  *p_out = orc_chisq_q(stat, 1.0);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Whole-gene driver = SkatTest::fit + CMCTest::fit + ZegginiTest::fit on one gene
 * (src/Model.h:2630-2720, 821-858, 1177-1215) with the null model supplied by the caller
 * (the reference caches it for SKAT, :2672-2699, and refits the identical OLS per gene for the
 * burden tests, :850/:1207 -- same numbers).
 *   G_raw : N x M col-major, as returned by DataConsolidator::getGenotype() (imputed, unflipped)
 *   af    : per ORIGINAL column allele frequency (GenotypeCounter::getAF, src/GenotypeCounter.h:46-52);
 *           weight i of the flipped/polymorphic matrix uses af[i]  -- the index quirk of
 *           src/Model.h:2644-2646 + src/DataConsolidator.cpp:527-529 is kept.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int m_poly;
  int status; /* 0 ok, 2 = no polymorphic variant (fit returns -1, output NA) */
  orc_skat_out skat;
  int cmc_nonref;
  double cmc_U, cmc_V, cmc_stat, cmc_p;
  int cmc_ok;
  double zeg_U, zeg_V, zeg_stat, zeg_p;
  int zeg_ok;
} orc_gene_out;

/* scratch for one gene: the flipped copy (N*M), keep[], w, v, col */
typedef struct {
  double* G; int* keep; double* w; double* v; double* col; int64_t capN; int capM;
} orc_scratch;
static void orc_scratch_init(orc_scratch* s, int64_t N, int M) {
  s->G = (double*)malloc(sizeof(double) * (size_t)N * M);
  s->keep = (int*)malloc(sizeof(int) * M);
  s->w = (double*)malloc(sizeof(double) * M);
  s->v = (double*)malloc(sizeof(double) * N);
  s->col = (double*)malloc(sizeof(double) * N);
  s->capN = N; s->capM = M;
}
static void orc_scratch_free(orc_scratch* s) { free(s->G); free(s->keep); free(s->w); free(s->v); free(s->col); }

static int orc_gene_with(orc_scratch* sc, int64_t N, int M, int C, const double* G_raw, const double* af, const double* X,
                         const double* resid, double sigma2, double beta1, double beta2,
                         orc_gene_out* out, double* out_lambda /* M */) {
  double* G = sc->G;
  int* keep = sc->keep;
  int mp = orc_flip_minor_polymorphic(N, M, G_raw, G, keep, NULL);
  memset(out, 0, sizeof(*out));
  out->m_poly = mp;
  if (mp == 0) {
    out->status = 2;
    return 0;
  }
  double* w = sc->w;
  for (int i = 0; i < mp; ++i) w[i] = orc_skat_weight(af[i], beta1, beta2, 1);
  double* v = sc->v;
  for (int64_t i = 0; i < N; ++i) v[i] = sigma2;
  orc_skat_reduced64(N, mp, C, G, X, resid, v, w, &out->skat, out_lambda);
  double* col = sc->col;
  orc_cmc_collapse(N, mp, G, col);
  out->cmc_nonref = orc_nonref_sites(N, col);
  out->cmc_ok = orc_score_test_1(N, C, X, resid, sigma2, col, &out->cmc_U, &out->cmc_V,
                                 &out->cmc_stat, &out->cmc_p) == 0;
  orc_zeggini_collapse(N, mp, G, col);
  out->zeg_ok = orc_score_test_1(N, C, X, resid, sigma2, col, &out->zeg_U, &out->zeg_V,
                                 &out->zeg_stat, &out->zeg_p) == 0;
  return 0;
}

ORC_API int orc_gene(int64_t N, int M, int C, const double* G_raw, const double* af, const double* X,
                     const double* resid, double sigma2, double beta1, double beta2,
                     orc_gene_out* out, double* out_lambda /* M */) {
  orc_scratch sc;
  orc_scratch_init(&sc, N, M);
  int rc = orc_gene_with(&sc, N, M, C, G_raw, af, X, resid, sigma2, beta1, beta2, out, out_lambda);
  orc_scratch_free(&sc);
  return rc;
}

/* ------------------------------------------------------------------------------------------
 * A6: the permutation p-value of SkatTest::fit (src/Model.h:2707-2717).
 *   permute()                    src/LinearAlgebra.h:8-21   i = n-1..1, j = rand() % (i+1), swap -- glibc rand(),
 *                                                           never seeded by the reference (default state == srand(1))
 *   Skat::GetQFromNewResidual    regression/Skat.cpp:107-116  float32: res cast to float, K_sqrt = W^1/2 G' in float,
 *                                                           (K_sqrt * res).squaredNorm()
 *   Permutation init/next/add/getPvalue   src/Permutation.h:69-98   `int threshold = 1.0*numPerm*alpha*2`
 * The shuffles are cumulative within a gene (permutedRes is shuffled again and again) and the rand() stream runs
 * on from gene to gene.  `reseed` != 0 calls srand(reseed) first (srand(1) == a fresh process).
 * G: the flipped/polymorphic N x mp matrix, w: the SQUARED Beta weights, obs: the observed statistic
 * (Skat::GetQ).  out[0..2] = ActualPerm, NumGreater, NumEqual; *pval = PermPvalue.  q_out (nullable): the
 * statistic of every permutation that ran.
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_skat_perm(int64_t N, int mp, const double* G, const double* w, const double* resid, double obs,
                          int nPerm, double alpha, unsigned reseed, int* out, double* pval, double* q_out) {
  if (reseed) srand(reseed);
  float* K = (float*)malloc(sizeof(float) * (size_t)N * mp);   /* K_sqrt, row j = sqrt(w_j) g_j' */
  float* rf = (float*)malloc(sizeof(float) * (size_t)N);
  double* v = (double*)malloc(sizeof(double) * (size_t)N);
  for (int j = 0; j < mp; ++j) {
    const float sw = (float)sqrt(w[j]);                         /* w_sqrt is a VectorXf (Skat.cpp:41-45) */
    for (int64_t i = 0; i < N; ++i) K[(size_t)j * N + i] = sw * (float)G[(size_t)j * N + i];
  }
  for (int64_t i = 0; i < N; ++i) v[i] = resid[i];              /* permutedRes = res */
  int actual = 0, numX = 0, numEq = 0;
  const int threshold = 1.0 * nPerm * alpha * 2;
  while (actual < nPerm && numX + numEq < threshold) {
    for (int64_t i = N - 1; i >= 1; --i) {
      int j = rand() % (i + 1);
      if (i != j) { double t = v[i]; v[i] = v[j]; v[j] = t; }
    }
    for (int64_t i = 0; i < N; ++i) rf[i] = (float)v[i];
    float q = 0.f;
    for (int j = 0; j < mp; ++j) {
      float d = 0.f;
      for (int64_t i = 0; i < N; ++i) d += K[(size_t)j * N + i] * rf[i];
      q += d * d;
    }
    const double s = (double)q;
    if (q_out) q_out[actual] = s;
    ++actual;
    if (s > obs) ++numX;
    if (s == obs) ++numEq;
  }
  out[0] = actual; out[1] = numX; out[2] = numEq;
  *pval = actual ? 1.0 * (numX + 0.5 * numEq) / actual : 1.0;
  free(K); free(rf); free(v);
  return 0;
}

/* gene-level wrapper: flip / drop monomorphic, weights (index quirk F9), then orc_skat_perm */
ORC_API int orc_gene_perm(int64_t N, int M, const double* G_raw, const double* af, const double* resid, double obs,
                          double beta1, double beta2, int nPerm, double alpha, unsigned reseed, int* out, double* pval,
                          double* q_out) {
  orc_scratch sc;
  orc_scratch_init(&sc, N, M);
  int mp = orc_flip_minor_polymorphic(N, M, G_raw, sc.G, sc.keep, NULL);
  int rc = -1;
  if (mp > 0) {
    for (int i = 0; i < mp; ++i) sc.w[i] = orc_skat_weight(af[i], beta1, beta2, 1);
    rc = orc_skat_perm(N, mp, sc.G, sc.w, resid, obs, nPerm, alpha, reseed, out, pval, q_out);
  }
  orc_scratch_free(&sc);
  return rc;
}

/* the next n values of glibc rand() (after srand(reseed) when reseed != 0): pins the device generator */
ORC_API void orc_glibc_rand(unsigned reseed, int64_t skip, int64_t n, int* out) {
  if (reseed) srand(reseed);
  for (int64_t i = 0; i < skip; ++i) (void)rand();
  for (int64_t i = 0; i < n; ++i) out[i] = rand();
}

/* Batch driver for the CPU baseline: genes laid out back to back (each N x M col-major doubles),
 * OpenMP over genes on `threads` host threads (the reference's own gene loop is serial,
 * src/Main.cpp:1221-1254; threads=1 reproduces that). */
ORC_API int orc_gene_batch(int64_t N, int M, int C, int n_genes, const double* G_all,
                           const double* af_all, const double* X, const double* resid,
                           double sigma2, double beta1, double beta2, orc_gene_out* out,
                           int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int g = 0; g < n_genes; ++g) {
    double* lam = (double*)malloc(sizeof(double) * M);
    orc_gene(N, M, C, G_all + (size_t)g * N * M, af_all + (size_t)g * M, X, resid, sigma2, beta1,
             beta2, &out[g], lam);
    free(lam);
  }
  return 0;
}

/* Same, but task t works on gene index[t] (lets a bounded set of distinct genes stand in for a
 * long gene list when timing the CPU baseline: every task still streams its 8*N*M bytes). */
ORC_API int orc_gene_batch_idx(int64_t N, int M, int C, int n_tasks, const int* index, const double* G_all,
                               const double* af_all, const double* X, const double* resid,
                               double sigma2, double beta1, double beta2, orc_gene_out* out,
                               int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel
#endif
  {
    /* per-thread scratch, allocated and first-touched once per thread and kept across calls
       (OpenMP keeps its worker threads): the timed loop does no 200 MB mallocs / page faults */
    static __thread orc_scratch sc;
    static __thread int sc_ready = 0;
    if (!sc_ready || sc.capN < N || sc.capM < M) {
      if (sc_ready) orc_scratch_free(&sc);
      orc_scratch_init(&sc, N, M);
      memset(sc.G, 0, sizeof(double) * (size_t)N * M);
      sc_ready = 1;
    }
    double* lam = (double*)malloc(sizeof(double) * M);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (int t = 0; t < n_tasks; ++t) {
      const int g = index[t];
      orc_gene_with(&sc, N, M, C, G_all + (size_t)g * N * M, af_all + (size_t)g * M, X, resid, sigma2, beta1,
                    beta2, &out[t], lam);
    }
    free(lam);
  }
  return 0;
}

/* Host twin of the device genotype generator (rvtests_b200/csrc/prep.cuh k_synth_rows): writes
 * rows x N doubles (row-major == each variant contiguous) for the given keys / thresholds. */
static inline uint64_t orc_mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
ORC_API void orc_synth_rows_f64(int64_t rows, int64_t N, const uint64_t* keys, const uint32_t* t0,
                                const uint32_t* t1, double* out, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(static)
#endif
  for (int64_t r = 0; r < rows; ++r) {
    const uint64_t key = keys[r];
    const uint32_t a = t0[r], b = t1[r];
    double* o = out + (size_t)r * N;
    for (int64_t i = 0; i < N; ++i) {
      uint32_t h = (uint32_t)(orc_mix64(key + (uint64_t)i * 0xD1B54A32D192ED03ull) >> 32);
      o[i] = (double)((h >= a) + (h >= b));
    }
  }
}

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
