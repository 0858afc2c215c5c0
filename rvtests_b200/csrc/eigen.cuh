// eigen.cuh -- eigenvalues of a small dense symmetric matrix (n <= 256), cooperative Jacobi.
//
// Replaces Eigen::SelfAdjointEigenSolver at regression/Skat.cpp:75-76 (float there; fp64 here)
// and regression/SkatO.cpp:350-352.  The reference tridiagonalises + QL on one CPU thread; on the
// GPU one CTA owns one gene's matrix in shared memory and runs a parallel-ordered (round-robin
// tournament) two-sided Jacobi: every round rotates n/2 disjoint (p,q) pairs at once --
//   phase A: rotation angles from the current a_pp, a_qq, a_pq
//   phase B: rows p,q  <- J^T A      (disjoint rows, all columns in parallel)
//   phase C: cols p,q  <- A J        (disjoint cols, all rows in parallel)
// Only eigenvalues are needed on this path, so no eigenvector accumulation.
#pragma once
#include "davies.cuh"

namespace rvt {

// a: n x n row-major with leading dimension lda (destroyed); cs: 2*((n+1)/2) doubles of scratch;
// red: scratch for the group's all-reduce.  Eigenvalues are left on the diagonal (unsorted).
template <class Par>
RVT_HDN int jacobi_eigenvalues(double* a, int n, int lda, double* cs, const Par& par) {
  if (n <= 1) return 0;
  const int ne = (n & 1) ? n + 1 : n;  // pad to even with a dummy player
  const int half = ne / 2;
  int sweeps = 0;
  double prev_off = -1.0;
  for (; sweeps < 60; ++sweeps) {
    // convergence: sum of squared off-diagonals vs diagonal
    double off = 0.0, dg = 0.0;
    for (int idx = par.tid(); idx < n * n; idx += par.nt()) {
      int i = idx / n, j = idx - i * n;
      double v = a[i * lda + j];
      if (i == j)
        dg += v * v;
      else
        off += v * v;
    }
    par.allreduce2(off, dg);
    // stop at ||off||_F <= 1e-15 ||diag||_F, or once the rounding floor is reached (no further
    // quadratic decrease while already below 1e-13): eigenvalue error is bounded by ||off||.
    if (off <= 1e-300 || off <= 1e-30 * dg) break;
    if (prev_off >= 0.0 && off <= 1e-26 * dg && off > 0.25 * prev_off) break;
    prev_off = off;
    for (int r = 0; r < ne - 1; ++r) {
      // phase A
      for (int i = par.tid(); i < half; i += par.nt()) {
        int p, q;
        if (i == 0) {
          p = ne - 1;
          q = r;
        } else {
          p = (r + i) % (ne - 1);
          q = (r - i + (ne - 1)) % (ne - 1);
        }
        double c = 1.0, s = 0.0;
        if (p < n && q < n) {
          double apq = a[p * lda + q];
          if (apq != 0.0) {
            double theta = (a[q * lda + q] - a[p * lda + p]) / (2.0 * apq);
            double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            c = 1.0 / sqrt(t * t + 1.0);
            s = t * c;
          }
        }
        cs[2 * i] = c;
        cs[2 * i + 1] = s;
      }
      par.sync();
      // phase B: rows
      for (int idx = par.tid(); idx < half * n; idx += par.nt()) {
        int i = idx / n, k = idx - i * n;
        int p, q;
        if (i == 0) {
          p = ne - 1;
          q = r;
        } else {
          p = (r + i) % (ne - 1);
          q = (r - i + (ne - 1)) % (ne - 1);
        }
        if (p < n && q < n) {
          double c = cs[2 * i], s = cs[2 * i + 1];
          double apk = a[p * lda + k], aqk = a[q * lda + k];
          a[p * lda + k] = c * apk - s * aqk;
          a[q * lda + k] = s * apk + c * aqk;
        }
      }
      par.sync();
      // phase C: columns
      for (int idx = par.tid(); idx < half * n; idx += par.nt()) {
        int i = idx / n, k = idx - i * n;
        int p, q;
        if (i == 0) {
          p = ne - 1;
          q = r;
        } else {
          p = (r + i) % (ne - 1);
          q = (r - i + (ne - 1)) % (ne - 1);
        }
        if (p < n && q < n) {
          double c = cs[2 * i], s = cs[2 * i + 1];
          double akp = a[k * lda + p], akq = a[k * lda + q];
          a[k * lda + p] = c * akp - s * akq;
          a[k * lda + q] = s * akp + c * akq;
        }
      }
      par.sync();
    }
  }
  return sweeps;
}

// -----------------------------------------------------------------------------------------------
// Faster path used by the kernels: Householder tridiagonalisation (parallel mat-vec and rank-2
// update over the group) followed by Sturm-sequence bisection, one eigenvalue per thread.
// ~n^3*4/3 flops and ~6n group barriers instead of ~10 Jacobi sweeps of 3(n-1) barriers each.
// Accuracy: backward stable, |d lambda| <~ n eps ||A|| -- the same class as Jacobi for this path
// (p-values depend on lambda / lambda_max).

// #{eigenvalues of the symmetric tridiagonal (d, e2 = e^2) that are < x}: sign changes of the
// Sturm sequence p_0 = 1, p_1 = d_0 - x, p_{i+1} = (d_i - x) p_i - e_{i-1}^2 p_{i-1}.  The usual
// ratio form q_i = p_i / p_{i-1} costs one fp64 DIVISION per step (a ~25-instruction dependent
// chain on the GPU; 52 bisection steps x n of them per eigenvalue was two thirds of the solver's
// time); the polynomial form costs two FMAs.  Its classical weakness -- over/underflow of p_i --
// is handled by rescaling the pair (p_i, p_{i-1}), which leaves every ratio and hence every sign
// unchanged.  An exact zero counts as a negative ratio, like q = -pivmin in the ratio form.
// The caller scales the matrix so that |d_i - x| <= 4 and e2 <= 1.
RVT_HD int sturm_count(const double* d, const double* e2, int n, double x) {
  double pp = 1.0;          // p_{i-1}
  double pm = d[0] - x;     // p_i
  if (pm == 0.0) pm = -1e-300;
#if defined(__CUDA_ARCH__)
  // Device form of the loop below, same arithmetic (hence the same counts), fewer instructions per term: this loop is
  // ~60 % of the instructions k_finalize executes (ncu source page, profiles/r02_finalize_hot.txt).  The magnitude test
  // first reads the exponent field (biased exponent in 360..1686 => 1e-200 < |p| < 1e200 for sure), the sign change is
  // an XOR of the sign bits.
  int hm = __double2hiint(pm);
  unsigned int cnt = (unsigned int)hm >> 31;
#pragma unroll 4
  for (int i = 1; i < n; ++i) {
    double p = (d[i] - x) * pm - e2[i - 1] * pp;
    int hp = __double2hiint(p);
    const unsigned int ex = ((unsigned int)hp >> 20) & 0x7FFu;
    if (ex - 360u > 1326u) {   // binades that straddle a bound take the exact test: same decisions as the host form
      const double ap = fabs(p);
      if (!(ap >= 1e-200 && ap <= 1e200)) {
        if (p == 0.0) {
          const double t = fmax(fabs(pm) * 1e-290, 1e-305);
          p = (pm < 0.0) ? t : -t;
        }
        const double sc = (ap > 1.0) ? 1e-200 : 1e200;
        p *= sc;
        pm *= sc;
        hp = __double2hiint(p);
        hm = __double2hiint(pm);
      }
    }
    cnt += (unsigned int)(hp ^ hm) >> 31;
    pp = pm;
    pm = p;
    hm = hp;
  }
  return (int)cnt;
#else
  int cnt = pm < 0.0;
  for (int i = 1; i < n; ++i) {
    double p = (d[i] - x) * pm - e2[i - 1] * pp;
    const double ap = fabs(p);
    if (!(ap >= 1e-200 && ap <= 1e200)) {
      if (p == 0.0) {
        // zero pivot: a tiny value of the sign opposite to p_i
        const double t = fmax(fabs(pm) * 1e-290, 1e-305);
        p = (pm < 0.0) ? t : -t;
      }
      const double sc = (ap > 1.0) ? 1e-200 : 1e200;
      p *= sc;
      pm *= sc;
    }
    cnt += (p < 0.0) != (pm < 0.0);
    pp = pm;
    pm = p;
  }
  return cnt;
#endif
}

// Householder tridiagonalisation alone: d[0..n) the diagonal, e[0..n-1) the sub-diagonal (e[n-1] = 0).  n >= 2.
template <class Par>
RVT_HDN void householder_tridiag(double* a, int n, int lda, double* d, double* e, double* v, double* p, const Par& par) {
  // thread (ti, tj) of a W-wide grid over the group: rows are dealt to ti, columns to tj, so the
  // inner loops run over consecutive addresses with no integer division
  const int W = par.nt() < 32 ? par.nt() : 32;
  const int ti = par.tid() / W, tj = par.tid() - ti * W, nti = par.nt() / W;
  for (int k = 0; k < n - 2; ++k) {
    const int m = n - k - 1;          // size of the trailing block; x = a[k+1.., k]
    double* x = a + (k + 1) * lda + k;  // stride lda
    double ss = 0.0, dummy = 0.0;
    for (int i = par.tid() + 1; i < m; i += par.nt()) ss += x[i * lda] * x[i * lda];
    par.allreduce2(ss, dummy);
    const double x0 = x[0];
    if (ss == 0.0) {  // already tridiagonal in this column
      if (par.tid() == 0) {
        d[k] = a[k * lda + k];
        e[k] = x0;
      }
      par.sync();
      continue;
    }
    const double alpha = (x0 >= 0.0 ? -1.0 : 1.0) * sqrt(x0 * x0 + ss);
    const double v0 = x0 - alpha;
    const double beta = 2.0 / (v0 * v0 + ss);
    for (int i = par.tid(); i < m; i += par.nt()) v[i] = (i == 0) ? v0 : x[i * lda];
    par.sync();
    // p = beta * A22 v : one row per thread (rows are lda apart: an odd lda keeps the banks distinct)
    double* a22 = a + (k + 1) * lda + (k + 1);
    double pv = 0.0;
    dummy = 0.0;
    for (int i = par.tid(); i < m; i += par.nt()) {
      double s0 = 0.0, s1 = 0.0;
      const double* row = a22 + i * lda;
      int j = 0;
      for (; j + 1 < m; j += 2) {
        s0 += row[j] * v[j];
        s1 += row[j + 1] * v[j + 1];
      }
      if (j < m) s0 += row[j] * v[j];
      const double pi = beta * (s0 + s1);
      p[i] = pi;
      pv += pi * v[i];
    }
    par.allreduce2(pv, dummy);   // (its barriers also publish p)
    const double kk = 0.5 * beta * pv;
    // w = p - kk v ;  A22 -= v w' + w v'   (w formed on the fly: p is read-only here)
    for (int i = ti; i < m; i += nti) {
      const double vi = v[i], wi = p[i] - kk * vi;
      double* row = a22 + i * lda;
      for (int j = tj; j < m; j += W) {
        const double vj = v[j], wj = p[j] - kk * vj;
        row[j] -= vi * wj + wi * vj;
      }
    }
    if (par.tid() == 0) {
      d[k] = a[k * lda + k];
      e[k] = alpha;
    }
    par.sync();
  }
  if (par.tid() == 0) {
    d[n - 2] = a[(n - 2) * lda + (n - 2)];
    d[n - 1] = a[(n - 1) * lda + (n - 1)];
    e[n - 2] = a[(n - 1) * lda + (n - 2)];
    e[n - 1] = 0.0;
  }
  par.sync();
}

// Eigenvalues of the symmetric tridiagonal (d, e) by Sturm bisection, one eigenvalue per thread; out[n] DESCENDING.
// v[n], p[n]: group-visible scratch.  n >= 2.  (Needs ~2 n doubles of shared memory and few registers: k_fin_sturm runs it
// at many CTAs per SM, where the all-in-one statistics kernel is held to 6 by the M x M matrix.)
template <class Par>
RVT_HDN void tridiag_eigenvalues(const double* d, const double* e, int n, double* v, double* p, double* out, const Par& par) {
  // Gershgorin interval and scale
  double glo = d[0] - fabs(e[0]), ghi = d[0] + fabs(e[0]);
  for (int i = 1; i < n; ++i) {
    const double r = fabs(e[i - 1]) + ((i < n - 1) ? fabs(e[i]) : 0.0);
    glo = fmin(glo, d[i] - r);
    ghi = fmax(ghi, d[i] + r);
  }
  const double tnorm = fmax(fabs(glo), fabs(ghi));
  if (!(tnorm > 0.0)) {   // the zero matrix
    for (int i = par.tid(); i < n; i += par.nt()) out[i] = 0.0;
    par.sync();
    return;
  }
  // work on T / tnorm (entries in [-1, 1], spectrum in [-3, 3]): v holds d / tnorm, p holds (e / tnorm)^2
  const double inv = 1.0 / tnorm;
  par.sync();
  for (int i = par.tid(); i < n; i += par.nt()) {
    v[i] = d[i] * inv;
    const double es = e[i] * inv;
    p[i] = es * es;
  }
  par.sync();
  const double slo = glo * inv - 2.0 * 2.3e-16 * n - 1e-290, shi = ghi * inv + 2.0 * 2.3e-16 * n + 1e-290;
  // eigenvalue with ascending index kidx: bisection on #{eigenvalues < x} (Sturm count)
  for (int kidx = par.tid(); kidx < n; kidx += par.nt()) {
    double lo = slo, hi = shi;
    for (int it = 0; it < 120; ++it) {
      const double mid = 0.5 * (lo + hi);
      if (mid <= lo || mid >= hi) break;  // interval is one ulp wide
      if (sturm_count(v, p, n, mid) <= kidx)
        lo = mid;
      else
        hi = mid;
      if (hi - lo <= 4.0e-16) break;
    }
    out[n - 1 - kidx] = 0.5 * (lo + hi) * tnorm;
  }
  par.sync();
}

// a: n x n symmetric, row-major, lda (destroyed).  d[n], e[n], v[n], p[n]: group-visible scratch.
// out[n]: eigenvalues in DESCENDING order.
template <class Par>
RVT_HDN void sym_eigenvalues_tridiag(double* a, int n, int lda, double* d, double* e, double* v, double* p,
                                     double* out, const Par& par) {
  if (n == 1) {
    if (par.tid() == 0) out[0] = a[0];
    par.sync();
    return;
  }
  householder_tridiag(a, n, lda, d, e, v, p, par);
  tridiag_eigenvalues(d, e, n, v, p, out, par);
}

// Rank-sort ev[0..n) into out[0..n) in DESCENDING order (ties keep index order).
template <class Par>
RVT_HDN void sort_descending(const double* ev, int n, double* out, const Par& par) {
  for (int i = par.tid(); i < n; i += par.nt()) {
    double v = ev[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      double u = ev[j];
      rank += (u > v) || (u == v && j < i);
    }
    out[rank] = v;
  }
  par.sync();
}

}  // namespace rvt
