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
