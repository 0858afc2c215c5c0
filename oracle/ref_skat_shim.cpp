// oracle/ref_skat_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" doors onto the REFERENCE's own per-gene statistics code, compiled UNMODIFIED from where
// it lies under /root/reference into oracle/_ref/libskat_ref.so by oracle/Makefile:
//   regression/Skat.cpp, SkatO.cpp, LinearRegression.cpp, LinearRegressionScoreTest.cpp,
//   EigenMatrixInterface.cpp, GSLIntegration.cpp, MixtureChiSquare.cpp (+ qfc.c), cdflib.cpp,
//   base/MathMatrix.cpp, base/MathVector.cpp (+ whatever of base/ they need to link),
// against (i) GSL 1.16 built from the tarball the reference vendors and (ii) oracle/eigen_standin --
// a from-scratch stand-in for the Eigen API those files use (Eigen itself is downloaded by the
// reference's build and is absent here; see eigen_standin/third/eigen/Eigen/Core for what that
// does and does not pin).  Nothing from the reference is copied into this repository: this file
// only calls its public classes (regression/Skat.h:7-45, SkatO.h:7-47, LinearRegression.h:10-60,
// LinearRegressionScoreTest.h:8-52).  All matrices cross the door column-major, like base/MathMatrix.h:34.
#include <cstring>

#include "base/MathMatrix.h"
#include "base/MathVector.h"
#include "regression/LinearRegression.h"
#include "regression/LinearRegressionScoreTest.h"
#include "regression/Skat.h"
#include "regression/SkatO.h"

namespace {
void to_matrix(const double* p, int r, int c, Matrix* m) {
  m->Dimension(r, c);
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i) (*m)(i, j) = p[(size_t)j * r + i];
}
void to_vector(const double* p, int n, Vector* v) {
  v->Dimension(n);
  for (int i = 0; i < n; ++i) (*v)[i] = p[i];
}
void from_matrix(const Matrix& m, double* p) {
  for (int j = 0; j < m.cols; ++j)
    for (int i = 0; i < m.rows; ++i) p[(size_t)j * m.rows + i] = m(i, j);
}
}  // namespace

extern "C" {
// Skat::Fit (regression/Skat.cpp:29-105) as SkatTest::fit calls it (src/Model.h:2700-2702);
// q_perm[k] = Skat::GetQFromNewResidual (Skat.cpp:107-116) for n_perm further residual vectors.
int ref_skat_fit(int N, int M, int C, const double* res, const double* v, const double* X,
                 const double* G, const double* w, double* Q, double* p, int n_perm,
                 const double* res_perm, double* q_perm) {
  Vector res_G, v_G, w_G;
  Matrix X_G, G_G;
  to_vector(res, N, &res_G);
  to_vector(v, N, &v_G);
  to_vector(w, M, &w_G);
  to_matrix(X, N, C, &X_G);
  to_matrix(G, N, M, &G_G);
  Skat skat;
  skat.Reset();
  const int rc = skat.Fit(res_G, v_G, X_G, G_G, w_G);
  *Q = skat.GetQ();
  *p = skat.GetPvalue();
  for (int k = 0; k < n_perm; ++k) {
    Vector r;
    to_vector(res_perm + (size_t)k * N, N, &r);
    q_perm[k] = skat.GetQFromNewResidual(r);
  }
  return rc;
}

// SkatO::Fit (regression/SkatO.cpp:497-516 -> :100-282) as SkatOTest::fit calls it (src/Model.h:2854-2858);
// type "C" (quantitative) or "D" (binary).
int ref_skato_fit(int N, int M, int C, const double* res, const double* v, const double* X,
                  const double* G, const double* w, const char* type, double* Q, double* rho,
                  double* p) {
  Vector res_G, v_G, w_G;
  Matrix X_G, G_G;
  to_vector(res, N, &res_G);
  to_vector(v, N, &v_G);
  to_vector(w, M, &w_G);
  to_matrix(X, N, C, &X_G);
  to_matrix(G, N, M, &G_G);
  SkatO skato;
  skato.Reset();
  const int rc = skato.Fit(res_G, v_G, X_G, G_G, w_G, type);
  *Q = skato.GetQ();
  *rho = skato.GetRho();
  *p = skato.GetPvalue();
  return rc;
}

// LinearRegression::FitLinearModel (regression/LinearRegression.cpp:20-69)
int ref_linear_fit(int N, int C, const double* X, const double* y, double* beta, double* resid,
                   double* predicted, double* sigma2, double* covB) {
  Matrix X_G;
  Vector y_G;
  to_matrix(X, N, C, &X_G);
  to_vector(y, N, &y_G);
  LinearRegression lr;
  const bool ok = lr.FitLinearModel(X_G, y_G);
  for (int i = 0; i < C; ++i) beta[i] = lr.GetCovEst()[i];
  for (int i = 0; i < N; ++i) {
    resid[i] = lr.GetResiduals()[i];
    predicted[i] = lr.GetPredicted()[i];
  }
  *sigma2 = lr.GetSigma2();
  from_matrix(lr.GetCovB(), covB);
  return ok ? 0 : -1;
}

// LinearRegressionScoreTest::FitNullModel + TestCovariate(Xnull, y, Xcol) -- the single-column form
// the burden tests use (regression/LinearRegressionScoreTest.cpp:27-113; callers src/Model.h CMCTest /
// ZegginiTest ::fit) -- and, when M > 1 or force_matrix, the matrix form (:173-263).
int ref_score_test(int N, int C, int M, const double* Xnull, const double* y, const double* Xcol,
                   int force_matrix, double* U, double* V, double* beta, double* stat, double* p,
                   double* sigma2, double* se_beta) {
  Matrix Xn;
  Vector y_G;
  to_matrix(Xnull, N, C, &Xn);
  to_vector(y, N, &y_G);
  LinearRegressionScoreTest st;
  if (!st.FitNullModel(Xn, y_G)) return -2;
  bool ok;
  if (M == 1 && !force_matrix) {
    Vector xc;
    to_vector(Xcol, N, &xc);
    ok = st.TestCovariate(Xn, y_G, xc);
  } else {
    Matrix xc;
    to_matrix(Xcol, N, M, &xc);
    ok = st.TestCovariate((const Matrix&)Xn, (const Vector&)y_G, (const Matrix&)xc);
  }
  from_matrix(st.GetU(), U);
  from_matrix(st.GetV(), V);
  from_matrix(st.GetBeta(), beta);
  *stat = st.GetStat();
  *p = st.GetPvalue();
  *sigma2 = st.GetSigma2();
  *se_beta = st.GetSEBeta(0);  // LinearRegressionScoreTest.cpp:365-376
  return ok ? 0 : -1;
}
}

// ------------------------------------------------------------------------------------------------
// A6: the permutation loop of SkatTest::fit (src/Model.h:2707-2717) with the reference's own
// permute() (src/LinearAlgebra.h:8-21: Fisher-Yates on glibc rand()), Permutation bookkeeping
// (src/Permutation.h:49-98: init/next/add/getPvalue, early stop at numPerm * alpha * 2) and
// Skat::GetQFromNewResidual (regression/Skat.cpp:107-116).
// ------------------------------------------------------------------------------------------------
#include <cmath>
#include <cstdlib>

#include "third/eigen/Eigen/Core"
#include "src/LinearAlgebra.h"
#include "src/Permutation.h"

extern "C" int ref_skat_perm(int N, int M, int C, const double* res, const double* v, const double* X,
                             const double* G, const double* w, int n_perm, double alpha, unsigned reseed,
                             double* stat, int* actual_perm, int* num_greater, int* num_equal,
                             double* p_perm, double* q_out) {
  Vector res_G, v_G, w_G;
  Matrix X_G, G_G;
  to_vector(res, N, &res_G);
  to_vector(v, N, &v_G);
  to_vector(w, M, &w_G);
  to_matrix(X, N, C, &X_G);
  to_matrix(G, N, M, &G_G);
  if (reseed) srand(reseed);  // a fresh process starts as after srand(1)
  Skat skat;
  skat.Reset();
  const int rc = skat.Fit(res_G, v_G, X_G, G_G, w_G);
  *stat = skat.GetQ();
  Permutation perm(n_perm, alpha);
  Vector permutedRes = res_G;
  perm.init(*stat);
  int k = 0, greater = 0, equal = 0;
  while (perm.next()) {
    permute(&permutedRes);
    const double s = skat.GetQFromNewResidual(permutedRes);
    perm.add(s);
    // Permutation keeps its counters private and only prints them: mirror add() for the caller
    if (s > *stat) ++greater;
    if (s == *stat) ++equal;
    if (q_out) q_out[k] = s;
    ++k;
  }
  *actual_perm = k;
  *num_greater = greater;
  *num_equal = equal;
  *p_perm = perm.getPvalue();
  return rc;
}

// ------------------------------------------------------------------------------------------------
// A13: FastLMM (regression/FastLMM.cpp): FitNullModel (:28-140: rotation by the kinship eigenvectors, the
// 101-point grid + Brent search for delta, beta / sigma2 / uResid / scaledK) and the score branch of
// TestCovariate (:215-249) for each of M variants.  U (N x N, column-major) and S (N) are the kinship
// eigenvectors / eigenvalues in float, as the reference holds them (regression/EigenMatrix.h:9-12).
// ------------------------------------------------------------------------------------------------
#include "regression/EigenMatrix.h"
#include "regression/FastLMM.h"

extern "C" int ref_fastlmm_score(int N, int C, int M, const double* X, const double* y, const float* U,
                                 const float* S, const double* G, double* delta, double* sigma_g2,
                                 double* beta, double* Ustat, double* Vstat, double* pvalue) {
  Matrix X_G, y_G;
  to_matrix(X, N, C, &X_G);
  to_matrix(y, N, 1, &y_G);
  EigenMatrix kU, kS;
  kU.mat.resize(N, N);
  kS.mat.resize(N, 1);
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < N; ++i) kU.mat(i, j) = U[(size_t)j * N + i];
  for (int i = 0; i < N; ++i) kS.mat(i, 0) = S[i];
  FastLMM lmm(FastLMM::SCORE, FastLMM::MLE);
  const int rc = lmm.FitNullModel(X_G, y_G, kU, kS);
  if (rc) return rc;
  *delta = lmm.GetDelta();
  *sigma_g2 = lmm.GetSigmaG2();
  Vector b;
  lmm.GetNullCovEst(&b);
  for (int i = 0; i < C && i < b.Length(); ++i) beta[i] = b[i];
  for (int k = 0; k < M; ++k) {
    Matrix g;
    to_matrix(G + (size_t)k * N, N, 1, &g);
    if (lmm.TestCovariate(X_G, y_G, g, kU, kS)) return -100 - k;
    Ustat[k] = lmm.GetUStat();
    Vstat[k] = lmm.GetVStat();
    pvalue[k] = lmm.GetPvalue();
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// A11: GenotypeCounter (src/GenotypeCounter.h:7-72, src/GenotypeCounter.cpp) with the exact HWE test
// (libsrc/snp_hwe.cpp:25-122) as MetaScoreTest::fitWithGivenGenotype uses it (src/Model.h:3210-3240).
// out = { nHomRef, nHet, nHomAlt, nMissing, callRate, AF, AC, HWE }
// ------------------------------------------------------------------------------------------------
#include "src/GenotypeCounter.h"

extern "C" void ref_genotype_counter(int N, const double* g, double* out) {
  GenotypeCounter c;
  for (int i = 0; i < N; ++i) c.add(g[i]);
  out[0] = c.getNumHomRef();
  out[1] = c.getNumHet();
  out[2] = c.getNumHomAlt();
  out[3] = c.getNumMissing();
  out[4] = c.getCallRate();
  out[5] = c.getAF();
  out[6] = c.getAC();
  out[7] = c.getHWE();
}

// ------------------------------------------------------------------------------------------------
// Binary traits: LogisticRegression::FitLogisticModel (regression/LogisticRegression.cpp:279-339)
// as SkatTest::fit uses it (src/Model.h:2673-2681: ynull = GetPredicted(), v = GetVariance()), and
// LogisticRegressionScoreTest::FitNullModel + TestCovariate(Matrix overload, :219-302) as CMCTest::fit uses it
// (src/Model.h:841-848).  The latter solves its m x m `SS` against a d x d identity (:292-295): with covariates
// (d > 1) that indexes out of bounds, so this door only runs it for d == 1 and returns -3 otherwise.
// ------------------------------------------------------------------------------------------------
#include "regression/LogisticRegression.h"
#include "regression/LogisticRegressionScoreTest.h"

extern "C" int ref_logistic_fit(int N, int C, const double* X, const double* y, int rounds, double* beta,
                                double* p, double* V, double* covB) {
  Matrix X_G;
  Vector y_G;
  to_matrix(X, N, C, &X_G);
  to_vector(y, N, &y_G);
  LogisticRegression lr;
  if (!lr.FitLogisticModel(X_G, y_G, rounds)) return -1;
  for (int i = 0; i < C; ++i) beta[i] = lr.GetCovEst()[i];
  for (int i = 0; i < N; ++i) {
    p[i] = lr.GetPredicted()[i];
    V[i] = lr.GetVariance()[i];
  }
  from_matrix(lr.GetCovB(), covB);
  return 0;
}

extern "C" int ref_logistic_score_test(int N, int C, const double* Xnull, const double* y, const double* Xcol,
                                       double* U, double* V, double* stat, double* p) {
  if (C != 1) return -3;
  Matrix Xn, xc;
  Vector y_G;
  to_matrix(Xnull, N, C, &Xn);
  to_vector(y, N, &y_G);
  to_matrix(Xcol, N, 1, &xc);
  LogisticRegressionScoreTest st;
  if (!st.FitNullModel(Xn, y_G, 100)) return -2;
  const bool ok = st.TestCovariate((const Matrix&)Xn, (const Vector&)y_G, (const Matrix&)xc);
  *U = st.GetU()(0, 0);
  *V = st.GetV()(0, 0);
  *stat = st.GetStat();
  *p = st.GetPvalue();
  return ok ? 0 : -1;
}
