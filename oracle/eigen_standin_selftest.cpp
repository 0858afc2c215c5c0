// oracle/eigen_standin_selftest.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" doors that run each operation of oracle/eigen_standin (the stand-in for the Eigen API the
// reference's sources use) on caller-supplied arrays, so that tests/test_eigen_standin.py can hold it against
// numpy.  The reference build (oracle/_ref/libskat_ref.so) is only as trustworthy as this header.
#include "third/eigen/Eigen/Dense"

using namespace Eigen;

template <class T>
static Matrix<T, Dynamic, Dynamic> load(const double* p, int r, int c) {
  Matrix<T, Dynamic, Dynamic> m(r, c);
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i) m(i, j) = (T)p[(size_t)j * r + i];
  return m;
}
template <class T>
static void store(const Dense<T>& m, double* p) {
  for (Index j = 0; j < m.cols(); ++j)
    for (Index i = 0; i < m.rows(); ++i) p[(size_t)j * m.rows() + i] = (double)m(i, j);
}

extern "C" {
// out = A' * diag(d) * B - C / s   (products, transpose, asDiagonal, scalar ops, unary minus)
void st_algebra(int n, int k, int m, const double* A, const double* d, const double* B, const double* C, double s,
                double* out) {
  MatrixXd a = load<double>(A, n, k), b = load<double>(B, n, m), c = load<double>(C, k, m);
  VectorXd dv = load<double>(d, n, 1);
  MatrixXd r = a.transpose() * dv.asDiagonal() * b + (-c) / s;
  store(r, out);
}
// inverse, LLT solve, LDLT solve, determinant, rank of a symmetric positive definite A (n x n) with rhs B (n x m)
void st_solvers(int n, int m, const double* A, const double* B, double* inv, double* llt, double* ldlt, double* L,
                double* det, int* rank) {
  MatrixXd a = load<double>(A, n, n), b = load<double>(B, n, m);
  store(a.inverse(), inv);
  store(a.llt().solve(b), llt);
  store(a.ldlt().solve(b), ldlt);
  LLT<MatrixXd> chol;
  chol.compute(a);
  store(chol.matrixL(), L);
  *det = a.determinant();
  *rank = (int)a.fullPivLu().rank();
}
// x = A.ldlt().solve(B) alone (A may be singular: Eigen applies the pseudo-inverse of D)
void st_ldlt(int n, int m, const double* A, const double* B, double* x) { store(load<double>(A, n, n).ldlt().solve(load<double>(B, n, m)), x); }
int st_rank(int r, int c, const double* A) { return (int)load<double>(A, r, c).fullPivLu().rank(); }
// eigenvalues (increasing) and eigenvectors of a symmetric matrix, in double and in float
void st_eigen(int n, const double* A, double* val64, double* vec64, double* val32) {
  MatrixXd a = load<double>(A, n, n);
  SelfAdjointEigenSolver<MatrixXd> es;
  es.compute(a);
  store(es.eigenvalues(), val64);
  store(es.eigenvectors(), vec64);
  MatrixXf af = a.cast<float>();
  SelfAdjointEigenSolver<MatrixXf> ef(af);
  store(ef.eigenvalues(), val32);
}
// reductions, broadcasts, arrays, blocks, comma initialiser, Map write-through, vector = row-vector transposition
void st_misc(int n, int m, const double* A, double* rowsum, double* colmean, double* centred, double* arr, double* blocks,
             double* comma, double* mapped, double* vec_from_row, double* scalars) {
  MatrixXd a = load<double>(A, n, m);
  store(a.rowwise().sum(), rowsum);
  store(a.colwise().mean(), colmean);
  RowVectorXd mean = a.colwise().mean();
  MatrixXd c = a.rowwise() - mean;
  store(c, centred);
  MatrixXd e = ((a.array() * a.array() - 1.0).square() / 2.0).matrix();
  store(e, arr);
  MatrixXd bl = MatrixXd::Zero(n, m);
  bl.col(0) = a.col(m - 1);
  bl.row(n - 1) = a.row(0);
  bl.diagonal() += a.col(0).eval().head(std::min(n, m));
  store(bl, blocks);
  MatrixXd cm(n, 2 * m);
  cm << a, c;
  store(cm, comma);
  Map<MatrixXd> mp(mapped, n, m);  // caller memory: assignment must write through
  mp = a * 2.0;
  VectorXd v;
  v = a.colwise().sum();  // 1 x m expression into a column vector
  store(v, vec_from_row);
  scalars[0] = a.sum();
  scalars[1] = a.squaredNorm();
  scalars[2] = a.norm();
  scalars[3] = a.trace();
  scalars[4] = (double)v.rows();
  scalars[5] = (double)v.cols();
  scalars[6] = a.minCoeff();
  scalars[7] = a.maxCoeff();
}
}

// ---- what regression/BoltLMM.cpp / BoltPlinkLoader.cpp use on top (blocks of blocks, write-through block arrays,
// projected column products, comparisons, thin SVD)
extern "C" void st_bolt_api(int n, int c, int k, const double* A /* (n+c) x k */, const double* B /* (n+c) x k */, const double* Zc /* n x c */,
                            double* projdot /* k */, double* projnorm /* k */, double* centred /* n */, double* proj /* (n+c) x k */,
                            double* colhead /* n */, int* all_lt, int* any_lt, double* sv /* c */, double* U /* n x c */, double* blkdiv /* n */) {
  using namespace Eigen;
  Map<const MatrixXd> a(A, n + c, k), b(B, n + c, k), z(Zc, n, c);
  MatrixXd v1 = a, v2 = b, Z = z;
  // projDot / projNorm2 exactly as BoltLMM.cpp:1088-1098, 1134-1137 spell them
  RowVectorXd pd = (v1.topRows(n).array() * v2.topRows(n).array()).eval().matrix().colwise().sum() -
                   (v1.bottomRows(c).array() * v2.bottomRows(c).array()).eval().matrix().colwise().sum();
  RowVectorXd pn = v1.topRows(n).cwiseAbs2().eval().colwise().sum() - v1.bottomRows(c).cwiseAbs2().eval().colwise().sum();
  for (int j = 0; j < k; ++j) { projdot[j] = pd(j); projnorm[j] = pn(j); }
  // preparePhenotype (BoltPlinkLoader.cpp:156-160): centre the top rows through a block array, then the covariate rows
  MatrixXd y = v1.col(0);
  double avg = y.topLeftCorner(n, 1).sum() / n;
  y.topLeftCorner(n, 1).array() -= avg;
  y.bottomLeftCorner(c, 1).noalias() = Z.transpose() * y.topLeftCorner(n, 1);
  for (int i = 0; i < n; ++i) centred[i] = y(i, 0);
  // projectCovariate (:266-271) on all columns
  MatrixXd m = v2;
  m.bottomRows(c).noalias() = Z.transpose() * m.topRows(n);
  for (int i = 0; i < (n + c) * k; ++i) proj[i] = m.data()[i];
  // g.col(0).head(N) = other.col(0)  (BoltLMM.cpp:449) and block / scalar (:569)
  MatrixXd g(n + c, 2);
  g.setZero();
  g.col(1).head(n) = v2.col(0).head(n);
  for (int i = 0; i < n; ++i) colhead[i] = g(i, 1);
  MatrixXd hd = v1.col(0) / 4.0;
  for (int i = 0; i < n; ++i) blkdiv[i] = hd(i, 0);
  *all_lt = (pn.array() < 1e300).all() ? 1 : 0;
  *any_lt = (pn.array() < -1e300).any() ? 1 : 0;
  BDCSVD<MatrixXd> svd(Z, ComputeThinU);
  for (int j = 0; j < c; ++j) sv[j] = svd.singularValues()[j];
  MatrixXd u = svd.matrixU().leftCols(c);
  for (int i = 0; i < n * c; ++i) U[i] = u.data()[i];
}
