// oracle/ref_bolt_shim.cpp -- TEST INFRASTRUCTURE ONLY.  extern "C" doors onto the REFERENCE's own BoltLMM
// (regression/BoltLMM.cpp + BoltPlinkLoader.cpp + libVcf/PlinkInputFile.cpp + base/IO.cpp ..., compiled UNMODIFIED from
// /root/reference by oracle/Makefile into oracle/_ref/libbolt_ref.so, against oracle/eigen_standin and the cnpy / samtools /
// bzip2 archives the reference vendors under third/).  Drives it the way src/Model.h's MetaScoreTest / MetaCovTest do for
// `--meta score,cov --boltPlink prefix`: FitNullModel(prefix, phenotype) once (src/Model.h:3560-3567), then
// TestCovariate(g) per variant (:3597-3606) and GetCovXX(g1, g2) per pair (src/Model.cpp:780-805).
// The null model's internals are private to BoltLMM::BoltLMMImpl; they come out the way the reference itself exports them:
// BOLTLMM_SAVE_NULL_MODEL=<file.npz> (BoltLMM.cpp:245-261: H_inv_y, H_inv_y_norm2, infStatCalibration, xVx_xx_ratio) and
// BOLTLMM_DEBUG=1 (the secant iterations "i = ..\tlogDelta = ..\tf = .." on stderr, :589-633), which this shim routes
// into a log file for the duration of the fit.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <unistd.h>
#include <fcntl.h>

#include "base/MathMatrix.h"
#include "regression/BoltLMM.h"
#include "regression/MatrixRef.h"
#include "regression/SaddlePointApproximation.h"

// useSaddlePoint is hard-wired to false (BoltLMM.cpp:160): these are linked but never reached.
SaddlePointApproximation::SaddlePointApproximation(const Eigen::MatrixXf& y, const Eigen::MatrixXf& mu, const Eigen::MatrixXf& resid)
    : y_(y), mu_(mu), resid_(resid) { abort(); }
int SaddlePointApproximation::calculatePvalue(const Eigen::MatrixXf&, float*) { abort(); return -1; }

static BoltLMM* g_bolt = NULL;

extern "C" {

void bolt_ref_free() { delete g_bolt; g_bolt = NULL; }

// prefix: PLINK fileset (+ optional prefix.covar); pheno: N values or NULL (then the .fam column, as pin_->getPheno());
// save_npz / log_path may be NULL.  Returns FitNullModel's return value.
int bolt_ref_fit(const char* prefix, const double* pheno, int n, const char* save_npz, const char* log_path, int binary) {
  bolt_ref_free();
  if (save_npz) setenv("BOLTLMM_SAVE_NULL_MODEL", save_npz, 1); else unsetenv("BOLTLMM_SAVE_NULL_MODEL");
  unsetenv("BOLTLMM_LOAD_NULL_MODEL");
  unsetenv("BOLTLMM_MINQUE");
  int saved = -1;
  if (log_path) {
    setenv("BOLTLMM_DEBUG", "1", 1);
    fflush(stderr);
    saved = dup(2);
    int fd = open(log_path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd >= 0) { dup2(fd, 2); close(fd); }
  } else {
    unsetenv("BOLTLMM_DEBUG");
  }
  g_bolt = new BoltLMM;
  if (binary) g_bolt->enableBinaryMode();
  int rc;
  if (pheno) {
    Matrix y;
    y.Dimension(n, 1);
    for (int i = 0; i < n; ++i) y(i, 0) = pheno[i];
    rc = g_bolt->FitNullModel(prefix, &y);
  } else {
    rc = g_bolt->FitNullModel(prefix, NULL);
  }
  if (saved >= 0) { fflush(stderr); dup2(saved, 2); close(saved); }
  return rc;
}

// out: af, U, V, effect, pvalue (BoltLMM.cpp:315-338)
int bolt_ref_test(const double* g, int n, double* out) {
  if (!g_bolt) return -1;
  Matrix x;
  x.Dimension(n, 1);
  for (int i = 0; i < n; ++i) x(i, 0) = g[i];
  int rc = g_bolt->TestCovariate(x);
  out[0] = g_bolt->GetAF();
  out[1] = g_bolt->GetU();
  out[2] = g_bolt->GetV();
  out[3] = g_bolt->GetEffect();
  out[4] = g_bolt->GetPvalue();
  return rc;
}

// both overloads of GetCovXX (BoltLMM.cpp:414-460): out[0] the vector<double> form, out[1] the FloatMatrixRef form
int bolt_ref_covxx(const double* g1, const double* g2, int n, double* out) {
  if (!g_bolt) return -1;
  std::vector<double> a(g1, g1 + n), b(g2, g2 + n);
  g_bolt->GetCovXX(a, b, &out[0]);
  std::vector<float> fa(a.begin(), a.end()), fb(b.begin(), b.end());
  FloatMatrixRef ra(fa.data(), n, 1), rb(fb.data(), n, 1);
  float f = 0;
  g_bolt->GetCovXX(ra, rb, &f);
  out[1] = f;
  return 0;
}

}  // extern "C"
