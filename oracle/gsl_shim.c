/* oracle/gsl_shim.c -- TEST INFRASTRUCTURE ONLY.
 * Doors onto the GSL 1.16 that the reference vendors as third/gsl-1.16.tar.gz (built offline by
 * oracle/Makefile from that tarball into a scratch dir; only the resulting oracle/_ref/libgsl_ref.so
 * is kept).  These are the GSL entry points on the hot path:
 *   gsl_ran_beta_pdf        src/Model.h:2652, :2807
 *   gsl_cdf_chisq_Q         regression/LinearRegressionScoreTest.cpp:259, regression/SkatO.cpp:423
 *   gsl_cdf_chisq_Qinv      regression/SkatO.cpp:431
 *   gsl_cdf_chisq_P         regression/SkatO.cpp:335
 *   gsl_ran_chisq_pdf       regression/SkatO.cpp:319, :335
 *   gsl_integration_qags    regression/GSLIntegration.cpp:37-49 (limit 1000, error handler off)
 */
#include <gsl/gsl_cdf.h>
#include <gsl/gsl_errno.h>
#include <gsl/gsl_integration.h>
#include <gsl/gsl_randist.h>

double ref_gsl_ran_beta_pdf(double x, double a, double b) { return gsl_ran_beta_pdf(x, a, b); }
double ref_gsl_cdf_chisq_Q(double x, double nu) { return gsl_cdf_chisq_Q(x, nu); }
double ref_gsl_cdf_chisq_P(double x, double nu) { return gsl_cdf_chisq_P(x, nu); }
double ref_gsl_cdf_chisq_Qinv(double q, double nu) { return gsl_cdf_chisq_Qinv(q, nu); }
double ref_gsl_ran_chisq_pdf(double x, double nu) { return gsl_ran_chisq_pdf(x, nu); }

/* Integration::integrateLU as configured by SkatO::Fit (regression/SkatO.cpp:236-242) */
int ref_gsl_qags(double (*f)(double, void*), void* params, double lb, double ub, double epsabs,
                 double epsrel, int limit, double* result, double* abserr, int* neval_intervals) {
  gsl_function F;
  F.function = f;
  F.params = params;
  gsl_set_error_handler_off();
  gsl_integration_workspace* w = gsl_integration_workspace_alloc(limit);
  int ret = gsl_integration_qags(&F, lb, ub, epsabs, epsrel, limit, w, result, abserr);
  if (neval_intervals) *neval_intervals = (int)w->size;
  gsl_integration_workspace_free(w);
  return ret;
}
