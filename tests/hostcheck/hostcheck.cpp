// tests/hostcheck/hostcheck.cpp -- TEST INFRASTRUCTURE ONLY.
// Compiles the product's __host__ __device__ math headers (rvtests_b200/csrc/*.cuh) with g++
// so that `pytest -m "not gpu"` can check their LOGIC against the oracle on the CPU.  The product
// never links or calls this; its GPU kernels instantiate the same templates with a CTA-wide Par.
#include <vector>
#include "../../rvtests_b200/csrc/eigen.cuh"
#include "../../rvtests_b200/csrc/skato_tail.cuh"
#include "../../rvtests_b200/csrc/skato_fast.cuh"
#include "../../rvtests_b200/csrc/permlogic.cuh"

extern "C" {
double hc_gamma_q(double a, double x) { return rvt::gamma_q(a, x); }
double hc_chisq_q(double x, double df) { return rvt::chisq_q(x, df); }
double hc_beta_weight(double f, double b1, double b2, int sq) { return rvt::beta_weight(f, b1, b2, sq != 0); }
double hc_liu(const double* lam, int n, double Q) { return rvt::liu_pvalue(lam, n, Q); }
double hc_mixchisq(const double* lam, int n, double Q, int* fault) {
  std::vector<int> th(n > 0 ? n : 1);
  rvt::SerialPar par;
  return rvt::mixchisq_pvalue(lam, n, Q, th.data(), fault, par);
}
double hc_qf(const double* lam, int n, double Q, int lim, double acc, int* fault) {
  std::vector<int> th(n > 0 ? n : 1);
  rvt::SerialPar par;
  return rvt::davies_qf(lam, n, Q, lim, acc, th.data(), fault, par);
}
// the serial product-form Davies of the SKAT-O quadrature (davies_fast.cuh): prepare once per spectrum, then evaluate
double hc_qf_fast(const double* lam, int n, double Q, int lim, double acc, int* fault) {
  std::vector<int> th(n > 0 ? n : 1);
  rvt::DaviesPre pre;
  rvt::davies_prepare(lam, n, lim, acc, th.data(), &pre);
  return rvt::davies_qf_fast(lam, pre, th.data(), Q, lim, acc, fault);
}
// many points c on one prepared spectrum (what the quadrature does)
void hc_qf_fast_many(const double* lam, int n, const double* Q, int nq, double* out, int* faults) {
  std::vector<int> th(n > 0 ? n : 1);
  rvt::DaviesPre pre;
  rvt::davies_prepare(lam, n, 10000, 0.000001, th.data(), &pre);
  for (int i = 0; i < nq; ++i) out[i] = rvt::davies_qf_fast(lam, pre, th.data(), Q[i], 10000, 0.000001, &faults[i]);
}
// SKAT-O through skato_prepare (trace moments) + the serial form of k_skato_qags.  out[0..3] = Q, rho, pvalue, ok;
// out[4] = 1 when the quadrature ran.  lam_min_w <= 0 forces the eigen-solve for every rho.
int hc_skato_fast(const double* Wm, int M, const double* v, double s2, double lam_min_w, double* out) {
  const int limit = 1000, lda = M;
  std::vector<double> Km((size_t)M * M), ev(M + 2), e(M + 2), vv(M + 2), pp(M + 2), lamz(M + 2), c(M + 2);
  std::vector<double> wa(limit), wb(limit), wr(limit), we(limit);
  std::vector<int> wo(limit), wl(limit), th(M + 2);
  rvt::QagsWork w{wa.data(), wb.data(), wr.data(), we.data(), wo.data(), wl.data(), limit};
  rvt::SerialPar par;
  rvt::SkatoJob job;
  const int run = rvt::skato_prepare(Wm, Km.data(), M, lda, v, s2, lam_min_w, ev.data(), e.data(), vv.data(), pp.data(), lamz.data(),
                                     c.data(), th.data(), &job, par);
  out[4] = run;
  if (!run) {
    out[0] = job.Q;
    out[1] = job.rho;
    out[2] = job.pvalue;
    out[3] = job.ok;
    return 0;
  }
  rvt::SkatoOut o;
  rvt::skato_quadrature_serial(job, w, th.data(), &o);
  out[0] = o.Q;
  out[1] = o.rho;
  out[2] = o.pvalue;
  out[3] = o.ok;
  return 0;
}

double hc_chisq_qinv(double q, double df) { return rvt::chisq_qinv(q, df); }

// QAGS state machine driven with a plain C callback (limit 1000 as GSLIntegration.cpp:7-15)
int hc_qags(double (*f)(double), double a, double b, double epsabs, double epsrel, double* result, double* abserr,
            int* n_intervals) {
  const int limit = 1000;
  std::vector<double> wa(limit), wb(limit), wr(limit), we(limit);
  std::vector<int> wo(limit), wl(limit);
  rvt::QagsWork w{wa.data(), wb.data(), wr.data(), we.data(), wo.data(), wl.data(), limit};
  rvt::QagsMachine m;
  m.init(w, a, b, epsabs, epsrel);
  double lo, hi, fv[21];
  while (m.want(&lo, &hi)) {
    const double c = 0.5 * (lo + hi), h = 0.5 * (hi - lo);
    for (int i = 0; i < 21; ++i) fv[i] = f(c + h * rvt::gk21_node(i));
    m.give(rvt::gk21_combine(fv, lo, hi));
  }
  *result = m.result;
  *abserr = m.abserr;
  if (n_intervals) *n_intervals = m.size;
  return m.status;
}

// SKAT-O tail on the M x M statistics.  out[0..3] = Q, rho, pvalue, ok
int hc_skato_tail(const double* Wm, int M, const double* v, double s2, double* out) {
  const int limit = 1000, lda = M;
  std::vector<double> Km((size_t)M * M), ev(M + 2), e(M + 2), vv(M + 2), pp(M + 2), lamz(M + 2), c(M + 2);
  std::vector<double> wa(limit), wb(limit), wr(limit), we(limit);
  std::vector<int> wo(limit), wl(limit), th(M + 2);
  rvt::QagsWork w{wa.data(), wb.data(), wr.data(), we.data(), wo.data(), wl.data(), limit};
  rvt::QagsMachine mach;
  double fv[21], bcast[3];
  rvt::SerialPar par;
  rvt::SkatoOut o = rvt::skato_tail(Wm, Km.data(), M, lda, v, s2, ev.data(), e.data(), vv.data(), pp.data(), lamz.data(),
                                    c.data(), &mach, w, fv, bcast, th.data(), 0, par);
  out[0] = o.Q;
  out[1] = o.rho;
  out[2] = o.pvalue;
  out[3] = o.ok;
  return 0;
}

// eigenvalues, descending, Householder + Sturm bisection (the path the kernels use)
int hc_eigen_tridiag(const double* a_in, int n, double* out) {
  std::vector<double> a(a_in, a_in + (size_t)n * n), d(n + 1), e(n + 1), v(n + 1), p(n + 1);
  rvt::SerialPar par;
  rvt::sym_eigenvalues_tridiag(a.data(), n, n, d.data(), e.data(), v.data(), p.data(), out, par);
  return 0;
}
// eigenvalues, descending, parallel-ordered Jacobi (cross-check)
int hc_eigen(const double* a_in, int n, double* out) {
  std::vector<double> a(a_in, a_in + (size_t)n * n), cs(n + 2), ev(n);
  rvt::SerialPar par;
  int sweeps = rvt::jacobi_eigenvalues(a.data(), n, n, cs.data(), par);
  for (int i = 0; i < n; ++i) ev[i] = a[(size_t)i * n + i];
  rvt::sort_descending(ev.data(), n, out, par);
  return sweeps;
}

// glibc rand() by polynomial jump-ahead (permlogic.cuh): the n values from stream position pos after srand(seed)
void hc_lfg_draws(unsigned seed, unsigned long long pos, int n, int* out) {
  uint32_t y0[2 * rvt::kLfgDeg - 1], w[2 * rvt::kLfgDeg - 1];
  rvt::lfg_seed_window(seed, y0);
  rvt::lfg_window_at(rvt::lfg_pow(pos + rvt::kLfgWarm), y0, w);
  uint32_t x[rvt::kLfgDeg];
  for (int k = 0; k < rvt::kLfgDeg; ++k) x[k] = w[k];
  for (int i = 0; i < n; ++i) {
    const int k = i % rvt::kLfgDeg;
    if (i >= rvt::kLfgDeg) x[k] += x[(k + rvt::kLfgDeg - 3) % rvt::kLfgDeg];
    out[i] = (int)(x[k] >> 1);
  }
}
// Fisher-Yates resolved as chains: root[i] for the shuffle driven by draws[0..n-2] (step s handles i = n-1-s)
void hc_fy_roots(const unsigned* draws, unsigned n, unsigned* root) {
  std::vector<uint32_t> head(n, rvt::kFyNil), link(n, rvt::kFyNil), j(n, 0);
  for (unsigned s = 0; s + 1 < n; ++s) {
    const unsigned i = n - 1 - s;
    j[i] = draws[s] % (i + 1);
    link[i] = head[j[i]];
    head[j[i]] = i;
  }
  for (unsigned i = 0; i < n; ++i) root[i] = rvt::fy_root(head.data(), link.data(), i, j[i]);
}
}
