// skato_fast.cuh -- SKAT-O at sweep-compatible speed (K4, second generation).
//
// Same function as skato_tail.cuh (regression/SkatO.cpp:101-281 on the M x M sufficient statistics), same numbers to
// rounding, organised around the two things that made the first version 40x slower than SKAT:
//
//  1. ELEVEN of its twelve eigen-solves only fed skato_moment(), i.e. the power sums  sum lambda^k, k = 1..4, of the
//     spectrum of  K_rho = a^2 W + a b (1 c' + c 1') + b^2 tau 1 1'  (W = Z1'Z1, c = W 1, tau = 1'W 1; skato_tail.cuh header).
//     Power sums are traces: with F = [1 c] and S = [[b^2 tau, a b], [a b, 0]],  K_rho = a^2 W + F S F'  and
//         tr K_rho^k  expands into  tr W^j  (j <= 4)  and 2 x 2 products of  S  with  G_j = F' W^j F,  j = 0..3,
//     whose entries are  n, tau, 1'W^2 1, .., 1'W^5 1.  ONE product W^2 (fully parallel, no barrier chain) and three
//     mat-vecs replace eleven Householder + bisection solves.  The reference drops eigenvalues below mean/1e5
//     (SkatO.cpp:350-382); the traces are used for a rho only when  lambda_min(W) (1 - rho)  -- a lower bound of
//     lambda_min(K_rho), K_rho being congruent to W through R_rho^1/2 -- proves that nothing is dropped; otherwise that
//     rho takes the eigen-solve as before (typically rho = 0.999 only).
//  2. the quadrature evaluates ~1 100 Davies p-values per gene.  k_skato_qags gives every Kronrod node of BOTH halves of
//     a bisection its own thread (42 of 64) running the serial product-form Davies of davies_fast.cuh -- no shuffles, no
//     barriers inside an evaluation, c-independent work done once per gene -- and many genes per SM (the kernel needs
//     ~3 KB of shared memory, where the statistics kernel needs ~50 KB).
#pragma once
#include "davies_fast.cuh"
#include "skato_tail.cuh"

namespace rvt {

constexpr int kSkatoMaxLam = 64;   // == kTileRows (common.cuh; asserted in finalize.cuh): genes of one tile

struct SkatoJob {
  int run;          // 1: the quadrature has to run (k_skato_qags); 0: the fields below are final
  int n_lam, min_index, ok;
  double Q, rho, pvalue;                  // final when run == 0
  double Qs[11], pvals[11], Qs_minP[11], taus[11], rhos[11];
  double MuQ, VarQ, VarZeta, Df, lam_sum, minP;
  double lam[kSkatoMaxLam];                // spectrum of Z(I-M)Z', kept, descending
};

// SkatO.cpp:383-416 from the power sums c0 = sum l, c1 = sum l^2, c2 = sum l^3, c3 = sum l^4
RVT_HDN SkatoMoment skato_moment_sums(double c0, double c1, double c2, double c3) {
  const double sigmaQ = sqrt(2 * c1);
  const double s1 = c2 / c1 / sqrt(c1);
  const double s2 = c3 / (c1 * c1);
  double l;
  if (s1 * s1 > s2) {
    const double a = 1 / (s1 - sqrt(s1 * s1 - s2));
    const double d = (s1 * a - 1.0 * a * a);
    l = a * a - 2 * d;
  } else {
    l = 1. / s2;
  }
  SkatoMoment m;
  m.muQ = c0;
  m.varQ = sigmaQ * sigmaQ;
  m.df = l;
  return m;
}

struct M2 {   // 2 x 2
  double a, b, c, d;
};
RVT_HD M2 m2mul(const M2& x, const M2& y) { return M2{x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d}; }
RVT_HD double m2tr(const M2& x) { return x.a + x.d; }

struct SkatoTraces {
  double n, tau, s2, s3, s4, s5;   // 1'W^j 1, j = 0..5
  double t1, t2, t3, t4;           // tr W^j
};

// power sums of the spectrum of  alpha W + F S F'  (header, item 1)
RVT_HDN void skato_trace_sums(const SkatoTraces& T, double alpha, double beta, double gamma, double* c /*4*/) {
  const M2 S{gamma, beta, beta, 0.0};
  const M2 A0 = m2mul(S, M2{T.n, T.tau, T.tau, T.s2});
  const M2 A1 = m2mul(S, M2{T.tau, T.s2, T.s2, T.s3});
  const M2 A2 = m2mul(S, M2{T.s2, T.s3, T.s3, T.s4});
  const M2 A3 = m2mul(S, M2{T.s3, T.s4, T.s4, T.s5});
  const M2 A00 = m2mul(A0, A0);
  const double a2 = alpha * alpha;
  c[0] = alpha * T.t1 + m2tr(A0);
  c[1] = a2 * T.t2 + 2.0 * alpha * m2tr(A1) + m2tr(A00);
  c[2] = a2 * alpha * T.t3 + 3.0 * a2 * m2tr(A2) + 3.0 * alpha * m2tr(m2mul(A0, A1)) + m2tr(m2mul(A00, A0));
  c[3] = a2 * a2 * T.t4 + 4.0 * a2 * alpha * m2tr(A3) + 4.0 * a2 * m2tr(m2mul(A0, A2)) + 2.0 * a2 * m2tr(m2mul(A1, A1)) +
         4.0 * alpha * m2tr(m2mul(A00, A1)) + m2tr(m2mul(A00, A00));
}

// Everything of SkatO::Fit before the quadrature.  Wm: M x M (lda) = Z1'Z1, kept intact.  Km: M x M scratch (lda).
// v[M] = w_j (g_j'r).  lam_min_w: smallest eigenvalue of Wm when ALL M are known to be positive, else <= 0 (then every
// rho takes the eigen-solve).  ev/e/vv/pp/lamz/c: group-visible scratch of >= M + 2 doubles each.  job: group-visible.
// Every thread returns the same value: 1 = job->run set (quadrature needed), 0 = job holds the final result.
template <class Par>
RVT_HDN int skato_prepare(const double* Wm, double* Km, int M, int lda, const double* v, double s2, double lam_min_w, double* ev,
                          double* e, double* vv, double* pp, double* lamz, double* c, int* th, SkatoJob* job, const Par& par) {
  if (par.tid() == 0) {
    job->run = 0;
    job->ok = 0;
    job->Q = 0;
    job->rho = 0;
    job->pvalue = -999.0;
    job->n_lam = 0;
    job->min_index = 0;
  }
  par.sync();
  if (M == 1) {   // FitSKAT, SkatO.cpp:60-99 / :118-120
    const double Q = v[0] * v[0] / s2 / 2.0;
    const double lam1 = Wm[0];
    if (lam1 > 0.0) {   // (else getEigen fails: numNonZero == 0)
      int fault = 0;
      const double p = mixchisq_pvalue(&lam1, 1, Q, th, &fault, par);
      if (par.tid() == 0) {
        job->Q = Q;
        job->rho = 0.0;
        job->pvalue = p;
        job->ok = 1;
      }
    }
    par.sync();
    return 0;
  }
  // c = W 1, d = W c, f = W d   (d in e[], f in vv[])
  for (int k = par.tid(); k < M; k += par.nt()) {
    double s = 0.0;
    for (int j = 0; j < M; ++j) s += Wm[j * lda + k];
    c[k] = s;
  }
  par.sync();
  for (int k = par.tid(); k < M; k += par.nt()) {
    double s = 0.0;
    for (int j = 0; j < M; ++j) s += Wm[k * lda + j] * c[j];
    e[k] = s;
  }
  par.sync();
  for (int k = par.tid(); k < M; k += par.nt()) {
    double s = 0.0;
    for (int j = 0; j < M; ++j) s += Wm[k * lda + j] * e[j];
    vv[k] = s;
  }
  par.sync();
  SkatoTraces T;
  double tot = 0.0, sv = 0.0, sv2 = 0.0, su2 = 0.0;
  T.s2 = T.s3 = T.s4 = T.s5 = T.t1 = 0.0;
  for (int k = 0; k < M; ++k) {
    tot += c[k];
    sv += v[k];
    sv2 += v[k] * v[k];
    su2 += (c[k] / M) * (c[k] / M);
    T.s2 += c[k] * c[k];
    T.s3 += c[k] * e[k];
    T.s4 += e[k] * e[k];
    T.s5 += e[k] * vv[k];
    T.t1 += Wm[k * lda + k];
  }
  T.n = (double)M;
  T.tau = tot;
  const double z_norm = tot / ((double)M * (double)M);
  // W^2 into Km; tr W^2 = |W|_F^2, tr W^3 = sum W^2 o W, tr W^4 = |W^2|_F^2
  {
    double p2 = 0.0, p3 = 0.0, p4 = 0.0, dummy = 0.0;
    for (int idx = par.tid(); idx < M * M; idx += par.nt()) {
      const int i = idx / M, k = idx - i * M;
      double s = 0.0;
      const double* wi = Wm + i * lda;
      for (int j = 0; j < M; ++j) s += wi[j] * Wm[j * lda + k];
      const double w = wi[k];
      p2 += w * w;
      p3 += s * w;
      p4 += s * s;
    }
    par.allreduce4(p2, p3, p4, dummy);
    T.t2 = p2;
    T.t3 = p3;
    T.t4 = p4;
  }
  SkatoMoment mom[11];
  double Qs[11], rhos[11], taus[11];
  for (int i = 0; i < 11; ++i) {
    const double rho_o = (double)i / 10;
    const double rho = (rho_o > 0.999) ? 0.999 : rho_o;   // capRhos, SkatO.cpp:436-446
    rhos[i] = rho;
    Qs[i] = ((1.0 - rho) * sv2 + rho * sv * sv) / s2 / 2.0;
    taus[i] = (double)M * (double)M * rho * z_norm + (1.0 - rho) * su2 / z_norm;
    const double a = sqrt(1.0 - rho), b = (sqrt(1.0 - rho + rho * M) - a) / M;
    double cs[4];
    skato_trace_sums(T, a * a, a * b, b * b * tot, cs);
    // all M eigenvalues of K_rho are kept when the smallest one is provably >= mean / 1e5 (factor 2 of margin)
    const bool all_kept = (lam_min_w > 0.0) && (cs[0] > 0.0) && (lam_min_w * (1.0 - rho) >= 2.0 * (cs[0] / M) / 100000);
    if (all_kept) {
      mom[i] = skato_moment_sums(cs[0], cs[1], cs[2], cs[3]);
    } else {
      par.sync();
      for (int idx = par.tid(); idx < M * M; idx += par.nt()) {
        const int j = idx / M, k = idx - j * M;
        Km[j * lda + k] = a * a * Wm[j * lda + k] + a * b * (c[j] + c[k]) + b * b * tot;
      }
      par.sync();
      sym_eigenvalues_tridiag(Km, M, lda, ev, e, vv, pp, lamz, par);
      const int keep = skato_keep(lamz, M);
      if (keep < 0) return 0;   // (uniform: lamz is shared)
      mom[i] = skato_moment(lamz, keep);
      par.sync();
    }
  }
  par.sync();
  // Z(I-M)Z' = Wm - (c c')/(M^2 z_norm)
  double vz_part = 0.0, dummy = 0.0;
  for (int idx = par.tid(); idx < M * M; idx += par.nt()) {
    const int j = idx / M, k = idx - j * M;
    const double zmz = (c[j] / M) * (c[k] / M) / z_norm;
    const double zimz = Wm[j * lda + k] - zmz;
    Km[j * lda + k] = zimz;
    vz_part += zmz * zimz;
  }
  par.allreduce2(vz_part, dummy);
  sym_eigenvalues_tridiag(Km, M, lda, ev, e, vv, pp, lamz, par);
  const int nl = skato_keep(lamz, M);
  if (nl < 0) return 0;
  const double VarZeta = 4.0 * vz_part;
  double l1 = 0, l2 = 0, l4 = 0;
  for (int i = 0; i < nl; ++i) {
    const double l = lamz[i];
    l1 += l;
    l2 += l * l;
    l4 += (l * l) * (l * l);
  }
  const double VarQ = 2.0 * l2 + VarZeta;
  const double KerQ = l4 / l2 / l2 * 12;
  // per-rho p-values by moment matching (one thread each), the minimum, and its quantiles (SkatO.cpp:206-233)
  for (int i = par.tid(); i < 11; i += par.nt()) {
    const double qn = (Qs[i] - mom[i].muQ) / sqrt(mom[i].varQ) * sqrt(2. * mom[i].df) + mom[i].df;
    job->pvals[i] = chisq_q(qn, mom[i].df);
  }
  par.sync();
  int minIndex = 0;
  double minP = job->pvals[0];
  for (int i = 1; i < 11; ++i)
    if (job->pvals[i] < minP) {
      minP = job->pvals[i];
      minIndex = i;
    }
  for (int i = par.tid(); i < 11; i += par.nt()) {
    const double q_org = chisq_qinv(minP, mom[i].df);
    job->Qs_minP[i] = (q_org - mom[i].df) / sqrt(2. * mom[i].df) * sqrt(mom[i].varQ) + mom[i].muQ;
    job->Qs[i] = Qs[i];
    job->taus[i] = taus[i];
    job->rhos[i] = rhos[i];
  }
  for (int i = par.tid(); i < nl; i += par.nt()) job->lam[i] = lamz[i];
  if (par.tid() == 0) {
    job->run = 1;
    job->n_lam = nl;
    job->min_index = minIndex;
    job->MuQ = l1;
    job->lam_sum = l1;
    job->VarQ = VarQ;
    job->VarZeta = VarZeta;
    job->Df = 12 / KerQ;
    job->minP = minP;
  }
  par.sync();
  return 1;
}

// the Davies integrand of SkatO.cpp:303-321 at one node, on the prepared spectrum
RVT_HDN double skato_node_davies(const SkatoParams& P, const DaviesPre& pre, const int* th, double x) {
  double kappa = 0.0;
  for (int i = 0; i < 11; ++i) {
    const double v = (P.Qs_minP[i] - P.taus[i] * x) / (1.0 - P.rhos[i]);
    if (i == 0 || v < kappa) kappa = v;
  }
  double temp;
  if (kappa > P.lam_sum * 10000) {
    temp = 0.0;
  } else {
    const double Q = (kappa - P.MuQ) * sqrt(P.VarQ - P.VarZeta) / sqrt(P.VarQ) + P.MuQ;
    int fault = 0;
    temp = (P.n_lam == 1) ? liu_pvalue(P.lam, 1, Q) : mixchisq_pvalue_fast(P.lam, pre, th, Q, &fault);
    if (temp <= 0.0 || temp == 1.0) temp = liu_pvalue(P.lam, P.n_lam, Q);
  }
  return (1.0 - temp) * chisq_pdf(x, 1.0);
}

// p-value from the integral, SkatO.cpp:257-277 (nRho = 11 -> multi = 3)
RVT_HDN double skato_final_p(double integral, double minP, const double* pvals) {
  double pvalue = 1.0 - integral;
  if (pvalue <= 0) {
    const double p3 = minP * 3;
    if (pvalue < p3) pvalue = p3;
  }
  if (pvalue == 0.0) {
    pvalue = pvals[0];
    for (int i = 1; i < 11; ++i)
      if (pvals[i] > 0 && pvals[i] < pvalue) pvalue = pvals[i];
  }
  return pvalue;
}

RVT_HDN void skato_params_from_job(const SkatoJob& job, const double* lam, SkatoParams* P) {
  for (int i = 0; i < 11; ++i) {
    P->Qs_minP[i] = job.Qs_minP[i];
    P->taus[i] = job.taus[i];
    P->rhos[i] = job.rhos[i];
  }
  P->MuQ = job.MuQ;
  P->VarQ = job.VarQ;
  P->VarZeta = job.VarZeta;
  P->Df = job.Df;
  P->lam_sum = job.lam_sum;
  P->lam = lam;
  P->n_lam = job.n_lam;
}

// Serial form of the quadrature (host-check build and documentation of what k_skato_qags does in parallel): both halves
// of a bisection are sampled before either is handed to the machine -- the machine cannot tell.
RVT_HDN void skato_quadrature_serial(const SkatoJob& job, const QagsWork& work, int* th /* n_lam ints */, SkatoOut* out) {
  SkatoParams P;
  skato_params_from_job(job, job.lam, &P);
  DaviesPre pre;
  pre.degenerate = 0;
  if (job.n_lam >= 2) davies_prepare(job.lam, job.n_lam, 10000, 0.000001, th, &pre);
  QagsMachine mach;
  double fv[42];
  int st = 0;
  for (int pass = 0; pass < 2; ++pass) {   // 0: Davies integrand, 1: Liu integrand (only when the first failed)
    mach.init(work, 0.0, 40.0, 1e-25, 0.0001220703);
    double lo, hi;
    while (mach.want(&lo, &hi)) {
      const int two = mach.stage == 1;
      const double lo2 = mach.a2, hi2 = mach.b2;
      for (int t = 0; t < (two ? 42 : 21); ++t) {
        const double l = t < 21 ? lo : lo2, h = t < 21 ? hi : hi2;
        const double x = 0.5 * (l + h) + 0.5 * (h - l) * gk21_node(t % 21);
        fv[t] = pass == 0 ? skato_node_davies(P, pre, th, x) : skato_integrand_liu(P, x);
      }
      mach.give(gk21_combine(fv, lo, hi));
      if (two) mach.give(gk21_combine(fv + 21, lo2, hi2));
    }
    st = mach.status;
    if (st == 0) break;
  }
  out->Q = job.Qs[job.min_index];
  out->rho = (job.rhos[job.min_index] >= 0.999) ? 1.0 : job.rhos[job.min_index];   // uncapRhos, :447-455
  out->pvalue = skato_final_p(mach.result, job.minP, job.pvals);
  out->ok = 1;
  out->timed_out = 0;
}

#if defined(__CUDACC__)
constexpr int kQagsThreads = 64;

// One CTA per gene; thread t < 42 owns Kronrod node t % 21 of half t / 21 of the current bisection.
#ifndef RVT_QAGS_MINBLOCKS
#define RVT_QAGS_MINBLOCKS 12
#endif
__global__ void __launch_bounds__(kQagsThreads, RVT_QAGS_MINBLOCKS)
k_skato_qags(const SkatoJob* __restrict__ jobs, int n_genes, QagsScratch* __restrict__ qags, rvt_gene_result* __restrict__ res,
             const int* __restrict__ out_index, long long wd_cycles) {
  __shared__ SkatoParams P;
  __shared__ double s_lam[kSkatoMaxLam];
  __shared__ int s_th[kSkatoMaxLam];
  __shared__ DaviesPre pre;
  __shared__ QagsMachine mach;
  __shared__ double fv[42], bc[5];
  const int g = blockIdx.x, tid = threadIdx.x;
  if (g >= n_genes) return;
  const SkatoJob& job = jobs[g];
  rvt_gene_result* dst = &res[out_index ? out_index[g] : g];
  if (job.run != 1) {
    if (tid == 0) {
      dst->skato_ok = job.ok;
      dst->skato_Q = job.Q;
      dst->skato_rho = job.rho;
      dst->skato_p = job.pvalue;
    }
    return;
  }
  const int nl = min(max(job.n_lam, 0), kSkatoMaxLam);
  for (int i = tid; i < nl; i += kQagsThreads) s_lam[i] = job.lam[i];
  if (tid == 0) skato_params_from_job(job, s_lam, &P);
  __syncthreads();
  if (tid == 0) {
    pre.degenerate = 0;
    if (nl >= 2) davies_prepare(s_lam, nl, 10000, 0.000001, s_th, &pre);
  }
  QagsWork work{qags[g].a, qags[g].b, qags[g].r, qags[g].e, qags[g].order, qags[g].level, kQagsLimit};
  work.deadline = wd_cycles > 0 ? clock64() + wd_cycles : 0;
  __syncthreads();
  int st = 0;
  for (int pass = 0; pass < 2; ++pass) {
    if (tid == 0) mach.init(work, 0.0, 40.0, 1e-25, 0.0001220703);
    __syncthreads();
    for (;;) {
      if (tid == 0) {
        if (mach.stage != 3 && mach.w.deadline != 0 && clock64() > mach.w.deadline) {   // watchdog, status 8
          mach.status = 8;
          mach.stage = 3;
        }
        double lo = 0, hi = 0;
        const bool more = mach.want(&lo, &hi);
        bc[0] = more ? (mach.stage == 1 ? 2.0 : 1.0) : 0.0;
        bc[1] = lo;
        bc[2] = hi;
        bc[3] = mach.a2;
        bc[4] = mach.b2;
      }
      __syncthreads();
      const int nint = (int)bc[0];
      if (nint == 0) break;
      if (tid < 21 * nint) {
        const int h = tid >= 21;
        const double lo = h ? bc[3] : bc[1], hi = h ? bc[4] : bc[2];
        const double x = 0.5 * (lo + hi) + 0.5 * (hi - lo) * gk21_node(tid - 21 * h);
        fv[tid] = pass == 0 ? skato_node_davies(P, pre, s_th, x) : skato_integrand_liu(P, x);
      }
      __syncthreads();
      if (tid == 0) {
        mach.give(gk21_combine(fv, bc[1], bc[2]));
        if (nint == 2) mach.give(gk21_combine(fv + 21, bc[3], bc[4]));
      }
      __syncthreads();
    }
    st = mach.status;
    __syncthreads();
    if (st == 0 || st == 8) break;
  }
  if (tid == 0) {
    if (st == 8) {
      dst->skato_ok = 0;
      dst->status = RVT_GENE_TIMEOUT;
    } else {
      dst->skato_ok = 1;
      dst->skato_Q = job.Qs[job.min_index];
      dst->skato_rho = (job.rhos[job.min_index] >= 0.999) ? 1.0 : job.rhos[job.min_index];
      dst->skato_p = skato_final_p(mach.result, job.minP, job.pvals);
    }
  }
}

// Packed form: the 42 nodes of a bisection fill 1.3 warps, so the kernel above keeps a third of its lanes idle.  Here a
// CTA of 128 threads runs THREE quadratures side by side -- slot s owns lanes 42 s .. 42 s + 41 -- and is persistent: a slot
// whose gene is finished takes the next one from a global counter, so the lanes stay busy until the queue is empty.
// All slots step together (sample -> barrier -> give -> barrier); the serial parts (machine update, per-gene set-up)
// are done by the slot's first lane while the others wait at the barrier (~1 % of a step).
constexpr int kQagsSlots = 3;
constexpr int kQagsPackThreads = 128;

__global__ void __launch_bounds__(kQagsPackThreads, 6)
k_skato_qags_packed(const SkatoJob* __restrict__ jobs, int n_genes, QagsScratch* __restrict__ qags /* [gridDim.x * kQagsSlots] */,
                    rvt_gene_result* __restrict__ res, const int* __restrict__ out_index, long long wd_cycles,
                    unsigned int* __restrict__ next /* zeroed by the host */) {
  struct Slot {
    SkatoParams P;
    double lam[kSkatoMaxLam];
    int th[kSkatoMaxLam];
    DaviesPre pre;
    QagsMachine mach;
    double fv[42], bc[5];
    int gene;   // -1: no more work
    int pass;   // 0 Davies integrand, 1 Liu integrand
    int nint;   // intervals to sample in this step (0: nothing)
  };
  __shared__ Slot sl[kQagsSlots];
  __shared__ int s_live;
  const int tid = threadIdx.x;
  const int s = tid / 42, t = tid - 42 * s;
  const bool leader = s < kQagsSlots && t == 0;
  if (leader) sl[s].gene = -2;   // -2: idle, fetch
  if (tid == 0) s_live = 1;
  __syncthreads();
  for (;;) {
    if (leader) {
      Slot& S = sl[s];
      // fetch + set up the next gene(s) until one needs the quadrature or the queue is empty
      while (S.gene == -2) {
        const unsigned int g = atomicAdd(next, 1u);
        if (g >= (unsigned int)n_genes) {
          S.gene = -1;
          break;
        }
        const SkatoJob& job = jobs[g];
        rvt_gene_result* dst = &res[out_index ? out_index[g] : (int)g];
        if (job.run != 1) {
          dst->skato_ok = job.ok;
          dst->skato_Q = job.Q;
          dst->skato_rho = job.rho;
          dst->skato_p = job.pvalue;
          continue;
        }
        const int nl = min(max(job.n_lam, 0), kSkatoMaxLam);
        for (int i = 0; i < nl; ++i) S.lam[i] = job.lam[i];
        skato_params_from_job(job, S.lam, &S.P);
        S.pre.degenerate = 0;
        if (nl >= 2) davies_prepare(S.lam, nl, 10000, 0.000001, S.th, &S.pre);
        QagsScratch& q = qags[(size_t)blockIdx.x * kQagsSlots + s];
        QagsWork work{q.a, q.b, q.r, q.e, q.order, q.level, kQagsLimit};
        work.deadline = wd_cycles > 0 ? clock64() + wd_cycles : 0;
        S.mach.init(work, 0.0, 40.0, 1e-25, 0.0001220703);
        S.gene = (int)g;
        S.pass = 0;
      }
      S.nint = 0;
      if (S.gene >= 0) {
        QagsMachine& m = S.mach;
        if (m.stage != 3 && m.w.deadline != 0 && clock64() > m.w.deadline) {   // watchdog, status 8
          m.status = 8;
          m.stage = 3;
        }
        double lo = 0, hi = 0;
        if (m.want(&lo, &hi)) {
          S.nint = m.stage == 1 ? 2 : 1;
          S.bc[1] = lo;
          S.bc[2] = hi;
          S.bc[3] = m.a2;
          S.bc[4] = m.b2;
        }
      }
    }
    if (tid == 0) {
      // (written before the barrier by thread 0 only; the leaders of slots 1, 2 publish through sl[].gene)
    }
    __syncthreads();
    if (tid == 0) {
      int live = 0;
      for (int k = 0; k < kQagsSlots; ++k) live |= (sl[k].gene != -1);
      s_live = live;
    }
    if (s < kQagsSlots) {
      Slot& S = sl[s];
      if (t < 21 * S.nint) {
        const int h = t >= 21;
        const double lo = h ? S.bc[3] : S.bc[1], hi = h ? S.bc[4] : S.bc[2];
        const double x = 0.5 * (lo + hi) + 0.5 * (hi - lo) * gk21_node(t - 21 * h);
        S.fv[t] = S.pass == 0 ? skato_node_davies(S.P, S.pre, S.th, x) : skato_integrand_liu(S.P, x);
      }
    }
    __syncthreads();
    if (!s_live) break;
    if (leader && sl[s].gene >= 0) {
      Slot& S = sl[s];
      QagsMachine& m = S.mach;
      if (S.nint >= 1) m.give(gk21_combine(S.fv, S.bc[1], S.bc[2]));
      if (S.nint == 2) m.give(gk21_combine(S.fv + 21, S.bc[3], S.bc[4]));
      if (m.stage == 3) {
        const int st = m.status;
        if (st != 0 && st != 8 && S.pass == 0) {   // the Davies integrand failed: Liu integrand (SkatO.cpp:243-255)
          QagsWork work = m.w;
          m.init(work, 0.0, 40.0, 1e-25, 0.0001220703);
          S.pass = 1;
        } else {
          const SkatoJob& job = jobs[S.gene];
          rvt_gene_result* dst = &res[out_index ? out_index[S.gene] : S.gene];
          if (st == 8) {
            dst->skato_ok = 0;
            dst->status = RVT_GENE_TIMEOUT;
          } else {
            dst->skato_ok = 1;
            dst->skato_Q = job.Qs[job.min_index];
            dst->skato_rho = (job.rhos[job.min_index] >= 0.999) ? 1.0 : job.rhos[job.min_index];
            dst->skato_p = skato_final_p(m.result, job.minP, job.pvals);
          }
          S.gene = -2;
        }
      }
    }
    // (the next step's leader section runs before any lane touches fv / bc again: ordered by the barrier at its end)
  }
}
#endif

}  // namespace rvt
