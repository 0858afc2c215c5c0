// finalize.cuh -- K2/K3: everything after the genotype sweep, one CTA per gene, O(M^3), fp64.
//
// From the exact integer sums of the sweep it rebuilds, per gene,
//   flip-to-minor + monomorphic removal   src/DataConsolidator.cpp:46-69, 94-142 (done
//                                         algebraically: g' = 2-g is an affine map of the sums)
//   Beta weights                          src/Model.h:2644-2661 (with the caller-order AF lookup
//                                         of src/DataConsolidator.cpp:527-529 when af is given)
//   Q = sum_j w_j (g_j'r)^2               regression/Skat.cpp:47-52
//   K = W^1/2 (G'VG - G'VX (X'VX)^-1 X'VG) W^1/2,  V = sigma2 I     Skat.cpp:55-76
//   eigenvalues, top-down > 1e-30         Skat.cpp:84-98
//   Davies -> (p<=0 or p==1) -> Liu       Skat.cpp:100-103, regression/MixtureChiSquare.cpp
//   CMC / Zeggini score tests             regression/LinearRegressionScoreTest.cpp:209-261,
//                                         NonRefSite src/Model.h:894-900
#pragma once
#include "../../include/rvtests_b200.h"
#include "common.cuh"
#include "eigen.cuh"
#include "skato_fast.cuh"

#include <type_traits>

namespace rvt {
static_assert(kSkatoMaxLam == kTileRows, "SkatoJob::lam holds the spectrum of a one-tile gene");

// Threads per gene: the per-gene tail is a chain of short dependent fp64 phases (latency-bound), so
// what buys throughput is MANY resident genes per SM, not many threads per gene: 64 threads x <= 128
// registers and ~30 KB of shared memory allow 8 genes per SM (it was 3 with 128 threads and 60 KB).
// SKAT-O keeps 128 threads (one warp per Kronrod node group in its quadrature).
constexpr int kFinThreads = 64;
constexpr int kFinThreadsSkato = 128;
// dynamic shared memory, sized by the widest gene of the launch (Mmax): K (fp64, Mmax x kld, kld odd
// so that rows AND columns are bank-conflict free); the reduced gene x digit sums De (int64,
// Mmax x ER) live in the same bytes (they are dead before K is built); SKAT-O adds Wm = Z1'Z1.
static inline int fin_kld(int Mmax) { return Mmax | 1; }
static inline int fin_uk_off(int Mmax, int ER, bool skato) {
  const int k = Mmax * fin_kld(Mmax) * 8, de = Mmax * ER * 8;
  return (k > de ? k : de) + (skato ? k : 0);
}
static inline int fin_smem(int Mmax, int ER, bool skato) { return fin_uk_off(Mmax, ER, skato) + Mmax * kMaxC * 8; }
constexpr int kFinPhases = 6;  // debug cycle counters per gene

__device__ __forceinline__ long long recombine4(const long long* d) {
  return d[0] + (d[1] << 8) + (d[2] << 16) + (d[3] << 24);
}

// Step 7 of the per-gene tail, shared by k_finalize and k_wide_finalize: the two burden score tests
// (LinearRegressionScoreTest::TestCovariate with m = 1, regression/LinearRegressionScoreTest.cpp:229-261) from
// bur[which] = {U, S'S, S'Z[0..C)} (which: 0 zeggini, 1 cmc), and the result record.
__device__ inline void burden_and_store(rvt_gene_result* dst, int Mp, int bad, double Q, double p_fin, double p_dav, double p_liu, int fault,
                                        int r, double lam_max, const SkatoOut& so, const double (*bur)[2 + kMaxC], int nonref,
                                        const NullModel* __restrict__ nm, int status_override) {
  const int C = nm->C;
  const double sigma2 = nm->sigma2;
  rvt_gene_result o;
  memset(&o, 0, sizeof(o));
  o.m_poly = Mp;
  o.status = (Mp == 0) ? RVT_GENE_NA : RVT_GENE_OK;
  if (bad == 1) o.status = RVT_GENE_BADFLAGS;
  if (bad == 2) o.status = RVT_GENE_BADVALUE;
  o.Q = Q;
  o.p_skat = p_fin;
  o.p_davies = p_dav;
  o.p_liu = p_liu;
  o.davies_fault = fault;
  o.n_lambda = r;
  o.lambda_max = lam_max;
  o.skato_ok = so.ok;
  o.skato_Q = so.Q;
  o.skato_rho = so.rho;
  o.skato_p = so.pvalue;
  for (int which = 0; which < 2; ++which) {  // 0 zeggini, 1 cmc
    const double U = bur[which][0];
    double SS = bur[which][1];
    const double* SZ = &bur[which][2];
    double q = 0.0;
    for (int l = 0; l < C; ++l)
      for (int m = 0; m < C; ++m) q += SZ[l] * nm->xtx_inv[l * C + m] * SZ[m];
    SS -= q;
    double V = SS * sigma2;
    double stat = U * ((1.0 / SS) / sigma2) * U;
    int ok = (Mp > 0) && !(stat < 0.0) && (stat == stat);
    double p = ok ? chisq_q(stat, 1.0) : nan("");
    if (which == 0) {
      o.zeg_U = U; o.zeg_V = V; o.zeg_stat = stat; o.zeg_p = p; o.zeg_ok = ok;
    } else {
      o.cmc_U = U; o.cmc_V = V; o.cmc_stat = stat; o.cmc_p = p; o.cmc_ok = ok;
      o.cmc_nonref = nonref;
    }
  }
  if (status_override) o.status = status_override;
  *dst = o;
}

// Split form of the statistics (option "fin_split", default on, SKAT-O off): the all-in-one kernel is held to 6 CTAs per
// SM by the M x M matrix in shared memory, and at that occupancy its serial fp64 chains -- the Sturm bisection above all,
// ~60 % of its instructions -- leave the SM idle three cycles in four (ncu: 24 % issue-active, 0.27 eligible warps per
// cycle).  Only the FRONT needs the matrix: split reduction, K, Householder.  It hands a FinMid record (the tridiagonal
// form and a few scalars, 1.7 KB) to k_fin_sturm (bisection: 2 KB of shared memory, few registers, many CTAs per SM) and
// k_fin_tail (Davies / Liu / burden score tests).
struct FinMid {
  int Mp, bad, nonref, status;   // Mp < 0: the front already wrote the record (bad-value gene)
  double Q;
  double bur[2][2 + kMaxC];
  double d[kTileRows], e[kTileRows];   // tridiagonal form of K (Mp >= 2)
  double lam[kTileRows];                // eigenvalues, descending (k_fin_sturm; Mp == 1: K itself, by the front)
};

template <bool SKATO, bool FRONT = false>
__global__ void __launch_bounds__(SKATO ? kFinThreadsSkato : kFinThreads)
k_finalize(const GeneDesc* __restrict__ genes, int n_genes, int kld /* fin_kld(Mmax of this launch) */,
           int wm_off /* byte offset of Wm inside the dynamic shared memory (SKAT-O) */, int uk_off /* ... of Uk */,
           const uint8_t* __restrict__ rowflags,
           const double* __restrict__ af, const RowCounts* __restrict__ counts,
           const NullModel* __restrict__ nm, EngineParams prm, int S,
           const SweepPartial* __restrict__ parts, rvt_gene_result* __restrict__ res,
           long long* __restrict__ dbg /* nullable: [n_genes][kFinPhases] cycle counters */,
           SkatoJob* __restrict__ jobs /* nullable: [n_genes]; non-null enables SKAT-O: everything before the quadrature runs
                                          here, the quadrature itself in k_skato_qags (skato_fast.cuh), launched next */,
           const TailInput* __restrict__ tin /* nullable: [n_genes] pre-digested statistics (dosage path);
                                                then `genes`/`parts` are unused and res is indexed through out_index */,
           const int* __restrict__ out_index, FinMid* __restrict__ mid = nullptr /* FRONT: [n_genes] */) {
  static_assert(!(SKATO && FRONT), "the split form serves the statistics without SKAT-O");
  extern __shared__ __align__(16) uint8_t dyn[];
  constexpr int NT = SKATO ? kFinThreadsSkato : kFinThreads;
  double* K = reinterpret_cast<double*>(dyn);                // [Mmax][kld]
  long long* De = reinterpret_cast<long long*>(dyn);         // [Mmax][ER] gene x digit sums, dead before K is written
  double* Wm = reinterpret_cast<double*>(dyn + wm_off);      // [Mmax][kld], SKAT-O only
  double* Uk = reinterpret_cast<double*>(dyn + uk_off);      // [Mmax][C]  (X'X)^-1 B_k
  struct SkatoShared {   // SKAT-O only
    double c[kTileRows + 2], lamz[kTileRows + 2];
  };
  __shared__ typename std::conditional<SKATO, SkatoShared, int>::type s_sk;
  __shared__ double s_vw[kTileRows];
  __shared__ long long s_ajj[kTileRows];
  __shared__ double s_red[64];
  __shared__ double s_e[kTileRows + 2], s_v[kTileRows + 2], s_p[kTileRows + 2];
  __shared__ double s_ev[kTileRows], s_lam[kTileRows];
  __shared__ int s_th[(NT / 32) * kTileRows];   // Davies order scratch, one slice per warp
  __shared__ long long s_coll[kCollapseN];
  __shared__ long long s_craw[kTileRows];
  __shared__ int s_idx[kTileRows], s_flip[kTileRows];
  __shared__ double s_s[kTileRows], s_sw[kTileRows], s_B[kTileRows][kMaxC];
  __shared__ int s_Mp, s_bad;
  __shared__ double s_Q;

  const int g = blockIdx.x;
  if (g >= n_genes) return;
  const int tid = threadIdx.x;
  GeneDesc gd;
  if (tin)
    memset(&gd, 0, sizeof(gd));
  else
    gd = genes[g];
  const int M = gd.M;
  const int64_t N = nm->N;
  const int C = nm->C, ER = nm->ER;
  const double sigma2 = nm->sigma2;
  BlockPar par{s_red};

  const SweepPartial* __restrict__ gp = tin ? nullptr : parts + (size_t)g * S;
  long long t_ph = clock64();
  auto phase = [&](int k) {
    if (dbg && tid == 0) {
      long long now = clock64();
      dbg[(size_t)g * kFinPhases + k] = now - t_ph;
      t_ph = now;
    }
  };
  __shared__ double s_bur[2][2 + kMaxC];   // per burden test: U, SS (raw), SZ[]
  __shared__ int s_nonref;
  if (tin) {
    // pre-digested statistics: load what steps 1-4 would have built
    const TailInput& ti = tin[g];
    if (tid == 0) {
      s_Mp = ti.Mp;
      s_bad = 0;
      s_Q = ti.Q;
      s_nonref = ti.nonref;
      s_bur[0][0] = ti.zegU; s_bur[0][1] = ti.zegSS;
      s_bur[1][0] = ti.cmcU; s_bur[1][1] = ti.cmcSS;
      for (int l = 0; l < C; ++l) { s_bur[0][2 + l] = ti.zegSZ[l]; s_bur[1][2 + l] = ti.cmcSZ[l]; }
    }
    for (int idx = tid; idx < ti.Mp * ti.Mp; idx += NT) {
      const int i = idx / ti.Mp, k = idx - i * ti.Mp;
      K[i * kld + k] = ti.K[idx];
    }
    if (tid < ti.Mp) s_vw[tid] = ti.vw[tid];
    __syncthreads();
  } else {
  // 1. reduce the splits (int64): gene x digit columns and the diagonal; the gene x gene block is
  //    summed on the fly where K is built (each entry is needed exactly once)
  for (int idx = tid; idx < M * ER; idx += NT) {
    const int i = idx / ER, e = idx - i * ER;
    long long s = 0;
    for (int sp = 0; sp < S; ++sp) s += gp[sp].d[i][kTileRows + e];
    De[i * ER + e] = s;
  }
  if (tid < M) {
    long long s = 0;
    for (int sp = 0; sp < S; ++sp) s += gp[sp].d[tid][tid];
    s_ajj[tid] = s;
  }
  for (int i = tid; i < 2 * (ER + 1); i += NT) {
    long long s = 0;
    for (int sp = 0; sp < S; ++sp) s += parts[(size_t)g * S + sp].coll[i];
    s_coll[i] = s;
  }
  if (tid == 0) s_bad = 0;
  __syncthreads();
  phase(0);

  // 2. per-variant counts -> flip / monomorphic, cross-checked with the flags the sweep used
  if (tid < M) {
    // intercept = vector 1 (X column 0): fixed-point image of 1.0 is 2^e exactly
    long long cint = recombine4(&De[tid * ER + 4]);
    double cd = (double)cint * nm->scale[1];
    long long c = llrint(cd);
    long long ajj = s_ajj[tid];
    long long n2 = (ajj - c) / 2, n1 = c - 2 * n2, n0 = N - n1 - n2;
    int flip = c > N;
    int mono = (n0 == N) || (n1 == N) || (n2 == N);
    uint8_t expect = mono ? kRowSkip : (flip ? kRowFlipped : kRowNormal);
    if (rowflags[gd.var0 + tid] != expect) atomicExch(&s_bad, 1);
    if ((ajj - c) & 1 || n0 < 0 || n1 < 0 || n2 < 0) atomicExch(&s_bad, 2);
    if (gd.counted && counts[gd.var0 + tid].bad > 0) atomicExch(&s_bad, 2);
    s_craw[tid] = c;
    s_flip[tid] = mono ? -1 : flip;
  }
  __syncthreads();
  if (s_bad == 2) {
    // values outside {0,1,2} (missing calls, dosages) reached the integer sweep: its sums are meaningless, and the
    // gene's record is (re)written by the fp64 path (dosage.cuh) or stays BADVALUE.  Do not run the O(M^3) tail --
    // least of all SKAT-O's quadrature -- on garbage (CTA-uniform exit: s_bad is shared and published above).
    if (tid == 0) {
      rvt_gene_result o;
      memset(&o, 0, sizeof(o));
      o.status = RVT_GENE_BADVALUE;
      o.p_skat = o.p_liu = 1.0;
      o.p_davies = -1.0;
      res[g] = o;
      if constexpr (FRONT) mid[g].Mp = -1;
      if constexpr (SKATO) {
        if (jobs) {   // nothing for k_skato_qags to do (an unset job would be read as garbage)
          jobs[g].run = 0;
          jobs[g].ok = 0;
          jobs[g].n_lam = 0;
          jobs[g].Q = jobs[g].rho = jobs[g].pvalue = 0.0;
        }
      }
    }
    return;
  }
  if (tid == 0) {
    int mp = 0;
    for (int j = 0; j < M; ++j)
      if (s_flip[j] >= 0) s_idx[mp++] = j;
    s_Mp = mp;
  }
  __syncthreads();
  const int Mp = s_Mp;

  // 3. real-valued score vector, covariate cross-products, weights (kept variants, minor-coded)
  if (tid < Mp) {
    const int j = s_idx[tid];
    const int fl = s_flip[j];
    long long sint = recombine4(&De[j * ER]);
    if (fl) sint = 2 * nm->vsum[0] - sint;
    s_s[tid] = (double)sint * nm->scale[0];
    for (int l = 0; l < C; ++l) {
      long long b = recombine4(&De[j * ER + 4 * (l + 1)]);
      if (fl) b = 2 * nm->vsum[l + 1] - b;
      s_B[tid][l] = (double)b * nm->scale[l + 1];
    }
    double freq = gd.has_af ? af[gd.var0 + tid] : (double)s_craw[j] / (2.0 * (double)N);
    double w = beta_weight(freq, prm.beta1, prm.beta2, true);
    s_sw[tid] = sqrt(w);
  }
  __syncthreads();
  if (tid == 0) {
    double q = 0.0;
    for (int i = 0; i < Mp; ++i) q += (s_sw[i] * s_sw[i]) * s_s[i] * s_s[i];
    s_Q = q;
  }
  // 4. K = W^1/2 sigma2 (A' - B' (X'X)^-1 B'^T) W^1/2   (upper triangle, mirrored)
  //    u_k = (X'X)^-1 B_k once per variant, then one dot product per entry
  if (tid < Mp) {
    for (int l = 0; l < C; ++l) {
      double u = 0.0;
      for (int m = 0; m < C; ++m) u += nm->xtx_inv[l * C + m] * s_B[tid][m];
      Uk[tid * C + l] = u;
    }
  }
  __syncthreads();
  //    rows r and Mp-1-r hold Mp+1 upper-triangle entries between them: one pass of the CTA per row pair
  for (int r = 0; r < (Mp + 1) / 2; ++r) {
    for (int t = tid; t < Mp + 1; t += NT) {
      int i, k;
      if (t < Mp - r) {
        i = r;
        k = r + t;
      } else {
        i = Mp - 1 - r;
        if (i == r) continue;
        k = i + (t - (Mp - r));
      }
      const int ji = s_idx[i], jk = s_idx[k];
      const int fi = s_flip[ji], fk = s_flip[jk];
      long long a = 0;
      for (int sp = 0; sp < S; ++sp) a += gp[sp].d[ji][jk];
      const long long ci = s_craw[ji], ck = s_craw[jk];
      if (fi && fk)
        a = 4 * N - 2 * ci - 2 * ck + a;
      else if (fi)
        a = 2 * ck - a;
      else if (fk)
        a = 2 * ci - a;
      double tt = 0.0;
      for (int l = 0; l < C; ++l) tt += s_B[i][l] * Uk[k * C + l];
      const double v = s_sw[i] * s_sw[k] * sigma2 * ((double)a - tt);
      K[i * kld + k] = v;
      K[k * kld + i] = v;
    }
  }
  __syncthreads();
  phase(1);
  if (tid == 0) {   // burden ingredients from the collapse digits
    for (int which = 0; which < 2; ++which) {
      const long long* cl = s_coll + which * (ER + 1);
      s_bur[which][0] = (double)recombine4(cl) * nm->scale[0];
      s_bur[which][1] = (double)cl[ER];
      for (int l = 0; l < C; ++l) s_bur[which][2 + l] = (double)recombine4(cl + 4 * (l + 1)) * nm->scale[l + 1];
    }
    s_nonref = (int)s_coll[(ER + 1) + ER];
  }
  __syncthreads();
  }  // !tin
  const int Mp = s_Mp;

  if (SKATO && jobs) {
    // SKAT-O: Z1'Z1 = W (G'G - G'X (X'X)^-1 X'G) W / 2 with the UN-squared weights = K / (2 sigma2)
    // (sqrt of the squared SKAT weight is the SKAT-O weight: src/Model.h:2652-2656 vs :2807-2809)
    const double sc = 0.5 / sigma2;
    for (int idx = tid; idx < Mp * Mp; idx += NT) {
      const int i = idx / Mp, k = idx - i * Mp;
      Wm[i * kld + k] = K[i * kld + k] * sc;
    }
    if (!tin && tid < Mp) s_vw[tid] = s_sw[tid] * s_s[tid];   // (pre-digested genes bring their own W G'r: TailInput::vw)
    __syncthreads();
  }

  if constexpr (FRONT) {
    FinMid* __restrict__ m = &mid[g];
    if (Mp >= 2) {
      householder_tridiag(K, Mp, kld, s_ev, s_e, s_v, s_p, par);
      for (int i = tid; i < Mp; i += NT) {
        m->d[i] = s_ev[i];
        m->e[i] = s_e[i];
      }
    }
    if (tid == 0) {
      m->Mp = Mp;
      m->bad = s_bad;
      m->nonref = s_nonref;
      m->status = tin ? tin[g].status : 0;
      m->Q = s_Q;
      for (int w = 0; w < 2; ++w)
        for (int l = 0; l < 2 + kMaxC; ++l) m->bur[w][l] = s_bur[w][l];
      if (Mp == 1) m->lam[0] = K[0];
    }
    return;
  }
  // 5. eigenvalues, descending, keep > 1e-30 from the top (Skat.cpp:84-98)
  double p_dav = -1.0, p_liu = 1.0, p_fin = 1.0, lam_max = 0.0;
  int fault = 0, r = 0;
  if (Mp > 0) {
    sym_eigenvalues_tridiag(K, Mp, kld, s_ev, s_e, s_v, s_p, s_lam, par);
    phase(2);
    const int r_ub = (N < (int64_t)Mp) ? (int)N : Mp;
    while (r < r_ub && s_lam[r] > 1e-30) ++r;
    lam_max = r ? s_lam[0] : 0.0;
    // 6. p-value
    const double Q = s_Q;
    p_dav = mixchisq_pvalue(s_lam, r, Q, s_th, &fault, par);
    phase(3);
    p_liu = liu_pvalue(s_lam, r, Q);
    p_fin = p_dav;
    if (p_fin <= 0.0 || p_fin == 1.0) p_fin = p_liu;
  }

  // 6b. SKAT-O (SkatO.cpp:101-281) on the same statistics: everything up to the quadrature; the record's skato_* fields
  //     are written by k_skato_qags
  SkatoOut so;
  so.ok = 0;
  so.timed_out = 0;
  so.Q = so.rho = so.pvalue = 0.0;
  if constexpr (SKATO) {
    if (jobs) {
      if (Mp > 0) {
        // ||r||^2/(N-1), SkatO.cpp:136-137; a binary trait takes s2 = 1 (SkatO.cpp:133-134, FitSKAT :72-75)
        const double s2 = nm->binary ? 1.0 : sigma2 * (double)N / (double)(N - 1);
        // Wm = K / (2 sigma2): its smallest eigenvalue is known from step 5 when all Mp are positive
        const double lam_min_w = (r == Mp) ? s_lam[Mp - 1] * (0.5 / sigma2) : 0.0;
        skato_prepare(Wm, K, Mp, kld, s_vw, s2, lam_min_w, s_ev, s_e, s_v, s_p, s_sk.lamz, s_sk.c, s_th, &jobs[g], par);
      } else if (tid == 0) {
        jobs[g].run = 0;
        jobs[g].ok = 0;
        jobs[g].Q = jobs[g].rho = jobs[g].pvalue = 0.0;
      }
      phase(5);
    }
  }

  // 7. burden score tests (m = 1)
  if (tid == 0) {
    burden_and_store(&res[out_index ? out_index[g] : g], Mp, s_bad, s_Q, p_fin, p_dav, p_liu, fault, r, lam_max, so, s_bur, s_nonref, nm,
                     tin ? tin[g].status : 0);
    phase(4);
  }
}

// bisection of the tridiagonal forms (one CTA per gene, one eigenvalue per thread)
__global__ void __launch_bounds__(kFinThreads, 16)
k_fin_sturm(FinMid* __restrict__ mid, int n_genes) {
  __shared__ double s_d[kTileRows + 2], s_e[kTileRows + 2], s_v[kTileRows + 2], s_p[kTileRows + 2], s_lam[kTileRows];
  const int g = blockIdx.x, tid = threadIdx.x;
  if (g >= n_genes) return;
  FinMid* __restrict__ m = &mid[g];
  const int Mp = m->Mp;
  if (Mp < 2) return;
  for (int i = tid; i < Mp; i += kFinThreads) {
    s_d[i] = m->d[i];
    s_e[i] = m->e[i];
  }
  __syncthreads();
  BlockPar par{nullptr};   // (no reduction in this phase)
  tridiag_eigenvalues(s_d, s_e, Mp, s_v, s_p, s_lam, par);
  for (int i = tid; i < Mp; i += kFinThreads) m->lam[i] = s_lam[i];
}

// steps 5b-7 of k_finalize from the FinMid record: kept eigenvalues, Davies -> Liu, burden score tests, the record
__global__ void __launch_bounds__(kFinThreads)
k_fin_tail(const FinMid* __restrict__ mid, int n_genes, const NullModel* __restrict__ nm, rvt_gene_result* __restrict__ res,
           const int* __restrict__ out_index) {
  __shared__ double s_lam[kTileRows], s_red[64];
  __shared__ int s_th[(kFinThreads / 32) * kTileRows];
  const int g = blockIdx.x, tid = threadIdx.x;
  if (g >= n_genes) return;
  const FinMid* __restrict__ m = &mid[g];
  const int Mp = m->Mp;
  if (Mp < 0) return;   // the front wrote the record
  const int64_t N = nm->N;
  for (int i = tid; i < Mp; i += kFinThreads) s_lam[i] = m->lam[i];
  __syncthreads();
  BlockPar par{s_red};
  double p_dav = -1.0, p_liu = 1.0, p_fin = 1.0, lam_max = 0.0;
  int fault = 0, r = 0;
  const double Q = m->Q;
  if (Mp > 0) {
    const int r_ub = (N < (int64_t)Mp) ? (int)N : Mp;
    while (r < r_ub && s_lam[r] > 1e-30) ++r;
    lam_max = r ? s_lam[0] : 0.0;
    p_dav = mixchisq_pvalue(s_lam, r, Q, s_th, &fault, par);
    p_liu = liu_pvalue(s_lam, r, Q);
    p_fin = p_dav;
    if (p_fin <= 0.0 || p_fin == 1.0) p_fin = p_liu;
  }
  if (tid == 0) {
    SkatoOut so;
    so.ok = 0;
    so.timed_out = 0;
    so.Q = so.rho = so.pvalue = 0.0;
    burden_and_store(&res[out_index ? out_index[g] : g], Mp, m->bad, Q, p_fin, p_dav, p_liu, fault, r, lam_max, so, m->bur, m->nonref, nm,
                     m->status);
  }
}

}  // namespace rvt
