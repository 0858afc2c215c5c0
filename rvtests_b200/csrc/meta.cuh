// meta.cuh -- K5/K6: single-variant meta-analysis statistics (--meta score,cov) from the same
// integer sweep sums as the gene tests.
//
//   MetaScoreTest::fitWithGivenGenotype + MetaUnrelatedQtl   src/Model.h:3188-3260, 3501-3555
//       GenotypeCounter (N_REF/N_HET/N_ALT, AF, call rate)   src/GenotypeCounter.h:14-59
//       Hardy-Weinberg exact test SNPHWE                     libsrc/snp_hwe.cpp:25-122
//       LinearRegressionScoreTest::TestCovariate (m = 1)     regression/LinearRegressionScoreTest.cpp:173-263
//       U_STAT = U/sigma2, SQRT_V_STAT = sqrt(V/sigma2^2), ALT_EFFSIZE, SE   src/Model.h:3543-3549,
//                                                            LinearRegressionScoreTest.cpp:365-376
//   MetaCovTest + MetaCovUnrelatedQtl                        src/Model.cpp:500-596, 844-1004
//       cov(i,j) = ( x~_i.x~_j/sigma2 - covXZ_i (Zc'Zc/sigma2)^+ covXZ_j' ) / N   for every queued
//       variant j within `windowSize` bp after i (src/Model.h:3954-3967, 3993-4020).  With the
//       intercept column centred to exactly zero Eigen's LDLT solve acts as a pseudo-inverse
//       (regression/EigenMatrixInterface.cpp:125-136), so this is (g_i'(I-H_X)g_j)/(sigma2 N)
//       = (A_ij - B_i (X'X)^-1 B_j') / (sigma2 N): an entry of the same projected Gram the SKAT
//       kernel builds, without weights.
//   MetaUnrelatedBinary / MetaCovUnrelatedBinary (binary trait)   src/Model.h:3669-3784, src/Model.cpp:695-778  (end of file)
// Every push is one tile (<= 64 consecutive variants); tile pairs (I,J) inside the window come from
// the PAIR mode of the tensor-core sweep (A tile = rows of I, B tile = rows of J).
#pragma once
#include "../../include/rvtests_b200.h"
#include "common.cuh"
#include "davies.cuh"

namespace rvt {

constexpr int kMetaThreads = 128;

// Hardy-Weinberg exact test, libsrc/snp_hwe.cpp:25-122 (Wigginton et al.), without the O(rare copies) table: the two
// recurrences away from the most likely heterozygote count `mid` are walked twice -- once for the normalising sum and the
// probability of the observed count, once to add up the outcomes that are no more likely than the observed one --
// with the outcomes dealt to the threads of a CTA.
// A variant with `rare` copies of its minor allele
// walks rare/2 outcomes four times; one thread per variant made the 64-variant statistics kernel wait ~100 ms for its
// most common variant at N = 500 000 (profiles/r02e_meta_time.log: 108 of 133 ms of a flush).  Both recurrences are running
// products pr_k = prod_{i<=k} r_i away from the mode, so a thread that owns outcomes [k0, k1) needs only the product of the
// ratios before k0: every thread multiplies up its own chunk (ratios <= 1 beyond the mode: no overflow; a tail that
// underflows contributes 0, as it does serially), thread 0 scans the chunk products, and the two passes of the serial
// routine -- normalising sum + probability of the observed count, then the sum of the outcomes no more likely than it --
// run chunk-parallel.  Sums differ from the serial order by rounding only (|dp| ~ 1e-13 relative; test: 1e-9).
constexpr int kHweThreads = 256;

struct HweBranch {   // one of the two walks away from the mode `mid`
  long long K;       // outcomes on this side
  long long homr0, homc0, mid;
  int up;            // 1: h = mid + 2k -> hets h + 2;  0: h = mid - 2k -> hets h - 2
  __device__ __forceinline__ double ratio(long long k) const {
    if (up) {
      const double h = (double)(mid + 2 * k);
      return 4.0 * (double)(homr0 - k) * (double)(homc0 - k) / ((h + 2.0) * (h + 1.0));
    }
    const double h = (double)(mid - 2 * k);
    return h * (h - 1.0) / (4.0 * ((double)(homr0 + k) + 1.0) * ((double)(homc0 + k) + 1.0));
  }
  __device__ __forceinline__ long long hets_after(long long k) const { return up ? mid + 2 * k + 2 : mid - 2 * k - 2; }
};

// the whole CTA (kHweThreads) calls this with the same counts; thread 0 writes *out
__device__ void hwe_cta(const long long obs_hets, const long long obs_hom1, const long long obs_hom2, double* __restrict__ out) {
  __shared__ double s_prod[2][kHweThreads], s_start[2][kHweThreads], s_red[kHweThreads];
  __shared__ double s_pobs, s_sum;
  const int tid = threadIdx.x;
  if (obs_hets < 0 || obs_hom1 < 0 || obs_hom2 < 0) {
    if (tid == 0) *out = 0.0;
    return;
  }
  const long long obs_homc = obs_hom1 < obs_hom2 ? obs_hom2 : obs_hom1;
  const long long obs_homr = obs_hom1 < obs_hom2 ? obs_hom1 : obs_hom2;
  const long long rare = 2 * obs_homr + obs_hets;
  const long long n = obs_hets + obs_homc + obs_homr;
  if (n == 0) {
    if (tid == 0) *out = 0.0;
    return;
  }
  long long mid = (long long)(1.0 * rare * (2 * n - rare) / (2 * n));
  if ((rare & 1) ^ (mid & 1)) mid++;
  HweBranch br[2];
  for (int b = 0; b < 2; ++b) {
    br[b].mid = mid;
    br[b].homr0 = (rare - mid) / 2;
    br[b].homc0 = n - mid - br[b].homr0;
    br[b].up = b;
  }
  br[0].K = mid > 1 ? mid / 2 : 0;                          // h = mid, mid - 2, .. > 1
  br[1].K = rare - 2 >= mid ? (rare - 2 - mid) / 2 + 1 : 0;   // h = mid, mid + 2, .. <= rare - 2
  // chunk products
  long long k0[2], k1[2];
  for (int b = 0; b < 2; ++b) {
    const long long per = (br[b].K + kHweThreads - 1) / kHweThreads;
    k0[b] = min((long long)tid * per, br[b].K);
    k1[b] = min(k0[b] + per, br[b].K);
    double pr = 1.0;
    for (long long k = k0[b]; k < k1[b]; ++k) pr *= br[b].ratio(k);
    s_prod[b][tid] = pr;
  }
  if (tid == 0) s_pobs = (obs_hets == mid) ? 1.0 : -1.0;
  __syncthreads();
  if (tid < 2) {
    double run = 1.0;
    for (int t = 0; t < kHweThreads; ++t) {
      s_start[tid][t] = run;
      run *= s_prod[tid][t];
    }
  }
  __syncthreads();
  // pass 1: normalising sum, probability of the observed count
  double part = 0.0;
  for (int b = 0; b < 2; ++b) {
    double pr = s_start[b][tid];
    for (long long k = k0[b]; k < k1[b]; ++k) {
      pr *= br[b].ratio(k);
      part += pr;
      if (br[b].hets_after(k) == obs_hets) s_pobs = pr;
    }
  }
  s_red[tid] = part;
  __syncthreads();
  if (tid == 0) {
    double sum = 1.0;
    for (int t = 0; t < kHweThreads; ++t) sum += s_red[t];
    s_sum = sum;
  }
  __syncthreads();
  const double sum = s_sum, p_obs = s_pobs;
  if (p_obs < 0.0) {   // observed count has the wrong parity (cannot happen for real counts)
    if (tid == 0) *out = 0.0;
    return;
  }
  // pass 2: the reference compares the NORMALISED probabilities (snp_hwe.cpp:113-116)
  const double thr = p_obs / sum;
  part = 0.0;
  if (tid == 0 && !(1.0 / sum > thr)) part += 1.0 / sum;
  for (int b = 0; b < 2; ++b) {
    double pr = s_start[b][tid];
    for (long long k = k0[b]; k < k1[b]; ++k) {
      pr *= br[b].ratio(k);
      if (!(pr / sum > thr)) part += pr / sum;
    }
  }
  __syncthreads();
  s_red[tid] = part;
  __syncthreads();
  if (tid == 0) {
    double p = 0.0;
    for (int t = 0; t < kHweThreads; ++t) p += s_red[t];
    *out = p > 1.0 ? 1.0 : p;
  }
}

__global__ void __launch_bounds__(kHweThreads)
k_meta_hwe(int64_t nv, rvt_variant_result* __restrict__ vout) {
  const int64_t v = blockIdx.x;
  if (v >= nv) return;
  hwe_cta(vout[v].n_het, vout[v].n_ref, vout[v].n_alt, &vout[v].hwe_p);
}
// cases and controls of a binary trait (MetaScoreTest prints all:case:control, src/Model.h:3300-3330)
__global__ void __launch_bounds__(kHweThreads)
k_meta_hwe_cc(int64_t nv, rvt_variant_cc* __restrict__ cc) {
  const int64_t v = blockIdx.x >> 1;
  const int w = blockIdx.x & 1;
  if (v >= nv) return;
  hwe_cta(cc[v].n_het[w], cc[v].n_ref[w], cc[v].n_alt[w], &cc[v].hwe_p[w]);
}

__device__ __forceinline__ long long recombine4m(const long long* d) {
  return d[0] + (d[1] << 8) + (d[2] << 16) + (d[3] << 24);
}

// One CTA per 64-variant tile: per-variant score statistics, the tile's B = G'X rows, and the
// within-tile covariance band.  tiles[t]: row0 = first variant row, M = variants in the tile.
__global__ void __launch_bounds__(kMetaThreads)
k_meta_block(const GeneDesc* __restrict__ tiles, int n_tiles, const NullModel* __restrict__ nm, int S, const SweepPartial* __restrict__ parts,
             const int* __restrict__ jmax /*[nv] last partner (variant index)*/, int wmax,
             rvt_variant_result* __restrict__ vout, double* __restrict__ Bmat /*[nv][kMaxC]*/,
             uint8_t* __restrict__ poly /*[nv]*/, double* __restrict__ band /*[nv][wmax+1] or null*/,
             double band_scale /* > 0: entries are (projected Gram) * band_scale instead of / (sigma2 N) */,
             double* __restrict__ Uraw = nullptr /* [nv] g'r (binary trait: k_metab_final finishes the statistics) */) {
  __shared__ long long De[kTileRows][kMaxER];
  __shared__ long long s_ajj[kTileRows];
  __shared__ double s_B[kTileRows][kMaxC];
  __shared__ int s_poly[kTileRows];
  const int t = blockIdx.x, tid = threadIdx.x;
  if (t >= n_tiles) return;
  const GeneDesc gd = tiles[t];
  const int M = gd.M;
  const int64_t N = nm->N;
  const int C = nm->C, ER = nm->ER;
  const double sigma2 = nm->sigma2;
  const SweepPartial* __restrict__ gp = parts + (size_t)t * S;
  const int64_t vbase = gd.var0;
  for (int idx = tid; idx < M * ER; idx += kMetaThreads) {
    const int i = idx / ER, e = idx - i * ER;
    long long s = 0;
    for (int sp = 0; sp < S; ++sp) s += gp[sp].d[i][kTileRows + e];
    De[i][e] = s;
  }
  if (tid < M) {
    long long s = 0;
    for (int sp = 0; sp < S; ++sp) s += gp[sp].d[tid][tid];
    s_ajj[tid] = s;
  }
  __syncthreads();
  if (tid < M) {
    const int64_t v = vbase + tid;
    const long long c = llrint((double)recombine4m(&De[tid][4]) * nm->scale[1]);
    const long long ajj = s_ajj[tid];
    const long long n2 = (ajj - c) / 2, n1 = c - 2 * n2, n0 = N - n1 - n2;
    const int mono = (n0 == N) || (n1 == N) || (n2 == N);   // isMonomorphicMarker, src/Model.h:3241-3244
    double Bi[kMaxC];
    for (int l = 0; l < C; ++l) {
      Bi[l] = (double)recombine4m(&De[tid][4 * (l + 1)]) * nm->scale[l + 1];
      s_B[tid][l] = Bi[l];
      Bmat[(size_t)v * kMaxC + l] = Bi[l];
    }
    const double U = (double)recombine4m(&De[tid][0]) * nm->scale[0];
    double q = 0.0;
    for (int l = 0; l < C; ++l)
      for (int m = 0; m < C; ++m) q += Bi[l] * nm->xtx_inv[l * C + m] * Bi[m];
    const double SS = (double)ajj - q;
    const double V = SS * sigma2;
    const double beta = U * (1.0 / SS);
    const double stat = U * ((1.0 / SS) / sigma2) * U;
    rvt_variant_result o;
    memset(&o, 0, sizeof(o));
    o.af = 0.5 * (double)c / (double)N;          // GenotypeCounter::getAF
    o.ac = (double)c;                             // getAC
    o.call_rate = 1.0;                            // hard calls, nothing missing
    o.n_ref = (int)n0;
    o.n_het = (int)n1;
    o.n_alt = (int)n2;
    o.hwe_p = 0.0;   // exact Hardy-Weinberg test: k_meta_hwe, one CTA per variant, launched next
    const int ok = !mono && !(stat < 0.0) && (stat == stat);
    o.ok = ok;
    o.polymorphic = !mono;
    if (ok) {
      o.U = U / sigma2;
      o.sqrtV = sqrt(V / sigma2 / sigma2);
      o.effect = (V != 0.0) ? beta : 0.0;
      o.effect_se = (V != 0.0) ? sigma2 / sqrt(V) : 0.0;
      o.pvalue = chisq_q(stat, 1.0);
    }
    vout[v] = o;
    if (Uraw) Uraw[v] = U;
    poly[v] = (uint8_t)!mono;
    s_poly[tid] = !mono;
  }
  __syncthreads();
  if (!band) return;
  // within-tile covariance band
  const double scale = band_scale > 0.0 ? band_scale : 1.0 / (sigma2 * (double)N);
  for (int idx = tid; idx < M * M; idx += kMetaThreads) {
    const int i = idx / M, j = idx - i * M;
    if (j < i) continue;
    const int64_t vi = vbase + i, vj = vbase + j;
    if (vj > jmax[vi]) continue;
    double val = nan("");
    if (s_poly[i] && s_poly[j]) {
      long long a = 0;
      for (int sp = 0; sp < S; ++sp) a += gp[sp].d[i][j];
      double tq = 0.0;
      for (int l = 0; l < C; ++l) {
        double u = 0.0;
        for (int m = 0; m < C; ++m) u += nm->xtx_inv[l * C + m] * s_B[j][m];
        tq += s_B[i][l] * u;
      }
      val = ((double)a - tq) * scale;
    }
    band[(size_t)vi * (wmax + 1) + (vj - vi)] = val;
  }
}

// One CTA per tile pair (I,J), J > I: the cross-tile part of the band.
__global__ void __launch_bounds__(kMetaThreads)
k_meta_pair(const GeneDesc* __restrict__ pairs, int n_pairs, const NullModel* __restrict__ nm, int S, const SweepPartial* __restrict__ parts,
            const int* __restrict__ jmax, int wmax, const double* __restrict__ Bmat,
            const uint8_t* __restrict__ poly, double* __restrict__ band, double band_scale) {
  const int p = blockIdx.x, tid = threadIdx.x;
  if (p >= n_pairs) return;
  const GeneDesc gd = pairs[p];
  const int Ma = gd.M, Mb = gd.Mb;
  const int64_t va = gd.var0, vb = gd.var0_b;
  const int C = nm->C;
  const double scale = band_scale > 0.0 ? band_scale : 1.0 / (nm->sigma2 * (double)nm->N);
  const SweepPartial* __restrict__ gp = parts + (size_t)p * S;
  for (int idx = tid; idx < Ma * Mb; idx += kMetaThreads) {
    const int i = idx / Mb, j = idx - i * Mb;
    const int64_t vi = va + i, vj = vb + j;
    if (vj > jmax[vi]) continue;
    double val = nan("");
    if (poly[vi] && poly[vj]) {
      long long a = 0;
      for (int sp = 0; sp < S; ++sp) a += gp[sp].d[i][j];
      double tq = 0.0;
      for (int l = 0; l < C; ++l) {
        double u = 0.0;
        for (int m = 0; m < C; ++m) u += nm->xtx_inv[l * C + m] * Bmat[(size_t)vj * kMaxC + m];
        tq += Bmat[(size_t)vi * kMaxC + l] * u;
      }
      val = ((double)a - tq) * scale;
    }
    band[(size_t)vi * (wmax + 1) + (vj - vi)] = val;
  }
}

// ---- binary trait: MetaUnrelatedBinary (src/Model.h:3669-3784) / MetaCovUnrelatedBinary (src/Model.cpp:695-778) ----------
//   U = g'(y - p),  V = g'Wg - g'WZ (Z'WZ)^-1 Z'Wg,  W = diag(p(1 - p))     LogisticRegressionScoreTest.cpp:219-302
//   cov(i, j) = ( g_i'W g_j - covXZ_i (Z'WZ)^-1 covXZ_j' ) / N,  covXZ_i = g_i'W Z   (raw genotypes: no centring)
// The weights stay on the integer tensor-core sweep: w_i is rounded to q_i 2^-28 (w <= 1/4, so q <= 2^26) and q_i written
// in four BALANCED base-128 digits d_k in [-64, 63], so that g d_k fits an s8 for g in {0, 1, 2}.  For each k the engine
// writes the tiles G o d_k (k_metab_scale), pairs them (A operand) with the plain tiles (B operand) in the PAIR sweep, and
//   g_i'W g_j = 2^-28 sum_k 128^k (G o d_k)_i' G_j,     g_i'W Z = 2^-28 sum_k 128^k (G o d_k)_i' E
// are exact integers (< 2^53) accumulated in doubles.  A fifth pass with d = y (0/1) gives the genotype counts among the
// cases from the diagonal pairs; the controls are all - cases.
constexpr int kMetabDigits = 4;
constexpr int kMetabShift = 28;

// digits of the weights and the case indicator: dig[k][i], k = 0..3 digits (least significant first), k = 4: y_i
__global__ void k_metab_digits(int64_t N, const double* __restrict__ vw, const double* __restrict__ resid, int8_t* __restrict__ dig, int64_t ldd,
                               unsigned long long* __restrict__ n_case) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int is_case = 0;
  if (i < ldd) {
    long long q = 0;
    if (i < N) {
      q = llrint(ldexp(vw[i], kMetabShift));
      is_case = resid[i] > 0.0;   // r = y - p with 0 < p < 1
    }
#pragma unroll
    for (int k = 0; k < kMetabDigits; ++k) {
      long long d = ((q + 64) & 127) - 64;   // balanced digit
      dig[(size_t)k * ldd + i] = (int8_t)d;
      q = (q - d) >> 7;
    }
    dig[(size_t)kMetabDigits * ldd + i] = (int8_t)is_case;
  }
  const unsigned b = __ballot_sync(0xFFFFFFFFu, is_case);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_case, (unsigned long long)__popc(b));
}

// out = g o d over one tiled block of M rows (layout: 128-sample chunks of M rows x 128 bytes).
// grid: (ceil(nchunk*32/256), M); one thread = one 4-sample word
__global__ void __launch_bounds__(256)
k_metab_scale(const int8_t* __restrict__ g, int M, int64_t N, const int8_t* __restrict__ d, int8_t* __restrict__ out) {
  const int r = blockIdx.y;
  const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nchunk = (N + 127) >> 7;
  if (wi >= nchunk * 32) return;
  const size_t off = ((size_t)(wi >> 5) * M + r) * 128 + (size_t)(wi & 31) * 4;
  const char4 gv = *reinterpret_cast<const char4*>(g + off);
  const char4 dv = *reinterpret_cast<const char4*>(d + wi * 4);   // (d is padded to a multiple of 128 samples with zeros)
  char4 o;
  o.x = (signed char)(gv.x * dv.x);
  o.y = (signed char)(gv.y * dv.y);
  o.z = (signed char)(gv.z * dv.z);
  o.w = (signed char)(gv.w * dv.w);
  *reinterpret_cast<char4*>(out + off) = o;
}

struct MetabVar {        // per variant, accumulated over the digit passes
  double xz[kMaxC];      // sum_k 128^k (G o d_k)' Z  (times 2^-28 = g'WZ)
  long long c_case, ajj_case;
};

// One CTA per PAIR unit of digit pass k (k = kMetabDigits: the case pass; diagonal pairs only).
// acc: [nv][wmax+1] (entry (vi, vj - vi)), exact integer sums in doubles.
__global__ void __launch_bounds__(kMetaThreads)
k_metab_acc(const GeneDesc* __restrict__ units, int n_units, const NullModel* __restrict__ nm, int S, const SweepPartial* __restrict__ parts,
            const int* __restrict__ jmax, int wmax, int k, double* __restrict__ acc, MetabVar* __restrict__ mv) {
  __shared__ long long De[kTileRows][kMaxER];
  const int p = blockIdx.x, tid = threadIdx.x;
  if (p >= n_units) return;
  const GeneDesc gd = units[p];
  const int Ma = gd.M, Mb = gd.Mb, C = nm->C, ER = nm->ER;
  const int64_t va = gd.var0, vb = gd.var0_b;
  const bool diag = va == vb;
  const SweepPartial* __restrict__ gp = parts + (size_t)p * S;
  const double f = ldexp(1.0, 7 * k);
  if (diag) {
    for (int idx = tid; idx < Ma * ER; idx += kMetaThreads) {
      const int i = idx / ER, e = idx - i * ER;
      long long s = 0;
      for (int sp = 0; sp < S; ++sp) s += gp[sp].d[i][kTileRows + e];
      De[i][e] = s;
    }
    __syncthreads();
    if (tid < Ma) {
      MetabVar* o = mv + va + tid;
      if (k == kMetabDigits) {
        long long s = 0;
        for (int sp = 0; sp < S; ++sp) s += gp[sp].d[tid][tid];
        o->ajj_case = s;
        o->c_case = llrint((double)recombine4m(&De[tid][4]) * nm->scale[1]);
      } else {
        for (int l = 0; l < C; ++l) {
          const double t = (double)recombine4m(&De[tid][4 * (l + 1)]) * nm->scale[l + 1] * f;
          o->xz[l] = (k == 0 ? 0.0 : o->xz[l]) + t;
        }
      }
    }
  }
  if (k == kMetabDigits) return;
  for (int idx = tid; idx < Ma * Mb; idx += kMetaThreads) {
    const int i = idx / Mb, j = idx - i * Mb;
    const int64_t vi = va + i, vj = vb + j;
    if (vj < vi || vj > jmax[vi]) continue;
    long long a = 0;
    for (int sp = 0; sp < S; ++sp) a += gp[sp].d[i][j];
    double* o = acc + (size_t)vi * (wmax + 1) + (vj - vi);
    *o = (k == 0 ? 0.0 : *o) + (double)a * f;
  }
}

// per variant: finish the score statistics and the case / control counts
__global__ void __launch_bounds__(128)
k_metab_final(int64_t nv, const NullModel* __restrict__ nm, const double* __restrict__ acc, int wmax, MetabVar* __restrict__ mv,
              const double* __restrict__ Uraw, const unsigned long long* __restrict__ n_case, rvt_variant_result* __restrict__ vout,
              rvt_variant_cc* __restrict__ cc, double* __restrict__ cov_xz /*[nv][kMaxC] = g'WZ */) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const int C = nm->C;
  const double w2 = ldexp(1.0, -kMetabShift);
  double xz[kMaxC];
  for (int l = 0; l < C; ++l) {
    xz[l] = mv[v].xz[l] * w2;
    cov_xz[(size_t)v * kMaxC + l] = xz[l];
  }
  rvt_variant_result o = vout[v];
  // cases / controls (GenotypeCounter on the case / control samples, src/Model.h:3220-3232)
  const long long Nc = (long long)*n_case, Nall = nm->N;
  const long long cj = mv[v].c_case, aj = mv[v].ajj_case;
  const long long n2c = (aj - cj) / 2, n1c = cj - 2 * n2c, n0c = Nc - n1c - n2c;
  rvt_variant_cc q;
  memset(&q, 0, sizeof(q));
  q.n[0] = (int)Nc;
  q.n[1] = (int)(Nall - Nc);
  q.n_ref[0] = (int)n0c;
  q.n_het[0] = (int)n1c;
  q.n_alt[0] = (int)n2c;
  q.n_ref[1] = o.n_ref - (int)n0c;
  q.n_het[1] = o.n_het - (int)n1c;
  q.n_alt[1] = o.n_alt - (int)n2c;
  cc[v] = q;
  // score statistics
  double proj = 0.0;
  for (int l = 0; l < C; ++l)
    for (int m = 0; m < C; ++m) proj += xz[l] * nm->xtx_inv[l * C + m] * xz[m];   // (xtx_inv holds (Z'WZ)^-1 for a binary trait)
  const double V = acc[(size_t)v * (wmax + 1)] * w2 - proj;
  const double U = Uraw[v];
  const double stat = U * U / V;
  const int ok = o.polymorphic && V > 0.0 && !(stat < 0.0) && (stat == stat);
  o.ok = ok;
  o.U = o.sqrtV = o.effect = o.effect_se = o.pvalue = 0.0;
  if (ok) {
    o.U = U;
    o.sqrtV = sqrt(V);
    o.effect = (U != 0.0) ? U / V : 0.0;       // MetaUnrelatedBinary::GetEffect
    o.effect_se = 1.0 / sqrt(V);               // GetEffectSE
    o.pvalue = chisq_q(stat, 1.0);
  }
  vout[v] = o;
}

// the band: (g_i'W g_j - covXZ_i (Z'WZ)^-1 covXZ_j') / N, NaN where either variant is monomorphic or j is outside i's window
__global__ void __launch_bounds__(128)
k_metab_band(int64_t nv, const NullModel* __restrict__ nm, const double* __restrict__ acc, const double* __restrict__ cov_xz,
             const uint8_t* __restrict__ poly, const int* __restrict__ jmax, int wmax, double* __restrict__ band) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nv * (int64_t)(wmax + 1)) return;
  const int64_t vi = idx / (wmax + 1), vj = vi + (idx - vi * (wmax + 1));
  double val = nan("");
  if (vj < nv && vj <= jmax[vi] && poly[vi] && poly[vj]) {
    const int C = nm->C;
    double proj = 0.0;
    for (int l = 0; l < C; ++l)
      for (int m = 0; m < C; ++m) proj += cov_xz[(size_t)vi * kMaxC + l] * nm->xtx_inv[l * C + m] * cov_xz[(size_t)vj * kMaxC + m];
    val = (acc[idx] * ldexp(1.0, -kMetabShift) - proj) / (double)nm->N;
  }
  band[idx] = val;
}

}  // namespace rvt
