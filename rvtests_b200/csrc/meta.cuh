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
// Every push is one tile (<= 64 consecutive variants); tile pairs (I,J) inside the window come from
// the PAIR mode of the tensor-core sweep (A tile = rows of I, B tile = rows of J).
#pragma once
#include "../../include/rvtests_b200.h"
#include "common.cuh"
#include "davies.cuh"

namespace rvt {

constexpr int kMetaThreads = 128;

// libsrc/snp_hwe.cpp:25-122 (Wigginton et al.) without the O(rare copies) table: the recurrences
// are walked twice -- once for the normalising sum and the probability of the observed
// heterozygote count, once to add up the outcomes that are no more likely than the observed one.
__device__ inline double snp_hwe(long long obs_hets, long long obs_hom1, long long obs_hom2) {
  const long long obs_homc = obs_hom1 < obs_hom2 ? obs_hom2 : obs_hom1;
  const long long obs_homr = obs_hom1 < obs_hom2 ? obs_hom1 : obs_hom2;
  const long long rare = 2 * obs_homr + obs_hets;
  const long long n = obs_hets + obs_homc + obs_homr;
  if (n == 0) return 0.0;
  long long mid = (long long)(1.0 * rare * (2 * n - rare) / (2 * n));
  if ((rare & 1) ^ (mid & 1)) mid++;
  double p_obs = -1.0, sum = 1.0;
  if (obs_hets == mid) p_obs = 1.0;
  {
    double pr = 1.0;
    long long homr = (rare - mid) / 2, homc = n - mid - homr;
    for (long long h = mid; h > 1; h -= 2) {
      pr = pr * h * (h - 1.0) / (4.0 * (homr + 1.0) * (homc + 1.0));
      sum += pr;
      if (h - 2 == obs_hets) p_obs = pr;
      homr++;
      homc++;
    }
    pr = 1.0;
    homr = (rare - mid) / 2;
    homc = n - mid - homr;
    for (long long h = mid; h <= rare - 2; h += 2) {
      pr = pr * 4.0 * homr * homc / ((h + 2.0) * (h + 1.0));
      sum += pr;
      if (h + 2 == obs_hets) p_obs = pr;
      homr--;
      homc--;
    }
  }
  if (p_obs < 0.0) return 0.0;  // observed count has the wrong parity (cannot happen for real counts)
  // the reference compares the NORMALISED probabilities (snp_hwe.cpp:113-116)
  const double thr = p_obs / sum;
  double p = 0.0;
  {
    double pr = 1.0;
    if (!(1.0 / sum > thr)) p += 1.0 / sum;
    long long homr = (rare - mid) / 2, homc = n - mid - homr;
    for (long long h = mid; h > 1; h -= 2) {
      pr = pr * h * (h - 1.0) / (4.0 * (homr + 1.0) * (homc + 1.0));
      if (!(pr / sum > thr)) p += pr / sum;
      homr++;
      homc++;
    }
    pr = 1.0;
    homr = (rare - mid) / 2;
    homc = n - mid - homr;
    for (long long h = mid; h <= rare - 2; h += 2) {
      pr = pr * 4.0 * homr * homc / ((h + 2.0) * (h + 1.0));
      if (!(pr / sum > thr)) p += pr / sum;
      homr--;
      homc--;
    }
  }
  return p > 1.0 ? 1.0 : p;
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
             uint8_t* __restrict__ poly /*[nv]*/, double* __restrict__ band /*[nv][wmax+1] or null*/) {
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
    o.hwe_p = (n0 < 0 || n1 < 0 || n2 < 0) ? 0.0 : snp_hwe(n1, n0, n2);
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
    poly[v] = (uint8_t)!mono;
    s_poly[tid] = !mono;
  }
  __syncthreads();
  if (!band) return;
  // within-tile covariance band
  const double scale = 1.0 / (sigma2 * (double)N);
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
            const uint8_t* __restrict__ poly, double* __restrict__ band) {
  const int p = blockIdx.x, tid = threadIdx.x;
  if (p >= n_pairs) return;
  const GeneDesc gd = pairs[p];
  const int Ma = gd.M, Mb = gd.Mb;
  const int64_t va = gd.var0, vb = gd.var0_b;
  const int C = nm->C;
  const double scale = 1.0 / (nm->sigma2 * (double)nm->N);
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

}  // namespace rvt
