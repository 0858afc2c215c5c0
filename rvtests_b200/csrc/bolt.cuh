// bolt.cuh -- A15: the BoltLMM null-model fit (regression/BoltLMM.cpp:169-299, 463-859, 926-1214;
// regression/BoltPlinkLoader.cpp:115-342) with the genotype panel resident on the GPU as PLINK 2-bit rows.
//
//   model        y = X beta + e,  beta ~ N(0, sigma2_g / M),  e ~ N(0, sigma2_e),  X = panel genotypes normalised to
//                (g - 2p)/sqrt(2p(1-p)) (missing -> 0), covariates projected out of everything
//   H            = X X' / M + delta I,  delta = sigma2_e / sigma2_g                      computeHx, BoltLMM.cpp:931-993
//   solve        multi-RHS conjugate gradients on H, BOLT's tolerance 5e-4, <= min(N, 250) iterations   :745-859
//   MC-REML      f(log delta) = log( (|beta^_data|^2/|e^_data|^2) / (sum_t |beta^_t|^2 / sum_t |e^_t|^2) ) over MCtrial
//                simulated phenotypes X beta_rand + sqrt(delta) e_rand; secant iteration on log delta, <= 7 evaluations
//                (evalREML :669-724, EstimateHeritabilityBolt :575-668)
//   calibration  30 random panel SNPs: prospective x'V^-1y^2 / x'V^-1x against the uncalibrated retrospective
//                statistic (EstimateInfStatCalibration :1141-1214)
//   output       H^-1 y / sigma2_g, its projected squared norm, infStatCalibration -- what BoltLMM::TestCovariate
//                (:315-338) needs, i.e. the inputs of rvt_set_null_residual + rvt_meta_flush (the A14 score step)
//
// The reference keeps every vector as [v ; Z'v] (N + C rows, Z = orthonormal covariate basis) so that projecting the
// covariates out becomes a sign flip on the last C rows of every inner product (projDot / projNorm2 :1064-1138).  The same
// layout is used here: vectors are (N + C) x R doubles, row-major (R right-hand sides side by side), the two O(N M R)
// products of computeHx stream the 2-bit panel from HBM (N M / 4 bytes per pass) and decode it through a 4-entry table per
// SNP held in shared memory.  The host drives the CG / secant logic (a few hundred scalars per iteration) and draws the
// random numbers with the reference's own generator (MT19937 seeded 12345 + polar Box-Muller, libsrc/Random.cpp) so that
// the Monte-Carlo path is the reference's path.  Arithmetic is fp64 where the reference uses float32.
#pragma once
#include <math.h>
#include <stdint.h>

#include <vector>

#include "common.cuh"

namespace rvt {

constexpr int kBoltMaxR = 32;        // right-hand sides per solve (MCtrial + 1 <= 16; 30 calibration SNPs)
constexpr int kBoltSnpBlock = 64;    // SNPs per CTA in the X'v product
constexpr int kBoltChunk = 1024;     // samples staged per step in the X'v product
constexpr int kBoltRowPad = kBoltChunk / 4 + 4;   // bytes per staged row (65 words: conflict-free across rows)

// libsrc/Random.cpp: MT19937 (InitMersenne :128-140, Next :146-183) and the polar Box-Muller Normal (:269-288)
struct BoltRandom {
  uint32_t mt[624];
  int mti;
  bool saved;
  double store;
  explicit BoltRandom(uint32_t s = 12345) { reset(s); }
  void reset(uint32_t s) {
    mt[0] = s;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    mti = 624;
    saved = false;
  }
  double next() {
    if (mti >= 624) {
      for (int kk = 0; kk < 624; ++kk) {
        const uint32_t y = (mt[kk] & 0x80000000u) | (mt[(kk + 1) % 624] & 0x7FFFFFFFu);
        mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
      }
      mti = 0;
    }
    uint32_t y = mt[mti++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    return (1.0 / 4294967296.0) * ((double)y + 0.5);
  }
  double normal() {
    if (saved) {
      saved = false;
      return store;
    }
    double v1, v2, rsq;
    do {
      v1 = 2.0 * next() - 1.0;
      v2 = 2.0 * next() - 1.0;
      rsq = v1 * v1 + v2 * v2;
    } while (rsq >= 1.0 || rsq == 0.0);
    const double fac = sqrt(-2.0 * log(rsq) / rsq);
    store = v1 * fac;
    saved = true;
    return v2 * fac;
  }
};

#if defined(__CUDACC__)
// PLINK code of sample i in a row: 00 -> 0, 10 -> 1, 11 -> 2, 01 -> missing (libVcf/PlinkInputFile.cpp:23-47)
__device__ __forceinline__ int bolt_code(const uint8_t* __restrict__ row, int64_t i) { return (row[i >> 2] >> (2 * (i & 3))) & 3; }

// One CTA per SNP: allele / missing counts -> the 4-entry table (BoltPlinkLoader::prepareGenotype, .cpp:164-234), then
// zg[c][m] = (Z'x_m)_c and gnorm2[m] = |x_m|^2 - |Z'x_m|^2 (:236-262).  tab[m][code].
__global__ void __launch_bounds__(256)
k_bolt_snp(const uint8_t* __restrict__ bed, int64_t stride, int64_t N, int C, const double* __restrict__ Z /*[N][C]*/,
           double* __restrict__ tab /*[M][4]*/, double* __restrict__ zg /*[M][C]*/, double* __restrict__ gnorm2) {
  __shared__ double s_red[8][kMaxC + 2];
  __shared__ double s_tab[4];
  __shared__ long long s_cnt[8][2];
  const int m = blockIdx.x, tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const uint8_t* __restrict__ row = bed + (size_t)m * stride;
  long long ac = 0, miss = 0;
  for (int64_t i = tid; i < N; i += 256) {
    const int c = bolt_code(row, i);
    ac += (c == 2) ? 1 : (c == 3) ? 2 : 0;
    miss += (c == 1);
  }
  for (int o = 16; o > 0; o >>= 1) {
    ac += __shfl_xor_sync(0xffffffffu, ac, o);
    miss += __shfl_xor_sync(0xffffffffu, miss, o);
  }
  if (l == 0) {
    s_cnt[w][0] = ac;
    s_cnt[w][1] = miss;
  }
  __syncthreads();
  if (tid == 0) {
    long long a = 0, mi = 0;
    for (int i = 0; i < 8; ++i) {
      a += s_cnt[i][0];
      mi += s_cnt[i][1];
    }
    const double af = (N - mi) > 0 ? 0.5 * (double)a / (double)(N - mi) : 0.0;
    const double mean = af + af, sd = sqrt(2.0 * af * (1.0 - af));
    const double inv = sd > 0.0 ? 1.0 / sd : 0.0;
    s_tab[0] = (0.0 - mean) * inv;   // 00 hom ref
    s_tab[1] = 0.0;                  // 01 missing
    s_tab[2] = (1.0 - mean) * inv;   // 10 het
    s_tab[3] = (2.0 - mean) * inv;   // 11 hom alt
    for (int k = 0; k < 4; ++k) tab[(size_t)m * 4 + k] = s_tab[k];
  }
  __syncthreads();
  double acc[kMaxC + 1];
#pragma unroll
  for (int c = 0; c <= kMaxC; ++c) acc[c] = 0.0;
  for (int64_t i = tid; i < N; i += 256) {
    const double x = s_tab[bolt_code(row, i)];
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      if (c < C) acc[c] += x * Z[(size_t)i * C + c];
    acc[kMaxC] += x * x;
  }
#pragma unroll
  for (int c = 0; c <= kMaxC; ++c) {
    double v = acc[c];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (l == 0) s_red[w][c] = v;
  }
  __syncthreads();
  if (tid == 0) {
    double n2 = 0.0, zz = 0.0;
    for (int i = 0; i < 8; ++i) n2 += s_red[i][kMaxC];
    for (int c = 0; c < C; ++c) {
      double v = 0.0;
      for (int i = 0; i < 8; ++i) v += s_red[i][c];
      zg[(size_t)m * C + c] = v;
      zz += v * v;
    }
    gnorm2[m] = n2 - zz;
  }
}

// One warp per SNP, 16 samples per 32-bit word: the allele / missing counts by popcount -> the 4-entry table, and
// sumsq[m] = sum_i x_mi^2 = n_00 t_00^2 + n_10 t_10^2 + n_11 t_11^2 (exact from the counts).  Z'x_m then is ONE pass of the
// X'v product with v = Z (rvt_bolt_fit_null), instead of a second scalar sweep per SNP (k_bolt_snp: 0.5 s of a 2.3 s fit at
// N = 10^6 x M = 16 384).  Rows are read as words: needs stride % 4 == 0 and a 4-byte aligned panel; padding bits are 00.
__global__ void __launch_bounds__(256)
k_bolt_count(const uint8_t* __restrict__ bed, int64_t stride, int64_t N, int M, double* __restrict__ tab /*[M][4]*/, double* __restrict__ sumsq) {
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5), l = threadIdx.x & 31;
  if (m >= M) return;
  const uint32_t* __restrict__ row = reinterpret_cast<const uint32_t*>(bed + (size_t)m * stride);
  const int64_t nwords = (N + 15) >> 4;
  long long het = 0, hom = 0, mis = 0;
  for (int64_t w = l; w < nwords; w += 32) {
    uint32_t x = row[w];
    if (w == nwords - 1 && (N & 15)) x &= (1u << (2 * (int)(N & 15))) - 1u;      // bits beyond sample N - 1 (00 by the format; masked anyway)
    const uint32_t lo = x & 0x55555555u, hi = (x >> 1) & 0x55555555u;
    mis += __popc(lo & ~hi);   // 01
    het += __popc(hi & ~lo);   // 10
    hom += __popc(hi & lo);    // 11
  }
  for (int o = 16; o > 0; o >>= 1) {
    het += __shfl_xor_sync(0xffffffffu, het, o);
    hom += __shfl_xor_sync(0xffffffffu, hom, o);
    mis += __shfl_xor_sync(0xffffffffu, mis, o);
  }
  if (l == 0) {
    const long long ac = het + 2 * hom, nobs = N - mis, ref = nobs - het - hom;
    const double af = nobs > 0 ? 0.5 * (double)ac / (double)nobs : 0.0;
    const double mean = af + af, sd = sqrt(2.0 * af * (1.0 - af));
    const double inv = sd > 0.0 ? 1.0 / sd : 0.0;
    const double t0 = (0.0 - mean) * inv, t2 = (1.0 - mean) * inv, t3 = (2.0 - mean) * inv;
    tab[(size_t)m * 4 + 0] = t0;
    tab[(size_t)m * 4 + 1] = 0.0;
    tab[(size_t)m * 4 + 2] = t2;
    tab[(size_t)m * 4 + 3] = t3;
    sumsq[m] = (double)ref * t0 * t0 + (double)het * t2 * t2 + (double)hom * t3 * t3;
  }
}
// zg = X'Z arrives in Xy ([M][C], the plain sum over the sample splits); gnorm2[m] = sumsq[m] - |zg[m]|^2
__global__ void k_bolt_gnorm(int M, int C, const double* __restrict__ Xy, double* __restrict__ zg, double* __restrict__ gnorm2 /* in: sumsq */) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double zz = 0.0;
  for (int c = 0; c < C; ++c) {
    const double v = Xy[(size_t)m * C + c];
    zg[(size_t)m * C + c] = v;
    zz += v * v;
  }
  gnorm2[m] -= zz;
}

// X'v over a split of the samples: part[split][m][r] = sum_{i in split} x_mi v[i][r].  One thread holds the R accumulators
// of TWO SNPs of a 128-SNP block (the loads of v are shared by both); the block's 2-bit rows are staged through shared
// memory chunk by chunk.  v rows are read with 16-byte loads when R is even.
constexpr int kBoltSnpPerThread = 2;
constexpr int kBoltXtvBlock = kBoltSnpBlock * kBoltSnpPerThread;   // SNPs per CTA
template <int RMAX>
__global__ void __launch_bounds__(kBoltSnpBlock)
k_bolt_xtv(const uint8_t* __restrict__ bed, int64_t stride, int64_t N, int M, const double* __restrict__ tab, const double* __restrict__ v,
           int R, int64_t split_len, double* __restrict__ part /*[splits][M][R]*/) {
  __shared__ __align__(16) uint8_t s_rows[kBoltXtvBlock][kBoltRowPad];
  const int mA = blockIdx.x * kBoltXtvBlock + threadIdx.x, mB = mA + kBoltSnpBlock;
  const int64_t i0 = (int64_t)blockIdx.y * split_len;
  int64_t i1 = i0 + split_len;
  if (i1 > N) i1 = N;
  double tA[4] = {0, 0, 0, 0}, tB[4] = {0, 0, 0, 0};
  for (int k = 0; k < 4; ++k) {
    if (mA < M) tA[k] = tab[(size_t)mA * 4 + k];
    if (mB < M) tB[k] = tab[(size_t)mB * 4 + k];
  }
  double accA[RMAX], accB[RMAX];
#pragma unroll
  for (int r = 0; r < RMAX; ++r) accA[r] = accB[r] = 0.0;
  const bool vec2 = (R == RMAX) && (RMAX % 2 == 0);   // full, even width: rows of v are 16-byte aligned
  for (int64_t c0 = i0; c0 < i1; c0 += kBoltChunk) {   // i0 and kBoltChunk are multiples of 4
    const int64_t n_here = (i1 - c0 < kBoltChunk) ? (i1 - c0) : kBoltChunk;
    const int nbytes = (int)((n_here + 3) >> 2);
    __syncthreads();
    for (int rr = 0; rr < kBoltXtvBlock; ++rr) {
      const int mm = blockIdx.x * kBoltXtvBlock + rr;
      if (mm >= M) break;
      const uint8_t* __restrict__ src = bed + (size_t)mm * stride + (c0 >> 2);
      for (int b = threadIdx.x; b < nbytes; b += kBoltSnpBlock) s_rows[rr][b] = src[b];
    }
    __syncthreads();
    if (mA < M) {
      const uint8_t* __restrict__ rowA = s_rows[threadIdx.x];
      const uint8_t* __restrict__ rowB = s_rows[threadIdx.x + kBoltSnpBlock];   // (stale bytes when mB >= M: tB is all zero)
      for (int64_t k4 = 0; k4 < n_here; k4 += 4) {
        const unsigned bA = rowA[k4 >> 2], bB = rowB[k4 >> 2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (k4 + q >= n_here) break;
          const double xA = tA[(bA >> (2 * q)) & 3], xB = tB[(bB >> (2 * q)) & 3];
          const double* __restrict__ vr = v + (size_t)(c0 + k4 + q) * R;   // warp-uniform address: broadcast loads
          if (vec2) {
#pragma unroll
            for (int r = 0; r < RMAX; r += 2) {
              const double2 vv = *reinterpret_cast<const double2*>(vr + r);
              accA[r] += xA * vv.x;
              accA[r + 1] += xA * vv.y;
              accB[r] += xB * vv.x;
              accB[r + 1] += xB * vv.y;
            }
          } else {
#pragma unroll
            for (int r = 0; r < RMAX; ++r)
              if (r < R) {
                const double vv = vr[r];
                accA[r] += xA * vv;
                accB[r] += xB * vv;
              }
          }
        }
      }
    }
  }
  if (mA < M) {
    double* o = part + ((size_t)blockIdx.y * M + mA) * R;
#pragma unroll
    for (int r = 0; r < RMAX; ++r)
      if (r < R) o[r] = accA[r];
  }
  if (mB < M) {
    double* o = part + ((size_t)blockIdx.y * M + mB) * R;
#pragma unroll
    for (int r = 0; r < RMAX; ++r)
      if (r < R) o[r] = accB[r];
  }
}

// Xy[m][r] = sum over splits (index order) of part - sum_c zg[m][c] vbot[c][r]     (X_minus' y, BoltLMM.cpp:948-958)
__global__ void k_bolt_xtv_finish(int M, int R, int C, int splits, const double* __restrict__ part, const double* __restrict__ zg,
                                  const double* __restrict__ vbot /*[C][R]*/, double scale, double* __restrict__ Xy) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)M * R) return;
  const int m = (int)(idx / R), r = (int)(idx - (int64_t)m * R);
  double s = 0.0;
  for (int sp = 0; sp < splits; ++sp) s += part[((size_t)sp * M + m) * R + r];
  for (int c = 0; c < C; ++c) s -= zg[(size_t)m * C + c] * vbot[(size_t)c * R + r];
  Xy[idx] = s * scale;
}

// out[i][r] = alpha * sum_m x_mi W[m][r] + beta * add[i][r]   (top rows; X_plus X_y / M + delta y, :960-975).
// One thread holds the R accumulators of SPT consecutive samples (SPT = 4: one whole byte of every SNP row per thread);
// W and the decode tables are staged per 64-SNP block.
template <int RMAX, int SPT>
__global__ void __launch_bounds__(256)
k_bolt_xw(const uint8_t* __restrict__ bed, int64_t stride, int64_t N, int M, const double* __restrict__ tab, const double* __restrict__ W /*[M][R]*/,
          int R, double alpha, double beta, const double* __restrict__ add, double* __restrict__ out) {
  static_assert(SPT == 1 || SPT == 4, "one sample or one byte per thread");
  __shared__ double s_W[kBoltSnpBlock][RMAX];
  __shared__ double s_tab[kBoltSnpBlock][4];
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * SPT;
  double acc[SPT][RMAX];
#pragma unroll
  for (int q = 0; q < SPT; ++q)
#pragma unroll
    for (int r = 0; r < RMAX; ++r) acc[q][r] = 0.0;
  for (int m0 = 0; m0 < M; m0 += kBoltSnpBlock) {
    const int nm = (M - m0 < kBoltSnpBlock) ? (M - m0) : kBoltSnpBlock;
    __syncthreads();
    for (int idx = threadIdx.x; idx < nm * RMAX; idx += 256) {
      const int mm = idx / RMAX, r = idx - mm * RMAX;
      s_W[mm][r] = (r < R) ? W[(size_t)(m0 + mm) * R + r] : 0.0;
    }
    for (int idx = threadIdx.x; idx < nm * 4; idx += 256) s_tab[idx >> 2][idx & 3] = tab[(size_t)m0 * 4 + idx];
    __syncthreads();
    if (i < N) {
      const uint8_t* __restrict__ col = bed + (size_t)m0 * stride + (i >> 2);
      const int sh = 2 * (int)(i & 3);   // 0 when SPT == 4
      for (int mm = 0; mm < nm; ++mm) {
        const unsigned b = col[(size_t)mm * stride] >> sh;
        double x[SPT];
#pragma unroll
        for (int q = 0; q < SPT; ++q) x[q] = s_tab[mm][(b >> (2 * q)) & 3];
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
          const double w = s_W[mm][r];
#pragma unroll
          for (int q = 0; q < SPT; ++q) acc[q][r] += x[q] * w;
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < SPT; ++q)
    if (i + q < N) {
#pragma unroll
      for (int r = 0; r < RMAX; ++r)
        if (r < R) out[(size_t)(i + q) * R + r] = alpha * acc[q][r] + (add ? beta * add[(size_t)(i + q) * R + r] : 0.0);
    }
}

// ---- second generation of the two panel products (option "bolt_kernels" = 2, the default) ----------------------------
// Both are bound by the fp64 pipe (2 N M R multiply-adds per pass against N M / 4 bytes), so what matters is how many
// non-DFMA instructions ride along with each DFMA.  k_bolt_xw2: a thread owns FOUR consecutive samples (one byte of every
// SNP row) and 16 right-hand sides -- 64 accumulators; per SNP it spends one byte load, four table look-ups (the four
// codes of a SNP sit in one 32-byte line of shared memory: one wavefront whatever the codes are) and 8 LDS.128 of W for
// 64 DFMA.  64-thread CTAs so that N / 256 CTAs fill the device from N ~ 40 000 on.  r0: first column of the window
// (a 30-column solve runs as two windows).
template <int RMAX>
__global__ void __launch_bounds__(64)
k_bolt_xw2(const uint8_t* __restrict__ bed, int64_t stride, int64_t N, int M, const double* __restrict__ tab, const double* __restrict__ W /*[M][R]*/,
           int R, int r0, double alpha, double beta, const double* __restrict__ add, double* __restrict__ out) {
  __shared__ __align__(16) double s_W[kBoltSnpBlock][RMAX];
  __shared__ __align__(16) double s_tab[kBoltSnpBlock][4];
  const int64_t i = ((int64_t)blockIdx.x * 64 + threadIdx.x) * 4;
  double acc[4][RMAX];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int r = 0; r < RMAX; ++r) acc[q][r] = 0.0;
  const int nr = (R - r0 < RMAX) ? (R - r0) : RMAX;
  for (int m0 = 0; m0 < M; m0 += kBoltSnpBlock) {
    const int nm = (M - m0 < kBoltSnpBlock) ? (M - m0) : kBoltSnpBlock;
    __syncthreads();
    for (int idx = threadIdx.x; idx < nm * RMAX; idx += 64) {
      const int mm = idx / RMAX, r = idx - mm * RMAX;
      s_W[mm][r] = (r < nr) ? W[(size_t)(m0 + mm) * R + r0 + r] : 0.0;
    }
    for (int idx = threadIdx.x; idx < nm * 4; idx += 64) s_tab[idx >> 2][idx & 3] = tab[(size_t)m0 * 4 + idx];
    __syncthreads();
    if (i < N) {
      const uint8_t* __restrict__ col = bed + (size_t)m0 * stride + (i >> 2);
#pragma unroll 2
      for (int mm = 0; mm < nm; ++mm) {
        const unsigned b = col[(size_t)mm * stride];
        const double x0 = s_tab[mm][b & 3], x1 = s_tab[mm][(b >> 2) & 3], x2 = s_tab[mm][(b >> 4) & 3], x3 = s_tab[mm][(b >> 6) & 3];
#pragma unroll
        for (int r = 0; r < RMAX; r += 2) {
          const double2 w = *reinterpret_cast<const double2*>(&s_W[mm][r]);
          acc[0][r] += x0 * w.x; acc[0][r + 1] += x0 * w.y;
          acc[1][r] += x1 * w.x; acc[1][r + 1] += x1 * w.y;
          acc[2][r] += x2 * w.x; acc[2][r + 1] += x2 * w.y;
          acc[3][r] += x3 * w.x; acc[3][r + 1] += x3 * w.y;
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (i + q < N) {
#pragma unroll
      for (int r = 0; r < RMAX; ++r)
        if (r < nr) out[(size_t)(i + q) * R + r0 + r] = alpha * acc[q][r] + (add ? beta * add[(size_t)(i + q) * R + r0 + r] : 0.0);
    }
}

// k_bolt_xtv2: a thread owns FOUR SNPs (rows t, t+64, t+128, t+192 of a 256-SNP block) and 16 right-hand sides; a row of v
// (warp-uniform address: 8 broadcast LDG.128) now serves 64 DFMA instead of 32, and the decode goes through a table in
// shared memory instead of a dynamically indexed register array (which the compiler keeps in local memory).
constexpr int kBoltXtv2Snps = 4;
constexpr int kBoltXtv2Block = kBoltSnpBlock * kBoltXtv2Snps;   // SNPs per CTA
constexpr int kBoltXtv2Chunk = 512;                            // samples staged per step
constexpr int kBoltXtv2RowPad = kBoltXtv2Chunk / 4 + 4;
template <int RMAX>
__global__ void __launch_bounds__(kBoltSnpBlock)
k_bolt_xtv2(const uint8_t* __restrict__ bed, int64_t stride, int64_t N, int M, const double* __restrict__ tab, const double* __restrict__ v,
            int R, int r0, int64_t split_len, double* __restrict__ part /*[splits][M][R]*/) {
  __shared__ __align__(16) uint8_t s_rows[kBoltXtv2Block][kBoltXtv2RowPad];
  __shared__ __align__(16) double s_tab[4][kBoltXtv2Block];   // [code][SNP]: lane t reads word t of a code's row -- no bank conflict whatever the codes
  const int mbase = blockIdx.x * kBoltXtv2Block;
  const int64_t i0 = (int64_t)blockIdx.y * split_len;
  int64_t i1 = i0 + split_len;
  if (i1 > N) i1 = N;
  for (int idx = threadIdx.x; idx < kBoltXtv2Block * 4; idx += kBoltSnpBlock) {
    const int mm = mbase + (idx >> 2);
    s_tab[idx & 3][idx >> 2] = (mm < M) ? tab[(size_t)mm * 4 + (idx & 3)] : 0.0;
  }
  const int nr = (R - r0 < RMAX) ? (R - r0) : RMAX;
  double acc[kBoltXtv2Snps][RMAX];
#pragma unroll
  for (int k = 0; k < kBoltXtv2Snps; ++k)
#pragma unroll
    for (int r = 0; r < RMAX; ++r) acc[k][r] = 0.0;
  const bool vec2 = (nr == RMAX) && (((size_t)R * 8) % 16 == 0) && (((size_t)r0 * 8) % 16 == 0);
  for (int64_t c0 = i0; c0 < i1; c0 += kBoltXtv2Chunk) {   // i0 and the chunk are multiples of 4
    const int64_t n_here = (i1 - c0 < kBoltXtv2Chunk) ? (i1 - c0) : kBoltXtv2Chunk;
    const int nbytes = (int)((n_here + 3) >> 2);
    __syncthreads();
    for (int rr = 0; rr < kBoltXtv2Block; ++rr) {
      const int mm = mbase + rr;
      if (mm >= M) {   // rows beyond the panel decode through an all-zero table; keep the bytes defined
        for (int b = threadIdx.x; b < nbytes; b += kBoltSnpBlock) s_rows[rr][b] = 0;
        continue;
      }
      const uint8_t* __restrict__ src = bed + (size_t)mm * stride + (c0 >> 2);
      for (int b = threadIdx.x; b < nbytes; b += kBoltSnpBlock) s_rows[rr][b] = src[b];
    }
    __syncthreads();
    const uint8_t* __restrict__ row0 = s_rows[threadIdx.x];
    for (int64_t k4 = 0; k4 < n_here; k4 += 4) {
      unsigned bb[kBoltXtv2Snps];
#pragma unroll
      for (int k = 0; k < kBoltXtv2Snps; ++k) bb[k] = row0[(size_t)k * kBoltSnpBlock * kBoltXtv2RowPad + (k4 >> 2)];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (k4 + q >= n_here) break;
        double x[kBoltXtv2Snps];
#pragma unroll
        for (int k = 0; k < kBoltXtv2Snps; ++k) x[k] = s_tab[(bb[k] >> (2 * q)) & 3][threadIdx.x + k * kBoltSnpBlock];
        const double* __restrict__ vr = v + (size_t)(c0 + k4 + q) * R + r0;   // warp-uniform address: broadcast loads
        if (vec2) {
#pragma unroll
          for (int r = 0; r < RMAX; r += 2) {
            const double2 vv = *reinterpret_cast<const double2*>(vr + r);
#pragma unroll
            for (int k = 0; k < kBoltXtv2Snps; ++k) {
              acc[k][r] += x[k] * vv.x;
              acc[k][r + 1] += x[k] * vv.y;
            }
          }
        } else {
#pragma unroll
          for (int r = 0; r < RMAX; ++r)
            if (r < nr) {
              const double vv = vr[r];
#pragma unroll
              for (int k = 0; k < kBoltXtv2Snps; ++k) acc[k][r] += x[k] * vv;
            }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kBoltXtv2Snps; ++k) {
    const int mm = mbase + threadIdx.x + k * kBoltSnpBlock;
    if (mm < M) {
      double* o = part + ((size_t)blockIdx.y * M + mm) * R + r0;
#pragma unroll
      for (int r = 0; r < RMAX; ++r)
        if (r < nr) o[r] = acc[k][r];
    }
  }
}

// ---- third generation (option "bolt_kernels" = 3, the default): the same thread decompositions with the global loads taken
// off the critical path.  ncu of generation 2 (profiles/r02t_bolt_ncu.txt): 70 % of the warp samples wait on a long
// scoreboard -- the strided byte loads of k_bolt_xw2 and the staging / v-row loads of k_bolt_xtv2 --, the fp64 pipe is 7-16 %
// busy.  Here every byte a CTA needs for the NEXT block arrives by cp.async into the other half of a double buffer while
// the current block is multiplied; the inner loops touch shared memory only.
__device__ __forceinline__ void bolt_cp4(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g));
}
__device__ __forceinline__ void bolt_cp8(void* smem, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g));
}
__device__ __forceinline__ void bolt_cp16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g));
}
__device__ __forceinline__ void bolt_cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void bolt_cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// X w: 128 threads, a thread owns four consecutive samples (one byte per SNP row) and RMAX right-hand sides; per 64-SNP block
// the CTA's 128-byte row segments, the block's rows of W and its decode tables are staged by cp.async (double buffer).
// Requires stride % 4 == 0 and a 4-byte aligned panel (the engine's own copy has a 16-byte pitch).
constexpr int kBoltXw3Threads = 128;
template <int RMAX>
struct BoltXw3Smem {
  uint8_t rows[2][kBoltSnpBlock][kBoltXw3Threads];
  double W[2][kBoltSnpBlock][RMAX];
  double tab[2][kBoltSnpBlock][4];
};
template <int RMAX>
__global__ void __launch_bounds__(kBoltXw3Threads)
k_bolt_xw3(const uint8_t* __restrict__ bed, int64_t stride, int64_t N, int M, const double* __restrict__ tab, const double* __restrict__ W /*[M][R]*/,
           int R, int r0, double alpha, double beta, const double* __restrict__ add, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char bolt_smem_raw[];
  BoltXw3Smem<RMAX>& sm = *reinterpret_cast<BoltXw3Smem<RMAX>*>(bolt_smem_raw);
  const int tid = threadIdx.x;
  const int64_t i0 = (int64_t)blockIdx.x * (kBoltXw3Threads * 4);
  const int64_t i = i0 + (int64_t)tid * 4;
  const int64_t byte0 = i0 >> 2;
  const int64_t rowb = (N + 3) >> 2;
  int words = (int)((rowb - byte0 + 3) >> 2);                   // 4-byte words of this CTA's segment inside a row
  if (words > kBoltXw3Threads / 4) words = kBoltXw3Threads / 4;
  const int nr = (R - r0 < RMAX) ? (R - r0) : RMAX;
  const int nblk = (M + kBoltSnpBlock - 1) / kBoltSnpBlock;
  auto stage = [&](int b, int buf) {
    const int m0 = b * kBoltSnpBlock;
    const int nm = (M - m0 < kBoltSnpBlock) ? (M - m0) : kBoltSnpBlock;
    for (int idx = tid; idx < kBoltSnpBlock * (kBoltXw3Threads / 4); idx += kBoltXw3Threads) {
      const int rr = idx / (kBoltXw3Threads / 4), wd = idx - rr * (kBoltXw3Threads / 4);
      if (rr < nm && wd < words) bolt_cp4(&sm.rows[buf][rr][wd * 4], bed + (size_t)(m0 + rr) * stride + byte0 + wd * 4);
    }
    for (int idx = tid; idx < nm * RMAX; idx += kBoltXw3Threads) {
      const int mm = idx / RMAX, r = idx - mm * RMAX;
      if (r < nr) bolt_cp8(&sm.W[buf][mm][r], W + (size_t)(m0 + mm) * R + r0 + r);
      else sm.W[buf][mm][r] = 0.0;
    }
    for (int idx = tid; idx < nm * 4; idx += kBoltXw3Threads) bolt_cp8(&sm.tab[buf][idx >> 2][idx & 3], tab + (size_t)m0 * 4 + idx);
    bolt_cp_commit();
  };
  double acc[4][RMAX];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int r = 0; r < RMAX; ++r) acc[q][r] = 0.0;
  stage(0, 0);
  for (int b = 0; b < nblk; ++b) {
    const int buf = b & 1;
    if (b + 1 < nblk) {
      stage(b + 1, buf ^ 1);
      bolt_cp_wait<1>();
    } else {
      bolt_cp_wait<0>();
    }
    __syncthreads();
    const int nm = (M - b * kBoltSnpBlock < kBoltSnpBlock) ? (M - b * kBoltSnpBlock) : kBoltSnpBlock;
    if (i < N) {
#pragma unroll 4
      for (int mm = 0; mm < nm; ++mm) {
        const unsigned bb = sm.rows[buf][mm][tid];
        const double x0 = sm.tab[buf][mm][bb & 3], x1 = sm.tab[buf][mm][(bb >> 2) & 3], x2 = sm.tab[buf][mm][(bb >> 4) & 3],
                     x3 = sm.tab[buf][mm][(bb >> 6) & 3];
#pragma unroll
        for (int r = 0; r < RMAX; r += 2) {
          const double2 w = *reinterpret_cast<const double2*>(&sm.W[buf][mm][r]);
          acc[0][r] += x0 * w.x; acc[0][r + 1] += x0 * w.y;
          acc[1][r] += x1 * w.x; acc[1][r + 1] += x1 * w.y;
          acc[2][r] += x2 * w.x; acc[2][r + 1] += x2 * w.y;
          acc[3][r] += x3 * w.x; acc[3][r + 1] += x3 * w.y;
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (i + q < N) {
#pragma unroll
      for (int r = 0; r < RMAX; ++r)
        if (r < nr) out[(size_t)(i + q) * R + r0 + r] = alpha * acc[q][r] + (add ? beta * add[(size_t)(i + q) * R + r0 + r] : 0.0);
    }
}

// X'v: 64 threads, a thread owns four SNPs (rows t, t+64, t+128, t+192 of a 256-SNP block) and RMAX right-hand sides; per chunk of
// CHUNK samples the 256 row segments (pitch CHUNK/4 + 4 bytes: conflict-free word reads) and the chunk's rows of v are staged by
// cp.async (double buffer); the decode table sits transposed in shared memory.  Same alignment requirement as k_bolt_xw3.
template <int RMAX, int CHUNK>
struct BoltXtv3Smem {
  double v[2][CHUNK][RMAX];
  double tab[4][kBoltXtv2Block];
  uint32_t rows[2][kBoltXtv2Block][CHUNK / 16 + 1];
};
template <int RMAX, int CHUNK>
__global__ void __launch_bounds__(kBoltSnpBlock)
k_bolt_xtv3(const uint8_t* __restrict__ bed, int64_t stride, int64_t N, int M, const double* __restrict__ tab, const double* __restrict__ v,
            int R, int r0, int64_t split_len, double* __restrict__ part /*[splits][M][R]*/) {
  extern __shared__ __align__(16) unsigned char bolt_smem_raw[];
  typedef BoltXtv3Smem<RMAX, CHUNK> Smem;
  Smem& sm = *reinterpret_cast<Smem*>(bolt_smem_raw);
  constexpr int kWords = CHUNK / 16;
  const int tid = threadIdx.x;
  const int mbase = blockIdx.x * kBoltXtv2Block;
  const int64_t i0 = (int64_t)blockIdx.y * split_len;          // a multiple of CHUNK
  int64_t i1 = i0 + split_len;
  if (i1 > N) i1 = N;
  const int64_t rowb = (N + 3) >> 2;
  const int nr = (R - r0 < RMAX) ? (R - r0) : RMAX;
  for (int idx = tid; idx < kBoltXtv2Block * 4; idx += kBoltSnpBlock) {
    const int mm = mbase + (idx >> 2);
    sm.tab[idx & 3][idx >> 2] = (mm < M) ? tab[(size_t)mm * 4 + (idx & 3)] : 0.0;
  }
  const int nchunk = (int)((i1 - i0 + CHUNK - 1) / CHUNK);
  auto stage = [&](int c, int buf) {
    const int64_t c0 = i0 + (int64_t)c * CHUNK;
    const int64_t n_here = (i1 - c0 < CHUNK) ? (i1 - c0) : CHUNK;
    const int64_t byte0 = c0 >> 2;
    int words = (int)((rowb - byte0 + 3) >> 2);
    if (words > kWords) words = kWords;
    for (int idx = tid; idx < kBoltXtv2Block * kWords; idx += kBoltSnpBlock) {
      const int rr = idx / kWords, wd = idx - rr * kWords;
      const int mm = mbase + rr;
      if (mm < M && wd < words) bolt_cp4(&sm.rows[buf][rr][wd], bed + (size_t)mm * stride + byte0 + wd * 4);
      else sm.rows[buf][rr][wd] = 0u;
    }
    for (int idx = tid; idx < (int)n_here * RMAX; idx += kBoltSnpBlock) {
      const int ii = idx / RMAX, r = idx - ii * RMAX;
      if (r < nr) bolt_cp8(&sm.v[buf][ii][r], v + (size_t)(c0 + ii) * R + r0 + r);
      else sm.v[buf][ii][r] = 0.0;
    }
    bolt_cp_commit();
  };
  double acc[kBoltXtv2Snps][RMAX];
#pragma unroll
  for (int k = 0; k < kBoltXtv2Snps; ++k)
#pragma unroll
    for (int r = 0; r < RMAX; ++r) acc[k][r] = 0.0;
  if (nchunk > 0) stage(0, 0);
  for (int c = 0; c < nchunk; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunk) {
      stage(c + 1, buf ^ 1);
      bolt_cp_wait<1>();
    } else {
      bolt_cp_wait<0>();
    }
    __syncthreads();
    const int64_t c0 = i0 + (int64_t)c * CHUNK;
    const int n_here = (int)((i1 - c0 < CHUNK) ? (i1 - c0) : CHUNK);
    for (int k16 = 0; k16 < n_here; k16 += 16) {
      uint32_t wv[kBoltXtv2Snps];
#pragma unroll
      for (int k = 0; k < kBoltXtv2Snps; ++k) wv[k] = sm.rows[buf][tid + k * kBoltSnpBlock][k16 >> 4];
      const int lim = (n_here - k16 < 16) ? (n_here - k16) : 16;
#pragma unroll 4
      for (int q = 0; q < 16; ++q) {
        if (q >= lim) break;
        double x[kBoltXtv2Snps];
#pragma unroll
        for (int k = 0; k < kBoltXtv2Snps; ++k) x[k] = sm.tab[(wv[k] >> (2 * q)) & 3][tid + k * kBoltSnpBlock];
#pragma unroll
        for (int r = 0; r < RMAX; r += 2) {
          const double2 vv = *reinterpret_cast<const double2*>(&sm.v[buf][k16 + q][r]);   // warp-uniform: broadcast
#pragma unroll
          for (int k = 0; k < kBoltXtv2Snps; ++k) {
            acc[k][r] += x[k] * vv.x;
            acc[k][r + 1] += x[k] * vv.y;
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < kBoltXtv2Snps; ++k) {
    const int mm = mbase + tid + k * kBoltSnpBlock;
    if (mm < M) {
      double* o = part + ((size_t)blockIdx.y * M + mm) * R + r0;
#pragma unroll
      for (int r = 0; r < RMAX; ++r)
        if (r < nr) o[r] = acc[k][r];
    }
  }
}

// bottom rows: out[c][r] = alpha * sum_m zg[m][c] W[m][r] + beta * add[c][r]   (one thread per (c, r), SNP order)
__global__ void k_bolt_bot(int M, int R, int C, const double* __restrict__ zg, const double* __restrict__ W, double alpha, double beta,
                           const double* __restrict__ add, double* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * R) return;
  const int c = idx / R, r = idx - c * R;
  double s = 0.0;
  for (int m = 0; m < M; ++m) s += zg[(size_t)m * C + c] * W[(size_t)m * R + r];
  out[idx] = alpha * s + (add ? beta * add[idx] : 0.0);
}

// bottom rows from top rows: vbot[c][r] = sum_i Z[i][c] v[i][r]  (projectCovariate, BoltPlinkLoader.cpp:266-271);
// one CTA per (c, r), fixed-order tree
__global__ void __launch_bounds__(256)
k_bolt_project(int64_t N, int R, int C, const double* __restrict__ Z, const double* __restrict__ v, double* __restrict__ vbot) {
  __shared__ double s[256];
  const int c = blockIdx.x / R, r = blockIdx.x - c * R;
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += 256) a += Z[(size_t)i * C + c] * v[(size_t)i * R + r];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) vbot[(size_t)c * R + r] = s[0];
}

// projected products of two (N + C) x R vectors: partial[cta][r] over the top rows (fixed grid, fixed-order trees);
// the host adds the CTAs in order and subtracts the bottom rows (projDot, BoltLMM.cpp:1064-1097)
constexpr int kBoltDotCtas = 256;
__global__ void __launch_bounds__(256)
k_bolt_dot(int64_t N, int R, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ partial /*[ctas][R]*/) {
  __shared__ double s[256];
  const int64_t total = N * R;
  // thread t of CTA k owns the elements e = (k*256 + t) + j * (ctas*256): with R | 256*ctas the column of e is fixed
  const int64_t step = (int64_t)gridDim.x * 256;
  for (int r = 0; r < R; ++r) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < N; i += step) acc += a[(size_t)i * R + r] * b[(size_t)i * R + r];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)blockIdx.x * R + r] = s[0];
    __syncthreads();
  }
  (void)total;
}

// elementwise column-scaled updates on all (N + C) rows:  y[i][r] = ca[r] * a[i][r] + cb[r] * b[i][r]
__global__ void k_bolt_axpby(int64_t rows, int R, const double* __restrict__ ca, const double* __restrict__ a, const double* __restrict__ cb,
                             const double* __restrict__ b, double* __restrict__ y) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * R) return;
  const int r = (int)(idx % R);
  y[idx] = ca[r] * a[idx] + (b ? cb[r] * b[idx] : 0.0);
}

// y[e] += beta * a[e]   (the "+ delta y" term of computeHx after the sum over ranks of the sharded product)
__global__ void k_bolt_axpy(int64_t n, double beta, const double* __restrict__ a, double* __restrict__ y) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) y[e] += beta * a[e];
}

// normalised genotype columns of chosen SNPs as top rows: out[i][k] = x_{idx[k], i}  (loadRandomSNPWithCov, .cpp:441-480);
// idx[k] < 0: the SNP lives on another rank, the column is zero here (the sum over ranks fills it)
__global__ void k_bolt_columns(const uint8_t* __restrict__ bed, int64_t stride, int64_t N, const int* __restrict__ idx, int K,
                               const double* __restrict__ tab, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * K) return;
  const int64_t i = e / K;
  const int k = (int)(e - i * K);
  const int m = idx[k];
  out[e] = m < 0 ? 0.0 : tab[(size_t)m * 4 + bolt_code(bed + (size_t)m * stride, i)];
}
// second stage of k_bolt_dot: out[r] = sum over CTAs (index order) - sum_c abot[c][r] bbot[c][r]  (abot null: plain sum)
__global__ void k_bolt_dot_finish(int ctas, int R, int C, const double* __restrict__ partial, const double* __restrict__ abot,
                                  const double* __restrict__ bbot, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  double s = 0.0;
  for (int k = 0; k < ctas; ++k) s += partial[(size_t)k * R + r];
  if (abot)
    for (int c = 0; c < C; ++c) s -= abot[(size_t)c * R + r] * bbot[(size_t)c * R + r];
  out[r] = s;
}

// dst[i][k] = scale * src[i][col] for every k < K, all rows (one column of a vector replicated K times)
__global__ void k_bolt_bcast(int64_t rows, int Rsrc, int col, int K, double scale, const double* __restrict__ src, double* __restrict__ dst) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * K) return;
  dst[e] = scale * src[(size_t)(e / K) * Rsrc + col];
}
#endif  // __CUDACC__

}  // namespace rvt
