// lmm.cuh -- A13: the FastLMM score step, FastLMM::Impl::TestCovariate score branch (regression/FastLMM.cpp:215-249):
//
//   u~ = U'(g - gbar)                                   U: the N x N eigenvectors of the kinship (kinshipU)
//   Ustat = sum_i u~_i uResid_i / (lambda_i + delta) / sigma2
//   Vstat = u~' scaledK u~ / sigma2,   scaledK = D - D ux (ux' D ux)^-1 ux' D,  D = diag(1/(lambda + delta))   (:131-138)
//   stat = Ustat^2 / Vstat, p = gsl_cdf_chisq_Q(stat, 1);  Vstat <= 0 -> stat = 0, p = 1
//
// The null fit (delta search, beta, sigma2: FastLMM.cpp:27-140) stays with the caller, who hands over U, lambda, delta,
// sigma2, uResid = U'y - U'X beta and ux = U'X, exactly the members FitNullModel leaves behind.
//
// The only O(N^2) piece is the rotation U'g: 2 N^2 flops per variant in the reference (float Eigen GEMV).  Here U is
// kept as the balanced base-256 digits of its 2^-29 fixed-point image (|U_si| <= 1), 16 eigenvectors x 4 digits = one
// 64-row tile in the engine's tiled layout, and U'G for a block of <= 64 variants is N/16 tile-PAIR units of the
// tensor-core sweep (kind::i8, exact integers): u~ comes out exact for the fixed-point image, i.e. more accurate than
// the reference's float32 product.  Centring is applied afterwards: U'(g - gbar 1) = U'g - gbar t, t = U'1.
//
// Covariance band of `--meta cov` for related samples -- MetaCovFamQtl (src/Model.cpp:437-498) on FastLMM::TransformCentered,
// GetCovXX, GetCovXZ, GetCovZZ (regression/FastLMM.cpp:538-625):
//   x~_v = U'(g_v - gbar_v)                                     TransformCentered
//   covXX(v, w) = x~_v' D x~_w / sigma2,  covXZ(v) = x~_v' D ux / sigma2,  covZZ = ux' D ux / sigma2
//   entry = (covXX - covXZ_v covZZ^-1 covXZ_w') / N             MetaCovTest::printCovariance (src/Model.cpp:990-996)
// The rotated vectors t_v = U'g_v come out of the same tensor-core units as the score step and are KEPT, scaled by
// sqrt(d_i), as rows of Y (nv x N doubles); the band is then a plain fp64 Gram of rows of Y over the tile pairs inside the
// window (k_lmm_gram) plus the centring / projection corrections, which need only per-variant scalars the score step already
// has (gbar, t'D u1, t'D ux):   x~_v' D x~_w = y_v.y_w - gbar_w s3_v - gbar_v s3_w + gbar_v gbar_w (u1' D u1).
#pragma once
#include "../../include/rvtests_b200.h"
#include "common.cuh"
#include "davies.cuh"

namespace rvt {

constexpr int kLmmShift = 29;          // U * 2^29 rounded: |R| <= 2^29 < 2^30, 4 balanced digits
constexpr int kLmmAcc = 3 + kMaxC;     // per variant: S1, S2, S3, S4[C]

struct LmmNull {
  int64_t N;       // samples == eigenvectors
  int32_t C;
  int32_t nb;      // eigenvector tiles (16 eigenvectors each)
  double delta, sigma2;
  const double* a;   // [16 nb]  uResid_i / (lambda_i + delta) / sigma2     (0 beyond N)
  const double* d;   // [16 nb]  1 / (lambda_i + delta)
  const double* t;   // [16 nb]  (U'1)_i
  const double* w;   // [16 nb][kMaxC]  d_i ux_il
  double ta, ttd, tw[kMaxC];        // sum_i t_i a_i, sum_i t_i^2 d_i, sum_i t_i w_il
  double xdx_inv[kMaxC * kMaxC];    // (ux' D ux)^-1
};

// One CTA per eigenvector of a column panel: digits of U[:, col] into its tile rows; t_col = sum_s R_s (exact).
// panel: ncols x N floats (column-major: column c contiguous), col0 = eigenvector index of its first column.
__global__ void __launch_bounds__(256)
k_lmm_digits(const float* __restrict__ panel, int64_t N, int64_t col0, int8_t* __restrict__ tiles, int64_t tile_bytes,
             long long* __restrict__ tsum /*[16 nb]*/) {
  __shared__ long long s_part[8];
  const int64_t col = col0 + blockIdx.x;
  const float* __restrict__ src = panel + (size_t)blockIdx.x * N;
  int8_t* __restrict__ tile = tiles + (size_t)(col >> 4) * tile_bytes;
  const int row0 = 4 * (int)(col & 15);
  long long loc = 0;
  for (int64_t s = threadIdx.x; s < N; s += blockDim.x) {
    long long R = llrint(ldexp((double)src[s], kLmmShift));
    loc += R;
    int8_t* p = tile + ((size_t)(s >> 7) * kTileRows + row0) * 128 + (s & 127);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long dg = ((R + 128) & 255) - 128;
      p[k * 128] = (int8_t)dg;
      R = (R - dg) >> 8;
    }
  }
  for (int o = 16; o > 0; o >>= 1) loc += __shfl_xor_sync(0xffffffffu, loc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = loc;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long s = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += s_part[i];
    tsum[col] = s;
  }
}

// per-eigenvector constants (one thread each) and their fixed-order sums (thread 0 of block 0 afterwards: k_lmm_consts2)
__global__ void k_lmm_consts(int64_t N, int C, const float* __restrict__ lambda, const float* __restrict__ uResid,
                             const float* __restrict__ ux /*N x C col-major*/, const long long* __restrict__ tsum, double delta,
                             double sigma2, double* __restrict__ a, double* __restrict__ d, double* __restrict__ t, double* __restrict__ w) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double di = 1.0 / ((double)fabsf(lambda[i]) + delta);   // lambda.cwiseAbs(), FastLMM.cpp:46-50
  d[i] = di;
  a[i] = (double)uResid[i] * di / sigma2;
  t[i] = ldexp((double)tsum[i], -kLmmShift);
  for (int l = 0; l < C; ++l) w[(size_t)i * kMaxC + l] = di * (double)ux[(size_t)l * N + i];
}
__global__ void k_lmm_consts2(LmmNull* nm) {   // <<<1, 32>>>: sums in index order, lane = quantity
  const int q = threadIdx.x, C = nm->C;
  if (q >= 2 + C) return;
  double s = 0.0;
  for (int64_t i = 0; i < nm->N; ++i) {
    const double ti = nm->t[i];
    s += (q == 0) ? ti * nm->a[i] : (q == 1) ? ti * ti * nm->d[i] : ti * nm->w[(size_t)i * kMaxC + (q - 2)];
  }
  if (q == 0) nm->ta = s;
  else if (q == 1) nm->ttd = s;
  else nm->tw[q - 2] = s;
}

// One CTA per sweep unit (variant tile x eigenvector tile b): this tile's share of the per-variant sums.
// acc: [n_units][64][kLmmAcc]
__global__ void __launch_bounds__(64)
k_lmm_reduce(const GeneDesc* __restrict__ units, int n_units, int S, const SweepPartial* __restrict__ parts, const LmmNull* __restrict__ nm,
             double* __restrict__ acc, double* __restrict__ Y /* nullable: [nv][ldY] rotated rows scaled by sqrt(d) */, int64_t ldY,
             int64_t v0 /* variant index of the tile's first row */) {
  const int u = blockIdx.x, j = threadIdx.x;
  if (u >= n_units) return;
  const GeneDesc gd = units[u];
  const int b = (int)gd.var0_b, C = nm->C;
  double s1 = 0, s2 = 0, s3 = 0, s4[kMaxC];
#pragma unroll
  for (int l = 0; l < kMaxC; ++l) s4[l] = 0;
  if (j < gd.M) {
    const SweepPartial* __restrict__ gp = parts + (size_t)u * S;
    for (int il = 0; il < 16; ++il) {
      long long dg[4] = {0, 0, 0, 0};
      for (int sp = 0; sp < S; ++sp)
        for (int k = 0; k < 4; ++k) dg[k] += gp[sp].d[j][4 * il + k];
      const double ut = ldexp((double)(dg[0] + (dg[1] << 8) + (dg[2] << 16) + (dg[3] << 24)), -kLmmShift);
      const int i = 16 * b + il;
      const double di = nm->d[i], ti = nm->t[i];
      if (Y) Y[(size_t)(v0 + j) * ldY + i] = ut * sqrt(di);
      s1 += ut * nm->a[i];
      s2 += ut * ut * di;
      s3 += ut * ti * di;
      for (int l = 0; l < C; ++l) s4[l] += ut * nm->w[(size_t)i * kMaxC + l];
    }
  }
  double* o = acc + ((size_t)u * kTileRows + j) * kLmmAcc;
  o[0] = s1;
  o[1] = s2;
  o[2] = s3;
  for (int l = 0; l < kMaxC; ++l) o[3 + l] = s4[l];
}

// One thread per variant of the tile: sum the eigenvector tiles in index order, centre, finish the statistics.
struct LmmVar {   // per variant, for the covariance band
  double gbar, s3, q[kMaxC];   // mean genotype, t'D u1, t'D ux - gbar u1'D ux
  int32_t poly, pad;
};

__global__ void __launch_bounds__(64)
k_lmm_final(int M, int nb, const double* __restrict__ acc /*[nb][64][kLmmAcc]*/, const RowCounts* __restrict__ counts, const LmmNull* __restrict__ nm,
            rvt_lmm_result* __restrict__ out, LmmVar* __restrict__ vars /* nullable: [M] */) {
  const int j = threadIdx.x;
  if (j >= M) return;
  const int C = nm->C;
  double s[kLmmAcc];
  for (int q = 0; q < kLmmAcc; ++q) s[q] = 0.0;
  for (int b = 0; b < nb; ++b) {
    const double* o = acc + ((size_t)b * kTileRows + j) * kLmmAcc;
    for (int q = 0; q < 3 + C; ++q) s[q] += o[q];
  }
  const double N = (double)nm->N;
  const double ac = (double)counts[j].n1 + 2.0 * (double)counts[j].n2;
  const double gbar = ac / N;                       // g.colwise().mean(), FastLMM.cpp:219-221
  const double U = s[0] - gbar * nm->ta;            // (already / sigma2 through a_i)
  const double quad = s[1] - 2.0 * gbar * s[2] + gbar * gbar * nm->ttd;
  double q[kMaxC], proj = 0.0;
  for (int l = 0; l < C; ++l) q[l] = s[3 + l] - gbar * nm->tw[l];
  for (int l = 0; l < C; ++l)
    for (int m = 0; m < C; ++m) proj += q[l] * nm->xdx_inv[l * C + m] * q[m];
  const double V = (quad - proj) / nm->sigma2;
  if (vars) {
    LmmVar lv;
    lv.gbar = gbar;
    lv.s3 = s[2];
    for (int l = 0; l < kMaxC; ++l) lv.q[l] = l < C ? q[l] : 0.0;
    const long long n1 = counts[j].n1, n2 = counts[j].n2, n0 = nm->N - n1 - n2;
    lv.poly = !(n0 == nm->N || n1 == nm->N || n2 == nm->N) && counts[j].bad == 0;
    lv.pad = 0;
    vars[j] = lv;
  }
  rvt_lmm_result r;
  memset(&r, 0, sizeof(r));
  r.af = 0.5 * ac / N;
  r.U = U;
  r.V = V;
  if (counts[j].bad > 0) {
    r.ok = 0;
    r.stat = 0.0;
    r.pvalue = nan("");
  } else if (V > 0.0) {
    r.ok = 1;
    r.stat = U * U / V;
    r.pvalue = chisq_q(r.stat, 1.0);
  } else {
    r.ok = 1;
    r.stat = 0.0;
    r.pvalue = 1.0;
  }
  out[j] = r;
}

// fp64 Gram of two 64-row tiles of Y: G[a][b] = sum_k Y[va + a][k] Y[vb + b][k].  One CTA (256 threads) per tile pair, each
// thread a 4 x 4 block of the 64 x 64 result; K walks the N eigen-coordinates 32 at a time through shared memory.
struct LmmPair {
  int64_t va, vb;
  int32_t Ma, Mb;
};
__global__ void __launch_bounds__(256)
k_lmm_gram(const LmmPair* __restrict__ pairs, int n_pairs, const double* __restrict__ Y, int64_t ldY, int64_t K, double* __restrict__ G /*[n_pairs][64][64]*/) {
  __shared__ double sA[32][kTileRows + 1], sB[32][kTileRows + 1];
  const int p = blockIdx.x, tid = threadIdx.x;
  if (p >= n_pairs) return;
  const LmmPair pr = pairs[p];
  const int ty = tid >> 4, tx = tid & 15;   // rows 4 ty .. 4 ty + 3 of A, rows 4 tx .. of B
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int64_t k0 = 0; k0 < K; k0 += 32) {
    for (int idx = tid; idx < 32 * kTileRows; idx += 256) {
      const int r = idx >> 5, kk = idx & 31;   // consecutive threads walk K: coalesced rows of Y
      const int64_t k = k0 + kk;
      sA[kk][r] = (r < pr.Ma && k < K) ? Y[(size_t)(pr.va + r) * ldY + k] : 0.0;
      sB[kk][r] = (r < pr.Mb && k < K) ? Y[(size_t)(pr.vb + r) * ldY + k] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = sA[kk][4 * ty + i];
        b[i] = sB[kk][4 * tx + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
  double* out = G + (size_t)p * kTileRows * kTileRows;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(4 * ty + i) * kTileRows + 4 * tx + j] = acc[i][j];
}

// band entries of the pairs: centring + covariate projection + the reference's scaling (header)
__global__ void __launch_bounds__(128)
k_lmm_band(const LmmPair* __restrict__ pairs, int n_pairs, const double* __restrict__ G, const LmmVar* __restrict__ vars, const LmmNull* __restrict__ nm,
           const int* __restrict__ jmax, int wmax, double* __restrict__ band) {
  const int p = blockIdx.x, tid = threadIdx.x;
  if (p >= n_pairs) return;
  const LmmPair pr = pairs[p];
  const int C = nm->C;
  const double scale = 1.0 / (nm->sigma2 * (double)nm->N);
  const double* g = G + (size_t)p * kTileRows * kTileRows;
  for (int idx = tid; idx < pr.Ma * pr.Mb; idx += 128) {
    const int a = idx / pr.Mb, b = idx - a * pr.Mb;
    const int64_t vi = pr.va + a, vj = pr.vb + b;
    if (vj < vi || vj > jmax[vi]) continue;
    double val = nan("");
    const LmmVar x = vars[vi], y = vars[vj];
    if (x.poly && y.poly) {
      const double xx = g[a * kTileRows + b] - y.gbar * x.s3 - x.gbar * y.s3 + x.gbar * y.gbar * nm->ttd;
      double proj = 0.0;
      for (int l = 0; l < C; ++l)
        for (int m = 0; m < C; ++m) proj += x.q[l] * nm->xdx_inv[l * C + m] * y.q[m];
      val = (xx - proj) * scale;
    }
    band[(size_t)vi * (wmax + 1) + (vj - vi)] = val;
  }
}

}  // namespace rvt
