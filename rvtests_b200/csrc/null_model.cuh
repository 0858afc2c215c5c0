// null_model.cuh -- linear null model on the device and its fixed-point image "E".
//
// Replaces LinearRegression::FitLinearModel (regression/LinearRegression.cpp:20-69):
//   (X'X)^-1 by Cholesky (:33-35), beta = (X'X)^-1 X'y (:42), resid = y - X beta (:51-52),
//   sigma2 = ||resid||^2 / n (:60, the MLE)
// as used once per run by SkatTest/SkatOTest (src/Model.h:2672-2699) and once per gene by
// CMCTest/ZegginiTest (src/Model.h:850, :1207 -- same fit, same numbers, done once here).
// All reductions are two-stage with a fixed grid so the result is run-to-run deterministic.
#pragma once
#include "common.cuh"

namespace rvt {

constexpr int kNullBlocks = 296;   // 2 x 148 SMs
constexpr int kNullThreads = 256;
constexpr int kNullAcc = kMaxC * (kMaxC + 1) / 2 + kMaxC + 1;  // X'X upper, X'y, sum|x0-1|

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i];
  return s;
}
__device__ __forceinline__ double block_max(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s = fmax(s, sh[i]);
  return s;
}

// stage 1: per-block partial sums of X'X (upper), X'y and sum |x_i0 - 1|
__global__ void __launch_bounds__(kNullThreads) k_null_moments(int64_t N, int C, const double* __restrict__ X,
                                                              const double* __restrict__ y,
                                                              double* __restrict__ part /*[blocks][kNullAcc]*/) {
  __shared__ double sh[32];
  double acc[kNullAcc];
#pragma unroll
  for (int a = 0; a < kNullAcc; ++a) acc[a] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    double x[kMaxC];
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) x[c] = (c < C) ? X[(size_t)c * N + i] : 0.0;
    double yi = y[i];
    int p = 0;
#pragma unroll
    for (int a = 0; a < kMaxC; ++a)
#pragma unroll
      for (int b = a; b < kMaxC; ++b) acc[p++] += x[a] * x[b];
#pragma unroll
    for (int a = 0; a < kMaxC; ++a) acc[p++] += x[a] * yi;
    acc[p] += fabs(x[0] - 1.0);
  }
  for (int a = 0; a < kNullAcc; ++a) {
    double s = block_sum(acc[a], sh);
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * kNullAcc + a] = s;
  }
}

// stage 2 (one thread): fixed-order sum of the partials, Cholesky, inverse, beta.
// status: 0 ok, 1 X'X not positive definite, 2 column 0 is not the intercept
__global__ void k_null_solve(int C, int nblocks, const double* __restrict__ part, NullModel* nm,
                             double* beta /*[kMaxC]*/, int* status) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double acc[kNullAcc];
  for (int a = 0; a < kNullAcc; ++a) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += part[(size_t)b * kNullAcc + a];
    acc[a] = s;
  }
  double xtx[kMaxC][kMaxC], xty[kMaxC], l[kMaxC][kMaxC];
  int p = 0;
  for (int a = 0; a < kMaxC; ++a)
    for (int b = a; b < kMaxC; ++b) {
      xtx[a][b] = xtx[b][a] = acc[p++];
    }
  for (int a = 0; a < kMaxC; ++a) xty[a] = acc[p++];
  *status = 0;
  if (acc[p] != 0.0) *status = 2;
  for (int j = 0; j < C; ++j) {
    double d = xtx[j][j];
    for (int k = 0; k < j; ++k) d -= l[j][k] * l[j][k];
    if (!(d > 0.0)) {
      *status = 1;
      return;
    }
    d = sqrt(d);
    l[j][j] = d;
    for (int i = j + 1; i < C; ++i) {
      double s = xtx[i][j];
      for (int k = 0; k < j; ++k) s -= l[i][k] * l[j][k];
      l[i][j] = s / d;
    }
  }
  for (int j = 0; j < C; ++j) {
    double col[kMaxC];
    for (int i = 0; i < C; ++i) col[i] = (i == j) ? 1.0 : 0.0;
    for (int i = 0; i < C; ++i) {
      double s = col[i];
      for (int k = 0; k < i; ++k) s -= l[i][k] * col[k];
      col[i] = s / l[i][i];
    }
    for (int i = C - 1; i >= 0; --i) {
      double s = col[i];
      for (int k = i + 1; k < C; ++k) s -= l[k][i] * col[k];
      col[i] = s / l[i][i];
    }
    for (int i = 0; i < C; ++i) nm->xtx_inv[i * C + j] = col[i];
  }
  for (int a = 0; a < C; ++a) {
    double s = 0.0;
    for (int b = 0; b < C; ++b) s += nm->xtx_inv[a * C + b] * xty[b];
    beta[a] = s;
  }
  // column 0 is the intercept: X'X[0][l] = sum_i x_il, X'y[0] = sum_i y_i
  double rs = xty[0];
  for (int l = 0; l < C; ++l) {
    nm->xsum[l] = xtx[0][l];
    rs -= beta[l] * xtx[0][l];
  }
  nm->rsum = rs;
}

// stage 3: residuals + per-block RSS and max-abs of r and of every covariate column
__global__ void __launch_bounds__(kNullThreads) k_null_resid(int64_t N, int C, const double* __restrict__ X,
                                                            const double* __restrict__ y,
                                                            const double* __restrict__ beta,
                                                            double* __restrict__ resid,
                                                            double* __restrict__ part /*[blocks][2+kMaxC]*/,
                                                            int keep_y /* 1: resid := y as given (caller-supplied null residual) */) {
  __shared__ double sh[32];
  double b[kMaxC];
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) b[c] = (c < C) ? beta[c] : 0.0;
  double rss = 0.0, mx[kMaxC + 1];
#pragma unroll
  for (int c = 0; c <= kMaxC; ++c) mx[c] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    double pred = 0.0;
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
      if (c < C) {
        double x = X[(size_t)c * N + i];
        pred += x * b[c];
        mx[c + 1] = fmax(mx[c + 1], fabs(x));
      }
    double r = keep_y ? y[i] : y[i] - pred;
    resid[i] = r;
    rss += r * r;
    mx[0] = fmax(mx[0], fabs(r));
  }
  double s = block_sum(rss, sh);
  if (threadIdx.x == 0) part[(size_t)blockIdx.x * (2 + kMaxC) + 0] = s;
  for (int c = 0; c <= kMaxC; ++c) {
    double m = block_max(mx[c], sh);
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * (2 + kMaxC) + 1 + c] = m;
  }
}

// stage 4 (one thread): sigma2 and the power-of-two fixed-point scales
__global__ void k_null_finish(int64_t N, int C, int nblocks, const double* __restrict__ part, NullModel* nm,
                              int* shift /*[kMaxC+1]*/, double sigma2_given /* < 0: use RSS/N */) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double rss = 0.0, mx[kMaxC + 1];
  for (int c = 0; c <= kMaxC; ++c) mx[c] = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    rss += part[(size_t)b * (2 + kMaxC)];
    for (int c = 0; c <= kMaxC; ++c) mx[c] = fmax(mx[c], part[(size_t)b * (2 + kMaxC) + 1 + c]);
  }
  nm->sigma2 = (sigma2_given >= 0.0) ? sigma2_given : rss / (double)N;
  for (int v = 0; v <= C; ++v) {
    int ex = 0;
    if (mx[v] > 0.0) frexp(mx[v], &ex);  // mx < 2^ex
    int e = 30 - ex;                      // |value| * 2^e < 2^30
    shift[v] = e;
    nm->scale[v] = ldexp(1.0, -e);
    nm->vsum[v] = 0;
  }
}

// stage 5: balanced base-256 digits of the fixed-point images; integer column sums
__global__ void __launch_bounds__(kNullThreads) k_build_E(int64_t N, int C, const double* __restrict__ X,
                                                         const double* __restrict__ resid,
                                                         const int* __restrict__ shift, int8_t* __restrict__ E,
                                                         int64_t ldE, NullModel* nm) {
  __shared__ unsigned long long ssum[kMaxC + 1];
  if (threadIdx.x <= kMaxC) ssum[threadIdx.x] = 0ull;
  __syncthreads();
  long long loc[kMaxC + 1];
#pragma unroll
  for (int v = 0; v <= kMaxC; ++v) loc[v] = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int v = 0; v <= kMaxC; ++v)
      if (v <= C) {
        double val = (v == 0) ? resid[i] : X[(size_t)(v - 1) * N + i];
        long long R = llrint(ldexp(val, shift[v]));
        loc[v] += R;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          long long d = ((R + 128) & 255) - 128;
          E[(size_t)(4 * v + k) * ldE + i] = (int8_t)d;
          R = (R - d) >> 8;
        }
      }
  }
#pragma unroll
  for (int v = 0; v <= kMaxC; ++v)
    if (v <= C) {
      long long s = loc[v];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(&ssum[v], (unsigned long long)s);
    }
  __syncthreads();
  if (threadIdx.x <= C)
    atomicAdd((unsigned long long*)&nm->vsum[threadIdx.x], ssum[threadIdx.x]);
}

// One Newton round of LogisticRegression::FitLogisticModel (regression/LogisticRegression.cpp:291-303) at the current beta:
// p = 1/(1+exp(-X beta)), V = p(1-p) (both stored), and per block, in fixed order, the partial sums of
// D = X'VX (C*C), r = X'(y-p) (C) and the log-likelihood of GetDeviance (:75-94).  part: [blocks][kLogitAcc]
constexpr int kLogitAcc = kMaxC * kMaxC + kMaxC + 1;
__global__ void __launch_bounds__(kNullThreads)
k_logit_round(int64_t N, int C, const double* __restrict__ X, const double* __restrict__ y, const double* __restrict__ beta,
              double* __restrict__ p_out, double* __restrict__ v_out, double* __restrict__ part) {
  __shared__ double sh[kNullThreads];
  double b[kMaxC];
  for (int l = 0; l < C; ++l) b[l] = beta[l];
  double acc[kLogitAcc];
#pragma unroll
  for (int q = 0; q < kLogitAcc; ++q) acc[q] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    double x[kMaxC], eta = 0.0;
#pragma unroll
    for (int l = 0; l < kMaxC; ++l)
      if (l < C) {
        x[l] = X[(size_t)l * N + i];
        eta += x[l] * b[l];
      }
    const double p = 1.0 / (1.0 + exp(-eta));
    const double v = p * (1.0 - p), yi = y[i];
    p_out[i] = p;
    v_out[i] = v;
#pragma unroll
    for (int l = 0; l < kMaxC; ++l)
      if (l < C) {
#pragma unroll
        for (int m = 0; m < kMaxC; ++m)
          if (m < C) acc[l * kMaxC + m] += v * x[l] * x[m];
        acc[kMaxC * kMaxC + l] += x[l] * (yi - p);
      }
    acc[kLogitAcc - 1] += yi * log(p) + (1.0 - yi) * log(1.0 - p);
  }
  for (int q = 0; q < kLogitAcc; ++q) {
    const int l = q / kMaxC, m = q % kMaxC;
    const bool used = (q == kLogitAcc - 1) || (q >= kMaxC * kMaxC ? (q - kMaxC * kMaxC) < C : (l < C && m < C));
    if (!used) continue;   // uniform
    double sres = block_sum(acc[q], sh);
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * kLogitAcc + q] = sres;
  }
}
// r = y - p into the buffer the linear path keeps its phenotype in
__global__ void k_logit_resid(int64_t N, const double* __restrict__ y, const double* __restrict__ p, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] = y[i] - p[i];
}

}  // namespace rvt
