// dosage.cuh -- generic fp64 path for genes whose genotypes are not all hard calls: dosages in
// [0,2] or the mean-imputed value 2p of DataConsolidator::imputeGenotypeToMean
// (src/DataConsolidator.cpp:217-245, the default --impute strategy).  Such a gene cannot use the
// packed int8 sweep; it is rare at the boundary (the H2D copy of the 8-byte Matrix dominates
// anyway), so this path favours generality over speed: two passes over the N x M doubles on CUDA
// cores, fp64 throughout.
//   pass 1  k_dosage_cols   column sums / min / max  -> flip-to-minor (sum > N) and monomorphic
//                           (min == max) decisions, src/DataConsolidator.cpp:46-69, 94-142
//   pass 2  k_dosage_stats  G'G, G'r, G'X on the raw values; the burden collapses on the flipped
//                           values with the reference's `(int)g > 0` rule (src/Model.cpp:73-130)
//   k_dosage_prepare        flip algebra (g' = 2-g is affine), compaction, weights, Q, K
//                           -> TailInput, consumed by the tail of k_finalize (finalize.cuh)
#pragma once
#include "../../include/rvtests_b200.h"
#include "common.cuh"
#include "mathdev.cuh"

namespace rvt {

struct DosageStats {
  double csum[kTileRows];
  double cw[kTileRows];   // sum_i v_i g_ij (binary trait; equals csum in exact arithmetic when v == 1)
  unsigned long long cmin[kTileRows], cmax[kTileRows];   // bit patterns of non-negative doubles (order preserving)
  double A[kTileRows][kTileRows];
  double s[kTileRows];
  double B[kTileRows][kMaxC];
  double zegU, zegSS, zegSZ[kMaxC], cmcU, cmcSS, cmcSZ[kMaxC];
  double nonref;
  int negative;   // a value < 0 (missing) reached fit(): never happens under mean/HWE imputation
};

constexpr int kDosThreads = 256;
constexpr int kDosTile = 32;   // samples per shared-memory tile

// grid: (blocks over samples, 1); G: N x M column-major doubles
__global__ void __launch_bounds__(kDosThreads)
k_dosage_cols(const double* __restrict__ G, int64_t N, int M, DosageStats* __restrict__ st) {
  for (int j = 0; j < M; ++j) {
    double s = 0.0, mn = 1e300, mx = -1e300;
    int neg = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
      const double g = G[(size_t)j * N + i];
      s += g;
      mn = fmin(mn, g);
      mx = fmax(mx, g);
      neg |= (g < 0.0);
    }
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      neg |= __shfl_xor_sync(0xffffffffu, neg, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&st->csum[j], s);
      if (mn < 1e299) atomicMin(&st->cmin[j], (unsigned long long)__double_as_longlong(fmax(mn, 0.0)));
      if (mx > -1e299) atomicMax(&st->cmax[j], (unsigned long long)__double_as_longlong(fmax(mx, 0.0)));
      if (neg) atomicExch(&st->negative, 1);
    }
  }
}

// grid: blocks over sample tiles (grid-stride).  Thread t owns A entries (t/4 + 64*?, ...): the
// 64 x 64 block is dealt as 16 entries per thread: row = t >> 2, columns (t & 3) + 4*c, c = 0..15.
__global__ void __launch_bounds__(kDosThreads)
k_dosage_stats(const double* __restrict__ G, int64_t N, int M, const double* __restrict__ X, int C,
               const double* __restrict__ resid, const double* __restrict__ vw /* per-sample variance, null: 1 */,
               DosageStats* __restrict__ st) {
  __shared__ double sg[kTileRows][kDosTile + 1];
  __shared__ double sr[kDosTile], sx[kMaxC][kDosTile], sv[kDosTile];
  __shared__ int sflip[kTileRows], smono[kTileRows];
  const int tid = threadIdx.x;
  if (tid < kTileRows) {
    sflip[tid] = (tid < M) ? (st->csum[tid] > (double)N) : 0;     // `s <= m.rows` keeps, else flips
    smono[tid] = (tid < M) ? (st->cmin[tid] == st->cmax[tid]) : 1;
  }
  const int row = tid >> 2, cb = tid & 3;
  double acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.0;
  double accs = 0.0, accw = 0.0, accB[kMaxC];
#pragma unroll
  for (int l = 0; l < kMaxC; ++l) accB[l] = 0.0;
  double bz[3 + kMaxC], bc[3 + kMaxC];   // U, SS, nonref, SZ[] for zeggini / cmc (thread < kDosTile only)
#pragma unroll
  for (int l = 0; l < 3 + kMaxC; ++l) bz[l] = bc[l] = 0.0;
  const int64_t ntiles = (N + kDosTile - 1) / kDosTile;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t i0 = t * kDosTile;
    __syncthreads();
    for (int idx = tid; idx < kTileRows * kDosTile; idx += kDosThreads) {
      const int j = idx / kDosTile, k = idx - j * kDosTile;
      const int64_t i = i0 + k;
      sg[j][k] = (j < M && i < N) ? G[(size_t)j * N + i] : 0.0;
    }
    if (tid < kDosTile) {
      const int64_t i = i0 + tid;
      sr[tid] = (i < N) ? resid[i] : 0.0;
      sv[tid] = (i < N) ? (vw ? vw[i] : 1.0) : 0.0;
      for (int l = 0; l < C; ++l) sx[l][tid] = (i < N) ? X[(size_t)l * N + i] : 0.0;
    }
    __syncthreads();
    if (row < M) {
      for (int k = 0; k < kDosTile; ++k) {
        const double g = sg[row][k];
        if (g == 0.0) continue;
        const double gv = g * sv[k];   // G'VG, G'VX (Skat.cpp:55-76 with V = diag(v)); G'r stays unweighted
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] += gv * sg[cb + 4 * c][k];
        if (cb == 0) {
          accs += g * sr[k];
          accw += gv;
          for (int l = 0; l < C; ++l) accB[l] += gv * sx[l][k];
        }
      }
    }
    if (tid < kDosTile && i0 + tid < N) {
      // cmcCollapse / zegginiCollapse on the flipped, polymorphic columns: `(int)g > 0`
      double z = 0.0;
      for (int j = 0; j < M; ++j) {
        if (smono[j]) continue;
        const double g = sflip[j] ? 2.0 - sg[j][tid] : sg[j][tid];
        if ((int)g > 0) z += 1.0;
      }
      const double c = (z > 0.0) ? 1.0 : 0.0;
      const double r = sr[tid], v = sv[tid];
      // U = S'r, SS = S'VS, SZ = S'VZ (LinearRegressionScoreTest.cpp:209-217 with v == 1; LogisticRegressionScoreTest.cpp:266-269)
      bz[0] += z * r;  bz[1] += v * z * z;  bz[2] += (z != 0.0);
      bc[0] += c * r;  bc[1] += v * c * c;  bc[2] += (c != 0.0);
      for (int l = 0; l < C; ++l) {
        bz[3 + l] += v * z * sx[l][tid];
        bc[3 + l] += v * c * sx[l][tid];
      }
    }
  }
  if (row < M) {
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (cb + 4 * c < M) atomicAdd(&st->A[row][cb + 4 * c], acc[c]);
    if (cb == 0) {
      atomicAdd(&st->s[row], accs);
      atomicAdd(&st->cw[row], accw);
      for (int l = 0; l < C; ++l) atomicAdd(&st->B[row][l], accB[l]);
    }
  }
  if (tid < 32) {   // kDosTile == 32: one warp holds the burden partials
#pragma unroll
    for (int l = 0; l < 3 + kMaxC; ++l)
      for (int o = 16; o > 0; o >>= 1) {
        bz[l] += __shfl_xor_sync(0xffffffffu, bz[l], o);
        bc[l] += __shfl_xor_sync(0xffffffffu, bc[l], o);
      }
    if (tid == 0) {
      atomicAdd(&st->zegU, bz[0]);  atomicAdd(&st->zegSS, bz[1]);
      atomicAdd(&st->cmcU, bc[0]);  atomicAdd(&st->cmcSS, bc[1]);
      atomicAdd(&st->nonref, bc[2]);
      for (int l = 0; l < C; ++l) {
        atomicAdd(&st->zegSZ[l], bz[3 + l]);
        atomicAdd(&st->cmcSZ[l], bc[3 + l]);
      }
    }
  }
}

// one CTA (64 threads) per dosage gene
__global__ void __launch_bounds__(64)
k_dosage_prepare(const DosageStats* __restrict__ st, int M, const double* __restrict__ af /*[M] or null*/,
                 const NullModel* __restrict__ nm, EngineParams prm, TailInput* __restrict__ out) {
  __shared__ int s_idx[kTileRows], s_flip[kTileRows];
  __shared__ double s_s[kTileRows], s_sw[kTileRows], s_B[kTileRows][kMaxC];
  __shared__ int s_Mp;
  const int tid = threadIdx.x;
  const int64_t N = nm->N;
  const int C = nm->C;
  const double sigma2 = nm->sigma2;
  if (tid == 0) {
    int mp = 0;
    for (int j = 0; j < M; ++j) {
      s_flip[j] = st->csum[j] > (double)N;
      if (st->cmin[j] != st->cmax[j]) s_idx[mp++] = j;
    }
    s_Mp = mp;
  }
  __syncthreads();
  const int Mp = s_Mp;
  if (tid < Mp) {
    const int j = s_idx[tid], fl = s_flip[j];
    s_s[tid] = fl ? 2.0 * nm->rsum - st->s[j] : st->s[j];
    for (int l = 0; l < C; ++l) s_B[tid][l] = fl ? 2.0 * (nm->binary ? nm->xsum_w[l] : nm->xsum[l]) - st->B[j][l] : st->B[j][l];
    // weight i of the kept columns uses af[i] in the caller's ORIGINAL order
    const double freq = af ? af[tid] : 0.5 * st->csum[j] / (double)N;
    s_sw[tid] = sqrt(beta_weight(freq, prm.beta1, prm.beta2, true));
  }
  __syncthreads();
  for (int idx = tid; idx < Mp * Mp; idx += 64) {
    const int i = idx / Mp, k = idx - i * Mp;
    const int ji = s_idx[i], jk = s_idx[k];
    const int fi = s_flip[ji], fk = s_flip[jk];
    double a = st->A[ji][jk];
    // g' = 2 - g under the weights v: (2-g_j)'V(2-g_k) = 4 sum v - 2 c^w_j - 2 c^w_k + A_jk, with c^w = G'v
    const double ci = nm->binary ? st->cw[ji] : st->csum[ji], ck = nm->binary ? st->cw[jk] : st->csum[jk];
    if (fi && fk)
      a = 4.0 * (nm->binary ? nm->vsum_w : (double)N) - 2.0 * ci - 2.0 * ck + a;
    else if (fi)
      a = 2.0 * ck - a;
    else if (fk)
      a = 2.0 * ci - a;
    double t = 0.0;
    for (int l = 0; l < C; ++l) {
      double u = 0.0;
      for (int m = 0; m < C; ++m) u += nm->xtx_inv[l * C + m] * s_B[k][m];
      t += s_B[i][l] * u;
    }
    out->K[idx] = s_sw[i] * s_sw[k] * sigma2 * (a - t);
  }
  if (tid < Mp) out->vw[tid] = s_sw[tid] * s_s[tid];
  if (tid == 0) {
    double q = 0.0;
    for (int i = 0; i < Mp; ++i) q += (s_sw[i] * s_sw[i]) * s_s[i] * s_s[i];
    out->Q = q;
    out->Mp = Mp;
    out->status = st->negative ? RVT_GENE_BADVALUE : ((Mp == 0) ? RVT_GENE_NA : RVT_GENE_OK);
    out->nonref = (int)llrint(st->nonref);
    out->zegU = st->zegU;  out->zegSS = st->zegSS;
    out->cmcU = st->cmcU;  out->cmcSS = st->cmcSS;
    for (int l = 0; l < C; ++l) {
      out->zegSZ[l] = st->zegSZ[l];
      out->cmcSZ[l] = st->cmcSZ[l];
    }
  }
}

}  // namespace rvt
