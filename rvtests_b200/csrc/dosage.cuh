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
__device__ inline void dosage_prepare(const DosageStats* __restrict__ st, int M, const double* __restrict__ af /*[M] or null*/,
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

__global__ void __launch_bounds__(64)
k_dosage_prepare(const DosageStats* __restrict__ st, int M, const double* __restrict__ af /*[M] or null*/,
                 const NullModel* __restrict__ nm, EngineParams prm, TailInput* __restrict__ out) {
  dosage_prepare(st, M, af, nm, prm, out);
}

// ---- genes whose values come from the engine's own hard-call tiles: missing calls (code 3) mean-imputed on the fly
// (DataConsolidator::imputeGenotypeToMean, src/DataConsolidator.cpp:217-245), and every gene of a binary-trait run.
// No N x M doubles are ever materialised: the statistics kernel reads the int8 tiles (1 byte per call) and many genes share
// one launch (blockIdx.y = gene).
struct TileGene {
  const int8_t* g;    // tiled block [chunk][M][128]
  int32_t M;
  int32_t has_af;
  int64_t var0;       // per-variant side arrays (counts, af)
  int32_t slot;       // index into the DosageStats / TailInput arrays of this launch
  int32_t allow_missing;   // 1: code 3 is a missing call to impute (2-bit pushes); 0: any value outside {0,1,2} is an error
};

// column sums / min / max of the imputed columns follow from the counts alone: no pass over the data
__global__ void k_tile_cols(const TileGene* __restrict__ genes, int n_genes, int64_t N, const RowCounts* __restrict__ counts,
                            DosageStats* __restrict__ st) {
  const int gi = blockIdx.x, j = threadIdx.x;
  if (gi >= n_genes) return;
  const TileGene tg = genes[gi];
  if (j >= tg.M) return;
  const RowCounts rc = counts[tg.var0 + j];
  const long long n1 = rc.n1, n2 = rc.n2, miss = rc.bad, n0 = N - n1 - n2 - miss, nobs = N - miss;
  const double ac = (double)(n1 + 2 * n2);
  const double fill = nobs > 0 ? 2.0 * (ac / (double)(2 * nobs)) : 0.0;     // 2 p^, p^ over the observed calls
  double mn = 1e300, mx = -1e300;
  if (n0 > 0) { mn = fmin(mn, 0.0); mx = fmax(mx, 0.0); }
  if (n1 > 0) { mn = fmin(mn, 1.0); mx = fmax(mx, 1.0); }
  if (n2 > 0) { mn = fmin(mn, 2.0); mx = fmax(mx, 2.0); }
  if (miss > 0) { mn = fmin(mn, fill); mx = fmax(mx, fill); }
  DosageStats* s = st + tg.slot;
  if (miss > 0 && !tg.allow_missing) s->negative = 1;   // reported as RVT_GENE_BADVALUE, never silently computed
  s->csum[j] = ac + (double)miss * fill;
  s->cmin[j] = (unsigned long long)__double_as_longlong(mn);
  s->cmax[j] = (unsigned long long)__double_as_longlong(mx);
}

// The statistics of k_dosage_stats on tiles, exploiting that rare-variant genotypes are almost all zero.  A thread owns 4
// consecutive samples.  Pass 1 walks the gene's rows with one 32-bit load per row (4 calls) and only RECORDS the non-zero
// calls in a short per-sample list; pass 2 does the arithmetic from the lists: products with r, v, X and with the other
// non-zero calls of the same sample.  (A row-synchronous loop that accumulated directly made all lanes of a warp hit the
// same shared-memory address at the same time: 435 us per gene; the lists decouple the lanes.)  Sums live in shared memory
// (fp64 atomics on scattered addresses) and are flushed once per CTA.  Zero calls matter only to the burden scores of FLIPPED
// rows (g' = 2 - g > 0): a per-gene constant F plus corrections at the non-zero calls.  A sample with more non-zero calls
// than the list holds (common variants) is handled by re-reading its bytes.  grid (blocks, genes), 256 threads.
constexpr int kSparseThreads = 256;
constexpr int kSparseList = 12;
__global__ void __launch_bounds__(kSparseThreads)
k_tile_sparse(const TileGene* __restrict__ genes, int64_t N, const RowCounts* __restrict__ counts, const double* __restrict__ X, int C,
              const double* __restrict__ resid, const double* __restrict__ vw, DosageStats* __restrict__ stats) {
  __shared__ double sA[kTileRows][kTileRows + 1];
  __shared__ double sS[kTileRows], sW[kTileRows], sB[kTileRows][kMaxC];
  __shared__ double sfill[kTileRows];
  __shared__ int8_t srole[kTileRows];   // 0 normal, 1 flipped, 2 monomorphic (ignored by the burden scores)
  __shared__ int sF;
  __shared__ double sbur[2][3 + kMaxC];
  const TileGene tg = genes[blockIdx.y];
  DosageStats* __restrict__ st = stats + tg.slot;
  const int M = tg.M, tid = threadIdx.x;
  for (int idx = tid; idx < kTileRows * (kTileRows + 1); idx += kSparseThreads) (&sA[0][0])[idx] = 0.0;
  for (int idx = tid; idx < kTileRows * kMaxC; idx += kSparseThreads) (&sB[0][0])[idx] = 0.0;
  if (tid < 2 * (3 + kMaxC)) (&sbur[0][0])[tid] = 0.0;
  if (tid < kTileRows) {
    sS[tid] = sW[tid] = 0.0;
    double f = 0.0;
    int role = 2;
    if (tid < M) {
      const RowCounts rc = counts[tg.var0 + tid];
      const long long nobs = N - rc.bad;
      f = nobs > 0 ? 2.0 * ((double)((long long)rc.n1 + 2ll * rc.n2) / (double)(2 * nobs)) : 0.0;
      role = (st->cmin[tid] == st->cmax[tid]) ? 2 : ((st->csum[tid] > (double)N) ? 1 : 0);
    }
    sfill[tid] = f;
    srole[tid] = (int8_t)role;
  }
  __syncthreads();
  if (tid == 0) {
    int F = 0;
    for (int j = 0; j < M; ++j) F += (srole[j] == 1);
    sF = F;
  }
  __syncthreads();
  const int F = sF;
  double bz[3 + kMaxC], bc[3 + kMaxC];   // per-thread burden partials: U, SS, nonref, SZ[]
#pragma unroll
  for (int l = 0; l < 3 + kMaxC; ++l) bz[l] = bc[l] = 0.0;
  const int64_t nquads = (N + 3) >> 2;   // 4-sample groups; a chunk of 128 samples holds 32 of them
  for (int64_t qd = (int64_t)blockIdx.x * kSparseThreads + tid; qd < nquads; qd += (int64_t)gridDim.x * kSparseThreads) {
    const int64_t i0 = qd << 2;
    const int8_t* __restrict__ base = tg.g + ((size_t)(i0 >> 7) * M) * 128 + (i0 & 127);   // row j at + j * 128
    uint8_t lj[4][kSparseList], lc[4][kSparseList];
    int cnt[4] = {0, 0, 0, 0};
    // pass 1: record.  The rows' words are fetched sixteen at a time BEFORE anything branches on them: with one load per
    // iteration behind `if (w == 0) continue` the loop ran at one HBM latency per row (ncu: 51 % long-scoreboard stalls;
    // 83 -> 63 us per 25 MB gene).  Tried and dropped (profiles/r02u_binary_path.txt): the per-variant sums in a separate
    // variant-major kernel without shared-memory atomics -- its gathers of r, v, X are as latency-bound as the atomics were.
    for (int j0 = 0; j0 < M; j0 += 16) {
      uint32_t wv[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) wv[k] = (j0 + k < M) ? __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)(j0 + k) * 128)) : 0u;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint32_t w = wv[k];
        if (w == 0) continue;
        const int j = j0 + k;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t code = (w >> (8 * q)) & 0xFFu;
          if (code == 0) continue;
          if (cnt[q] < kSparseList) {
            lj[q][cnt[q]] = (uint8_t)j;
            lc[q][cnt[q]] = (uint8_t)code;
          }
          ++cnt[q];
        }
      }
    }
    // pass 2: arithmetic, sample by sample
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t i = i0 + q;
      if (i >= N) continue;
      const int n = cnt[q];
      if (n == 0 && F == 0) continue;
      const double r = resid[i], v = vw ? vw[i] : 1.0;
      double x[kMaxC];
#pragma unroll
      for (int l = 0; l < kMaxC; ++l)
        if (l < C) x[l] = X[(size_t)l * N + i];
      int zc = 0;
      const bool listed = n <= kSparseList;
      // iterate the non-zero calls of this sample: from the list, or (overflow) by re-reading the sample's bytes
      int a = 0, ja = -1;
      for (;;) {
        int code;
        if (listed) {
          if (a >= n) break;
          ja = lj[q][a];
          code = lc[q][a];
        } else {
          do { ++ja; } while (ja < M && base[(size_t)ja * 128 + q] == 0);
          if (ja >= M) break;
          code = base[(size_t)ja * 128 + q];
        }
        const double g = (code == 3) ? sfill[ja] : (double)code;
        if (g != 0.0) {
          const double gv = g * v;
          atomicAdd(&sS[ja], g * r);
          atomicAdd(&sW[ja], gv);
#pragma unroll
          for (int l = 0; l < kMaxC; ++l)
            if (l < C) atomicAdd(&sB[ja][l], gv * x[l]);
          atomicAdd(&sA[ja][ja], gv * g);
          if (listed) {
            for (int b = a + 1; b < n; ++b) {
              const int jb = lj[q][b], cb = lc[q][b];
              const double g2 = (cb == 3) ? sfill[jb] : (double)cb;
              if (g2 != 0.0) atomicAdd(&sA[ja][jb], gv * g2);
            }
          } else {
            for (int jb = ja + 1; jb < M; ++jb) {
              const int cb = base[(size_t)jb * 128 + q];
              if (cb == 0) continue;
              const double g2 = (cb == 3) ? sfill[jb] : (double)cb;
              if (g2 != 0.0) atomicAdd(&sA[ja][jb], gv * g2);
            }
          }
          // burden indicator `(int)g' > 0` (src/Model.cpp:82-83) relative to a zero call of the same row
          const int role = srole[ja];
          if (role == 0) zc += ((int)g > 0);
          else if (role == 1) zc -= !((int)(2.0 - g) > 0);
        }
        ++a;
      }
      const double z = (double)(F + zc);
      if (z == 0.0) continue;
      bz[0] += z * r;  bz[1] += v * z * z;  bz[2] += 1.0;
      bc[0] += r;      bc[1] += v;          bc[2] += 1.0;
#pragma unroll
      for (int l = 0; l < kMaxC; ++l)
        if (l < C) {
          bz[3 + l] += v * z * x[l];
          bc[3 + l] += v * x[l];
        }
    }
  }
#pragma unroll
  for (int l = 0; l < 3 + kMaxC; ++l) {
    if (l >= 3 + C) break;
    double a = bz[l], b = bc[l];
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((tid & 31) == 0) {
      atomicAdd(&sbur[0][l], a);
      atomicAdd(&sbur[1][l], b);
    }
  }
  __syncthreads();
  for (int idx = tid; idx < M * M; idx += kSparseThreads) {
    const int a = idx / M, b = idx - a * M;
    const double val = (a <= b) ? sA[a][b] : sA[b][a];   // only the upper triangle was accumulated
    if (val != 0.0) atomicAdd(&st->A[a][b], val);
  }
  if (tid < M) {
    if (sS[tid] != 0.0) atomicAdd(&st->s[tid], sS[tid]);
    if (sW[tid] != 0.0) atomicAdd(&st->cw[tid], sW[tid]);
    for (int l = 0; l < C; ++l)
      if (sB[tid][l] != 0.0) atomicAdd(&st->B[tid][l], sB[tid][l]);
  }
  if (tid == 0) {
    atomicAdd(&st->zegU, sbur[0][0]);  atomicAdd(&st->zegSS, sbur[0][1]);
    atomicAdd(&st->cmcU, sbur[1][0]);  atomicAdd(&st->cmcSS, sbur[1][1]);
    atomicAdd(&st->nonref, sbur[1][2]);
    for (int l = 0; l < C; ++l) {
      atomicAdd(&st->zegSZ[l], sbur[0][3 + l]);
      atomicAdd(&st->cmcSZ[l], sbur[1][3 + l]);
    }
  }
}

// k_dosage_prepare for a batch of tile genes (blockIdx.x = gene)
__global__ void __launch_bounds__(64)
k_tile_prepare(const TileGene* __restrict__ genes, int n_genes, const DosageStats* __restrict__ st, const double* __restrict__ af,
               const NullModel* __restrict__ nm, EngineParams prm, TailInput* __restrict__ out) {
  if ((int)blockIdx.x >= n_genes) return;
  const TileGene tg = genes[blockIdx.x];
  dosage_prepare(st + tg.slot, tg.M, tg.has_af ? af + tg.var0 : nullptr, nm, prm, out + tg.slot);
}

}  // namespace rvt
