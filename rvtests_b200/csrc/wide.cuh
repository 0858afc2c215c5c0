// wide.cuh -- genes wider than one tensor-core tile (M > 64 variants, up to kWideMaxM).
//
// The reference has no width limit: Skat::Fit (regression/Skat.cpp:29-105), SkatO::Fit
// (regression/SkatO.cpp:101-281), cmcCollapse / zegginiCollapse (src/Model.cpp:73-130) take whatever N x M
// matrix the gene file selects, and MixtureChiSquare grows its lambda array on demand
// (regression/MixtureChiSquare.h:26-52).  A wide gene is cut into T = ceil(M/64) tiles of consecutive variants, each
// staged as its own tiled block; its M x M Gram comes from the T diagonal sweeps plus the T(T-1)/2 tile pairs of the
// PAIR mode of the tensor-core sweep (the same units `--meta cov` uses), still exact integers.  The burden collapses
// are per-SAMPLE functions of all M variants (an OR and a count), so they take one extra pass over the tiles
// (k_wide_collapse).  The O(M^3) tail is the same device code as k_finalize (eigen.cuh, davies.cuh, skato_tail.cuh)
// run by one CTA per gene on a workspace in global memory instead of shared memory.
#pragma once
#include "../../include/rvtests_b200.h"
#include "common.cuh"
#include "finalize.cuh"

namespace rvt {

constexpr int kWideMaxM = 2048;     // workspace = 2 M^2 x 8 B = 64 MB per gene at the cap
constexpr int kWideThreads = 256;
constexpr int kWideGatherThreads = 128;
constexpr int kWideCollapseThreads = 128;   // one thread per sample of a 128-sample chunk

// per-gene workspace, carved from one allocation by wide_job_make()
struct WideJob {
  int M;               // variants of the gene (all tiles)
  int out_index;       // slot of the gene in the result array
  int64_t var0;        // first slot of the gene in the per-variant side arrays (flags, af, counts)
  int has_af, counted;
  int imp, pad;        // imp = 1: a gene with MISSING calls (mean-imputed, G = H + Mi diag(delta)): A_raw is the (2M x 2M) Gram of the rows
                       // [H ; Mi] (H = hard calls with the fill 0 / 2 at missing entries, Mi = 0/1 indicators), De their (2M x ER)
                       // digit sums; the workspace is the one of a 2M-variant gene
  long long* A_raw;    // [M][M] raw G'G (both triangles)
  long long* De;       // [M][ER] gene x digit sums
  long long* coll;     // [kCollapseN] burden sums, SweepPartial::coll layout
  long long* craw;     // [M] column sums
  double* K;           // [M][M]
  double* Wm;          // [M][M], SKAT-O only: aliases A_raw (dead once K is built)
  double* dws;         // doubles: see wide_dws_*
  int* iws;            // ints: idx[M], flip[M], th[8][M]
};
static inline size_t wide_dws_count(int M) { return (size_t)M * (3 + 2 * kMaxC) + 7 * (size_t)(M + 2); }
static inline size_t wide_iws_count(int M) { return (size_t)M * (2 + kWideThreads / 32); }
static inline size_t wide_ws_bytes(int M) {
  size_t b = 0;
  b += (size_t)M * M * 8 * 2;               // A_raw, K
  b += (size_t)M * kMaxER * 8;              // De
  b += (size_t)kCollapseN * 8 + (size_t)M * 8;   // coll, craw
  b += wide_dws_count(M) * 8 + wide_iws_count(M) * 4;
  return (b + 255) & ~(size_t)255;
}
static inline WideJob wide_job_make(uint8_t* ws, int M) {
  WideJob j;
  memset(&j, 0, sizeof(j));
  j.M = M;
  uint8_t* p = ws;
  j.A_raw = (long long*)p;  p += (size_t)M * M * 8;
  j.K = (double*)p;         p += (size_t)M * M * 8;
  j.De = (long long*)p;     p += (size_t)M * kMaxER * 8;
  j.coll = (long long*)p;   p += (size_t)kCollapseN * 8;
  j.craw = (long long*)p;   p += (size_t)M * 8;
  j.dws = (double*)p;       p += wide_dws_count(M) * 8;
  j.iws = (int*)p;
  j.Wm = (double*)j.A_raw;
  return j;
}

// One CTA per sweep unit (a diagonal tile or a tile pair) of ONE wide gene: sum the splits into the gene's M x M
// integer Gram; diagonal tiles also deliver the gene x digit columns.
__global__ void __launch_bounds__(kWideGatherThreads)
k_wide_gather(const GeneDesc* __restrict__ units, int n_units, int64_t var_base, int M, int ER, int S,
              const SweepPartial* __restrict__ parts, long long* __restrict__ A_raw, long long* __restrict__ De, int pair) {
  const int u = blockIdx.x, tid = threadIdx.x;
  if (u >= n_units) return;
  const GeneDesc gd = units[u];
  const int ri = (int)(gd.var0 - var_base), rj = (int)(gd.var0_b - var_base), Ma = gd.M, Mb = gd.Mb;
  const SweepPartial* __restrict__ gp = parts + (size_t)u * S;
  for (int idx = tid; idx < Ma * Mb; idx += kWideGatherThreads) {
    const int i = idx / Mb, j = idx - i * Mb;
    long long a = 0;
    for (int sp = 0; sp < S; ++sp) a += gp[sp].d[i][j];
    A_raw[(size_t)(ri + i) * M + (rj + j)] = a;
    if (pair) A_raw[(size_t)(rj + j) * M + (ri + i)] = a;
  }
  if (!pair)
    for (int idx = tid; idx < Ma * ER; idx += kWideGatherThreads) {
      const int i = idx / ER, e = idx - i * ER;
      long long s = 0;
      for (int sp = 0; sp < S; ++sp) s += gp[sp].d[i][kTileRows + e];
      De[(size_t)(ri + i) * ER + e] = s;
    }
}

// cmcCollapse / zegginiCollapse over ALL tiles of a wide gene (src/Model.cpp:73-89, 115-130) and their dot products
// with the null-model digit rows: thread = one sample of a 128-sample chunk, grid-stride over chunks.  Exact integers,
// so the atomics make the result independent of the schedule.  coll[]: SweepPartial::coll layout.
__global__ void __launch_bounds__(kWideCollapseThreads)
k_wide_collapse(const GeneDesc* __restrict__ tiles, int T, int64_t var_base, int M, const uint8_t* __restrict__ rowflags,
                const NullModel* __restrict__ nm, long long* __restrict__ coll) {
  __shared__ uint8_t s_flag[kWideMaxM];
  __shared__ unsigned long long s_red[kCollapseN];
  const int tid = threadIdx.x;
  const int64_t N = nm->N, ldE = nm->ldE;
  const int ER = nm->ER;
  const int8_t* __restrict__ E = nm->E;
  for (int j = tid; j < M; j += kWideCollapseThreads) s_flag[j] = rowflags[var_base + j];
  if (tid < kCollapseN) s_red[tid] = 0ull;
  __syncthreads();
  long long az[kMaxER + 1], ac[kMaxER + 1];
#pragma unroll
  for (int e = 0; e <= kMaxER; ++e) az[e] = ac[e] = 0;
  const int64_t nchunks = (N + 127) >> 7;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int64_t i = (c << 7) + tid;
    if (i >= N) continue;   // the padding of the last chunk is zero and would count under a flipped row
    int z = 0;
    for (int t = 0; t < T; ++t) {
      const int Mt = tiles[t].M, r0 = (int)(tiles[t].var0 - var_base);
      const int8_t* __restrict__ p = tiles[t].g + ((size_t)c * Mt) * 128 + tid;
      for (int r = 0; r < Mt; ++r) {
        const int g = p[(size_t)r * 128];
        const uint8_t f = s_flag[r0 + r];
        z += (f == kRowNormal) ? (g > 0) : (f == kRowFlipped) ? (g < 2) : 0;
      }
    }
    const int cm = z > 0;
#pragma unroll
    for (int e = 0; e < kMaxER; ++e)
      if (e < ER) {
        const int d = E[(size_t)e * ldE + i];
        az[e] += (long long)z * d;
        ac[e] += cm * d;
      }
    az[kMaxER] += (long long)z * z;
    ac[kMaxER] += cm;
  }
#pragma unroll
  for (int e = 0; e <= kMaxER; ++e) {
    if (e < ER || e == kMaxER) {
      long long a = az[e], b = ac[e];
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
      }
      if ((tid & 31) == 0) {
        const int slot = (e == kMaxER) ? ER : e;
        atomicAdd(&s_red[slot], (unsigned long long)a);
        atomicAdd(&s_red[(ER + 1) + slot], (unsigned long long)b);
      }
    }
  }
  __syncthreads();
  if (tid < 2 * (ER + 1)) atomicAdd(reinterpret_cast<unsigned long long*>(coll) + tid, s_red[tid]);
}

// ---- binary trait (or any weighted Gram): the fp64 statistics of a wide gene from its int8 tiles ---------------------------
// The p(1-p)-weighted Gram G'VG is not an integer product, so a wide gene of a binary-trait run cannot use the pair sweep as
// it is; it gets what k_tile_sparse does for a single tile, generalised to T tiles and to accumulators in global memory:
// a thread owns one sample, records the non-zero calls of that sample over ALL M variants (rare variants: a handful), and
// adds their products with r, v, v x_l and with each other -- A (upper triangle), S, CW, B -- by fp64 atomics in global memory
// (native on sm_100a); a sample with more calls than the list holds re-reads its bytes.  Missing calls (code 3) are imputed on
// the fly to 2 p^ of the observed calls (DataConsolidator.cpp:217-245).  The burden scores follow src/Model.cpp:73-130 on the
// minor-coded matrix as in k_tile_sparse.  WideJob fields re-used as doubles in this mode (imp == 2): A_raw -> A[M][M],
// De -> per variant {S, CW, B[0..C)} at stride kMaxER, craw -> imputed column sums, coll -> burden sums [2][3 + kMaxC].
constexpr int kWideSparseList = 48;
__global__ void __launch_bounds__(kWideCollapseThreads)
k_wide_sparse(const GeneDesc* __restrict__ tiles, int T, int64_t var_base, int M, const uint8_t* __restrict__ rowflags,
              const RowCounts* __restrict__ counts, const NullModel* __restrict__ nm, const double* __restrict__ X,
              const double* __restrict__ vw, double* __restrict__ A, double* __restrict__ SB, double* __restrict__ bur) {
  extern __shared__ __align__(16) unsigned char ws_smem[];
  double* s_fill = reinterpret_cast<double*>(ws_smem);                   // [M]
  uint8_t* s_role = reinterpret_cast<uint8_t*>(s_fill + M);              // [M] 0 normal, 1 flipped, 2 monomorphic
  uint8_t* s_tile = s_role + M;                                          // [M] tile of variant j
  uint8_t* s_row = s_tile + M;                                           // [M] its row inside the tile
  __shared__ int s_F;
  __shared__ double s_bur[2][3 + kMaxC];
  const int tid = threadIdx.x;
  const int64_t N = nm->N;
  const int C = nm->C;
  const double* __restrict__ resid = nm->resid;
  for (int t = 0; t < T; ++t) {
    const int r0 = (int)(tiles[t].var0 - var_base), Mt = tiles[t].M;
    for (int r = tid; r < Mt; r += kWideCollapseThreads) {
      const int j = r0 + r;
      s_tile[j] = (uint8_t)t;
      s_row[j] = (uint8_t)r;
      const RowCounts rc = counts[var_base + j];
      const long long nobs = N - rc.bad;
      s_fill[j] = nobs > 0 ? 2.0 * ((double)((long long)rc.n1 + 2ll * rc.n2) / (double)(2 * nobs)) : 0.0;
      const uint8_t f = rowflags[var_base + j];
      s_role[j] = (f == kRowNormal) ? 0 : (f == kRowFlipped) ? 1 : 2;
    }
  }
  if (tid < 2 * (3 + kMaxC)) (&s_bur[0][0])[tid] = 0.0;
  __syncthreads();
  if (tid == 0) {
    int F = 0;
    for (int j = 0; j < M; ++j) F += (s_role[j] == 1);
    s_F = F;
  }
  __syncthreads();
  const int F = s_F;
  double bz[3 + kMaxC], bc[3 + kMaxC];
#pragma unroll
  for (int l = 0; l < 3 + kMaxC; ++l) bz[l] = bc[l] = 0.0;
  const int64_t nchunks = (N + 127) >> 7;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int64_t i = (c << 7) + tid;
    if (i >= N) continue;
    uint16_t lj[kWideSparseList];
    uint8_t lc[kWideSparseList];
    int n = 0;
    for (int t = 0; t < T; ++t) {
      const int Mt = tiles[t].M, r0 = (int)(tiles[t].var0 - var_base);
      const int8_t* __restrict__ p = tiles[t].g + ((size_t)c * Mt) * 128 + tid;
      for (int r = 0; r < Mt; ++r) {
        const int code = p[(size_t)r * 128];
        if (code == 0) continue;
        if (n < kWideSparseList) {
          lj[n] = (uint16_t)(r0 + r);
          lc[n] = (uint8_t)code;
        }
        ++n;
      }
    }
    if (n == 0 && F == 0) continue;
    const double r = resid[i], v = vw ? vw[i] : 1.0;
    double x[kMaxC];
#pragma unroll
    for (int l = 0; l < kMaxC; ++l)
      if (l < C) x[l] = X[(size_t)l * N + i];
    const bool listed = n <= kWideSparseList;
    auto code_at = [&](int j) -> int { const int t = s_tile[j]; return tiles[t].g[((size_t)c * tiles[t].M + s_row[j]) * 128 + tid]; };
    int zc = 0, a = 0, ja = -1;
    for (;;) {
      int code;
      if (listed) {
        if (a >= n) break;
        ja = lj[a];
        code = lc[a];
      } else {
        do { ++ja; } while (ja < M && code_at(ja) == 0);
        if (ja >= M) break;
        code = code_at(ja);
      }
      const double g = (code == 3) ? s_fill[ja] : (double)code;
      if (g != 0.0) {
        const double gv = g * v;
        double* sb = SB + (size_t)ja * kMaxER;
        atomicAdd(&sb[0], g * r);
        atomicAdd(&sb[1], gv);
#pragma unroll
        for (int l = 0; l < kMaxC; ++l)
          if (l < C) atomicAdd(&sb[2 + l], gv * x[l]);
        atomicAdd(&A[(size_t)ja * M + ja], gv * g);
        if (listed) {
          for (int b = a + 1; b < n; ++b) {
            const int jb = lj[b], cb = lc[b];
            const double g2 = (cb == 3) ? s_fill[jb] : (double)cb;
            if (g2 != 0.0) atomicAdd(&A[(size_t)ja * M + jb], gv * g2);
          }
        } else {
          for (int jb = ja + 1; jb < M; ++jb) {
            const int cb = code_at(jb);
            if (cb == 0) continue;
            const double g2 = (cb == 3) ? s_fill[jb] : (double)cb;
            if (g2 != 0.0) atomicAdd(&A[(size_t)ja * M + jb], gv * g2);
          }
        }
        const int role = s_role[ja];
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
#pragma unroll
  for (int l = 0; l < 3 + kMaxC; ++l) {
    if (l >= 3 + C) break;
    double a = bz[l], b = bc[l];
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((tid & 31) == 0) {
      atomicAdd(&s_bur[0][l], a);
      atomicAdd(&s_bur[1][l], b);
    }
  }
  __syncthreads();
  if (tid < 2 * (3 + kMaxC)) {
    const double val = (&s_bur[0][0])[tid];
    if (val != 0.0) atomicAdd(&bur[tid], val);
  }
}

// ---- a wide gene whose genotypes are REAL dosages (N x M column-major doubles, e.g. from BGEN): dense fp64 statistics ------
// k_wide_dos_cols: one CTA per column -- sum, min, max -> flag (all values equal drops, column sum above N flips; a negative
// value poisons the sum: RVT_GENE_BADVALUE), and the dot products S = g'r, CW = g'v, B = g'(v x_l).
constexpr int kWideDosThreads = 256;
__global__ void __launch_bounds__(kWideDosThreads)
k_wide_dos_cols(const double* __restrict__ G, int64_t N, int M, int64_t var0, const NullModel* __restrict__ nm, const double* __restrict__ X,
                const double* __restrict__ vw, double* __restrict__ SB, double* __restrict__ csum, uint8_t* __restrict__ rowflags) {
  __shared__ double s_red[kWideDosThreads];
  const int j = blockIdx.x, tid = threadIdx.x;
  const int C = nm->C;
  const double* __restrict__ g = G + (size_t)j * N;
  const double* __restrict__ resid = nm->resid;
  double acc[5 + kMaxC];   // sum, min, max, S, CW, B[]
  acc[0] = 0.0; acc[1] = 1e300; acc[2] = -1e300; acc[3] = acc[4] = 0.0;
#pragma unroll
  for (int l = 0; l < kMaxC; ++l) acc[5 + l] = 0.0;
  for (int64_t i = tid; i < N; i += kWideDosThreads) {
    const double x = g[i], v = vw ? vw[i] : 1.0, xv = x * v;
    acc[0] += x;
    acc[1] = fmin(acc[1], x);
    acc[2] = fmax(acc[2], x);
    acc[3] += x * resid[i];
    acc[4] += xv;
#pragma unroll
    for (int l = 0; l < kMaxC; ++l)
      if (l < C) acc[5 + l] += xv * X[(size_t)l * N + i];
  }
  double out[5 + kMaxC];
  for (int q = 0; q < 5 + C; ++q) {
    s_red[tid] = acc[q];
    __syncthreads();
    for (int o = kWideDosThreads / 2; o > 0; o >>= 1) {
      if (tid < o) s_red[tid] = (q == 1) ? fmin(s_red[tid], s_red[tid + o]) : (q == 2) ? fmax(s_red[tid], s_red[tid + o]) : s_red[tid] + s_red[tid + o];
      __syncthreads();
    }
    out[q] = s_red[0];
    __syncthreads();
  }
  if (tid == 0) {
    const double sum = out[0], mn = out[1], mx = out[2];
    csum[j] = (mn < 0.0) ? nan("") : sum;
    rowflags[var0 + j] = (mn == mx) ? kRowSkip : ((sum > (double)N) ? kRowFlipped : kRowNormal);
    double* sb = SB + (size_t)j * kMaxER;
    sb[0] = out[3];
    sb[1] = out[4];
    for (int l = 0; l < C; ++l) sb[2 + l] = out[5 + l];
  }
}
// k_wide_dos_gram: A[ja..ja+64)[jb..jb+64) += sum over a sample range of v g_a g_b; grid (block pairs a <= b, sample splits),
// 256 threads, each a 4 x 4 patch of the 64 x 64 block; the two column tiles of 32 samples staged in shared memory.
__global__ void __launch_bounds__(256)
k_wide_dos_gram(const double* __restrict__ G, int64_t N, int M, const double* __restrict__ vw, int nblk, int64_t split_len, double* __restrict__ A) {
  __shared__ double sa[32][64 + 1], sb[32][64 + 1];
  // unrank the pair index into (a, b), a <= b
  int a = 0, rem = blockIdx.x;
  while (rem >= nblk - a) { rem -= nblk - a; ++a; }
  const int b = a + rem;
  const int ja = a * 64, jb = b * 64;
  const int tid = threadIdx.x, tr = (tid >> 4) * 4, tc = (tid & 15) * 4;
  const int64_t i0 = (int64_t)blockIdx.y * split_len;
  int64_t i1 = i0 + split_len;
  if (i1 > N) i1 = N;
  double acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[p][q] = 0.0;
  for (int64_t c0 = i0; c0 < i1; c0 += 32) {
    for (int idx = tid; idx < 32 * 64; idx += 256) {
      const int col = idx >> 5, ii = idx & 31;                 // consecutive threads: consecutive samples of one column
      const int64_t i = c0 + ii;
      const bool in = i < i1;
      const double v = in ? (vw ? vw[i] : 1.0) : 0.0;
      sa[ii][col] = (in && ja + col < M) ? G[(size_t)(ja + col) * N + i] * v : 0.0;
      sb[ii][col] = (in && jb + col < M) ? G[(size_t)(jb + col) * N + i] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int ii = 0; ii < 32; ++ii) {
      double xa[4], xb[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) { xa[p] = sa[ii][tr + p]; xb[p] = sb[ii][tc + p]; }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] += xa[p] * xb[q];
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = ja + tr + p, c = jb + tc + q;
      if (r < M && c < M && r <= c && acc[p][q] != 0.0) atomicAdd(&A[(size_t)r * M + c], acc[p][q]);
    }
}
// k_wide_dos_burden: thread = one sample over all M columns (coalesced across samples): the collapses of src/Model.cpp:73-130 on
// the minor-coded matrix, `(int)g' > 0`, and their sums with r, v, v x_l.  bur: [2][3 + kMaxC] as k_wide_sparse.
__global__ void __launch_bounds__(kWideDosThreads)
k_wide_dos_burden(const double* __restrict__ G, int64_t N, int M, int64_t var0, const uint8_t* __restrict__ rowflags, const NullModel* __restrict__ nm,
                  const double* __restrict__ X, const double* __restrict__ vw, double* __restrict__ bur) {
  __shared__ double s_bur[2][3 + kMaxC];
  extern __shared__ uint8_t s_roleb[];   // [M]
  const int tid = threadIdx.x;
  const int C = nm->C;
  const double* __restrict__ resid = nm->resid;
  for (int j = tid; j < M; j += kWideDosThreads) s_roleb[j] = rowflags[var0 + j];
  if (tid < 2 * (3 + kMaxC)) (&s_bur[0][0])[tid] = 0.0;
  __syncthreads();
  double bz[3 + kMaxC], bc[3 + kMaxC];
#pragma unroll
  for (int l = 0; l < 3 + kMaxC; ++l) bz[l] = bc[l] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kWideDosThreads + tid; i < N; i += (int64_t)gridDim.x * kWideDosThreads) {
    int zi = 0;
    for (int j = 0; j < M; ++j) {
      const uint8_t f = s_roleb[j];
      if (f == kRowSkip) continue;
      const double g = G[(size_t)j * N + i];
      zi += (f == kRowFlipped) ? ((int)(2.0 - g) > 0) : ((int)g > 0);
    }
    if (zi == 0) continue;
    const double z = (double)zi, r = resid[i], v = vw ? vw[i] : 1.0;
    bz[0] += z * r;  bz[1] += v * z * z;  bz[2] += 1.0;
    bc[0] += r;      bc[1] += v;          bc[2] += 1.0;
#pragma unroll
    for (int l = 0; l < kMaxC; ++l)
      if (l < C) {
        const double x = X[(size_t)l * N + i];
        bz[3 + l] += v * z * x;
        bc[3 + l] += v * x;
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
      atomicAdd(&s_bur[0][l], a);
      atomicAdd(&s_bur[1][l], b);
    }
  }
  __syncthreads();
  if (tid < 2 * (3 + kMaxC)) {
    const double val = (&s_bur[0][0])[tid];
    if (val != 0.0) atomicAdd(&bur[tid], val);
  }
}

// Flags of a wide gene with missing calls from its counts, with the imputation folded in -- the decisions k_tile_cols +
// k_aug_flags take for a single tile (DataConsolidator.cpp:46-142 on the mean-imputed matrix): the column sum above N flips,
// all values equal drops.  They steer k_split_hm (fill 0 / 2), the collapse and the tail alike.
__global__ void k_wide_imp_flags(int64_t var0, int M, int64_t N, const RowCounts* __restrict__ counts, uint8_t* __restrict__ rowflags) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const RowCounts rc = counts[var0 + j];
  const long long n1 = rc.n1, n2 = rc.n2, miss = rc.bad, n0 = N - n1 - n2 - miss, nobs = N - miss;
  const double ac = (double)(n1 + 2 * n2);
  const double fill = nobs > 0 ? 2.0 * (ac / (double)(2 * nobs)) : 0.0;
  double mn = 1e300, mx = -1e300;
  if (n0 > 0) { mn = fmin(mn, 0.0); mx = fmax(mx, 0.0); }
  if (n1 > 0) { mn = fmin(mn, 1.0); mx = fmax(mx, 1.0); }
  if (n2 > 0) { mn = fmin(mn, 2.0); mx = fmax(mx, 2.0); }
  if (miss > 0) { mn = fmin(mn, fill); mx = fmax(mx, fill); }
  const double csum = ac + (double)miss * fill;
  rowflags[var0 + j] = (mn == mx) ? kRowSkip : ((csum > (double)N) ? kRowFlipped : kRowNormal);
}

// One CTA per wide gene: steps 2-7 of k_finalize on the global-memory workspace.
template <bool SKATO>
__global__ void __launch_bounds__(kWideThreads)
k_wide_finalize(const WideJob* __restrict__ jobs, int n_jobs, const uint8_t* __restrict__ rowflags, const double* __restrict__ af,
                const RowCounts* __restrict__ counts, const NullModel* __restrict__ nm, EngineParams prm,
                rvt_gene_result* __restrict__ res, QagsScratch* __restrict__ qags) {
  struct SkatoShared {
    QagsMachine mach;
    double fv[21], bcast[3];
  };
  __shared__ typename std::conditional<SKATO, SkatoShared, int>::type s_sk;
  __shared__ double s_red[64];
  __shared__ double s_bur[2][2 + kMaxC];
  __shared__ int s_Mp, s_bad, s_nonref;
  __shared__ double s_Q;
  const int g = blockIdx.x, tid = threadIdx.x;
  if (g >= n_jobs) return;
  constexpr int NT = kWideThreads;
  const WideJob jb = jobs[g];
  const int M = jb.M;
  const int64_t N = nm->N;
  const int C = nm->C, ER = nm->ER;
  const double sigma2 = nm->sigma2;
  BlockPar par{s_red};
  double* s_s = jb.dws;
  double* s_sw = s_s + M;
  double* s_vw = s_sw + M;
  double* s_B = s_vw + M;              // [M][kMaxC]
  double* Uk = s_B + (size_t)M * kMaxC;   // [M][kMaxC]
  double* s_ev = Uk + (size_t)M * kMaxC;
  double* s_e = s_ev + (M + 2);
  double* s_v = s_e + (M + 2);
  double* s_p = s_v + (M + 2);
  double* s_lam = s_p + (M + 2);
  double* s_c = s_lam + (M + 2);
  double* s_lamz = s_c + (M + 2);
  int* s_idx = jb.iws;
  int* s_flip = s_idx + M;
  int* s_th = s_flip + M;              // [NT/32][M]
  long long* s_craw = jb.craw;
  const long long* __restrict__ De = jb.De;
  double* K = jb.K;
  const int kld = M;
  const bool imp = jb.imp == 1;
  const bool dos = jb.imp == 3;             // real dosages: as f64, with the column sums (craw slots) and the flags computed by k_wide_dos_cols
  const bool f64 = jb.imp == 2 || dos;      // fp64 statistics (k_wide_sparse: binary trait; k_wide_dos_*: dosages): A, {S, CW, B}, burden sums as doubles
  const double* __restrict__ Ad = reinterpret_cast<const double*>(jb.A_raw);
  const double* __restrict__ SBd = reinterpret_cast<const double*>(jb.De);
  const double* __restrict__ burd = reinterpret_cast<const double*>(jb.coll);
  const int lda = imp ? 2 * M : M;          // leading dimension of A_raw
  double* s_delta = s_lamz + (M + 2);       // imp: delta_j = fill_j - (flipped ? 2 : 0); the workspace of a 2M gene has the room
  // column sums of the imputed matrix: imp -- behind s_delta (2M workspace); f64 -- the craw slots (the workspace is M-sized there)
  double* s_csum = f64 ? reinterpret_cast<double*>(jb.craw) : s_delta + M;

  if (tid == 0) s_bad = 0;
  __syncthreads();
  // 2. per-variant counts -> flip / monomorphic, cross-checked with the flags the collapse used
  if (dos) {
    for (int j = tid; j < M; j += NT) {
      const uint8_t f = rowflags[jb.var0 + j];
      if (!(s_csum[j] == s_csum[j])) atomicExch(&s_bad, 2);   // a negative value (a missing call that was never imputed) reached fit()
      s_flip[j] = (f == kRowSkip) ? -1 : (f == kRowFlipped);
    }
  } else if (imp || f64) {
    for (int j = tid; j < M; j += NT) {
      const RowCounts rc = counts[jb.var0 + j];
      const long long n1 = rc.n1, n2 = rc.n2, miss = rc.bad, n0 = N - n1 - n2 - miss, nobs = N - miss;
      const double ac = (double)(n1 + 2 * n2);
      const double fill = nobs > 0 ? 2.0 * (ac / (double)(2 * nobs)) : 0.0;   // imputeGenotypeToMean: 2 p^ (as k_tile_cols)
      double mn = 1e300, mx = -1e300;
      if (n0 > 0) { mn = fmin(mn, 0.0); mx = fmax(mx, 0.0); }
      if (n1 > 0) { mn = fmin(mn, 1.0); mx = fmax(mx, 1.0); }
      if (n2 > 0) { mn = fmin(mn, 2.0); mx = fmax(mx, 2.0); }
      if (miss > 0) { mn = fmin(mn, fill); mx = fmax(mx, fill); }
      const double csum = ac + (double)miss * fill;
      const int mono = mn == mx, flip = csum > (double)N;
      const uint8_t expect = mono ? kRowSkip : (flip ? kRowFlipped : kRowNormal);
      if (rowflags[jb.var0 + j] != expect) atomicExch(&s_bad, 1);
      // the integer sums must be those of the rows [H ; Mi]: H'H_jj = n1 + 4 n2 (+ 4 miss in a flipped row), Mi'Mi_jj = miss
      if (imp) {
        const long long hjj = jb.A_raw[(size_t)j * lda + j], mjj = jb.A_raw[(size_t)(M + j) * lda + (M + j)];
        if (mjj != miss || hjj != n1 + 4 * n2 + (flip && !mono ? 4 * miss : 0)) atomicExch(&s_bad, 2);
      }
      s_csum[j] = csum;
      if (imp) {
        s_craw[j] = 0;
        s_delta[j] = fill - (flip && !mono ? 2.0 : 0.0);
      }
      s_flip[j] = mono ? -1 : flip;
    }
  } else
  for (int j = tid; j < M; j += NT) {
    const long long cint = recombine4(&De[(size_t)j * ER + 4]);
    const long long c = llrint((double)cint * nm->scale[1]);
    const long long ajj = jb.A_raw[(size_t)j * M + j];
    const long long n2 = (ajj - c) / 2, n1 = c - 2 * n2, n0 = N - n1 - n2;
    const int flip = c > N;
    const int mono = (n0 == N) || (n1 == N) || (n2 == N);
    const uint8_t expect = mono ? kRowSkip : (flip ? kRowFlipped : kRowNormal);
    if (rowflags[jb.var0 + j] != expect) atomicExch(&s_bad, 1);
    if ((ajj - c) & 1 || n0 < 0 || n1 < 0 || n2 < 0) atomicExch(&s_bad, 2);
    if (jb.counted && counts[jb.var0 + j].bad > 0) atomicExch(&s_bad, 2);
    s_craw[j] = c;
    s_flip[j] = mono ? -1 : flip;
  }
  __syncthreads();
  if (tid == 0) {
    int mp = 0;
    for (int j = 0; j < M; ++j)
      if (s_flip[j] >= 0) s_idx[mp++] = j;
    s_Mp = mp;
  }
  __syncthreads();
  const int Mp = s_Mp;
  // 3. score vector, covariate cross-products, weights (kept variants, minor-coded)
  for (int t = tid; t < Mp; t += NT) {
    const int j = s_idx[t];
    const int fl = s_flip[j];
    if (f64) {
      // (2 - g)'r = 2 sum r - g'r, (2 - g)'V x_l = 2 sum v x_l - g'V x_l   (dosage_prepare)
      const double* sb = SBd + (size_t)j * kMaxER;
      s_s[t] = fl ? 2.0 * nm->rsum - sb[0] : sb[0];
      for (int l = 0; l < C; ++l) s_B[t * kMaxC + l] = fl ? 2.0 * (nm->binary ? nm->xsum_w[l] : nm->xsum[l]) - sb[2 + l] : sb[2 + l];
    } else if (imp) {
      // G'v = H'v + delta (Mi'v) for v = r, X_l; a flipped column is 2 - g
      const double dj = s_delta[j];
      double sv = ((double)recombine4(&De[(size_t)j * ER]) + dj * (double)recombine4(&De[(size_t)(M + j) * ER])) * nm->scale[0];
      if (fl) sv = 2.0 * (double)nm->vsum[0] * nm->scale[0] - sv;
      s_s[t] = sv;
      for (int l = 0; l < C; ++l) {
        double b = ((double)recombine4(&De[(size_t)j * ER + 4 * (l + 1)]) + dj * (double)recombine4(&De[(size_t)(M + j) * ER + 4 * (l + 1)])) *
                   nm->scale[l + 1];
        if (fl) b = 2.0 * (double)nm->vsum[l + 1] * nm->scale[l + 1] - b;
        s_B[t * kMaxC + l] = b;
      }
    } else {
    long long sint = recombine4(&De[(size_t)j * ER]);
    if (fl) sint = 2 * nm->vsum[0] - sint;
    s_s[t] = (double)sint * nm->scale[0];
    for (int l = 0; l < C; ++l) {
      long long b = recombine4(&De[(size_t)j * ER + 4 * (l + 1)]);
      if (fl) b = 2 * nm->vsum[l + 1] - b;
      s_B[t * kMaxC + l] = (double)b * nm->scale[l + 1];
    }
    }
    // weight t of the kept columns uses af[t] in the caller's ORIGINAL order
    const double freq = jb.has_af ? af[jb.var0 + t] : ((imp || f64) ? s_csum[j] : (double)s_craw[j]) / (2.0 * (double)N);
    s_sw[t] = sqrt(beta_weight(freq, prm.beta1, prm.beta2, true));
  }
  __syncthreads();
  if (tid == 0) {
    double q = 0.0;
    for (int i = 0; i < Mp; ++i) q += (s_sw[i] * s_sw[i]) * s_s[i] * s_s[i];
    s_Q = q;
  }
  // 4. K = W^1/2 sigma2 (A' - B' (X'X)^-1 B'^T) W^1/2
  for (int t = tid; t < Mp; t += NT)
    for (int l = 0; l < C; ++l) {
      double u = 0.0;
      for (int m = 0; m < C; ++m) u += nm->xtx_inv[l * C + m] * s_B[t * kMaxC + m];
      Uk[t * kMaxC + l] = u;
    }
  __syncthreads();
  for (size_t idx = tid; idx < (size_t)Mp * Mp; idx += NT) {
    const int i = (int)(idx / Mp), k = (int)(idx - (size_t)i * Mp);
    if (k < i) continue;
    const int ji = s_idx[i], jk = s_idx[k];
    const int fi = s_flip[ji], fk = s_flip[jk];
    double ad;
    if (f64) {
      // the weighted Gram (upper triangle accumulated); g' = 2 - g under the weights: 4 sum v - 2 c^w_j - 2 c^w_k + A_jk, c^w = G'v
      ad = Ad[(size_t)(ji < jk ? ji : jk) * M + (ji < jk ? jk : ji)];
      const double ci = nm->binary ? SBd[(size_t)ji * kMaxER + 1] : s_csum[ji], ck = nm->binary ? SBd[(size_t)jk * kMaxER + 1] : s_csum[jk];
      if (fi && fk)
        ad = 4.0 * (nm->binary ? nm->vsum_w : (double)N) - 2.0 * ci - 2.0 * ck + ad;
      else if (fi)
        ad = 2.0 * ck - ad;
      else if (fk)
        ad = 2.0 * ci - ad;
    } else if (imp) {
      // G'G = H'H + H'Mi D + D Mi'H + D Mi'Mi D   (D = diag delta), then the same flip identities on the imputed column sums
      const double di = s_delta[ji], dk = s_delta[jk];
      ad = (double)jb.A_raw[(size_t)ji * lda + jk] + dk * (double)jb.A_raw[(size_t)ji * lda + (M + jk)] +
           di * (double)jb.A_raw[(size_t)(M + ji) * lda + jk] + di * dk * (double)jb.A_raw[(size_t)(M + ji) * lda + (M + jk)];
      // (H carries the fill 2 of a flipped row, i.e. ad is the Gram of the UNflipped imputed columns only for normal rows:
      //  for a flipped row H + Mi delta = g with g_missing = 2 + (fill - 2) = fill as well -- both cases are the raw imputed g)
      const double ci = s_csum[ji], ck = s_csum[jk];
      if (fi && fk)
        ad = 4.0 * (double)N - 2.0 * ci - 2.0 * ck + ad;
      else if (fi)
        ad = 2.0 * ck - ad;
      else if (fk)
        ad = 2.0 * ci - ad;
    } else {
    long long a = jb.A_raw[(size_t)ji * M + jk];
    const long long ci = s_craw[ji], ck = s_craw[jk];
    if (fi && fk)
      a = 4 * N - 2 * ci - 2 * ck + a;
    else if (fi)
      a = 2 * ck - a;
    else if (fk)
      a = 2 * ci - a;
    ad = (double)a;
    }
    double tt = 0.0;
    for (int l = 0; l < C; ++l) tt += s_B[i * kMaxC + l] * Uk[k * kMaxC + l];
    const double v = s_sw[i] * s_sw[k] * sigma2 * (ad - tt);
    K[(size_t)i * kld + k] = v;
    K[(size_t)k * kld + i] = v;
  }
  if (tid == 0 && f64) {
    for (int which = 0; which < 2; ++which) {
      const double* b = burd + which * (3 + kMaxC);          // U, SS, count, SZ[]
      s_bur[which][0] = b[0];
      s_bur[which][1] = b[1];
      for (int l = 0; l < C; ++l) s_bur[which][2 + l] = b[3 + l];
    }
    s_nonref = (int)llrint(burd[(3 + kMaxC) + 2]);
  } else if (tid == 0) {
    for (int which = 0; which < 2; ++which) {
      const long long* cl = jb.coll + which * (ER + 1);
      s_bur[which][0] = (double)recombine4(cl) * nm->scale[0];
      s_bur[which][1] = (double)cl[ER];
      for (int l = 0; l < C; ++l) s_bur[which][2 + l] = (double)recombine4(cl + 4 * (l + 1)) * nm->scale[l + 1];
    }
    s_nonref = (int)jb.coll[(ER + 1) + ER];
  }
  __syncthreads();   // every read of A_raw is done: Wm may overwrite it

  if (SKATO && qags) {
    const double sc = 0.5 / sigma2;
    for (size_t idx = tid; idx < (size_t)Mp * Mp; idx += NT) {
      const int i = (int)(idx / Mp), k = (int)(idx - (size_t)i * Mp);
      jb.Wm[(size_t)i * kld + k] = K[(size_t)i * kld + k] * sc;
    }
    for (int t = tid; t < Mp; t += NT) s_vw[t] = s_sw[t] * s_s[t];
    __syncthreads();
  }
  // 5./6. eigenvalues, Davies / Liu
  double p_dav = -1.0, p_liu = 1.0, p_fin = 1.0, lam_max = 0.0;
  int fault = 0, r = 0;
  if (Mp > 0) {
    sym_eigenvalues_tridiag(K, Mp, kld, s_ev, s_e, s_v, s_p, s_lam, par);
    const int r_ub = (N < (int64_t)Mp) ? (int)N : Mp;
    while (r < r_ub && s_lam[r] > 1e-30) ++r;
    lam_max = r ? s_lam[0] : 0.0;
    p_dav = mixchisq_pvalue(s_lam, r, s_Q, s_th, &fault, par);
    p_liu = liu_pvalue(s_lam, r, s_Q);
    p_fin = p_dav;
    if (p_fin <= 0.0 || p_fin == 1.0) p_fin = p_liu;
  }
  SkatoOut so;
  so.ok = 0;
  so.timed_out = 0;
  so.Q = so.rho = so.pvalue = 0.0;
  if constexpr (SKATO) {
    if (qags && Mp > 0) {
      QagsWork work{qags[g].a, qags[g].b, qags[g].r, qags[g].e, qags[g].order, qags[g].level, kQagsLimit};
      work.deadline = prm.wd_cycles > 0 ? clock64() + prm.wd_cycles : 0;
      const double s2 = nm->binary ? 1.0 : sigma2 * (double)N / (double)(N - 1);   // SkatO.cpp:133-137 (type "D": s2 = 1)
      so = skato_tail(jb.Wm, K, Mp, kld, s_vw, s2, s_ev, s_e, s_v, s_p, s_lamz, s_c, &s_sk.mach, work, s_sk.fv, s_sk.bcast, s_th, M, par);
    }
  }
  if (tid == 0)
    burden_and_store(&res[jb.out_index], Mp, s_bad, s_Q, p_fin, p_dav, p_liu, fault, r, lam_max, so, s_bur, s_nonref, nm, so.timed_out ? RVT_GENE_TIMEOUT : 0);
}

}  // namespace rvt
