// perm.cuh -- A6: the permutation p-value of `--kernel skat` (the reference's DEFAULT: nPerm = 10000, alpha = 0.05).
//
//   SkatTest::fit           src/Model.h:2707-2717   permutedRes = res; while (perm.next()) { permute(&permutedRes);
//                                                   s = skat.GetQFromNewResidual(permutedRes); perm.add(s); }
//   permute                 src/LinearAlgebra.h:8-21   Fisher-Yates, i = n-1..1, j = rand() % (i + 1), swap(v[i], v[j])
//   Permutation             src/Permutation.h:49-98    adaptive stop: numX + numEqual >= (int)(nPerm * alpha * 2)
//   Skat::GetQFromNewResidual  regression/Skat.cpp:107-116   Q' = || W^1/2 G' r_pi ||^2
//
// The reference never calls srand(): the shuffles consume glibc's default rand() stream (seed 1) in order,
// gene after gene, and every shuffle is applied to the ALREADY shuffled residual.  To reproduce the same
// permutations -- not merely equally distributed ones -- three serial-looking steps are made parallel:
//
//  1. rand().  glibc's TYPE_3 generator is the additive lagged Fibonacci recurrence y_m = y_{m-31} + y_{m-3}
//     (mod 2^32), output y >> 1.  It is LINEAR, so the state at any stream position n is
//     z^n mod (z^31 - z^28 - 1) applied to the initial window: the host supplies the window at the start of a
//     batch plus two tables of jump polynomials, each CTA jumps to its block of draws and each thread to its
//     run of 248 draws, which it then generates with the 31-word state held in registers.
//  2. Fisher-Yates.  Position i receives its final value at step i and is never touched again, so
//     final[i] = "what position j_i held just before step i".  What a position q holds before step t is the
//     value written by the most recent earlier step that targeted q (the smallest i' > t with j_i' = q), which
//     wrote what position i' held before step i' -- and so on up a chain of strictly increasing step numbers
//     that ends at a position nobody wrote to (its original content).  The steps targeting a position form a
//     short list (expected length ln(n/q)), built with one atomicExch per step; every output element then walks
//     its chain independently.  One thread per (permutation, sample).
//  3. Q'.  A shuffled residual is a shuffled copy of the 4 int8 digit rows of r in the null-model image E, so a
//     group of 16 permutations is one 64-row tile in the engine's tiled layout, and G' r_pi for all of them is
//     ONE tile-pair unit of the tensor-core sweep (PAIR mode, exact integers): 16 permutations per pass over
//     the gene's genotypes.
//
// Q' is evaluated in exact integer arithmetic + fp64 where the reference uses float32 (Skat.cpp:111-113); the
// comparison s > obs can therefore differ from the reference only for a permuted statistic within ~1e-6 of
// the observed one.
#pragma once
#include "common.cuh"
#include "permlogic.cuh"

namespace rvt {

#if defined(__CUDACC__)
constexpr int kLfgThreads = 256;
constexpr int kLfgRun = 8 * kLfgDeg;                  // draws per thread
constexpr int kLfgBlock = kLfgThreads * kLfgRun;      // draws per CTA

struct LfgTables {
  LfgPoly zblock[32];            // z^(kLfgBlock * 2^k)
  LfgPoly zthread[kLfgThreads];  // z^(t * kLfgRun)
};

// draws[d] = the (pos0 + d)-th value rand() returns after srand(seed), d < cnt.  w0: window y_{pos0+341 .. +60}.
__global__ void __launch_bounds__(kLfgThreads)
k_lfg_draws(const LfgTables* __restrict__ tab, const uint32_t* __restrict__ w0 /*[61]*/, uint64_t cnt, uint32_t* __restrict__ draws) {
  __shared__ LfgPoly s_acc;
  __shared__ uint32_t s_prod[2 * kLfgDeg - 1];
  __shared__ uint32_t s_w[2 * kLfgDeg - 1];
  const int tid = threadIdx.x;
  const unsigned b = blockIdx.x;
  // 1. C_b = z^(b * kLfgBlock) = product of zblock[k] over the set bits of b
  if (tid < kLfgDeg) s_acc.c[tid] = (tid == 0) ? 1u : 0u;
  __syncthreads();
  for (int k = 0; k < 32; ++k) {
    if (!((b >> k) & 1u)) continue;   // uniform
    if (tid < 2 * kLfgDeg - 1) {
      uint32_t s = 0;
      const int lo = tid < kLfgDeg ? 0 : tid - kLfgDeg + 1, hi = tid < kLfgDeg ? tid : kLfgDeg - 1;
      for (int i = lo; i <= hi; ++i) s += s_acc.c[i] * tab->zblock[k].c[tid - i];
      s_prod[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
      for (int d = 2 * kLfgDeg - 2; d >= kLfgDeg; --d) {
        s_prod[d - 3] += s_prod[d];
        s_prod[d - kLfgDeg] += s_prod[d];
      }
    }
    __syncthreads();
    if (tid < kLfgDeg) s_acc.c[tid] = s_prod[tid];
    __syncthreads();
  }
  // 2. the CTA's window, extended to 61 values
  if (tid < kLfgDeg) {
    uint32_t s = 0;
    for (int j = 0; j < kLfgDeg; ++j) s += s_acc.c[j] * w0[j + tid];
    s_w[tid] = s;
  }
  __syncthreads();
  if (tid == 0)
    for (int k = kLfgDeg; k < 2 * kLfgDeg - 1; ++k) s_w[k] = s_w[k - 31] + s_w[k - 3];
  __syncthreads();
  // 3. the thread's own 31-word state, then kLfgRun draws with the state in registers
  const uint64_t d0 = (uint64_t)b * kLfgBlock + (uint64_t)tid * kLfgRun;
  if (d0 >= cnt) return;
  uint32_t x[kLfgDeg];
  {
    const LfgPoly& zt = tab->zthread[tid];
#pragma unroll
    for (int k = 0; k < kLfgDeg; ++k) x[k] = 0;
    for (int j = 0; j < kLfgDeg; ++j) {
      const uint32_t cj = zt.c[j];
#pragma unroll
      for (int k = 0; k < kLfgDeg; ++k) x[k] += cj * s_w[j + k];
    }
  }
  for (int rep = 0; rep < kLfgRun / kLfgDeg; ++rep) {
    const uint64_t d = d0 + (uint64_t)rep * kLfgDeg;
#pragma unroll
    for (int k = 0; k < kLfgDeg; ++k)
      if (d + k < cnt) draws[d + k] = x[k] >> 1;
    // next 31 values: slot k holds y_{m-31}; y_{m-3} sits in slot k-3 (already advanced) or k+28
#pragma unroll
    for (int k = 0; k < kLfgDeg; ++k) x[k] += x[(k + kLfgDeg - 3) % kLfgDeg];
  }
}

// one thread per (permutation p, step): step s = 0..N-2 handles position i = N-1-s with the p*(N-1)+s-th draw
__global__ void __launch_bounds__(256)
k_fy_link(const uint32_t* __restrict__ draws, uint32_t N, int P, uint32_t* __restrict__ head, uint32_t* __restrict__ link) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t per = (uint64_t)N - 1;
  if (idx >= per * (uint64_t)P) return;
  const uint64_t p = idx / per;
  const uint32_t s = (uint32_t)(idx - p * per), i = N - 1 - s;
  const uint32_t j = draws[idx] % (i + 1);
  link[p * N + i] = atomicExch(&head[p * N + j], i);
}

// root[p][i] = index, in the vector before shuffle p, of the value shuffle p leaves at position i
__global__ void __launch_bounds__(256)
k_fy_root(const uint32_t* __restrict__ draws, uint32_t N, int P, const uint32_t* __restrict__ head, const uint32_t* __restrict__ link,
          uint32_t* __restrict__ root) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (uint64_t)N * (uint64_t)P) return;
  const uint64_t p = idx / N;
  const uint32_t i = (uint32_t)(idx - p * N);
  const uint32_t j = (i == 0) ? 0u : draws[p * ((uint64_t)N - 1) + (N - 1 - i)] % (i + 1);
  root[idx] = fy_root(head + p * N, link + p * N, i, j);
}

// the 4 balanced base-256 digits of the null residual (rows 0..3 of E), one packed word per sample
__global__ void __launch_bounds__(256)
k_perm_init(const NullModel* __restrict__ nm, uint32_t* __restrict__ R) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nm->N) return;
  uint32_t w = 0;
  for (int k = 0; k < 4; ++k) w |= (uint32_t)(uint8_t)nm->E[(size_t)k * nm->ldE + i] << (8 * k);
  R[i] = w;
}

// R_out[i] = R_in[root[i]]; the 4 digits also go to rows row0..row0+3 of a 64-row tile in the tiled layout
__global__ void __launch_bounds__(256)
k_perm_gather(const uint32_t* __restrict__ R_in, const uint32_t* __restrict__ root, uint32_t N, uint32_t* __restrict__ R_out,
              int8_t* __restrict__ tile, int row0) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t w = R_in[root[i]];
  R_out[i] = w;
  int8_t* p = tile + ((size_t)(i >> 7) * kTileRows + row0) * 128 + (i & 127);
#pragma unroll
  for (int k = 0; k < 4; ++k) p[k * 128] = (int8_t)(w >> (8 * k));
}

// One CTA per sweep unit (gene tile x 16-permutation tile): sint[p][variant] = G_variant . r_pi (fixed point)
__global__ void __launch_bounds__(128)
k_perm_sint(const GeneDesc* __restrict__ units, int n_units, int64_t var_base, int M, int S, const SweepPartial* __restrict__ parts,
            long long* __restrict__ sint /*[P][M]*/) {
  const int u = blockIdx.x, tid = threadIdx.x;
  if (u >= n_units) return;
  const GeneDesc gd = units[u];
  const int r0 = (int)(gd.var0 - var_base), Ma = gd.M;
  const int pg = (int)gd.var0_b;   // permutation group of the B tile
  const SweepPartial* __restrict__ gp = parts + (size_t)u * S;
  for (int idx = tid; idx < Ma * 16; idx += 128) {
    const int i = idx >> 4, pl = idx & 15;
    long long d[4] = {0, 0, 0, 0};
    for (int sp = 0; sp < S; ++sp)
      for (int k = 0; k < 4; ++k) d[k] += gp[sp].d[i][4 * pl + k];
    sint[(size_t)(pg * 16 + pl) * M + (r0 + i)] = d[0] + (d[1] << 8) + (d[2] << 16) + (d[3] << 24);
  }
}

// Q'_p = sum over kept variants of w_t (s'_t)^2, in the variant order of k_finalize (one CTA, thread = permutation)
__global__ void __launch_bounds__(256)
k_perm_q(int P, int M, int64_t var_base, int has_af, const long long* __restrict__ sint, const uint8_t* __restrict__ rowflags,
         const double* __restrict__ af, const RowCounts* __restrict__ counts, const NullModel* __restrict__ nm, EngineParams prm,
         double* __restrict__ w /*[M] scratch*/, double* __restrict__ Qout /*[P]*/) {
  const int tid = threadIdx.x;
  if (tid == 0) {   // kept index t of variant j, and its weight (af[t] in the caller's ORIGINAL order: SURVEY.md F9)
    int t = 0;
    for (int j = 0; j < M; ++j) {
      if (rowflags[var_base + j] == kRowSkip) {
        w[j] = -1.0;
        continue;
      }
      const double freq = has_af ? af[var_base + t]
                                 : (double)((long long)counts[var_base + j].n1 + 2ll * counts[var_base + j].n2) / (2.0 * (double)nm->N);
      w[j] = beta_weight(freq, prm.beta1, prm.beta2, true);
      ++t;
    }
  }
  __syncthreads();
  for (int p = tid; p < P; p += blockDim.x) {
    double q = 0.0;
    for (int j = 0; j < M; ++j) {
      if (w[j] < 0.0) continue;
      long long s = sint[(size_t)p * M + j];
      if (rowflags[var_base + j] == kRowFlipped) s = 2 * nm->vsum[0] - s;   // sum_i r_pi(i) = sum_i r_i
      const double sd = (double)s * nm->scale[0];
      const double sw = sqrt(w[j]);
      q += (sw * sw) * sd * sd;
    }
    Qout[p] = q;
  }
}
// ---- genes with missing calls (mean-imputed: G = H + M diag(delta), sweep_aug.cuh) -------------------------------------
// k_split_hm writes the two operand tiles the permuted products need -- H (hard calls, missing -> fill) and M (0/1 missing
// indicators) -- as tiled blocks; the sweep units then pair both with the permutation tiles, sint holds H'r_pi in columns
// 0..M-1 and M'r_pi in columns M..2M-1, and k_perm_q_aug combines them: s = H'r_pi + delta (M'r_pi).
// grid: (ceil(nchunk*32/256), M); one thread = one 4-sample word of one row of one 128-sample chunk
__global__ void __launch_bounds__(256)
k_split_hm(const int8_t* __restrict__ g, int M, int64_t N, const uint8_t* __restrict__ rowflags /* of this gene */,
           int8_t* __restrict__ H, int8_t* __restrict__ Mt) {
  const int r = blockIdx.y;
  const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // word index along the samples
  const int64_t nchunk = (N + 127) >> 7;
  if (wi >= nchunk * 32) return;
  const size_t off = ((size_t)(wi >> 5) * M + r) * 128 + (size_t)(wi & 31) * 4;
  const uint32_t w = *reinterpret_cast<const uint32_t*>(g + off);
  const uint32_t m = w & (w >> 1) & 0x01010101u;
  const uint32_t fill2 = (rowflags[r] == kRowFlipped) ? 0x02020202u : 0u;
  *reinterpret_cast<uint32_t*>(H + off) = (w & ~(m | (m << 1))) | ((m << 1) & fill2);
  *reinterpret_cast<uint32_t*>(Mt + off) = m;
}

__global__ void __launch_bounds__(256)
k_perm_q_aug(int P, int M, int64_t var_base, int has_af, const long long* __restrict__ sint /*[P][2M]*/, const uint8_t* __restrict__ rowflags,
             const double* __restrict__ af, const RowCounts* __restrict__ counts, const NullModel* __restrict__ nm, EngineParams prm,
             double* __restrict__ w /*[2M] scratch: weights, then deltas*/, double* __restrict__ Qout /*[P]*/) {
  const int tid = threadIdx.x;
  const double N = (double)nm->N;
  if (tid == 0) {
    int t = 0;
    for (int j = 0; j < M; ++j) {
      const RowCounts rc = counts[var_base + j];
      const double nobs = N - (double)rc.bad, ac = (double)((long long)rc.n1 + 2ll * rc.n2);
      const double fill = nobs > 0 ? 2.0 * (ac / (2.0 * nobs)) : 0.0;          // imputeGenotypeToMean (as k_tile_cols)
      w[M + j] = fill - ((rowflags[var_base + j] == kRowFlipped) ? 2.0 : 0.0);
      if (rowflags[var_base + j] == kRowSkip) {
        w[j] = -1.0;
        continue;
      }
      const double freq = has_af ? af[var_base + t] : 0.5 * (ac + (double)rc.bad * fill) / N;   // dosage_prepare's frequency
      w[j] = beta_weight(freq, prm.beta1, prm.beta2, true);
      ++t;
    }
  }
  __syncthreads();
  const double rsum = (double)nm->vsum[0] * nm->scale[0];   // sum_i r_pi(i) = sum_i r_i
  for (int p = tid; p < P; p += blockDim.x) {
    double q = 0.0;
    for (int j = 0; j < M; ++j) {
      if (w[j] < 0.0) continue;
      double sd = ((double)sint[(size_t)p * 2 * M + j] + w[M + j] * (double)sint[(size_t)p * 2 * M + M + j]) * nm->scale[0];
      if (rowflags[var_base + j] == kRowFlipped) sd = 2.0 * rsum - sd;
      const double sw = sqrt(w[j]);
      q += (sw * sw) * sd * sd;
    }
    Qout[p] = q;
  }
}
#endif  // __CUDACC__

}  // namespace rvt
