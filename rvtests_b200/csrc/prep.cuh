// prep.cuh -- K0: getting genotypes into the packed device layout, plus the synthetic cohort.
//
//   k_synth_rows     counter-based HWE genotype generator; host twin:
//                    oracle/oracle.py synth_genotypes().  Writes int8 rows + per-row counts.
//   k_count_rows     per-row counts (#1, #2, #invalid) of caller-supplied int8 blocks
//   k_pack_f64       reference boundary: N x M column-major doubles (base/MathMatrix.h:33-41)
//                    -> int8 rows + counts; non-{0,1,2} values are counted as invalid
//   k_flags_from_counts  flip-to-minor / monomorphic flags + allele frequency from the counts:
//                    convertToMinorAlleleCount (colsum > N => flip, src/DataConsolidator.cpp:46-69),
//                    isMonomorphicMarker (:94-116), GenotypeCounter::getAF (src/GenotypeCounter.h:46-52)
#pragma once
#include "common.cuh"

namespace rvt {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// tiled address of (row r of an M-row gene, sample i): [chunk][M][128]
__device__ __forceinline__ size_t tiled_off(int M, int r, int64_t i) { return ((size_t)(i >> 7) * M + r) * 128 + (i & 127); }

// grid: (ceil(npad/16/256), rows), npad = N rounded up to 128.  One thread = 16 consecutive samples
// of one row; row = gene*M + r, genes of equal M stacked back to back in the tiled layout.
__global__ void __launch_bounds__(256)
k_synth_rows(int8_t* __restrict__ arena, int M, int64_t gene_bytes, int64_t N, const unsigned long long* __restrict__ keys,
             const uint32_t* __restrict__ t0, const uint32_t* __restrict__ t1, RowCounts* __restrict__ counts) {
  const int64_t row = blockIdx.y;
  const int64_t c16 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i0 = c16 * 16;
  const int64_t npad = (N + 127) & ~(int64_t)127;
  int n1 = 0, n2 = 0;
  if (i0 < npad) {
    const unsigned long long key = keys[row];
    const uint32_t a = t0[row], b = t1[row];
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t word = 0;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int64_t i = i0 + 4 * q + s;
        uint32_t gq = 0;
        if (i < N) {
          uint32_t h = (uint32_t)(mix64(key + (unsigned long long)i * 0xD1B54A32D192ED03ull) >> 32);
          gq = (h >= a) + (h >= b);
        }
        n1 += (gq == 1);
        n2 += (gq == 2);
        word |= gq << (8 * s);
      }
      w[q] = word;
    }
    const int64_t gene = row / M;
    const int r = (int)(row - gene * M);
    *reinterpret_cast<uint4*>(arena + (size_t)gene * gene_bytes + tiled_off(M, r, i0)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    n1 += __shfl_xor_sync(0xffffffffu, n1, o);
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
  }
  if ((threadIdx.x & 31) == 0 && (n1 | n2)) {
    atomicAdd(&counts[row].n1, n1);
    atomicAdd(&counts[row].n2, n2);
  }
}

// grid: (ceil(N/16/256), rows); base + row*ld must be 16-byte aligned.
__global__ void __launch_bounds__(256)
k_count_rows(const int8_t* __restrict__ base, int64_t ld, int64_t N, RowCounts* __restrict__ counts) {
  const int64_t row = blockIdx.y;
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  int n1 = 0, n2 = 0, bad = 0;
  if (i0 < N) {
    const int8_t* p = base + (size_t)row * ld + i0;
    if (i0 + 16 <= N) {
      uint4 v = *reinterpret_cast<const uint4*>(p);
      uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          uint32_t gq = (w[q] >> (8 * s)) & 0xFF;
          n1 += (gq == 1);
          n2 += (gq == 2);
          bad += (gq > 2);
        }
    } else {
      for (int64_t i = i0; i < N; ++i) {
        uint32_t gq = (uint8_t)p[i - i0];
        n1 += (gq == 1);
        n2 += (gq == 2);
        bad += (gq > 2);
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    n1 += __shfl_xor_sync(0xffffffffu, n1, o);
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0 && (n1 | n2 | bad)) {
    atomicAdd(&counts[row].n1, n1);
    atomicAdd(&counts[row].n2, n2);
    atomicAdd(&counts[row].bad, bad);
  }
}

// grid: (ceil(npad/4/256), M).  src: N x M column-major doubles; dst: tiled int8 block (zero padded).
// Values outside {0,1,2} are stored as code 3 and their range per row is kept in frac[row] = {min, max} (bit patterns of
// positive doubles order like the values; anything else -- negative, NaN -- sets max to all ones).  The Matrix that
// ModelFitter::fit() sees has ALREADY been through DataConsolidator::imputeGenotypeToMean (src/DataConsolidator.cpp:217-245): a
// column whose non-integer entries all equal 2 p^ of its observed calls is a hard-call column with missing calls, and the host
// then treats the gene exactly like a 2-bit push with code 01 (augmented sweep / wide operand tiles) instead of as dosages.
__global__ void __launch_bounds__(256)
k_pack_f64(const double* __restrict__ src, int64_t N, int8_t* __restrict__ dst, int M,
           RowCounts* __restrict__ counts, unsigned long long* __restrict__ frac /* [rows][2] or null */) {
  const int64_t row = blockIdx.y;
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int64_t npad = (N + 127) & ~(int64_t)127;
  int n1 = 0, n2 = 0, bad = 0;
  unsigned long long fmn = ~0ull, fmx = 0ull;
  if (i0 < npad) {
    uint32_t word = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int64_t i = i0 + s;
      uint32_t gq = 0;
      if (i < N) {
        double v = src[(size_t)row * N + i];
        if (v == 0.0) gq = 0;
        else if (v == 1.0) gq = 1;
        else if (v == 2.0) gq = 2;
        else {
          gq = 3;
          bad += 1;
          const unsigned long long b = (v > 0.0 && v < 1e300) ? (unsigned long long)__double_as_longlong(v) : ~0ull;
          fmn = b < fmn ? b : fmn;
          fmx = b > fmx ? b : fmx;
        }
      }
      n1 += (gq == 1);
      n2 += (gq == 2);
      word |= gq << (8 * s);
    }
    *reinterpret_cast<uint32_t*>(dst + tiled_off(M, (int)row, i0)) = word;
  }
  for (int o = 16; o > 0; o >>= 1) {
    n1 += __shfl_xor_sync(0xffffffffu, n1, o);
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
    const unsigned long long a = __shfl_xor_sync(0xffffffffu, fmn, o), b = __shfl_xor_sync(0xffffffffu, fmx, o);
    fmn = a < fmn ? a : fmn;
    fmx = b > fmx ? b : fmx;
  }
  if (frac && (threadIdx.x & 31) == 0 && bad) {
    atomicMin(&frac[2 * row], fmn);
    atomicMax(&frac[2 * row + 1], fmx);
  }
  if ((threadIdx.x & 31) == 0 && (n1 | n2 | bad)) {
    atomicAdd(&counts[row].n1, n1);
    atomicAdd(&counts[row].n2, n2);
    atomicAdd(&counts[row].bad, bad);
  }
}

// grid: (ceil(npad/16/256), M).  src: int8 [M][ld_src] variant-major; dst: tiled block; counts rows.
__global__ void __launch_bounds__(256)
k_tile_rows(const int8_t* __restrict__ src, int64_t ld_src, int64_t N, int8_t* __restrict__ dst, int M,
            RowCounts* __restrict__ counts) {
  const int64_t row = blockIdx.y;
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  const int64_t npad = (N + 127) & ~(int64_t)127;
  int n1 = 0, n2 = 0, bad = 0;
  if (i0 < npad) {
    uint32_t w[4] = {0, 0, 0, 0};
    for (int s = 0; s < 16; ++s) {
      const int64_t i = i0 + s;
      uint32_t gq = 0;
      if (i < N) gq = (uint8_t)src[(size_t)row * ld_src + i];
      n1 += (gq == 1);
      n2 += (gq == 2);
      bad += (gq > 2);
      w[s >> 2] |= gq << (8 * (s & 3));
    }
    *reinterpret_cast<uint4*>(dst + tiled_off(M, (int)row, i0)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    n1 += __shfl_xor_sync(0xffffffffu, n1, o);
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0 && (n1 | n2 | bad)) {
    atomicAdd(&counts[row].n1, n1);
    atomicAdd(&counts[row].n2, n2);
    atomicAdd(&counts[row].bad, bad);
  }
}

// PLINK .bed SNP-major rows -> tiled int8 block + counts.  Encoding as the reference decodes it
// (libVcf/PlinkInputFile.cpp:23-47, PlinkInputFile.h:206-209): sample p sits in bits 2(p&3)..2(p&3)+1
// of byte p>>2 of its variant's row; 00 -> 0 (HOM_REF), 10 -> 1 (HET), 11 -> 2 (HOM_ALT), 01 -> missing.
// Missing calls are stored as 3 and counted in `bad`; a gene that has any is mean-imputed
// (k_impute_tiled_f64) and takes the fp64 path at flush.
// grid: (ceil(npad/16/256), M); src rows are 4-byte aligned (pitch a multiple of 4).
__global__ void __launch_bounds__(256)
k_unpack_bed(const uint8_t* __restrict__ src, int64_t pitch, int64_t N, int8_t* __restrict__ dst, int M,
             RowCounts* __restrict__ counts) {
  const int64_t row = blockIdx.y;
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  const int64_t npad = (N + 127) & ~(int64_t)127;
  int n1 = 0, n2 = 0, bad = 0;
  if (i0 < npad) {
    uint32_t packed = 0;
    if (i0 < N) packed = *reinterpret_cast<const uint32_t*>(src + (size_t)row * pitch + (i0 >> 2));
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t x = (packed >> (8 * q)) & 0xFFu;
      const uint32_t t = (x | (x << 6) | (x << 12) | (x << 18)) & 0x03030303u;   // byte k = code of sample 4q+k
      const uint32_t b0 = t & 0x01010101u, b1 = (t >> 1) & 0x01010101u;
      const uint32_t miss = b0 & ~b1;
      uint32_t g = b1 + b0 + 2u * miss;                                         // 0, 3 (missing), 1, 2
      const int64_t rem = N - (i0 + 4 * q);
      const uint32_t vm = rem >= 4 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - (int)rem))));
      g &= vm;
      const uint32_t lo = g & 0x01010101u, hi = (g >> 1) & 0x01010101u;
      n1 += __popc(lo & ~hi);
      n2 += __popc(hi & ~lo);
      bad += __popc(hi & lo);
      w[q] = g;
    }
    *reinterpret_cast<uint4*>(dst + tiled_off(M, (int)row, i0)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    n1 += __shfl_xor_sync(0xffffffffu, n1, o);
    n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  // one global atomic per CTA and counter (a warp each made 984 warps of a row queue on three addresses: the launch list
  // showed 60 us per 25 MB gene, ten times its HBM time)
  __shared__ int s_cnt[3];
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && (n1 | n2 | bad)) {
    atomicAdd(&s_cnt[0], n1);
    atomicAdd(&s_cnt[1], n2);
    atomicAdd(&s_cnt[2], bad);
  }
  __syncthreads();
  if (threadIdx.x == 0 && (s_cnt[0] | s_cnt[1] | s_cnt[2])) {
    atomicAdd(&counts[row].n1, s_cnt[0]);
    atomicAdd(&counts[row].n2, s_cnt[1]);
    atomicAdd(&counts[row].bad, s_cnt[2]);
  }
}

// DataConsolidator::imputeGenotypeToMean (src/DataConsolidator.cpp:217-245) for a gene that arrived
// as hard calls with missing entries (code 3): column j gets g = 2 * ac / an over its non-missing
// calls (p = 0 when nothing is called).  dst: N x M column-major doubles, the layout the fp64 path eats.
// grid: (ceil(N/256), M)
__global__ void __launch_bounds__(256)
k_impute_tiled_f64(const int8_t* __restrict__ src, int M, int64_t N, const RowCounts* __restrict__ counts,
                   double* __restrict__ dst) {
  const int row = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const long long ac = (long long)counts[row].n1 + 2ll * counts[row].n2;
  const long long an = 2ll * (N - counts[row].bad);
  const double fill = an == 0 ? 0.0 : 2.0 * (1.0 * (double)ac / (double)an);
  const int g = src[tiled_off(M, row, i)];
  dst[(size_t)row * N + i] = (g == 3) ? fill : (double)g;
}

// grid: (ceil(N/16/256), rows).  Tiled blocks of equal M back to back -> plain [rows][N] (tests).
__global__ void __launch_bounds__(256)
k_untile(const int8_t* __restrict__ arena, int M, int64_t gene_bytes, int64_t row_first, int64_t N, int8_t* __restrict__ out) {
  const int64_t lr = blockIdx.y, row = row_first + lr;
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i0 >= N) return;
  const int64_t gene = row / M;
  const int r = (int)(row - gene * M);
  const int8_t* p = arena + (size_t)gene * gene_bytes + tiled_off(M, r, i0);
  for (int s = 0; s < 16 && i0 + s < N; ++s) out[(size_t)lr * N + i0 + s] = p[s];
}

// one thread per row
__global__ void k_flags_from_counts(int64_t n_rows, int64_t N, const RowCounts* __restrict__ counts,
                                    uint8_t* __restrict__ flags, double* __restrict__ af_out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const long long n1 = counts[r].n1, n2 = counts[r].n2, n0 = N - n1 - n2;
  const long long c = n1 + 2 * n2;
  uint8_t f = (c > N) ? kRowFlipped : kRowNormal;
  if (n0 == N || n1 == N || n2 == N) f = kRowSkip;
  flags[r] = f;
  if (af_out) af_out[r] = 0.5 * (double)c / (double)N;
}

}  // namespace rvt
