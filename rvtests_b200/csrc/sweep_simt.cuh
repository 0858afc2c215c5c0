// sweep_simt.cuh -- K1, CUDA-core (dp4a) version of the single pass over one gene's genotypes.
//
// One work unit = (gene, split): a contiguous range of samples of one gene.  For its range the
// CTA accumulates, in exact integer arithmetic,
//   d[i][j]  = sum_s G[i][s] * T[j][s]      i < 64 gene rows, T = gene rows followed by E rows
//            -> G'G (regression/Skat.cpp:47,76 K_sqrt P0 K_sqrt' needs G'VG),
//               G'r (Skat.cpp:50-52), G'X (Skat.cpp:63-66 X'V G), column sums
//   coll[]   = dot products of the per-sample burden scores with the E rows
//            -> cmcCollapse / zegginiCollapse (src/Model.cpp:73-89,115-130) followed by
//               U = S'r, SS = S'S, SZ = S'Z (regression/LinearRegressionScoreTest.cpp:209-217)
// This kernel is the reference implementation of the integer pipeline on the GPU: it handles any
// alignment/shape the C ABI accepts, and the tcgen05 kernel (sweep_tc.cuh) must reproduce its
// SweepPartial bit for bit.  Roofline class: HBM sweep (1 byte per genotype), but dp4a-bound in
// practice (64 x NC int8 MACs per sample on CUDA cores) -- see DESIGN.md section 4.
#pragma once
#include "common.cuh"

namespace rvt {

constexpr int kSimtThreads = 256;
constexpr int kSimtKT = 256;                 // samples per smem tile
constexpr int kSimtPitch = kSimtKT + 16;     // bytes; 68 words == 4 mod 32 -> conflict-free LDS.128
constexpr int kSimtRows = kTileRows + kMaxER;
constexpr int kSimtSmem = 2 * kSimtRows * kSimtPitch;

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// collapse indicator of 4 packed genotypes: (g>0) for a normal row, (g<2) for a flipped row
// (g' = 2-g > 0), 0 for a skipped (monomorphic) row.  xf/mn/en are per-row byte masks.
__device__ __forceinline__ uint32_t collapse_ind(uint32_t w, uint32_t xf, uint32_t mn, uint32_t en) {
  return (((w >> 1) ^ xf) | (w & mn)) & en;
}

__global__ void __launch_bounds__(kSimtThreads)
k_sweep_simt(const GeneDesc* __restrict__ genes, int n_genes, const uint8_t* __restrict__ rowflags,
             const NullModel* __restrict__ nm, int S, int64_t chunk, SweepPartial* __restrict__ out,
             unsigned int* __restrict__ counter) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ int s_unit;
  __shared__ uint32_t s_xf[kTileRows], s_mn[kTileRows], s_en[kTileRows];
  __shared__ unsigned long long s_coll[kCollapseN];

  const int tid = threadIdx.x;
  const int64_t N = nm->N;
  const int ER = nm->ER;
  const int NC = kTileRows + ER;
  const int8_t* __restrict__ E = nm->E;
  const int64_t ldE = nm->ldE;
  const int n_units = n_genes * S;
  const int ti = tid >> 4, tj = tid & 15;

  for (;;) {
    if (tid == 0) s_unit = (int)atomicAdd(counter, 1u);
    __syncthreads();
    const int u = s_unit;
    if (u >= n_units) break;
    const int gi = u / S, sp = u - gi * S;
    const GeneDesc gd = genes[gi];
    const int M = gd.M;
    const int64_t k0 = (int64_t)sp * chunk;
    int64_t k1 = k0 + chunk;
    if (k1 > N) k1 = N;
    if (tid < kTileRows) {
      uint8_t f = (tid < M) ? rowflags[gd.var0 + tid] : (uint8_t)kRowSkip;
      s_xf[tid] = (f == kRowFlipped) ? 0x01010101u : 0u;
      s_mn[tid] = (f == kRowNormal) ? 0x01010101u : 0u;
      s_en[tid] = (f == kRowSkip) ? 0u : 0x01010101u;
    }
    if (tid < kCollapseN) s_coll[tid] = 0ull;

    int acc[4][6];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) acc[i][j] = 0;
    int cz[kMaxER + 1], cc[kMaxER + 1];
#pragma unroll
    for (int e = 0; e <= kMaxER; ++e) cz[e] = cc[e] = 0;

    const int ntiles = (k1 > k0) ? (int)((k1 - k0 + kSimtKT - 1) / kSimtKT) : 0;
    auto issue_tile = [&](int t, int buf) {
      const int64_t kb = k0 + (int64_t)t * kSimtKT;
      uint8_t* base = smem + (size_t)buf * kSimtRows * kSimtPitch;
      const int nchunks = NC * (kSimtKT / 16);
      for (int c = tid; c < nchunks; c += kSimtThreads) {
        const int row = c >> 4, kc = c & 15;
        const int64_t k = kb + kc * 16;
        int64_t rem = k1 - k;
        int nb = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
        const int8_t* src;
        if (row < kTileRows) {
          if (row >= M) nb = 0;
          src = geno_ptr(gd, row < M ? row : 0, nb ? k : 0);
        } else {
          src = E + (size_t)(row - kTileRows) * ldE + (nb ? k : 0);
        }
        cp_async16_zfill(base + (size_t)row * kSimtPitch + kc * 16, src, nb);
      }
      cp_async_commit();
    };
    if (ntiles > 0) issue_tile(0, 0);
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      if (t + 1 < ntiles) {
        issue_tile(t + 1, buf ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const uint8_t* base = smem + (size_t)buf * kSimtRows * kSimtPitch;
      // ---- integer Gram: rows 4*ti..4*ti+3 x columns tj+16*jj
      if (4 * ti < M) {
#pragma unroll 2
        for (int kq = 0; kq < kSimtKT / 16; ++kq) {
          uint4 a[4], b[6];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            a[i] = *reinterpret_cast<const uint4*>(base + (size_t)(4 * ti + i) * kSimtPitch + kq * 16);
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            const int col = tj + 16 * j;
            if (col < NC) b[j] = *reinterpret_cast<const uint4*>(base + (size_t)col * kSimtPitch + kq * 16);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) {
              const int col = tj + 16 * j;
              if (col < NC) {
                int v = acc[i][j];
                v = __dp4a((int)a[i].x, (int)b[j].x, v);
                v = __dp4a((int)a[i].y, (int)b[j].y, v);
                v = __dp4a((int)a[i].z, (int)b[j].z, v);
                v = __dp4a((int)a[i].w, (int)b[j].w, v);
                acc[i][j] = v;
              }
            }
        }
      }
      // ---- burden collapse: one 4-sample word column per thread
      if (tid < kSimtKT / 4) {
        const int64_t ks = k0 + (int64_t)t * kSimtKT + 4 * tid;
        uint32_t z = 0;
        for (int r = 0; r < M; ++r) {
          uint32_t w = *reinterpret_cast<const uint32_t*>(base + (size_t)r * kSimtPitch + 4 * tid);
          z += collapse_ind(w, s_xf[r], s_mn[r], s_en[r]);
        }
        // samples at/after k1 are zero-filled but a flipped row would count them: mask
        int64_t rem = k1 - ks;
        uint32_t vm = rem >= 4 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - (int)rem))));
        z &= vm;
        uint32_t c = ((z + 0x7F7F7F7Fu) >> 7) & 0x01010101u;
#pragma unroll
        for (int e = 0; e < kMaxER; ++e)
          if (e < ER) {
            int ew = *reinterpret_cast<const int*>(base + (size_t)(kTileRows + e) * kSimtPitch + 4 * tid);
            cz[e] = __dp4a((int)z, ew, cz[e]);
            cc[e] = __dp4a((int)c, ew, cc[e]);
          }
        cz[kMaxER] = __dp4a((int)z, (int)z, cz[kMaxER]);
        cc[kMaxER] = __dp4a((int)c, (int)c, cc[kMaxER]);
      }
      __syncthreads();
    }
    // ---- write the unit's partial
    SweepPartial* o = out + u;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int col = tj + 16 * j;
        if (col < NC) o->d[4 * ti + i][col] = (4 * ti < M) ? acc[i][j] : 0;
      }
    if (tid < 64) {  // two full warps hold the collapse accumulators
#pragma unroll
      for (int e = 0; e <= kMaxER; ++e) {
        long long a = cz[e], b = cc[e];
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
          a += __shfl_xor_sync(0xffffffffu, a, o2);
          b += __shfl_xor_sync(0xffffffffu, b, o2);
        }
        if ((tid & 31) == 0 && (e < ER || e == kMaxER)) {
          const int slot = (e == kMaxER) ? ER : e;
          atomicAdd(&s_coll[slot], (unsigned long long)a);
          atomicAdd(&s_coll[(ER + 1) + slot], (unsigned long long)b);
        }
      }
    }
    __syncthreads();
    if (tid < 2 * (ER + 1)) o->coll[tid] = (long long)s_coll[tid];
    __syncthreads();
  }
}

}  // namespace rvt
