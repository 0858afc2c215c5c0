// sweep_aug.cuh -- K1 for genes WITH MISSING CALLS: the integer tensor-core sweep on an augmented tile.
//
// DataConsolidator::imputeGenotypeToMean (src/DataConsolidator.cpp:217-245) replaces a missing call of variant j by the mean
// mu_j of the observed calls, so the imputed genotype matrix is   G = H + M diag(delta),   H = the hard calls with a FILL
// value f_j in {0, 2} at the missing entries, M = the 0/1 missing indicators, delta_j = mu_j - f_j.  Every statistic the
// tests need is then an exact integer sum over the augmented rows [H ; M]:
//     G'G = H'H + H'M D + D M'H + D M'M D,    G'r = H'r + D M'r,    G'X = H'X + D M'X          (D = diag(delta))
// and the burden collapses use `(int)g > 0` (src/Model.cpp:82-83): an imputed value (mean of a minor-coded column < 1)
// never counts, which the fill reproduces -- 0 in a normal column, 2 in a flipped one (g' = 2 - g = 0).
// Round 1 sent such genes to a sparse CUDA-core kernel (k_tile_sparse: ~80 us per gene at 500 000 x 50, 20 x the HBM time);
// real cohorts always have missing calls.
//
// Same machinery as sweep_tc.cuh (TMA -> mbarrier ring -> tcgen05.mma.kind::i8 -> TMEM -> tcgen05.ld) with
//   * ONE genotype tile per box from HBM (bytes 0/1/2, 3 = missing): the traffic of the hard-call sweep;
//   * the consumer warps, which read every word for the collapse anyway, write the indicator tile M next to it in shared
//     memory, patch the missing bytes of H to their fill, and add the two burden rows (ZC scheme);
//   * one UMMA of M = 128 (rows [H ; M]) x N = 128 + ER (columns [H ; M ; E]) per 32-sample slice;
//   * three SweepPartials per (gene, split): {H'H, H'E, burden sums}, {H'M}, {M'M, M'E}.
// 2 boxes per stage (36 KB): 6 stages of ring for ER = 16.  Needs M <= 62 (two spare rows), like ZC.
#pragma once
#include "sweep_tc.cuh"

namespace rvt {

template <int ER, int STAGES>
struct AugCfg {
  static constexpr int kBoxes = 2;
  static constexpr int kStageK = kBoxes * kTcBoxK;
  static constexpr int kHOff = 0;
  static constexpr int kMOff = kTileRows * 128;
  static constexpr int kEOff = 2 * kTileRows * 128;
  static constexpr int kBoxBytes = kEOff + ER * 128;
  static constexpr int kStageBytes = kBoxes * kBoxBytes;
  static constexpr int kSmem = STAGES * kStageBytes + 1024 /*align*/ + 1024 /*barriers*/;
  static constexpr int kN = 2 * kTileRows + ER;   // UMMA N: columns [H ; M ; E]
  static constexpr int kAccCols = 256;            // TMEM columns between the two accumulators
  static constexpr int kTmemCols = 512;
};

constexpr int kAugParts = 3;   // SweepPartials per (gene, split)

template <int ER, int STAGES>
__global__ void __launch_bounds__(kTcThreads, 1)
k_sweep_aug(const CUtensorMap* __restrict__ maps_g, const __grid_constant__ CUtensorMap map_e, const GeneDesc* __restrict__ genes,
            int n_genes, const uint8_t* __restrict__ rowflags, int64_t N, int S, int64_t chunk, SweepPartial* __restrict__ out) {
  using Cfg = AugCfg<ER, STAGES>;
  constexpr int kBoxes = Cfg::kBoxes;
  constexpr int kGroups = kTcConsumerWarps / kBoxes;   // 4 consumer groups, one stage each in turn
  constexpr int kPeriod = tc_lcm(STAGES, kGroups);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + STAGES * Cfg::kStageBytes);
  uint64_t* full = bars;                          // [kGroups][STAGES]  (one per (group, stage): see sweep_tc.cuh)
  uint64_t* empty = full + kGroups * STAGES;      // [STAGES]  the UMMAs have read the stage
  uint64_t* ready = empty + STAGES;               // [STAGES]  the consumers have written M, the fills and the burden rows
  uint64_t* tfull = ready + STAGES;               // [2]
  uint64_t* tempty = tfull + 2;                   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  static_assert((kGroups * STAGES + 2 * STAGES + 4) * 8 + 16 <= 1024, "barrier area");
  auto full_bar = [&](uint32_t it_) { return &full[(it_ % kGroups) * STAGES + (it_ % STAGES)]; };
  auto full_ph = [&](uint32_t it_) { return (it_ / (uint32_t)kPeriod) & 1u; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = n_genes * S;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < kGroups * STAGES; ++s) mbar_init(&full[s], 1);
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&empty[s], 1);
        mbar_init(&ready[s], kBoxes);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull[a], 1);
        mbar_init(&tempty[a], kTcConsumerWarps);
      }
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto unit_steps = [&](int u, int64_t* k0_out, int64_t* k1_out) {
    const int gi = u / S, sp = u - gi * S;
    const int64_t k0 = (int64_t)sp * chunk;
    int64_t k1 = k0 + chunk;
    if (k1 > N) k1 = N;
    *k0_out = k0;
    *k1_out = k1;
    return (k1 > k0) ? (int)((k1 - k0 + Cfg::kStageK - 1) / Cfg::kStageK) : 0;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t it = 0;
    const int nchunks = (int)((N + 127) >> 7);
    constexpr int kOobRow = 0x7FFF0000;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int gi = u / S;
      const int row0 = (int)genes[gi].row0, Mg = genes[gi].M;
      const CUtensorMap* mg = maps_g + (Mg - 1);
      const uint32_t stage_tx = (uint32_t)(kBoxes * (Mg + ER) * 128);
      int64_t k0, k1;
      const int nsteps = unit_steps(u, &k0, &k1);
      for (int ks = 0; ks < nsteps; ++ks, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        __syncwarp();
        if (elect_one_sync()) {
          uint64_t* fb = full_bar(it);
          mbar_expect_tx(fb, stage_tx);
          uint8_t* st = tiles + (size_t)s * Cfg::kStageBytes;
          const int kb = (int)(k0 + (int64_t)ks * Cfg::kStageK);
#pragma unroll
          for (int b = 0; b < kBoxes; ++b) {
            const int ch = (kb >> 7) + b;
            tma_load_2d(st + b * Cfg::kBoxBytes + Cfg::kHOff, mg, 0, ch < nchunks ? row0 + ch * Mg : kOobRow, fb, kEvictFirst);
            tma_load_2d(st + b * Cfg::kBoxBytes + Cfg::kEOff, &map_e, kb + b * kTcBoxK, 0, fb, kEvictLast);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // D = s32, A = B = s8 K-major; M = 128 rows [H ; M], N = 128 + ER columns [H ; M ; E]
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cfg::kN >> 3) << 17) | ((uint32_t)((2 * kTileRows) >> 4) << 24);
    uint32_t it = 0, ui = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ui) {
      int64_t k0, k1;
      const int nsteps = unit_steps(u, &k0, &k1);
      const int a = ui & 1;
      mbar_wait(&tempty[a], ((ui >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(a * Cfg::kAccCols);
      for (int ks = 0; ks < nsteps; ++ks, ++it) {
        const int s = it % STAGES;
        mbar_wait(full_bar(it), full_ph(it));
        mbar_wait(&ready[s], (it / STAGES) & 1);
        tc_fence_after();
        __syncwarp();
        if (elect_one_sync()) {
          const uint32_t st = smem_u32(tiles + (size_t)s * Cfg::kStageBytes);
#pragma unroll
          for (int b = 0; b < kBoxes; ++b) {
            const uint64_t d0 = umma_desc_sw128(st + b * Cfg::kBoxBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_i8(tmem_d, d0 + (uint64_t)(2 * k), d0 + (uint64_t)(2 * k), idesc, (ks | b | k) ? 1u : 0u);
          }
          umma_commit(&empty[s]);
          if (ks == nsteps - 1) umma_commit(&tfull[a]);
        }
        __syncwarp();
      }
      if (nsteps == 0 && elect_one_sync()) umma_commit(&tfull[a]);
      __syncwarp();
    }
  } else {
    // ===================== consumers: indicator tile + fills + burden rows, epilogue =====================
    const int cw = (warp - 2) % kBoxes;    // box of this warp inside its stages
    const int grp = (warp - 2) / kBoxes;   // stages with it % kGroups == grp
    const int egrp = (warp - 2) >> 2;      // epilogue: which of the quadrant's two warps
    const int q = warp & 3;                // TMEM lane quadrant this warp may read
    uint32_t it = 0, ui = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ui) {
      const int gi = u / S;
      const GeneDesc gd = genes[gi];
      const int M = gd.M;
      int64_t k0, k1;
      const int nsteps = unit_steps(u, &k0, &k1);
      const uint8_t f0 = (lane < M) ? rowflags[gd.var0 + lane] : (uint8_t)kRowSkip;
      const uint8_t f1 = (lane + 32 < M) ? rowflags[gd.var0 + lane + 32] : (uint8_t)kRowSkip;
      const unsigned long long fmask = (unsigned long long)__ballot_sync(0xffffffffu, f0 == kRowFlipped) |
                                       ((unsigned long long)__ballot_sync(0xffffffffu, f1 == kRowFlipped) << 32);
      const unsigned long long emask = (unsigned long long)__ballot_sync(0xffffffffu, f0 != kRowSkip) |
                                       ((unsigned long long)__ballot_sync(0xffffffffu, f1 != kRowSkip) << 32);
      const int Mr8 = (M + 7) & ~7;
      uint32_t woff[8];   // lane-constant swizzled offsets of word `lane` in rows j = 0..7 of an 8-row group
#pragma unroll
      for (int j = 0; j < 8; ++j) woff[j] = (uint32_t)(j * 128 + ((((lane >> 2) ^ j) << 4) | ((lane & 3) << 2)));
      for (int ks = 0; ks < nsteps; ++ks, ++it) {
        if ((int)(it % (uint32_t)kGroups) != grp) continue;
        const int s = it % STAGES;
        mbar_wait(full_bar(it), full_ph(it));
        uint8_t* hbox = tiles + (size_t)s * Cfg::kStageBytes + cw * Cfg::kBoxBytes + Cfg::kHOff;
        uint8_t* mbox = hbox + Cfg::kMOff;
        const int64_t ksamp = k0 + (int64_t)ks * Cfg::kStageK + cw * kTcBoxK + 4 * lane;
        uint32_t z = 0;
        // eight rows per pass: the eight loads are issued before anything depends on them (a row-at-a-time loop ran
        // LDS -> ALU -> STS serially, ~100 clk per row, and made the consumers -- not the UMMAs -- the limit).  Rows >= M of
        // the last pass are disabled in `emask`; their stores land in rows nothing reads (the burden rows M, M+1 are
        // written after this loop).
        for (int r0 = 0; r0 < Mr8; r0 += 8) {
          uint32_t w[8];
          const uint32_t gbase = (uint32_t)(r0 >> 3) * 1024u;
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = *reinterpret_cast<const uint32_t*>(hbox + gbase + woff[j]);
          const uint32_t fb = (uint32_t)(fmask >> r0) & 0xFFu, eb = (uint32_t)(emask >> r0) & 0xFFu;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t m = w[j] & (w[j] >> 1) & 0x01010101u;        // code 3 = missing
            const uint32_t xf = ((fb >> j) & 1u) * 0x01010101u;
            const uint32_t en = ((eb >> j) & 1u) * 0x01010101u;
            // missing bytes: 3 -> fill (0, or 2 in a flipped row)
            const uint32_t h = (w[j] & ~(m | (m << 1))) | ((m << 1) & (xf << 1));
            *reinterpret_cast<uint32_t*>(mbox + gbase + woff[j]) = m;
            if (m) *reinterpret_cast<uint32_t*>(hbox + gbase + woff[j]) = h;
            z += collapse_ind(h, xf, en & ~xf, en);
          }
        }
        // samples at/after k1 are zero-filled but a flipped row would count them: mask
        const int64_t rem = k1 - ksamp;
        const uint32_t vm = rem >= 4 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - (int)rem))));
        z &= vm;
        const uint32_t c = ((z + 0x7F7F7F7Fu) >> 7) & 0x01010101u;
        *reinterpret_cast<uint32_t*>(hbox + sw128_word_off(M, lane)) = z;
        *reinterpret_cast<uint32_t*>(hbox + sw128_word_off(M + 1, lane)) = c;
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[s]);
      }
      // ---- epilogue: accumulator of this unit -> three SweepPartials
      const int a = ui & 1;
      mbar_wait(&tfull[a], (ui >> 1) & 1);
      tc_fence_after();
      SweepPartial* o = out + (size_t)kAugParts * u;
      const uint32_t taddr = tmem_base + (uint32_t)(a * Cfg::kAccCols) + ((uint32_t)(q * 32) << 16);
      const bool hrow = q < 2;                       // lanes 0..63 = rows of H, 64..127 = rows of M
      const int row = 32 * (q & 1) + lane;           // row inside its half
      constexpr int nG = (2 * kTileRows + ER) / 16;  // column groups of 16
#pragma unroll
      for (int g = 0; g < nG; ++g) {
        if ((g & 1) != egrp) continue;               // the quadrant's two warps alternate column groups
        if (!hrow && g < 4) continue;                // M'H is the transpose of H'M
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)(16 * g), v);
        tmem_ld_wait();
        SweepPartial* op;
        int dcol;
        if (g < 4) {
          op = o;                 // H'H
          dcol = 16 * g;
        } else if (g < 8) {
          op = hrow ? o + 1 : o + 2;   // H'M / M'M
          dcol = 16 * (g - 4);
        } else {
          op = hrow ? o : o + 2;       // H'E / M'E
          dcol = kTileRows + 16 * (g - 8);
        }
        if (row < M) {
          int4* dst = reinterpret_cast<int4*>(&op->d[row][dcol]);
          dst[0] = make_int4((int)v[0], (int)v[1], (int)v[2], (int)v[3]);
          dst[1] = make_int4((int)v[4], (int)v[5], (int)v[6], (int)v[7]);
          dst[2] = make_int4((int)v[8], (int)v[9], (int)v[10], (int)v[11]);
          dst[3] = make_int4((int)v[12], (int)v[13], (int)v[14], (int)v[15]);
        }
        // rows M / M+1 of H are the Zeggini / CMC burden rows: their digit columns and their own diagonal entry
        if (hrow && (row == M || row == M + 1)) {
          const int base = (row == M) ? 0 : (ER + 1);
          if (g >= 8) {
#pragma unroll
            for (int i = 0; i < 16; ++i) o->coll[base + 16 * (g - 8) + i] = nsteps ? (long long)(int)v[i] : 0ll;
          } else if (g < 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (16 * g + i == row) o->coll[base + ER] = nsteps ? (long long)(int)v[i] : 0ll;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(Cfg::kTmemCols));
  }
}

// ---- augmented sums -> the fp64 statistics of the imputed matrix (DosageStats, dosage.cuh), one CTA per gene ---------
// csum / cmin / cmax come from k_tile_cols (counts alone); flags[] are the ones the sweep used.
struct AugGene {
  int32_t M, slot;        // slot: index into the DosageStats array of this flush
  int64_t var0;
};

__global__ void __launch_bounds__(128)
k_aug_stats(const AugGene* __restrict__ genes, int n_genes, int S, const SweepPartial* __restrict__ parts, const uint8_t* __restrict__ rowflags,
            const RowCounts* __restrict__ counts, const NullModel* __restrict__ nm, DosageStats* __restrict__ stats) {
  __shared__ double s_delta[kTileRows];
  const int gi = blockIdx.x, tid = threadIdx.x;
  if (gi >= n_genes) return;
  const AugGene ag = genes[gi];
  const int M = ag.M;
  const int64_t N = nm->N;
  const int C = nm->C, ER = nm->ER;
  DosageStats* __restrict__ st = stats + ag.slot;
  const SweepPartial* __restrict__ gp = parts + (size_t)gi * S * kAugParts;
  if (tid < kTileRows) {
    double d = 0.0;
    if (tid < M) {
      const RowCounts rc = counts[ag.var0 + tid];
      const long long nobs = N - rc.bad;
      const double ac = (double)((long long)rc.n1 + 2ll * rc.n2);
      const double fill = nobs > 0 ? 2.0 * (ac / (double)(2 * nobs)) : 0.0;   // imputeGenotypeToMean: 2 p^ (as k_tile_cols)
      d = fill - ((rowflags[ag.var0 + tid] == kRowFlipped) ? 2.0 : 0.0);
    }
    s_delta[tid] = d;
  }
  __syncthreads();
  for (int idx = tid; idx < M * M; idx += 128) {
    const int a = idx / M, b = idx - a * M;
    long long hh = 0, hm = 0, mh = 0, mm = 0;
    for (int sp = 0; sp < S; ++sp) {
      const SweepPartial* p = gp + (size_t)sp * kAugParts;
      hh += p[0].d[a][b];
      hm += p[1].d[a][b];
      mh += p[1].d[b][a];
      mm += p[2].d[a][b];
    }
    st->A[a][b] = (double)hh + s_delta[b] * (double)hm + s_delta[a] * (double)mh + s_delta[a] * s_delta[b] * (double)mm;
  }
  if (tid < M) {
    const int a = tid;
    for (int v = 0; v <= C; ++v) {   // v = 0: the null residual, v = 1..C: the covariate columns
      long long dh[4] = {0, 0, 0, 0}, dm[4] = {0, 0, 0, 0};
      for (int sp = 0; sp < S; ++sp) {
        const SweepPartial* p = gp + (size_t)sp * kAugParts;
        for (int k = 0; k < 4; ++k) {
          dh[k] += p[0].d[a][kTileRows + 4 * v + k];
          dm[k] += p[2].d[a][kTileRows + 4 * v + k];
        }
      }
      const long long hv = dh[0] + (dh[1] << 8) + (dh[2] << 16) + (dh[3] << 24);
      const long long mv = dm[0] + (dm[1] << 8) + (dm[2] << 16) + (dm[3] << 24);
      const double val = ((double)hv + s_delta[a] * (double)mv) * nm->scale[v];
      if (v == 0)
        st->s[a] = val;
      else
        st->B[a][v - 1] = val;
    }
    st->cw[a] = st->csum[a];
  }
  if (tid == 0) {
    // burden sums from the collapse digits (as k_finalize step 4b)
    long long cl[2][kMaxER + 1];
    for (int w = 0; w < 2; ++w)
      for (int e = 0; e <= ER; ++e) {
        long long s = 0;
        for (int sp = 0; sp < S; ++sp) s += gp[(size_t)sp * kAugParts].coll[w * (ER + 1) + e];
        cl[w][e] = s;
      }
    auto rec = [&](const long long* d) { return d[0] + (d[1] << 8) + (d[2] << 16) + (d[3] << 24); };
    st->zegU = (double)rec(&cl[0][0]) * nm->scale[0];
    st->zegSS = (double)cl[0][ER];
    st->cmcU = (double)rec(&cl[1][0]) * nm->scale[0];
    st->cmcSS = (double)cl[1][ER];
    for (int l = 0; l < C; ++l) {
      st->zegSZ[l] = (double)rec(&cl[0][4 * (l + 1)]) * nm->scale[l + 1];
      st->cmcSZ[l] = (double)rec(&cl[1][4 * (l + 1)]) * nm->scale[l + 1];
    }
    st->nonref = (double)cl[1][ER];
  }
}

// flags of a gene with missing calls, from the statistics k_tile_cols derived from the counts: the same decisions
// dosage_prepare takes (csum > N flips, cmin == cmax drops), so that the sweep's collapse and the tail agree
__global__ void k_aug_flags(const AugGene* __restrict__ genes, int n_genes, int64_t N, const DosageStats* __restrict__ stats,
                            uint8_t* __restrict__ rowflags) {
  const int gi = blockIdx.x, j = threadIdx.x;
  if (gi >= n_genes) return;
  const AugGene ag = genes[gi];
  if (j >= ag.M) return;
  const DosageStats* st = stats + ag.slot;
  uint8_t f = (st->csum[j] > (double)N) ? kRowFlipped : kRowNormal;
  if (st->cmin[j] == st->cmax[j]) f = kRowSkip;
  rowflags[ag.var0 + j] = f;
}

inline int aug_init(char* err, size_t errlen) {
  cudaError_t e = cudaFuncSetAttribute(k_sweep_aug<16, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, AugCfg<16, 6>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_aug<32, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, AugCfg<32, 5>::kSmem);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "cudaFuncSetAttribute(k_sweep_aug): %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

// genes: device array of the descriptors of the genes to sweep (all in segment `seg`, tiled, M <= 62)
inline int aug_launch(TcSegments* tc, int seg, const GeneDesc* d_genes, const GeneDesc* h_genes, int n, const uint8_t* d_flags, int64_t N,
                      int ER, int S, int64_t chunk, SweepPartial* d_parts, int sm_count, cudaStream_t st, char* err, size_t errlen) {
  int rc = tc_prepare_maps(tc, seg, h_genes, n, st, err, errlen);
  if (rc) return rc;
  const int grid = std::min(n * S, sm_count);
  if (ER == 16)
    k_sweep_aug<16, 6><<<grid, kTcThreads, AugCfg<16, 6>::kSmem, st>>>(tc->seg[seg].d_maps, tc->map_e, d_genes, n, d_flags, N, S, chunk, d_parts);
  else
    k_sweep_aug<32, 5><<<grid, kTcThreads, AugCfg<32, 5>::kSmem, st>>>(tc->seg[seg].d_maps, tc->map_e, d_genes, n, d_flags, N, S, chunk, d_parts);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, errlen, "k_sweep_aug launch: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

}  // namespace rvt
