// sweep_tc.cuh -- K1 on the 5th-generation tensor cores: TMA -> SMEM -> tcgen05.mma (kind::i8,
// s32 accumulators in TMEM) with the burden collapse riding on the same SMEM tiles.
//
// Same contract as sweep_simt.cuh (one SweepPartial per (gene, split), bit-identical numbers):
//   D[64 x (64+ER)] += T_gene[64 x K] * [T_gene ; E]^T        regression/Skat.cpp:47-76 G'VG, G'r,
//                                                             X'VG; LinearRegressionScoreTest.cpp:209-217
// Why kind::i8 and not kind::f16: genotypes are 0/1/2 and the null-model vectors are carried as
// base-256 digits (common.cuh), so the contraction is an exact integer GEMM.  sm_100a has int8
// UMMA at twice the bf16 rate, the operands need no conversion pass (TMA lands them ready to use)
// and the s32 accumulation is exact -- results do not depend on tiling or split order.
//
// CTA = 6 warps, persistent over units u = blockIdx.x + i*gridDim.x (static: units cost the same):
//   warp 0      TMA producer: per stage 4 boxes of 128 samples, each box = gene tile M x 128 B
//               (SWIZZLE_128B; ONE contiguous M*128-byte run of HBM thanks to the tiled layout)
//               followed at a fixed offset by the E tile ER x 128 B, so that the B operand of one
//               UMMA is the contiguous (64+ER)-row tile.  OOB rows/samples are zero-filled.
//   warp 1      MMA issuer: 4 UMMAs (K = 32 bytes) per box; tcgen05.commit frees the stage and,
//               after the unit's last stage, publishes the TMEM accumulator (double-buffered).
//   warps 2..9  consumers, two groups of four: (a) burden collapse -- group g takes the stages with
//               (it & 1) == g, warp (w & 3) the box of that number, straight from the swizzled SMEM
//               tile (the collapse, not HBM, was the limiter with four warps: see profiles/);
//               (b) epilogue: tcgen05.ld the accumulator of the finished unit (UMMA M=64 layout:
//               row m at lane (m%16)+32*(m/16); the two warps of a TMEM lane quadrant split the
//               column groups) and store the SweepPartial.
// Roofline class: HBM sweep, 1 byte per genotype: the tensor pipe needs
// 128*(64+ER)/256 = 40 cycles per 32-sample slice, i.e. ~64 B/clk/SM, ~3x what HBM can deliver.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "sweep_simt.cuh"

namespace rvt {

constexpr int kTcConsumerWarps = 8;             // two groups of 4: group g collapses the stages with (it & 1) == g
constexpr int kTcThreads = 64 + 32 * kTcConsumerWarps;
constexpr int kTcMaxStages = 5;
constexpr int kTcBoxK = 128;                // samples per box (one 128-byte swizzle row)
constexpr int kTcChunkAlign = 512;          // split boundaries are multiples of this (>= any stage width)

// PAIR units are walked SPLIT-MAJOR (all pairs of sample chunk 0, then chunk 1, ...): a tile takes part in up to 2 W pairs
// of the band (W partner tiles), and with the pairs of one chunk adjacent in time its chunk (64 x N/S bytes) stays in the
// 126 MB L2 between its uses instead of being re-read from HBM for every partner.
// PAIR = the A operand (rows of block I) and the B operand (rows of block J) are different tiles:
// the cross-block Gram of the meta-analysis covariance (src/Model.cpp:534-554 calculateXX for
// every pair of variants in the sliding window).  Box = [A tile][B tile][E tile].
// WIDE = one UMMA covers TWO boxes (256 samples x K=32 slices of each): A = [G box0 ; G box1] (M = 128),
// B = [G box0 ; G box1 ; E box0 ; E box1] (N = 128 + 2 ER).  A tcgen05.mma.kind::i8 costs ~65 clk plus
// ~0.35 clk per column of N whatever M is (tools/umma_bench.cu, profiles/r01_umma_bench.txt): 76 clk for
// the 64 x 80 slice of one box but only 118 clk for the 128 x 160 slice of two, i.e. 59 clk per box
// slice -- below the ~72 clk per slice at which HBM delivers the bytes.  The cross-box blocks of D are
// garbage and ignored; lanes 0-63 hold the even boxes' sums, lanes 64-127 the odd boxes', written
// as two SweepPartials per unit (exact integers: the split of the sum is invisible downstream).
__host__ __device__ constexpr int tc_gcd(int a, int b) { return b == 0 ? a : tc_gcd(b, a % b); }
__host__ __device__ constexpr int tc_lcm(int a, int b) { return a / tc_gcd(a, b) * b; }

template <int ER, int STAGES, bool PAIR = false, int BOXES = 4, bool WIDE = false, bool ZC = false>
struct TcCfg {
  static_assert(!(PAIR && WIDE), "block pairs use the single-box UMMA");
  static_assert(!WIDE || (BOXES % 2 == 0), "WIDE pairs boxes");
  static constexpr int kBoxes = BOXES;          // boxes (of 128 samples) per pipeline stage
  static constexpr int kStageK = BOXES * kTcBoxK;
  static constexpr int kAOff = 0;
  static constexpr int kBOff = PAIR ? kTileRows * 128 : 0;
  static constexpr int kEOff = kBOff + kTileRows * 128;
  static constexpr int kBoxBytes = kEOff + ER * 128;
  static constexpr int kPairBytes = 2 * kBoxBytes;       // WIDE: [G b0][G b1][E b0][E b1]
  static constexpr int kStageBytes = BOXES * kBoxBytes;
  static constexpr int kSmem = STAGES * kStageBytes + 1024 /*align*/ + 1024 /*barriers*/;
  static constexpr int kNC = kTileRows + ER;
  static constexpr int kAccCols = WIDE ? 2 * kNC : 128;  // TMEM columns of one accumulator
  static constexpr int kTmemCols = WIDE ? 512 : 256;     // two accumulators, power of two
  // byte offsets of box b's genotype tile / digit tile inside a stage
  __host__ __device__ static constexpr int g_off(int b) {
    return WIDE ? (b >> 1) * kPairBytes + (b & 1) * kTileRows * 128 : b * kBoxBytes + kAOff;
  }
  __host__ __device__ static constexpr int e_off(int b) {
    return WIDE ? (b >> 1) * kPairBytes + 2 * kTileRows * 128 + (b & 1) * ER * 128 : b * kBoxBytes + kEOff;
  }
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// Device watchdog: a wait that does not complete within kMbarBudgetCycles (~2 s; a whole sweep launch is ~10 ms) can only
// be a protocol bug.  It ends the kernel with `trap` -- the next CUDA call of the host reports a launch failure -- instead
// of spinning until the box is killed ("no watchdog anywhere").  The whole wait is ONE opaque PTX
// block, as before: the spin stays a tight try_wait loop and the clock is read once per 4096 failed tries.
// -DRVT_MBAR_WATCHDOG=0 builds the plain spin (A/B timing of the watchdog's cost).
#ifndef RVT_MBAR_WATCHDOG
#define RVT_MBAR_WATCHDOG 1
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if RVT_MBAR_WATCHDOG
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      ".reg .u32 n, m;\n"
      ".reg .u64 t0, t1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "mov.u32 n, 0;\n"
      "mov.u64 t0, %%clock64;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "add.u32 n, n, 1;\n"
      "and.b32 m, n, 4095;\n"
      "setp.ne.u32 q, m, 0;\n"
      "@q bra WAIT_LOOP;\n"
      "mov.u64 t1, %%clock64;\n"
      "sub.u64 t1, t1, t0;\n"
      "setp.lt.u64 q, t1, 4000000000;\n"
      "@q bra WAIT_LOOP;\n"
      "trap;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
#else
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
#endif
}
// L2 eviction-priority policies (createpolicy encodings): genotypes stream through once,
// the null-model digits E are re-read by every gene of the batch.
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;
// elect.sync: exactly one lane of the (converged) warp gets `true`.  Unlike `lane == 0` the compiler
// KNOWS a single thread is active behind it, so tcgen05.mma / cp.async.bulk.tensor are emitted as bare
// UTCIMMA / UTMALDG instead of being wrapped in an ELECT..BRA.U.ANY loop with R2UR.BROADCAST per
// instruction (measured: ~110 instead of ~40 clk per 64x80x32 UMMA -- tools/umma_bench.cu).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n.reg .b32 %%rx;\n.reg .pred %%px;\nelect.sync %%rx|%%px, %2;\n@%%px mov.s32 %1, 1;\nmov.s32 %0, %%rx;\n}\n"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, version 1):
// LBO = 16 B (unused for a single swizzle row of K), SBO = 1024 B between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t hi = 64u | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

// byte offset of 4-byte word `w` (0..31) of row `r` inside a SWIZZLE_128B tile with 128-byte rows
__device__ __forceinline__ uint32_t sw128_word_off(int r, int w) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((w >> 2) ^ (r & 7)) << 4) | ((w & 3) << 2)));
}

// ZC = the burden scores ride on the tensor core: the consumer warps write the per-sample Zeggini
// count z and CMC indicator c of a box as two extra int8 ROWS (M and M+1) of the genotype tile, so
// that D[M][64+e] = z.E_e, D[M][M] = z.z (rows M+1 likewise for c) come out of the same UMMAs.  The
// consumers then never read the digit tile (16 LDS + 34 dp4a per lane and box less, a fifth of
// their shared-memory wavefronts); the price is an extra hop TMA -> consumers -> UMMA per stage.
// Needs M <= 62 and splits short enough for z.E_e to stay inside int32 (host checks both).
template <int ER, int kTcStages, bool PAIR, int kTcBoxes, bool WIDE = false, bool ZC = false>
__global__ void __launch_bounds__(kTcThreads, 1)
k_sweep_tc(const CUtensorMap* __restrict__ maps_g /* [64]: box rows = index+1 */,
           const CUtensorMap* __restrict__ maps_b /* PAIR: the maps of the segment the B tiles live in (may equal maps_g) */,
           const __grid_constant__ CUtensorMap map_e,
           const GeneDesc* __restrict__ genes, int n_genes, const uint8_t* __restrict__ rowflags, int64_t N, int S,
           int64_t chunk, SweepPartial* __restrict__ out, int dbg_skip /* timing experiments only: 1 no collapse,
           2 no MMA, 4 no E loads; results are then meaningless */,
           const RowCounts* __restrict__ counts /* nullable.  Gene sweeps: a gene whose rows hold values outside {0,1,2}
           (missing calls of a 2-bit push, counted at push time) is not swept here at all -- its units run zero stages; the
           statistics kernel reports it from the same counts and the augmented sweep (sweep_aug.cuh) computes it */) {
  using Cfg = TcCfg<ER, kTcStages, PAIR, kTcBoxes, WIDE, ZC>;
  static_assert(!(ZC && PAIR), "block pairs carry no burden scores");
  constexpr int kTcStageK = Cfg::kStageK;
  constexpr int kGroups = kTcConsumerWarps / kTcBoxes;   // consumer groups, one stage each in turn
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B (TMA destination and UMMA descriptors)
  uint8_t* tiles = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + kTcStages * Cfg::kStageBytes);
  // "stage s holds the data of iteration `it`": ONE BARRIER PER (consumer group, stage).  When the ring depth is not a
  // multiple of the group count, successive uses of a stage belong to alternating groups; with a single barrier per
  // stage a consumer warp would see only every other phase, and a parity wait cannot tell "use it landed" from "use
  // it - 2*stages landed, use it - stages still in flight".  With its own barrier per group every waiter (the MMA
  // warp included) observes every phase of the barrier it waits on.  Uses of (group g, stage s) recur every
  // lcm(stages, groups) iterations.  PAIR: one group (every consumer warp passes through every stage).
  constexpr int kFullGroups = PAIR ? 1 : kGroups;
  constexpr int kFullPeriod = tc_lcm(kTcStages, kFullGroups);
  uint64_t* full = bars;                                  // [kFullGroups][kTcStages]
  uint64_t* empty = bars + kFullGroups * kTcStages;       // [kTcStages]
  uint64_t* tfull = empty + kTcStages;                    // [2]
  auto full_bar = [&](uint32_t it_) { return &full[(it_ % kFullGroups) * kTcStages + (it_ % kTcStages)]; };
  auto full_ph = [&](uint32_t it_) { return (it_ / (uint32_t)kFullPeriod) & 1u; };
  static_assert((kFullGroups * kTcStages + 2 * kTcStages + 4) * 8 + 16 <= 1024, "barrier area");
  uint64_t* tempty = tfull + 2;             // [2]
  uint64_t* ready = tempty + 2;             // [kTcStages] ZC: the consumers have added the z / c rows
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ready + kTcStages);
  __shared__ unsigned long long s_coll[2][kCollapseN];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_units = n_genes * S;
  // warp-uniform (every lane of the calling warp must be active): does gene gi hold non-hard-call values?
  auto gene_skipped = [&](int gi) -> bool {
    if (PAIR || counts == nullptr || !genes[gi].counted) return false;
    const int Mg = genes[gi].M;
    const int64_t v0 = genes[gi].var0;
    const int b0 = (lane < Mg) ? counts[v0 + lane].bad : 0, b1 = (lane + 32 < Mg) ? counts[v0 + lane + 32].bad : 0;
    return __any_sync(0xffffffffu, (b0 | b1) != 0);
  };

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < kFullGroups * kTcStages; ++s) mbar_init(&full[s], 1);
      for (int s = 0; s < kTcStages; ++s) {
        // ZC: the UMMAs (which wait for the consumers) are the last readers.  PAIR: every consumer warp passes
        // through every stage (see the consumer loop), so all of them release it.
        mbar_init(&empty[s], PAIR ? 1 + kTcConsumerWarps : (ZC ? 1 : 1 + kTcBoxes));
        mbar_init(&ready[s], kTcBoxes);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull[a], 1);
        mbar_init(&tempty[a], kTcConsumerWarps);
      }
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)),
                 "n"(Cfg::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  if (threadIdx.x < 2 * kCollapseN) (&s_coll[0][0])[threadIdx.x] = 0ull;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // warp-uniform loop; one elected lane issues the bulk copies of a stage
    {
      uint32_t it = 0;
      const int nchunks = (int)((N + 127) >> 7);
      constexpr int kOobRow = 0x7FFF0000;   // beyond any arena (< 2^31 rows of 128 B): TMA zero-fills
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int gi = PAIR ? u % n_genes : u / S, sp = PAIR ? u / n_genes : u - gi * S;   // PAIR: split-major (L2 reuse, below)
        const int row0 = (int)genes[gi].row0;
        const int Mg = genes[gi].M;
        // the box of this gene holds exactly its M rows (no bytes of the neighbouring gene):
        // rows M..63 of the smem tile keep stale data, which only feeds ignored rows/columns of D
        const CUtensorMap* mg = maps_g + (Mg - 1);
        const int row0b = (int)genes[gi].row0_b;
        const int Mgb = genes[gi].Mb;
        const CUtensorMap* mgb = maps_b + (Mgb - 1);
        const uint32_t stage_tx = (uint32_t)(kTcBoxes * (Mg + (PAIR ? Mgb : 0) + ((dbg_skip & 4) ? 0 : ER)) * 128);
        const int64_t k0 = (int64_t)sp * chunk;
        int64_t k1 = k0 + chunk;
        if (k1 > N) k1 = N;
        const int nsteps = (k1 > k0 && !gene_skipped(gi)) ? (int)((k1 - k0 + kTcStageK - 1) / kTcStageK) : 0;
        for (int ks = 0; ks < nsteps; ++ks, ++it) {
          const int s = it % kTcStages;
          const uint32_t ph = (it / kTcStages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          __syncwarp();
          if (elect_one_sync()) {
            uint64_t* fb = full_bar(it);
            mbar_expect_tx(fb, stage_tx);
            uint8_t* st = tiles + (size_t)s * Cfg::kStageBytes;
            const int kb = (int)(k0 + (int64_t)ks * kTcStageK);
#pragma unroll
            for (int b = 0; b < kTcBoxes; ++b) {
              // tiled genotype layout: chunk c of a gene is the contiguous run of M 128-byte rows
              // starting at arena row  row0 + c*M  (arena viewed as [bytes/128][128])
              // a stage may run past the gene's last chunk: such a box is fetched from beyond the
              // arena (row kOobRow), i.e. zero-filled by TMA, instead of from the next gene's block
              const int ch = (kb >> 7) + b;
              const bool in = ch < nchunks;
              tma_load_2d(st + Cfg::g_off(b), mg, 0, in ? row0 + ch * Mg : kOobRow, fb, kEvictFirst);
              if (PAIR)
                tma_load_2d(st + b * Cfg::kBoxBytes + Cfg::kBOff, mgb, 0, in ? row0b + ch * Mgb : kOobRow, fb, kEvictFirst);
              if (!(dbg_skip & 4)) tma_load_2d(st + Cfg::e_off(b), &map_e, kb + b * kTcBoxK, 0, fb, kEvictLast);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D = s32, A = B = signed int8, both K-major, M = 64, N = 64 + ER
    // WIDE: M = 128, N = 2 * (64 + ER)
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((WIDE ? 2 * Cfg::kNC : Cfg::kNC) >> 3) << 17) |
                           ((uint32_t)((WIDE ? 2 * kTileRows : kTileRows) >> 4) << 24);
    uint32_t it = 0, ui = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ui) {
      const int gi = PAIR ? u % n_genes : u / S, sp = PAIR ? u / n_genes : u - gi * S;   // PAIR: split-major (L2 reuse, below)
      const int64_t k0 = (int64_t)sp * chunk;
      int64_t k1 = k0 + chunk;
      if (k1 > N) k1 = N;
      const int nsteps = (k1 > k0 && !gene_skipped(gi)) ? (int)((k1 - k0 + kTcStageK - 1) / kTcStageK) : 0;
      const int a = ui & 1;
      mbar_wait(&tempty[a], ((ui >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(a * Cfg::kAccCols);
      for (int ks = 0; ks < nsteps; ++ks, ++it) {
        const int s = it % kTcStages;
        const uint32_t ph = (it / kTcStages) & 1;
        mbar_wait(full_bar(it), full_ph(it));
        if (ZC) mbar_wait(&ready[s], ph);
        tc_fence_after();
        __syncwarp();
        if (elect_one_sync()) {
          const uint32_t st = smem_u32(tiles + (size_t)s * Cfg::kStageBytes);
          if constexpr (WIDE) {
#pragma unroll
            for (int p = 0; p < ((dbg_skip & 2) ? 0 : kTcBoxes / 2); ++p) {
              // A = rows 0..127 of the pair (two genotype tiles), B = all 128 + 2 ER rows of it
              const uint64_t d0 = umma_desc_sw128(st + p * Cfg::kPairBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_i8(tmem_d, d0 + (uint64_t)(2 * k), d0 + (uint64_t)(2 * k), idesc, (ks | p | k) ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int b = 0; b < ((dbg_skip & 2) ? 0 : kTcBoxes); ++b) {
              const uint64_t da0 = umma_desc_sw128(st + b * Cfg::kBoxBytes + Cfg::kAOff);
              const uint64_t db0 = umma_desc_sw128(st + b * Cfg::kBoxBytes + Cfg::kBOff);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // advance 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
                umma_i8(tmem_d, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, (ks | b | k) ? 1u : 0u);
              }
            }
          }
          umma_commit(&empty[s]);
          if (ks == nsteps - 1) umma_commit(&tfull[a]);
        }
        __syncwarp();
      }
      if (nsteps == 0 && elect_one_sync()) umma_commit(&tfull[a]);  // degenerate unit: publish (stale) accumulator
      __syncwarp();
    }
  } else {
    // ===================== consumers: collapse + epilogue =====================
    const int cw = (warp - 2) % kTcBoxes;   // box handled in this warp's stages
    const int grp = (warp - 2) / kTcBoxes;  // which stages: it % kGroups == grp
    const int egrp = (warp - 2) >> 2;       // epilogue: which of the quadrant's two warps
    const int q = warp & 3;          // TMEM lane quadrant this warp may read
    uint32_t it = 0, ui = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ui) {
      const int gi = PAIR ? u % n_genes : u / S, sp = PAIR ? u / n_genes : u - gi * S;   // PAIR: split-major (L2 reuse, below)
      const GeneDesc gd = genes[gi];
      const int M = gd.M;
      const int64_t k0 = (int64_t)sp * chunk;
      int64_t k1 = k0 + chunk;
      if (k1 > N) k1 = N;
      const int nsteps = (k1 > k0 && !gene_skipped(gi)) ? (int)((k1 - k0 + kTcStageK - 1) / kTcStageK) : 0;
      // per-row flags as two 64-bit masks held in registers (bit r: row r flipped / row r enabled)
      const uint8_t f0 = (lane < M) ? rowflags[gd.var0 + lane] : (uint8_t)kRowSkip;
      const uint8_t f1 = (lane + 32 < M) ? rowflags[gd.var0 + lane + 32] : (uint8_t)kRowSkip;
      const unsigned long long fmask = (unsigned long long)__ballot_sync(0xffffffffu, f0 == kRowFlipped) |
                                       ((unsigned long long)__ballot_sync(0xffffffffu, f1 == kRowFlipped) << 32);
      const unsigned long long emask = (unsigned long long)__ballot_sync(0xffffffffu, f0 != kRowSkip) |
                                       ((unsigned long long)__ballot_sync(0xffffffffu, f1 != kRowSkip) << 32);
      const int Mr8 = (M + 7) & ~7, M8 = M & ~7;
      // warp-uniform: every row of this gene is a normal (unflipped, polymorphic) row
      const bool plain = (fmask == 0ull) && (emask == ((M >= 64) ? ~0ull : ((1ull << M) - 1ull)));
      // lane-constant swizzled offsets of word `lane` in rows j = 0..7 of an 8-row group
      uint32_t woff[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) woff[j] = (uint32_t)(j * 128 + ((((lane >> 2) ^ j) << 4) | ((lane & 3) << 2)));
      int cz[ER + 1], cc[ER + 1];
#pragma unroll
      for (int e = 0; e <= ER; ++e) cz[e] = cc[e] = 0;
      for (int ks = 0; ks < nsteps; ++ks, ++it) {
        // Another consumer group owns this stage -- except in PAIR mode, where the consumers have nothing to do but
        // release the stage: there every warp waits on, and releases, every stage (one barrier group, see `full`).
        // (Found as a hang at N = 200 000 x 16 pair units on 64 CTAs when the barrier was still per stage only: idle
        // consumers raced ahead and a parity wait fell through while the previous use of the stage was in flight.)
        if (!PAIR && (int)(it % (uint32_t)kGroups) != grp) continue;
        const int s = it % kTcStages;
        mbar_wait(full_bar(it), full_ph(it));
        const uint8_t* box = tiles + (size_t)s * Cfg::kStageBytes + Cfg::g_off(cw);
        const int64_t ksamp = k0 + (int64_t)ks * kTcStageK + cw * kTcBoxK + 4 * lane;
        uint32_t z = 0;
        if (PAIR || (dbg_skip & 1)) {
          // block pairs need the Gram only (no burden collapse)
        } else if (plain) {
          // common case (no flipped / monomorphic row): indicator = (g | g>>1) & 1 per byte,
          // 8 independent LDS in flight, 3 ALU ops per row
          for (int r0 = 0; r0 < M8; r0 += 8) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = *reinterpret_cast<const uint32_t*>(box + (r0 >> 3) * 1024 + woff[j]);
#pragma unroll
            for (int j = 0; j < 8; j += 2)
              z += (((w[j] >> 1) | w[j]) & 0x01010101u) + (((w[j + 1] >> 1) | w[j + 1]) & 0x01010101u);
          }
          for (int r = M8; r < M; ++r) {
            const uint32_t w = *reinterpret_cast<const uint32_t*>(box + (r >> 3) * 1024 + woff[r & 7]);
            z += ((w >> 1) | w) & 0x01010101u;
          }
        } else {
          for (int r0 = 0; r0 < Mr8; r0 += 8) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = *reinterpret_cast<const uint32_t*>(box + (r0 >> 3) * 1024 + woff[j]);
            const uint32_t fb = (uint32_t)(fmask >> r0) & 0xFFu, eb = (uint32_t)(emask >> r0) & 0xFFu;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t en = ((eb >> j) & 1u) * 0x01010101u;   // rows >= M are disabled: stale smem never counts
              const uint32_t xf = ((fb >> j) & 1u) * 0x01010101u;
              z += collapse_ind(w[j], xf, en & ~xf, en);
            }
          }
        }
        if (ZC) {
          // samples at/after k1 are zero-filled but a flipped row would count them: mask
          int64_t rem = k1 - ksamp;
          uint32_t vm = rem >= 4 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - (int)rem))));
          z &= vm;
          const uint32_t c = ((z + 0x7F7F7F7Fu) >> 7) & 0x01010101u;
          // rows M and M+1 of this box's genotype tile (never written by TMA: its box has M rows)
          uint8_t* wbox = tiles + (size_t)s * Cfg::kStageBytes + Cfg::g_off(cw);
          *reinterpret_cast<uint32_t*>(wbox + sw128_word_off(M, lane)) = z;
          *reinterpret_cast<uint32_t*>(wbox + sw128_word_off(M + 1, lane)) = c;
          // generic-proxy stores -> visible to the async proxy (UMMA operand fetch)
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&ready[s]);
          continue;
        }
        if (!PAIR && !(dbg_skip & 1)) {
          int64_t rem = k1 - ksamp;
          uint32_t vm = rem >= 4 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : (0xFFFFFFFFu >> (8 * (4 - (int)rem))));
          z &= vm;
          const uint32_t c = ((z + 0x7F7F7F7Fu) >> 7) & 0x01010101u;
          const uint8_t* ebox = tiles + (size_t)s * Cfg::kStageBytes + Cfg::e_off(cw);
#pragma unroll
          for (int e = 0; e < ER; ++e) {
            int ew = *reinterpret_cast<const int*>(ebox + (e >> 3) * 1024 + woff[e & 7]);
            cz[e] = __dp4a((int)z, ew, cz[e]);
            cc[e] = __dp4a((int)c, ew, cc[e]);
          }
          cz[ER] = __dp4a((int)z, (int)z, cz[ER]);
          cc[ER] = __dp4a((int)c, (int)c, cc[ER]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      // ---- collapse sums of this unit: warp reduce (int64), combine the 4 warps through smem
      const int cb = ui & 1;
      if constexpr (!ZC) {
#pragma unroll
        for (int e = 0; e <= ER; ++e) {
          long long a = cz[e], b = cc[e];
          for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
          }
          if (lane == 0) {
            atomicAdd(&s_coll[cb][e], (unsigned long long)a);
            atomicAdd(&s_coll[cb][(ER + 1) + e], (unsigned long long)b);
          }
        }
      }
      // ---- epilogue: accumulator of this unit -> SweepPartial
      const int a = ui & 1;
      mbar_wait(&tfull[a], (ui >> 1) & 1);
      tc_fence_after();
      SweepPartial* o = out + (WIDE ? 2 * (size_t)u : (PAIR ? (size_t)gi * S + sp : (size_t)u));
      // ZC: rows M / M+1 of D are the Zeggini / CMC burden sums of this unit (digit columns, and the
      // sum of squares on their own diagonal entry) -> coll[] of the partial, same slots as the dp4a path
      auto zc_store = [&](SweepPartial* op, int row, int dcol, const uint32_t (&v)[16]) {
        if (!ZC || (row != M && row != M + 1)) return;
        const int base = (row == M) ? 0 : (ER + 1);
        if (dcol >= kTileRows) {
#pragma unroll
          for (int i = 0; i < 16; ++i) op->coll[base + dcol - kTileRows + i] = nsteps ? (long long)(int)v[i] : 0ll;
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (dcol + i == row) op->coll[base + ER] = nsteps ? (long long)(int)v[i] : 0ll;
        }
      };
      if constexpr (WIDE) {
        // UMMA M=128: row m of D lives in TMEM lane m.  Quadrants 0,1 = even boxes (rows 0..63 of the
        // gene), quadrants 2,3 = odd boxes; each half goes to its own SweepPartial.
        const int h = q >> 1;
        const int row = 32 * (q & 1) + lane;
        SweepPartial* oh = o + h;
        const uint32_t taddr = tmem_base + (uint32_t)(a * Cfg::kAccCols) + ((uint32_t)(q * 32) << 16);
        constexpr int nG = kTileRows / 16, nE = ER / 16;
#pragma unroll
        for (int gi2 = 0; gi2 < nG + nE; ++gi2) {
          if ((gi2 & 1) != egrp) continue;   // the quadrant's two warps alternate column groups
          const int col = gi2 < nG ? kTileRows * h + 16 * gi2 : 2 * kTileRows + ER * h + 16 * (gi2 - nG);
          const int dcol = gi2 < nG ? 16 * gi2 : kTileRows + 16 * (gi2 - nG);
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)col, v);
          tmem_ld_wait();
          if (row < M) {
            int4* dst = reinterpret_cast<int4*>(&oh->d[row][dcol]);
            dst[0] = make_int4((int)v[0], (int)v[1], (int)v[2], (int)v[3]);
            dst[1] = make_int4((int)v[4], (int)v[5], (int)v[6], (int)v[7]);
            dst[2] = make_int4((int)v[8], (int)v[9], (int)v[10], (int)v[11]);
            dst[3] = make_int4((int)v[12], (int)v[13], (int)v[14], (int)v[15]);
          }
          zc_store(oh, row, dcol, v);
        }
      } else {
      const uint32_t taddr = tmem_base + (uint32_t)(a * Cfg::kAccCols) + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c0 = 16 * egrp; c0 < Cfg::kNC; c0 += 32) {   // the quadrant's two warps alternate column groups
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (lane < 16) {
          const int row = 16 * q + lane;   // UMMA M=64: row m lives at lane (m%16) + 32*(m/16)
          int4* dst = reinterpret_cast<int4*>(&o->d[row][c0]);
          if (row < M) {
            dst[0] = make_int4((int)v[0], (int)v[1], (int)v[2], (int)v[3]);
            dst[1] = make_int4((int)v[4], (int)v[5], (int)v[6], (int)v[7]);
            dst[2] = make_int4((int)v[8], (int)v[9], (int)v[10], (int)v[11]);
            dst[3] = make_int4((int)v[12], (int)v[13], (int)v[14], (int)v[15]);
          }
          zc_store(o, row, c0, v);
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[a]);
      // all consumer warps have added their collapse sums -> one warp writes them out
      if constexpr (!ZC) asm volatile("bar.sync 1, %0;\n" ::"n"(32 * kTcConsumerWarps) : "memory");
      if (!ZC && warp == 2) {
        for (int i = lane; i < 2 * (ER + 1); i += 32) {
          o->coll[i] = (long long)s_coll[cb][i];
          if (WIDE) o[1].coll[i] = 0ll;   // the collapse sums of a unit live in its first partial
          s_coll[cb][i] = 0ull;
        }
      }
      // s_coll[cb] is reused two units later; the bar.sync of the next unit orders that reuse
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "n"(Cfg::kTmemCols));
  }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcSegments {
  void* encode = nullptr;   // cuTensorMapEncodeTiled through the runtime's driver entry point
  int stages = 5;           // ring depth of the ZC gene sweep: 5 (200 KB), 4 (160 KB) or 3 (120 KB of shared memory)
  int boxes = 4;            // 128-sample boxes per pipeline stage for ER=16: 4 (5 stages) or 2 (10 stages)
  int l2promo = 2;          // CUtensorMapL2promotion: 0 none, 1 64B, 2 128B, 3 256B
  int dbg_skip = 0;         // timing experiments (see k_sweep_tc)
  bool wide = false;        // M=128 two-box UMMAs (TcCfg WIDE; 2 partials per (gene, split)): measured no faster, the
                            // sweep is bound by the shared-memory data pipe, not by UMMA issue (profiles/)
  bool zc = true;           // burden scores through the UMMA (TcCfg ZC) when every gene has M <= 62
  char why[128] = "";
  bool have_e = false;
  int ER = 0;
  CUtensorMap map_e;
  static constexpr int kMaxSeg = 5;
  bool have_seg[kMaxSeg] = {false, false, false, false, false};
  // one tensor map per box height M = 1..64 over the same arena (encoded lazily, kept in HBM)
  struct Seg {
    const int8_t* base = nullptr;
    int64_t rows = 0;   // arena bytes / 128
    bool have_m[kTileRows] = {};
    CUtensorMap* d_maps = nullptr;   // [64] in device memory
  } seg[kMaxSeg];
};

inline int tc_init(TcSegments* tc, char* err, size_t errlen) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    snprintf(tc->why, sizeof(tc->why), "cuTensorMapEncodeTiled unavailable");
    tc->encode = nullptr;
    (void)cudaGetLastError();
    return 0;  // the dp4a engine still works; an explicit engine=tc request fails loudly
  }
  tc->encode = fn;
  e = cudaFuncSetAttribute(k_sweep_tc<16, 5, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<16, 5, false, 4>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tc<16, 10, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<16, 10, false, 2>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tc<32, 4, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<32, 4, false, 4>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tc<16, 5, false, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<16, 5, false, 4, true>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tc<32, 4, false, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<32, 4, false, 4, true>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tc<16, 5, false, 4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<16, 5, false, 4, false, true>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tc<16, 5, false, 4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<16, 5, false, 4, true, true>::kSmem);
  // shallower rings (option "tc_stages"): 160 / 120 KB instead of 200 KB, so that the statistics kernels of the previous
  // batch fit on the same SMs and run UNDER the sweep (rvt_api.cu: launch_range, option "overlap")
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tc<16, 4, false, 4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<16, 4, false, 4, false, true>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tc<16, 3, false, 4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<16, 3, false, 4, false, true>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tc<32, 3, false, 4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<32, 3, false, 4, false, true>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tc<32, 4, false, 4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<32, 4, false, 4, false, true>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tc<32, 4, false, 4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<32, 4, false, 4, true, true>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tc<16, 3, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<16, 3, true, 4>::kSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sweep_tc<32, 2, true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<32, 2, true, 4>::kSmem);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "cudaFuncSetAttribute(k_sweep_tc): %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

inline void tc_destroy(TcSegments* tc) {
  for (auto& sg : tc->seg)
    if (sg.d_maps) {
      cudaFree(sg.d_maps);
      sg.d_maps = nullptr;
    }
}

// 2-D map over [rows][row_bytes] with pitch ld: the genotype arena is [bytes/128][128] (ld = 128),
// the null-model digits E are [ER][N] (ld = ldE).
inline int tc_make_map(TcSegments* tc, CUtensorMap* map, const void* base, int64_t rows, int64_t N, int64_t ld, int box_rows,
                       char* err, size_t errlen) {
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld};
  cuuint32_t box[2] = {(cuuint32_t)kTcBoxK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((PFN_encodeTiled)tc->encode)(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box,
                                             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                             (CUtensorMapL2promotion)tc->l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (%d) rows=%lld N=%lld ld=%lld", (int)r, (long long)rows, (long long)N,
             (long long)ld);
    return -2;
  }
  return 0;
}

inline int tc_bind_null(TcSegments* tc, const int8_t* E, int ER, int64_t N, int64_t ldE, char* err, size_t errlen) {
  tc->have_e = false;
  for (bool& b : tc->have_seg) b = false;  // segments are tied to N
  if (!tc->encode) return 0;
  if (ER != 16 && ER != 32) {
    snprintf(tc->why, sizeof(tc->why), "ER=%d unsupported by the tensor-core sweep", ER);
    return 0;
  }
  int rc = tc_make_map(tc, &tc->map_e, E, ER, N, ldE, ER, err, errlen);
  if (rc) return rc;
  tc->have_e = true;
  tc->ER = ER;
  return 0;
}

inline int tc_bind_segment(TcSegments* tc, int seg, const int8_t* base, int64_t bytes, char* err, size_t errlen) {
  if (!tc->encode || seg < 0 || seg >= TcSegments::kMaxSeg) return 0;
  tc->have_seg[seg] = false;
  if (bytes <= 0) return 0;
  if (bytes / 128 >= ((int64_t)1 << 31)) {
    snprintf(tc->why, sizeof(tc->why), "segment larger than 2^31 TMA rows");
    return 0;
  }
  TcSegments::Seg& sg = tc->seg[seg];
  if (sg.base == base && sg.rows == bytes / 128 && sg.d_maps) {   // unchanged: keep the encoded maps
    tc->have_seg[seg] = true;
    return 0;
  }
  sg.base = base;
  sg.rows = bytes / 128;
  for (bool& b : sg.have_m) b = false;
  if (!sg.d_maps) {
    cudaError_t e = cudaMalloc((void**)&sg.d_maps, sizeof(CUtensorMap) * kTileRows);
    if (e != cudaSuccess) {
      snprintf(err, errlen, "cudaMalloc(tensor maps): %s", cudaGetErrorString(e));
      return -2;
    }
  }
  tc->have_seg[seg] = true;
  return 0;
}

// make sure the segment has a tensor map for every box height present in this batch
inline int tc_prepare_maps(TcSegments* tc, int seg, const GeneDesc* h_genes, int n, cudaStream_t st, char* err, size_t errlen,
                           int seg_b = -1 /* PAIR: segment of the B tiles when it is not `seg` */) {
  if (seg_b >= 0 && seg_b != seg) {
    int rc = tc_prepare_maps(tc, seg_b, h_genes, n, st, err, errlen, -1);   // (encodes a few unused heights; harmless)
    if (rc) return rc;
  }
  TcSegments::Seg& sg = tc->seg[seg];
  for (int i = 0; i < 2 * n; ++i) {
    const int M = (i < n) ? h_genes[i].M : h_genes[i - n].Mb;
    if (M < 1 || M > kTileRows || sg.have_m[M - 1]) continue;
    CUtensorMap m;
    int rc = tc_make_map(tc, &m, sg.base, sg.rows, 128, 128, M, err, errlen);
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(sg.d_maps + (M - 1), &m, sizeof(m), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // `m` is a stack temporary
    if (e != cudaSuccess) {
      snprintf(err, errlen, "tensor map upload: %s", cudaGetErrorString(e));
      return -2;
    }
    sg.have_m[M - 1] = true;
  }
  return 0;
}

// SweepPartials written per (gene, split) by the gene sweep
inline int tc_parts_per_unit(const TcSegments* tc) { return (tc->wide && tc->boxes == 4) ? 2 : 1; }

inline bool tc_usable(TcSegments* tc, const GeneDesc* h_genes, int n) {
  if (!tc->encode) return false;
  if (!tc->have_e) {
    if (!tc->why[0]) snprintf(tc->why, sizeof(tc->why), "null-model tensor map missing");
    return false;
  }
  const int seg = h_genes[0].seg;
  if (seg < 0 || seg >= TcSegments::kMaxSeg || !tc->have_seg[seg]) {
    snprintf(tc->why, sizeof(tc->why), "genes are not in a TMA-mapped segment");
    return false;
  }
  for (int i = 0; i < n; ++i)
    if (h_genes[i].seg != seg || !h_genes[i].tiled) {
      snprintf(tc->why, sizeof(tc->why), "genes span several segments or are not in the tiled layout");
      return false;
    }
  return true;
}

inline int tc_launch(TcSegments* tc, const GeneDesc* d_genes, const GeneDesc* h_genes, int n, const uint8_t* d_flags,
                     const NullModel* /*d_nm*/, int64_t N, int ER, int S, int64_t chunk, SweepPartial* d_parts,
                     unsigned int* /*counter*/, int sm_count, cudaStream_t st, char* err, size_t errlen,
                     bool pair = false, bool wide = false, int seg_b = -1, const RowCounts* d_counts = nullptr) {
  const int seg = h_genes[0].seg;
  if (seg_b < 0) seg_b = seg;
  const int grid = std::min(n * S, sm_count);
  if (chunk % kTcChunkAlign != 0) {
    snprintf(err, errlen, "internal: chunk %lld is not a multiple of the TMA stage (%d)", (long long)chunk, kTcChunkAlign);
    return -3;
  }
  int rc = tc_prepare_maps(tc, seg, h_genes, n, st, err, errlen, seg_b);
  if (rc) return rc;
#define RVT_TC_LAUNCH(ER_, ST_, PAIR_, BX_, ...)                                                                             \
  k_sweep_tc<ER_, ST_, PAIR_, BX_, ##__VA_ARGS__><<<grid, kTcThreads, TcCfg<ER_, ST_, PAIR_, BX_, ##__VA_ARGS__>::kSmem, st>>>( \
      tc->seg[seg].d_maps, tc->seg[seg_b].d_maps, tc->map_e, d_genes, n, d_flags, N, S, chunk, d_parts, tc->dbg_skip, d_counts)
  // burden scores through the UMMA: two spare tile rows and |z . digit| sums that fit int32
  bool zc = tc->zc && !pair && tc->boxes == 4 && chunk <= 262144;
  for (int i = 0; i < n && zc; ++i) zc = h_genes[i].M <= kTileRows - 2;
  if (zc && wide && ER == 16)
    RVT_TC_LAUNCH(16, 5, false, 4, true, true);
  else if (zc && wide)
    RVT_TC_LAUNCH(32, 4, false, 4, true, true);
  else if (zc && ER == 16 && tc->stages == 4)
    RVT_TC_LAUNCH(16, 4, false, 4, false, true);
  else if (zc && ER == 16 && tc->stages == 3)
    RVT_TC_LAUNCH(16, 3, false, 4, false, true);
  else if (zc && ER == 16)
    RVT_TC_LAUNCH(16, 5, false, 4, false, true);
  else if (zc && tc->stages <= 3)
    RVT_TC_LAUNCH(32, 3, false, 4, false, true);
  else if (zc)
    RVT_TC_LAUNCH(32, 4, false, 4, false, true);
  else if (wide && !pair && ER == 16)
    RVT_TC_LAUNCH(16, 5, false, 4, true);
  else if (wide && !pair)
    RVT_TC_LAUNCH(32, 4, false, 4, true);
  else if (pair && ER == 16)
    RVT_TC_LAUNCH(16, 3, true, 4);
  else if (pair)
    RVT_TC_LAUNCH(32, 2, true, 4);
  else if (ER == 16 && tc->boxes == 2)
    RVT_TC_LAUNCH(16, 10, false, 2);
  else if (ER == 16)
    RVT_TC_LAUNCH(16, 5, false, 4);
  else
    RVT_TC_LAUNCH(32, 4, false, 4);
#undef RVT_TC_LAUNCH
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, errlen, "k_sweep_tc launch: %s", cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

}  // namespace rvt
