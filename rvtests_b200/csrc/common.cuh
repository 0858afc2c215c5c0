// common.cuh -- shared definitions of the B200 gene engine (device data layout + helpers).
//
// DATA LAYOUT IN HBM
//   genotype block of one gene : int8, hard calls 0/1/2, one byte per call, in the engine's TILED
//       variant-major layout  [chunk c = sample/128][variant j][128 samples]  (zero padded to a
//       whole chunk).  One sweep stage of 4 chunks is then ONE contiguous 4*M*128-byte run of HBM
//       (a plain variant-major [M][N] block would scatter it over M DRAM pages, 128-512 B each,
//       which capped the sweep at ~64 % of the HBM roofline).  All genes of a segment sit in one byte
//       arena viewed as [bytes/128][128], so one family of TMA tensor maps (box = M x 128 B) serves
//       every gene.  Caller-owned device blocks may stay plain variant-major [M][ld] (zero-copy,
//       dp4a engine only).
//   null-model digits "E"      : int8 [ER rows][ldE], ER = 4*(C+1) rounded up to 16.  Row 4*v+k is
//       base-256 balanced digit k of the fixed-point image of vector v, v=0 the null residual r,
//       v=1..C the covariate columns (column 0 = intercept).  value_i = (sum_k d_ik 256^k) 2^-e_v.
//       With G in {0,1,2} and digits in [-128,127] every dot product the tests need
//       (G'r, G'X, G'G, collapse'r ...) is an EXACT integer sum -> bit-reproducible whatever the
//       split / kernel / reduction order, and tensor-core friendly (kind::i8, s32 accumulate).
//   sweep partials             : per (gene, split) one SweepPartial (int32 tile + int64 collapse
//       sums), reduced by the finalize kernel in int64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace rvt {

constexpr int kMaxM = 64;       // variants per gene handled by the single-tile sweep kernels
constexpr int kMaxC = 7;        // covariate columns incl. intercept (ER <= 32)
constexpr int kMaxER = 32;      // rows of E
constexpr int kTileRows = 64;   // gene rows per tile (UMMA M)
constexpr int kMaxNC = kTileRows + kMaxER;  // output columns per tile (gene + digits) <= 96
constexpr int kCollapseN = 2 * (kMaxER + 1);  // zeggini then cmc: ER dot products + sum of squares

// per-variant flag bytes consumed by the collapse pass
enum : uint8_t { kRowNormal = 0, kRowFlipped = 1, kRowSkip = 2 };

struct GeneDesc {
  const int8_t* g;   // row 0 of the gene's block
  int64_t ld;        // bytes between rows (multiple of 16)
  int32_t M;         // rows (variants) in the block, 1..kMaxM
  int32_t seg;       // tensor-map segment (TC path), -1 if none
  int64_t row0;      // first row inside the segment
  int64_t var0;      // offset of this gene's per-variant side arrays (flags, af)
  int32_t has_af;    // caller supplied allele frequencies (F9 quirk path)
  int32_t counted;   // the engine counted this gene's rows itself (RowCounts valid)
  int64_t row0_b;    // meta-cov block pairs: first row of the B-operand tile (== row0 for a gene)
  int32_t Mb;        // rows of the B tile (== M for a gene)
  int32_t tiled;     // 1: [chunk][M][128] layout (row0 = byte offset / 128 inside the segment); 0: [M][ld]
  int64_t var0_b;    // meta: variant index of the B tile's first row
};

// address of (variant row r, sample k) inside a gene block
__device__ __forceinline__ const int8_t* geno_ptr(const GeneDesc& gd, int r, int64_t k) {
  return gd.tiled ? gd.g + ((size_t)(k >> 7) * gd.M + r) * 128 + (k & 127) : gd.g + (size_t)r * gd.ld + k;
}
static inline int64_t tiled_bytes(int64_t N, int M) { return ((N + 127) / 128) * (int64_t)M * 128; }

struct RowCounts {
  int n1, n2, bad, pad;  // #het, #hom-alt, #values outside {0,1,2}
};

struct SweepPartial {
  int32_t d[kTileRows][kMaxNC];  // d[i][j]: j<64 gene x gene, j>=64 gene x digit row (j-64)
  long long coll[kCollapseN];    // [0..ER) zeg.digit, [ER] sum zeg^2, then the same for cmc
  long long pad[2];
};

struct NullModel {
  int64_t N;
  int32_t C;
  int32_t ER;            // rows of E (16 or 32)
  int64_t ldE;           // bytes between rows of E
  const int8_t* E;       // digits
  const double* resid;   // N
  double sigma2;         // RSS / N   (regression/LinearRegression.cpp:60)
  double xtx_inv[kMaxC * kMaxC];  // (X'X)^-1 row-major
  double scale[kMaxC + 1];        // 2^-e_v per vector v
  long long vsum[kMaxC + 1];      // sum_i fixed-point value of vector v (for flipped columns)
  double rsum;                    // sum_i r_i  and  sum_i x_il  in fp64 (dosage path)
  double xsum[kMaxC];
  // binary trait (logistic null, regression/LogisticRegression.cpp:279-339): r = y - p, per-sample variance
  // v_i = p_i (1 - p_i), xtx_inv holds (X'VX)^-1 and sigma2 is 1 so that the quantitative formulas carry over
  int32_t binary, pad_;
  const double* vw;               // N variances (null for a quantitative trait: v == 1)
  double vsum_w;                  // sum_i v_i
  double xsum_w[kMaxC];           // sum_i v_i x_il
};

// Pre-digested per-gene statistics handed to the tail of k_finalize by the dosage path
// (dosage.cuh): everything steps 1-4 of k_finalize would have produced.
struct TailInput {
  int Mp, status, nonref, pad;
  double Q;
  double K[kTileRows * kTileRows];   // Mp x Mp, row-major with leading dimension Mp
  double vw[kTileRows];              // w_j (unsquared) * g_j'r   (SKAT-O)
  double zegU, zegSS, zegSZ[kMaxC];  // burden score-test ingredients (before the covariate projection)
  double cmcU, cmcSS, cmcSZ[kMaxC];
};

struct EngineParams {
  double beta1, beta2;   // Beta weight parameters (src/ModelManager.cpp:169-175 defaults 1, 25)
  long long wd_cycles;   // device watchdog of the per-gene tail (SKAT-O quadrature), SM cycles; 0 = off
};

#define RVT_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      snprintf(ctx->err, sizeof(ctx->err), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, \
               cudaGetErrorString(_e));                                                \
      return RVT_E_CUDA;                                                               \
    }                                                                                  \
  } while (0)

// CTA-wide cooperative group used by the finalize kernel (davies.cuh / eigen.cuh `Par`).
struct BlockPar {
  double* red;  // 2 * 32 doubles of shared scratch
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int nt() const { return blockDim.x; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ void allreduce2(double& a, double& b) const {
    // fixed-order tree: warp shuffles, then warp 0 over the per-warp partials; result broadcast.
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) {
      red[w] = a;
      red[32 + w] = b;
    }
    __syncthreads();
    double sa = 0.0, sb = 0.0;
    for (int i = 0; i < nw; ++i) {
      sa += red[i];
      sb += red[32 + i];
    }
    a = sa;
    b = sb;
  }
  __device__ __forceinline__ void allreduce4(double& a, double& b, double& c, double& d) const {
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
      d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) {
      red[w] = a;
      red[8 + w] = b;
      red[16 + w] = c;
      red[24 + w] = d;
    }
    __syncthreads();
    double sa = 0.0, sb = 0.0, sc = 0.0, sd = 0.0;
    for (int i = 0; i < nw; ++i) {
      sa += red[i];
      sb += red[8 + i];
      sc += red[16 + i];
      sd += red[24 + i];
    }
    a = sa;
    b = sb;
    c = sc;
    d = sd;
  }
};

}  // namespace rvt
