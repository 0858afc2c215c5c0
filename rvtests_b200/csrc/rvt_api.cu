// rvt_api.cu -- the C ABI (include/rvtests_b200.h) of the B200 gene engine: context, null model,
// gene queue, flush = [K0 flags] -> [K1 sweep] -> [K2/K3 finalize], and the entry points of the rows built around it
// (wide genes, permutation test, meta score/cov, FastLMM score step, BoltLMM null fit, binary traits).
// Host code here is allocation, launches, copies and CONTROL FLOW on O(1)..O(C^2)..O(R) scalars: the per-gene and
// per-variant statistics are computed on the GPU.  The places where the host does arithmetic are the ones whose inputs are
// a handful of scalars per step: the Cholesky solve of the C x C logistic Newton step, the conjugate-gradient / secant
// bookkeeping and the random draws of the Bolt fit (the reference's own MT19937 stream), its covariate basis (Gram-Schmidt
// on the N x C covariates, as BoltPlinkLoader does on the host), (ux' D ux)^-1 of the FastLMM score step, the polynomial
// jump of the rand() stream and the adaptive stop rule of the permutation test.  There is no CPU fallback of any kernel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <new>
#include <string>
#include <vector>

#include "../../include/rvtests_b200.h"
#include "common.cuh"
#include "dosage.cuh"
#include "finalize.cuh"
#include "bolt.cuh"
#include "lmm.cuh"
#include "meta.cuh"
#include "null_model.cuh"
#include "perm.cuh"
#include "prep.cuh"
#include "sweep_simt.cuh"
#include "sweep_tc.cuh"
#include "sweep_aug.cuh"
#include "wide.cuh"

using namespace rvt;

namespace {
struct DosGene {     // a pushed gene with non-hard-call values: handled by the fp64 path (dosage.cuh)
  int gene_index;
  int M;
  double* dG;        // N x M column-major doubles, device (null: expand the gene's hard-call tiles at flush)
  bool has_af;
  bool from_bed = false;   // arrived as 2-bit rows: code 3 in its tiles is a missing call to impute
  std::vector<double> af;
};
struct WideGene {    // a pushed gene of more than kMaxM variants: T tiles of consecutive variants (wide.cuh)
  int gene_index;    // its slot in the pending list (a placeholder descriptor = tile 0 sits there)
  int M;
  int64_t var0;
  bool has_af;
  bool imputed = false;   // pushed as 2-bit rows and holds missing calls: mean-imputed through the rows [H ; Mi] (run_wide)
  double* dG = nullptr;   // real dosages: the pushed N x M doubles, kept on the device for the dense statistics (owned)
  std::vector<GeneDesc> tiles;
};
struct TilePlan {    // where the tiles of one pushed gene were staged
  int T = 0;
  std::vector<int64_t> off;
  std::vector<int> rows, r0;
};
}  // namespace

constexpr int kSegLoaded = 0;  // the synthetic / loaded cohort arena
constexpr int kSegStaged = 1;  // genes staged from host buffers
constexpr int kSegLmm = 3;     // eigenvector digit tiles of the mixed-model score step (lmm.cuh)
constexpr int kSegPerm = 2;    // 16-permutation digit tiles of the permutation test (perm.cuh)
constexpr int kSegAux = 4;     // H / M operand tiles of a gene with missing calls (permutation test)

struct rvt_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  std::vector<cudaEvent_t> evpool;
  int pending_timing_batches = 0;
  char err[512] = {0};
  int sm_count = 148;
  // options
  double beta1 = 1.0, beta2 = 25.0;
  int engine = RVT_ENGINE_AUTO;
  int splits = 0;
  // null model
  bool have_null = false;
  int64_t N = 0;
  int C = 0, ER = 0;
  int64_t ldE = 0;
  double *dX = nullptr, *dy = nullptr, *dresid = nullptr, *dnull_part = nullptr, *dbeta = nullptr;
  int8_t* dE = nullptr;
  NullModel* d_nm = nullptr;
  NullModel h_nm;
  int *d_shift = nullptr, *d_status = nullptr;
  // pending genes
  std::vector<GeneDesc> genes;
  std::vector<uint8_t> userflags;  // per variant, 0xFF = derive from counts
  std::vector<double> af;          // per variant (valid when gene.has_af)
  std::vector<int64_t> count_slot; // per gene: offset into d_counts or -1
  std::vector<DosGene> dos;        // pending genes that need the dosage path
  std::vector<WideGene> wide;      // pending genes of more than kMaxM variants
  std::vector<int> slots;          // per gene: per-variant slots it owns (== M except for the placeholder of a wide gene)
  uint8_t* d_zero_flags = nullptr; // "every row normal" flags for the tile sweeps of wide genes
  size_t cap_zero_flags = 0;
  // permutation test (perm.cuh): options, rand() stream position, scratch, records of the last flush
  int perm_n = 0, perm_batch = 512;   // (28.6 k perm/s at 256, 37.4 k at 1024 for 500 000 x 50: profiles/r01d_perm_time.txt)
  double perm_alpha = 0.05;
  uint64_t perm_pos = 0;           // rand() values consumed so far (the reference's process-wide stream)
  uint32_t perm_seed = 1;          // glibc's default state == srand(1); the reference never seeds
  LfgTables* d_lfg = nullptr;
  uint8_t* d_perm = nullptr;       // one scratch allocation, carved per gene
  size_t cap_perm = 0;
  std::vector<rvt_perm_result> perm_out;
  std::vector<char> is_dos;        // per pending gene: took the fp64 path
  // fp64 path: per-gene statistics / tail inputs, kept across flushes (cudaMalloc / cudaFree per flush cost 30-170 ms)
  void *d_dos_st = nullptr, *d_dos_tin = nullptr, *d_dos_idx = nullptr, *d_dos_afd = nullptr, *d_dos_tg = nullptr;
  size_t cap_dos_st = 0, cap_dos_tin = 0, cap_dos_idx = 0, cap_dos_afd = 0, cap_dos_tg = 0;
  // scratch of the meta / mixed-model flushes, kept across calls for the same reason
  double h_xvx[kMaxC * kMaxC] = {};   // binary trait: X'VX of the logistic null model
  int64_t metab_nv = 0;               // variants of the last binary-trait meta flush (rvt_meta_binary_extras)
  void* d_metab[6] = {};              // its scratch: digits, scaled tiles, band sums, per-variant sums, case/control counts, covXZ
  size_t cap_metab[6] = {};
  void* d_scr[8] = {};
  size_t cap_scr[8] = {};
  // binary trait (logistic null model)
  bool binary = false;
  double *d_p = nullptr, *d_vw = nullptr;
  // FastLMM score step (lmm.cuh)
  bool have_lmm = false;
  LmmNull* d_lmm = nullptr;
  LmmNull h_lmm;
  int8_t* d_lmm_tiles = nullptr;
  double* d_lmm_vec = nullptr;     // a, d, t (16 nb each), w (16 nb x kMaxC)
  long long* d_lmm_tsum = nullptr;
  bool perm_log = false;           // option "debug_perm_q": keep every permuted statistic of the last flush
  std::vector<double> perm_q_log;
  std::vector<int> bed_genes;      // pending genes pushed as PLINK 2-bit rows: checked for missing calls at flush
  // binary trait: tile genes whose fp64 statistics + tail were enqueued right behind their copies (launch_range), so that they
  // run under the PCIe transfers of the following genes instead of all at flush: 0 no, 1 done (hard calls / missing unknown yet)
  std::vector<char> bin_streamed;
  std::vector<char> is_bed;        // per pending gene: pushed as 2-bit rows (code 3 = a missing call to impute)
  std::vector<TileGene> bs_tg;     // host staging of one streamed batch
  std::vector<int> bs_idx;
  cudaStream_t bin_s = nullptr;
  cudaEvent_t ev_bin_in = nullptr, ev_bin_out = nullptr;
  bool bin_pending = false;
  void *d_bs_st = nullptr, *d_bs_tin = nullptr, *d_bs_idx = nullptr, *d_bs_tg = nullptr;
  size_t cap_bs_st = 0, cap_bs_tin = 0, cap_bs_idx = 0, cap_bs_tg = 0;
  std::vector<int> unsupported;    // pending genes the flush cannot compute (wide + missing calls): record status only
  int launched = 0;                // pending genes [0, launched) already have their kernels enqueued (stream_batch)
  int stream_batch = 0;            // option: enqueue sweep + statistics every this many host pushes (0 = only at flush)
  bool flush_open = false;         // the timing window of the current flush has started
  // host pushes land through a ring of buffers on a copy stream: the H2D copy of gene k+1 overlaps the
  // unpack / re-tile kernel of gene k and the sweeps enqueued by stream_batch
  static constexpr int kLandRing = 6;
  cudaStream_t copy_stream = nullptr;
  int8_t* d_land[kLandRing] = {};
  size_t cap_land = 0;
  cudaEvent_t ev_landed[kLandRing] = {}, ev_unpacked[kLandRing] = {};
  unsigned long long land_seq = 0;
  int64_t n_var = 0;
  // device side arrays (grown on demand)
  GeneDesc* d_genes = nullptr;
  size_t cap_genes = 0;
  uint8_t *d_flags = nullptr, *d_userflags = nullptr;
  double* d_af = nullptr;
  RowCounts* d_counts = nullptr;
  size_t cap_var = 0;
  SweepPartial* d_parts = nullptr;
  size_t cap_parts = 0;
  rvt_gene_result* d_res = nullptr;
  size_t cap_res = 0;
  unsigned int* d_counter = nullptr;
  QagsScratch* d_qags = nullptr;   // SKAT-O interval lists, one per gene of a batch
  size_t cap_qags = 0;
  SkatoJob* d_jobs = nullptr;      // SKAT-O: what k_finalize hands to k_skato_qags, one per gene of a batch
  size_t cap_jobs = 0;
  // statistics (K2/K3/K4) of batch i on a second stream UNDER the sweep of batch i+1 (option "overlap"): needs the sweep to
  // leave shared memory free ("tc_stages" 4 or 3) and the partials double-buffered
  int overlap = 0;                 // 0 off, else the number of batches a flush is cut into
  cudaStream_t fin_stream = nullptr;
  cudaEvent_t ev_swept[2] = {nullptr, nullptr}, ev_fin_done[2] = {nullptr, nullptr};
  bool fin_busy[2] = {false, false};
  unsigned long long batch_seq = 0;
  int fin_split = 1;               // statistics as front / bisection / tail kernels (finalize.cuh); 0: the all-in-one kernel
  FinMid* d_mid = nullptr;
  size_t cap_mid = 0;
  int aug = 1;                     // genes with missing calls on the augmented tensor-core sweep (sweep_aug.cuh); 0: sparse kernel
  double meta_cov_scale = 0.0;     // > 0: mixed-model (Bolt) covariance band, see rvt_meta_flush
  int qags_pack = 1;               // SKAT-O quadrature: 1 = three genes per persistent 128-thread CTA (k_skato_qags_packed)
  long long wd_cycles = 8000000000ll;   // device watchdog of the per-gene tail, SM cycles (option "watchdog_ms"; ~4 s)
  bool skato = false;
  bool bin_stream = true;      // option "binary_stream": binary-trait tile genes are computed in launch_range (behind their copies), not at flush
  bool f64_imputed = true;     // option "f64_imputed": rvt_gene_push_f64 recognises mean-imputed hard calls and routes them like a 2-bit push with missing calls
  unsigned long long* d_frac = nullptr;   // per-row {min, max} of the non-integer values of the last rvt_gene_push_f64
  bool perm_stream_lost = false;   // a gene the reference would have permuted was skipped: later stream positions are not the reference's
  bool bolt_binary = false;    // option "bolt_binary": the null fit of a binary trait (BoltLMM::enableBinaryMode: the phenotype is not centred)
  int bolt_kernels = 3;        // generation of the panel-product kernels of rvt_bolt_fit_null (bolt.cuh)
  bool skato_binary = true;    // SKAT-O for a binary trait (SkatO::Fit type "D"); on by default, see include/rvtests_b200.h
  long long* d_dbg = nullptr;   // optional finalize phase counters (rvt_set_option "debug_phases")
  size_t cap_dbg = 0;
  bool want_dbg = false;
  // staging: one growable byte arena of tiled gene blocks, TMA segment kSegStaged
  int8_t* d_stage = nullptr;
  int64_t stage_cap = 0, stage_used = 0;
  double* d_stage64 = nullptr;
  size_t cap_stage64 = 0;
  // loaded synthetic cohort
  int8_t* d_loaded = nullptr;
  int64_t loaded_rows = 0, loaded_ld = 0;
  int loaded_genes = 0, loaded_M = 0;
  std::vector<uint8_t> loaded_flags;
  std::vector<double> loaded_af;
  TcSegments tc;
  // timing
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double t_sweep = 0, t_fin = 0, t_total = 0, n_launch = 0;
  int last_engine = 0, last_S = 0, last_aug = 0;
  int64_t last_parts = 0;
  int last_n = 0;
};

static void pending_reset(rvt_ctx* ctx) {
  if (ctx->bin_pending && ctx->bin_s) cudaStreamSynchronize(ctx->bin_s);   // (a failed flush drops the queue: nothing may still run on it)
  ctx->bin_pending = false;
  for (auto& w : ctx->wide)
    if (w.dG) cudaFree(w.dG);
  ctx->genes.clear();
  ctx->userflags.clear();
  ctx->af.clear();
  ctx->count_slot.clear();
  ctx->bed_genes.clear();
  ctx->bin_streamed.clear();
  ctx->is_bed.clear();
  ctx->unsupported.clear();
  ctx->wide.clear();
  ctx->slots.clear();
  ctx->n_var = 0;
  ctx->stage_used = 0;
  ctx->launched = 0;
  ctx->pending_timing_batches = 0;
  ctx->flush_open = false;
}

#define CTX_FAIL(code, ...)                              \
  do {                                                   \
    snprintf(ctx->err, sizeof(ctx->err), __VA_ARGS__);   \
    return (code);                                       \
  } while (0)

static int ensure(rvt_ctx* ctx, void** p, size_t* cap, size_t need, size_t elem) {
  if (need <= *cap) return RVT_OK;
  size_t ncap = std::max(need, *cap * 2);
  void* np = nullptr;
  RVT_CUDA_OK(cudaMalloc(&np, ncap * elem));
  if (*p) {
    RVT_CUDA_OK(cudaMemcpyAsync(np, *p, *cap * elem, cudaMemcpyDeviceToDevice, ctx->stream));
    RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    RVT_CUDA_OK(cudaFree(*p));
  }
  *p = np;
  *cap = ncap;
  return RVT_OK;
}

// scratch slot k of the context, at least `bytes` large; contents are NOT preserved on growth
static int scratch(rvt_ctx* ctx, int k, size_t bytes, void** out) {
  if (bytes > ctx->cap_scr[k]) {
    if (ctx->d_scr[k]) {
      RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->d_scr[k]);
    }
    ctx->d_scr[k] = nullptr;
    ctx->cap_scr[k] = 0;
    RVT_CUDA_OK(cudaMalloc(&ctx->d_scr[k], bytes));
    ctx->cap_scr[k] = bytes;
  }
  *out = ctx->d_scr[k];
  return RVT_OK;
}

// per-variant device arrays share one capacity
static int ensure_var(rvt_ctx* ctx, size_t need) {
  if (need <= ctx->cap_var) return RVT_OK;
  size_t ncap = std::max(need, ctx->cap_var * 2 + 1024);
  auto grow = [&](void** p, size_t elem) -> int {
    void* np = nullptr;
    RVT_CUDA_OK(cudaMalloc(&np, ncap * elem));
    RVT_CUDA_OK(cudaMemsetAsync(np, 0, ncap * elem, ctx->stream));
    if (*p) {
      RVT_CUDA_OK(cudaMemcpyAsync(np, *p, ctx->cap_var * elem, cudaMemcpyDeviceToDevice, ctx->stream));
      RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
      RVT_CUDA_OK(cudaFree(*p));
    }
    *p = np;
    return RVT_OK;
  };
  int rc;
  if ((rc = grow((void**)&ctx->d_flags, 1))) return rc;
  if ((rc = grow((void**)&ctx->d_userflags, 1))) return rc;
  if ((rc = grow((void**)&ctx->d_af, sizeof(double)))) return rc;
  if ((rc = grow((void**)&ctx->d_counts, sizeof(RowCounts)))) return rc;
  ctx->cap_var = ncap;
  return RVT_OK;
}

// bytes for one staged (tiled) gene block; grows the arena (device-to-device copy, pending
// descriptors re-based).  Offsets are multiples of 128 so that they are whole TMA rows.
static int stage_alloc(rvt_ctx* ctx, int64_t bytes, int64_t* off) {
  if (ctx->stage_used + bytes > ctx->stage_cap) {
    int64_t ncap = std::max<int64_t>(ctx->stage_used + bytes, ctx->stage_cap * 2);
    ncap = std::max<int64_t>(ncap, (int64_t)64 << 20);
    int8_t* np = nullptr;
    RVT_CUDA_OK(cudaMalloc((void**)&np, (size_t)ncap));
    if (ctx->d_stage) {
      RVT_CUDA_OK(cudaMemcpyAsync(np, ctx->d_stage, (size_t)ctx->stage_used, cudaMemcpyDeviceToDevice, ctx->stream));
      RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
      RVT_CUDA_OK(cudaFree(ctx->d_stage));
    }
    for (auto& g : ctx->genes)
      if (g.seg == kSegStaged) g.g = np + (size_t)g.row0 * 128;
    for (auto& w : ctx->wide)
      for (auto& g : w.tiles) g.g = np + (size_t)g.row0 * 128;
    ctx->d_stage = np;
    ctx->stage_cap = ncap;
  }
  *off = ctx->stage_used;
  ctx->stage_used += bytes;
  return RVT_OK;
}

extern "C" {

int rvt_ctx_create(int device, rvt_ctx** out) {
  if (!out) return RVT_E_BADARG;
  *out = nullptr;
  rvt_ctx* ctx = new (std::nothrow) rvt_ctx();
  if (!ctx) return RVT_E_CUDA;
  *out = ctx;  // returned even on failure so that rvt_last_error() can explain
  ctx->device = device;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    CTX_FAIL(RVT_E_CUDA, "no CUDA device available (%s): the engine has no CPU fallback",
             cudaGetErrorString(e));
  RVT_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  RVT_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  if (prop.major != 10)
    CTX_FAIL(RVT_E_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device,
             prop.major, prop.minor);
  RVT_CUDA_OK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  for (auto& ev : ctx->ev) RVT_CUDA_OK(cudaEventCreate(&ev));
  RVT_CUDA_OK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  RVT_CUDA_OK(cudaStreamCreateWithFlags(&ctx->fin_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    RVT_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_swept[i], cudaEventDisableTiming));
    RVT_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_fin_done[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < rvt_ctx::kLandRing; ++i) {
    RVT_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_landed[i], cudaEventDisableTiming));
    RVT_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_unpacked[i], cudaEventDisableTiming));
  }
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->d_nm, sizeof(NullModel)));
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->d_counter, sizeof(unsigned int) * 4));
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->d_shift, sizeof(int) * (kMaxC + 1)));
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->d_status, sizeof(int)));
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->dbeta, sizeof(double) * kMaxC));
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->dnull_part, sizeof(double) * kNullBlocks * kNullAcc));
  RVT_CUDA_OK(cudaFuncSetAttribute(k_sweep_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, kSimtSmem));
  RVT_CUDA_OK(cudaFuncSetAttribute(k_finalize<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fin_smem(kTileRows, kMaxER, false)));
  RVT_CUDA_OK(cudaFuncSetAttribute(k_finalize<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fin_smem(kTileRows, kMaxER, true)));
  RVT_CUDA_OK(cudaFuncSetAttribute(k_finalize<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fin_smem(kTileRows, kMaxER, false)));
  int rc = tc_init(&ctx->tc, ctx->err, sizeof(ctx->err));
  if (rc) return rc;
  if (ctx->tc.encode && (rc = aug_init(ctx->err, sizeof(ctx->err)))) return rc;
  return RVT_OK;
}

void rvt_ctx_destroy(rvt_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  void* ptrs[] = {ctx->dX, ctx->dy, ctx->dresid, ctx->dnull_part, ctx->dbeta, ctx->dE, ctx->d_nm,
                  ctx->d_shift, ctx->d_status, ctx->d_genes, ctx->d_flags, ctx->d_userflags, ctx->d_af,
                  ctx->d_counts, ctx->d_parts, ctx->d_res, ctx->d_counter, ctx->d_stage64, ctx->d_loaded, ctx->d_stage, ctx->d_dbg, ctx->d_qags, ctx->d_jobs, ctx->d_mid, ctx->d_zero_flags, ctx->d_lfg, ctx->d_perm, ctx->d_lmm, ctx->d_lmm_tiles, ctx->d_lmm_vec, ctx->d_lmm_tsum, ctx->d_p, ctx->d_vw, ctx->d_dos_st, ctx->d_dos_tin, ctx->d_dos_idx, ctx->d_dos_afd, ctx->d_dos_tg, ctx->d_bs_st, ctx->d_bs_tin, ctx->d_bs_idx, ctx->d_bs_tg, ctx->d_frac};
  tc_destroy(&ctx->tc);
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (void* p : ctx->d_metab)
    if (p) cudaFree(p);
  for (void* p : ctx->d_scr)
    if (p) cudaFree(p);
  for (auto& ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& e : ctx->evpool) cudaEventDestroy(e);
  if (ctx->copy_stream) {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
  }
  if (ctx->bin_s) {
    cudaStreamSynchronize(ctx->bin_s);
    cudaStreamDestroy(ctx->bin_s);
  }
  if (ctx->ev_bin_in) cudaEventDestroy(ctx->ev_bin_in);
  if (ctx->ev_bin_out) cudaEventDestroy(ctx->ev_bin_out);
  if (ctx->fin_stream) {
    cudaStreamSynchronize(ctx->fin_stream);
    cudaStreamDestroy(ctx->fin_stream);
  }
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_swept[i]) cudaEventDestroy(ctx->ev_swept[i]);
    if (ctx->ev_fin_done[i]) cudaEventDestroy(ctx->ev_fin_done[i]);
  }
  for (int i = 0; i < rvt_ctx::kLandRing; ++i) {
    if (ctx->d_land[i]) cudaFree(ctx->d_land[i]);
    if (ctx->ev_landed[i]) cudaEventDestroy(ctx->ev_landed[i]);
    if (ctx->ev_unpacked[i]) cudaEventDestroy(ctx->ev_unpacked[i]);
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* rvt_last_error(const rvt_ctx* ctx) { return ctx ? ctx->err : "null context"; }

int rvt_set_option(rvt_ctx* ctx, const char* key, double value) {
  if (!ctx || !key) return RVT_E_BADARG;
  std::string k(key);
  if (k == "beta1") ctx->beta1 = value;
  else if (k == "beta2") ctx->beta2 = value;
  else if (k == "engine") {
    int e = (int)value;
    if (e < 0 || e > 2) CTX_FAIL(RVT_E_BADARG, "engine must be 0 (auto), 1 (simt) or 2 (tc)");
    ctx->engine = e;
  } else if (k == "splits") {
    if (value < 0 || value > 64) CTX_FAIL(RVT_E_BADARG, "splits must be in 0..64");
    ctx->splits = (int)value;
  } else if (k == "skato") {
    ctx->skato = value != 0;
  } else if (k == "overlap") {
    if (value < 0 || value > 64) CTX_FAIL(RVT_E_BADARG, "overlap must be in 0..64 (batches per flush; 0 = off)");
    ctx->overlap = (int)value;
  } else if (k == "tc_stages") {
    if (value != 3 && value != 4 && value != 5) CTX_FAIL(RVT_E_BADARG, "tc_stages must be 3, 4 or 5");
    ctx->tc.stages = (int)value;
  } else if (k == "fin_split") {
    ctx->fin_split = value != 0;
  } else if (k == "aug") {
    ctx->aug = value != 0;
  } else if (k == "meta_cov_scale") {
    if (value < 0) CTX_FAIL(RVT_E_BADARG, "meta_cov_scale must be >= 0");
    ctx->meta_cov_scale = value;
  } else if (k == "qags_pack") {
    ctx->qags_pack = value != 0;
  } else if (k == "watchdog_ms") {
    if (value < 0 || value > 3.6e6) CTX_FAIL(RVT_E_BADARG, "watchdog_ms must be in 0..3600000 (0 = off)");
    ctx->wd_cycles = (long long)(value * 2.0e6);   // ~2 GHz SM clock
  } else if (k == "f64_imputed") {
    ctx->f64_imputed = value != 0;
  } else if (k == "binary_stream") {
    ctx->bin_stream = value != 0;
  } else if (k == "bolt_binary") {
    ctx->bolt_binary = value != 0;
  } else if (k == "bolt_kernels") {
    ctx->bolt_kernels = (value >= 3.0) ? 3 : (value >= 2.0) ? 2 : 1;
  } else if (k == "skato_binary") {
    ctx->skato_binary = value != 0;
  } else if (k == "tc_boxes") {
    if (value != 2 && value != 4) CTX_FAIL(RVT_E_BADARG, "tc_boxes must be 2 or 4");
    ctx->tc.boxes = (int)value;
  } else if (k == "stream_batch") {
    if (value < 0 || value > 65536) CTX_FAIL(RVT_E_BADARG, "stream_batch must be in 0..65536");
    ctx->stream_batch = (int)value;
  } else if (k == "tc_zc") {
    ctx->tc.zc = value != 0;
  } else if (k == "tc_wide") {
    ctx->tc.wide = value != 0;
  } else if (k == "tc_debug_skip") {
    ctx->tc.dbg_skip = (int)value;
  } else if (k == "tc_l2promo") {
    if (value < 0 || value > 3) CTX_FAIL(RVT_E_BADARG, "tc_l2promo must be 0..3");
    ctx->tc.l2promo = (int)value;
    for (auto& sg : ctx->tc.seg)
      for (bool& b : sg.have_m) b = false;  // re-encode lazily
  } else if (k == "perm") {
    if (value < 0 || value > 1e9) CTX_FAIL(RVT_E_BADARG, "perm (nPerm) must be in 0..1e9");
    ctx->perm_n = (int)value;
  } else if (k == "perm_alpha") {
    ctx->perm_alpha = value;
  } else if (k == "perm_batch") {
    if (value < 16 || value > 4096 || ((int)value & 15)) CTX_FAIL(RVT_E_BADARG, "perm_batch must be a multiple of 16 in 16..4096");
    ctx->perm_batch = (int)value;
  } else if (k == "perm_stream_pos") {
    if (value < 0) CTX_FAIL(RVT_E_BADARG, "perm_stream_pos must be >= 0");
    ctx->perm_pos = (uint64_t)value;
    ctx->perm_stream_lost = false;
  } else if (k == "perm_seed") {
    ctx->perm_seed = (uint32_t)value;
    ctx->perm_pos = 0;
    ctx->perm_stream_lost = false;
  } else if (k == "debug_perm_q") {
    ctx->perm_log = value != 0;
  } else if (k == "debug_phases") {
    ctx->want_dbg = value != 0;
  } else
    CTX_FAIL(RVT_E_BADARG, "unknown option '%s'", key);
  return RVT_OK;
}

int rvt_set_stream(rvt_ctx* ctx, void* cuda_stream) {
  if (!ctx) return RVT_E_BADARG;
  RVT_CUDA_OK(cudaSetDevice(ctx->device));
  RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return RVT_OK;
}

double rvt_get_info(const rvt_ctx* ctx, const char* key) {
  if (!ctx || !key) return -1;
  std::string k(key);
  if (k == "sm_count") return ctx->sm_count;
  if (k == "last_engine") return ctx->last_engine;
  if (k == "last_splits") return ctx->last_S;
  if (k == "last_aug") return ctx->last_aug;
  if (k == "N") return (double)ctx->N;
  if (k == "C") return ctx->C;
  if (k == "ER") return ctx->ER;
  if (k == "loaded_ld") return (double)ctx->loaded_ld;
  if (k == "tc_available") return ctx->tc.encode ? 1 : 0;
  if (k == "perm_stream_pos") return (double)ctx->perm_pos;
  return -1;
}

static int null_model_run(rvt_ctx* ctx, bool keep_resid = false, double sigma2_given = -1.0) {
  const int64_t N = ctx->N;
  const int C = ctx->C;
  ctx->ER = ((4 * (C + 1)) + 15) & ~15;  // kind::i8 UMMA needs N = 64 + ER to be a multiple of 16
  ctx->ldE = (N + 127) & ~(int64_t)127;
  if (ctx->dresid) cudaFree(ctx->dresid);
  if (ctx->dE) cudaFree(ctx->dE);
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->dresid, sizeof(double) * N));
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->dE, (size_t)ctx->ER * ctx->ldE));
  RVT_CUDA_OK(cudaMemsetAsync(ctx->dE, 0, (size_t)ctx->ER * ctx->ldE, ctx->stream));
  NullModel h;
  memset(&h, 0, sizeof(h));
  h.N = N;
  h.C = C;
  h.ER = ctx->ER;
  h.ldE = ctx->ldE;
  h.E = ctx->dE;
  h.resid = ctx->dresid;
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->d_nm, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  k_null_moments<<<kNullBlocks, kNullThreads, 0, ctx->stream>>>(N, C, ctx->dX, ctx->dy, ctx->dnull_part);
  k_null_solve<<<1, 32, 0, ctx->stream>>>(C, kNullBlocks, ctx->dnull_part, ctx->d_nm, ctx->dbeta, ctx->d_status);
  k_null_resid<<<kNullBlocks, kNullThreads, 0, ctx->stream>>>(N, C, ctx->dX, ctx->dy, ctx->dbeta, ctx->dresid,
                                                             ctx->dnull_part, keep_resid ? 1 : 0);
  k_null_finish<<<1, 32, 0, ctx->stream>>>(N, C, kNullBlocks, ctx->dnull_part, ctx->d_nm, ctx->d_shift, sigma2_given);
  k_build_E<<<kNullBlocks, kNullThreads, 0, ctx->stream>>>(N, C, ctx->dX, ctx->dresid, ctx->d_shift, ctx->dE,
                                                          ctx->ldE, ctx->d_nm);
  RVT_CUDA_OK(cudaGetLastError());
  int status = 0;
  RVT_CUDA_OK(cudaMemcpyAsync(&status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  RVT_CUDA_OK(cudaMemcpyAsync(&ctx->h_nm, ctx->d_nm, sizeof(NullModel), cudaMemcpyDeviceToHost, ctx->stream));
  RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (status == 1) CTX_FAIL(RVT_E_NUMERIC, "null model: X'X is not positive definite");
  if (status == 2)
    CTX_FAIL(RVT_E_BADARG, "null model: column 0 of X must be the intercept (all ones), as produced by "
                           "copyCovariateAndIntercept (src/ModelUtil.h:102-130)");
  ctx->have_null = true;
  int rc = tc_bind_null(&ctx->tc, ctx->dE, ctx->ER, N, ctx->ldE, ctx->err, sizeof(ctx->err));
  if (rc) return rc;
  return RVT_OK;
}

static int null_model_alloc(rvt_ctx* ctx, int64_t N, int C) {
  if (N <= 0 || N > ((int64_t)1 << 31) - 256) CTX_FAIL(RVT_E_BADARG, "N out of range");
  if (C < 1 || C > kMaxC) CTX_FAIL(RVT_E_UNSUPPORTED, "C=%d covariate columns (incl. intercept); this build supports 1..%d", C, kMaxC);
  if (!ctx->genes.empty()) CTX_FAIL(RVT_E_STATE, "flush pending genes before changing the null model");
  RVT_CUDA_OK(cudaSetDevice(ctx->device));
  if (ctx->dX) cudaFree(ctx->dX);
  if (ctx->dy) cudaFree(ctx->dy);
  ctx->dX = ctx->dy = nullptr;
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->dX, sizeof(double) * N * C));
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->dy, sizeof(double) * N));
  ctx->N = N;
  ctx->C = C;
  ctx->have_null = false;
  ctx->binary = false;
  return RVT_OK;
}

// LogisticRegression::FitLogisticModel (regression/LogisticRegression.cpp:279-339): Newton rounds on the device (one
// fixed-order reduction per round), the C x C solve and the stopping rule on the host.  Leaves p, v = p(1-p) of the LAST
// round evaluated -- i.e. at the beta before the final update, exactly what GetPredicted / GetVariance return -- and
// covB = (X'VX)^-1 of that round.
static int logistic_null(rvt_ctx* ctx, double* covB /*C*C*/, double* vsum, double* xsum_w, double* rsum) {
  const int64_t N = ctx->N;
  const int C = ctx->C;
  cudaStream_t st = ctx->stream;
  for (double** p : {&ctx->d_p, &ctx->d_vw}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    RVT_CUDA_OK(cudaMalloc((void**)p, sizeof(double) * N));
  }
  double* d_part = nullptr;
  RVT_CUDA_OK(cudaMalloc((void**)&d_part, sizeof(double) * kNullBlocks * kLogitAcc));
  std::vector<double> part((size_t)kNullBlocks * kLogitAcc), beta(C, 0.0), D(C * C), r(C), L(C * C), dlt(C);
  int rounds = 0;
  double last = -99999, cur = -9999;
  bool ok = false;
  while (rounds < 100) {
    RVT_CUDA_OK(cudaMemcpyAsync(ctx->dbeta, beta.data(), sizeof(double) * C, cudaMemcpyHostToDevice, st));
    RVT_CUDA_OK(cudaMemsetAsync(d_part, 0, sizeof(double) * kNullBlocks * kLogitAcc, st));
    k_logit_round<<<kNullBlocks, kNullThreads, 0, st>>>(N, C, ctx->dX, ctx->dy, ctx->dbeta, ctx->d_p, ctx->d_vw, d_part);
    RVT_CUDA_OK(cudaMemcpyAsync(part.data(), d_part, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, st));
    RVT_CUDA_OK(cudaStreamSynchronize(st));
    double ll = 0.0;
    for (int l = 0; l < C; ++l) {
      r[l] = 0.0;
      for (int m = 0; m < C; ++m) D[l * C + m] = 0.0;
    }
    for (int b = 0; b < kNullBlocks; ++b) {
      const double* pb = part.data() + (size_t)b * kLogitAcc;
      for (int l = 0; l < C; ++l) {
        for (int m = 0; m < C; ++m) D[l * C + m] += pb[l * kMaxC + m];
        r[l] += pb[kMaxC * kMaxC + l];
      }
      ll += pb[kLogitAcc - 1];
    }
    // delta_beta = D.llt().solve(r)
    for (int i = 0; i < C; ++i)
      for (int j = 0; j <= i; ++j) {
        double sacc = D[i * C + j];
        for (int k = 0; k < j; ++k) sacc -= L[i * C + k] * L[j * C + k];
        if (i == j) {
          if (!(sacc > 0.0)) { cudaFree(d_part); CTX_FAIL(RVT_E_NUMERIC, "logistic null model: X'VX is not positive definite"); }
          L[i * C + i] = sqrt(sacc);
        } else
          L[i * C + j] = sacc / L[j * C + j];
      }
    for (int i = 0; i < C; ++i) {
      double sacc = r[i];
      for (int k = 0; k < i; ++k) sacc -= L[i * C + k] * dlt[k];
      dlt[i] = sacc / L[i * C + i];
    }
    for (int i = C - 1; i >= 0; --i) {
      double sacc = dlt[i];
      for (int k = i + 1; k < C; ++k) sacc -= L[k * C + i] * dlt[k];
      dlt[i] = sacc / L[i * C + i];
    }
    for (int l = 0; l < C; ++l) beta[l] += dlt[l];
    cur = -2.0 * ll;
    if (rounds > 1 && fabs(cur - last) < 1e-3) {
      ok = true;
      break;
    }
    if (!std::isnormal(cur)) break;   // "probably separation happens"
    last = cur;
    ++rounds;
  }
  cudaFree(d_part);
  if (!ok) CTX_FAIL(RVT_E_NUMERIC, "logistic null model did not converge (separation, or more than 100 rounds)");
  // covB = D^-1 through the Cholesky factor
  for (int col = 0; col < C; ++col) {
    std::vector<double> e(C, 0.0);
    e[col] = 1.0;
    for (int i = 0; i < C; ++i) {
      double sacc = e[i];
      for (int k = 0; k < i; ++k) sacc -= L[i * C + k] * e[k];
      e[i] = sacc / L[i * C + i];
    }
    for (int i = C - 1; i >= 0; --i) {
      double sacc = e[i];
      for (int k = i + 1; k < C; ++k) sacc -= L[k * C + i] * e[k];
      e[i] = sacc / L[i * C + i];
    }
    for (int i = 0; i < C; ++i) covB[i * C + col] = e[i];
  }
  for (int l = 0; l < C * C; ++l) ctx->h_xvx[l] = D[l];   // Z'WZ of that round = covZZ of MetaCovUnrelatedBinary
  *rsum = r[0];                                 // sum_i (y_i - p_i): not zero, p is one Newton step stale
  *vsum = D[0];                                 // column 0 of X is the intercept: D[0][l] = sum v x_l
  for (int l = 0; l < C; ++l) xsum_w[l] = D[l];
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->dbeta, beta.data(), sizeof(double) * C, cudaMemcpyHostToDevice, st));
  return RVT_OK;
}

int rvt_set_null_model(rvt_ctx* ctx, int64_t N, int C, const double* X, const double* y, int binary) {
  if (!ctx || !X || !y) return RVT_E_BADARG;
  int rc = null_model_alloc(ctx, N, C);
  if (rc) return rc;
  ctx->binary = false;
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->dX, X, sizeof(double) * N * C, cudaMemcpyHostToDevice, ctx->stream));
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->dy, y, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
  if (!binary) return null_model_run(ctx);
  // binary trait: r = y - p and v = p(1-p) from the logistic fit; the linear machinery then runs on r with sigma2 = 1 and
  // (X'VX)^-1 in place of (X'X)^-1 (src/Model.h:2673-2681: ynull = GetPredicted(), v = GetVariance())
  for (int64_t i = 0; i < N; ++i)
    if (y[i] != 0.0 && y[i] != 1.0) CTX_FAIL(RVT_E_BADARG, "binary trait: phenotype values must be 0 or 1");
  double covB[kMaxC * kMaxC], vsum = 0.0, xsw[kMaxC], rsum = 0.0;
  if ((rc = logistic_null(ctx, covB, &vsum, xsw, &rsum))) return rc;
  std::vector<double> beta(C);
  RVT_CUDA_OK(cudaMemcpy(beta.data(), ctx->dbeta, sizeof(double) * C, cudaMemcpyDeviceToHost));
  k_logit_resid<<<(unsigned)((N + 255) / 256), 256, 0, ctx->stream>>>(N, ctx->dy, ctx->d_p, ctx->dy);
  if ((rc = null_model_run(ctx, true, 1.0))) return rc;
  NullModel h = ctx->h_nm;
  memcpy(h.xtx_inv, covB, sizeof(double) * C * C);
  h.binary = 1;
  h.rsum = rsum;   // (the linear fit reports the sum of ITS residuals, ~0; the flip algebra of the fp64 path needs sum_i r_i)
  h.vw = ctx->d_vw;
  h.vsum_w = vsum;
  for (int l = 0; l < C; ++l) h.xsum_w[l] = xsw[l];
  RVT_CUDA_OK(cudaMemcpy(ctx->d_nm, &h, sizeof(h), cudaMemcpyHostToDevice));
  RVT_CUDA_OK(cudaMemcpy(ctx->dbeta, beta.data(), sizeof(double) * C, cudaMemcpyHostToDevice));   // (the linear fit overwrote it)
  ctx->h_nm = h;
  ctx->binary = true;
  return RVT_OK;
}

int rvt_set_null_residual(rvt_ctx* ctx, int64_t N, int C, const double* X, const double* resid, double sigma2) {
  if (!ctx || !X || !resid) return RVT_E_BADARG;
  if (!(sigma2 > 0.0)) CTX_FAIL(RVT_E_BADARG, "sigma2 must be positive");
  int rc = null_model_alloc(ctx, N, C);
  if (rc) return rc;
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->dX, X, sizeof(double) * N * C, cudaMemcpyHostToDevice, ctx->stream));
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->dy, resid, sizeof(double) * N, cudaMemcpyHostToDevice, ctx->stream));
  rc = null_model_run(ctx, true, sigma2);
  if (rc) return rc;
  // with a caller-supplied residual sum_i r_i is what it is (the linear fit reports the sum of ITS residuals, ~0): the flip
  // algebra of the fp64 path needs it.  It is the intercept component of X'r, which k_null_moments already reduced.
  double rs = 0.0;
  for (int64_t i = 0; i < N; ++i) rs += resid[i];
  NullModel h = ctx->h_nm;
  h.rsum = rs;
  RVT_CUDA_OK(cudaMemcpy(ctx->d_nm, &h, sizeof(h), cudaMemcpyHostToDevice));
  ctx->h_nm = h;
  return RVT_OK;
}

int rvt_set_null_model_dev(rvt_ctx* ctx, int64_t N, int C, const double* dX, const double* dy) {
  if (!ctx || !dX || !dy) return RVT_E_BADARG;
  int rc = null_model_alloc(ctx, N, C);
  if (rc) return rc;
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->dX, dX, sizeof(double) * N * C, cudaMemcpyDeviceToDevice, ctx->stream));
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->dy, dy, sizeof(double) * N, cudaMemcpyDeviceToDevice, ctx->stream));
  return null_model_run(ctx);
}

int rvt_get_null_model(rvt_ctx* ctx, double* resid, double* sigma2, double* xtx_inv) {
  if (!ctx) return RVT_E_BADARG;
  if (!ctx->have_null) CTX_FAIL(RVT_E_STATE, "no null model set");
  if (sigma2) *sigma2 = ctx->h_nm.sigma2;
  if (xtx_inv) memcpy(xtx_inv, ctx->h_nm.xtx_inv, sizeof(double) * ctx->C * ctx->C);
  if (resid) {
    RVT_CUDA_OK(cudaMemcpyAsync(resid, ctx->dresid, sizeof(double) * ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
    RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  }
  return RVT_OK;
}

int rvt_get_null_beta(rvt_ctx* ctx, double* beta) {
  if (!ctx || !beta) return RVT_E_BADARG;
  if (!ctx->have_null) CTX_FAIL(RVT_E_STATE, "no null model set");
  RVT_CUDA_OK(cudaMemcpyAsync(beta, ctx->dbeta, sizeof(double) * ctx->C, cudaMemcpyDeviceToHost, ctx->stream));
  RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return RVT_OK;
}

// common tail of every push: append the descriptor and the per-variant side data
static int push_common(rvt_ctx* ctx, const int8_t* dG, int M, int64_t ld, const double* af, const uint8_t* flags,
                       bool counted, int seg, int64_t row0, bool tiled, int slots = 0 /* per-variant slots owned; 0: M */) {
  if (slots <= 0) slots = M;
  GeneDesc gd;
  memset(&gd, 0, sizeof(gd));
  gd.g = dG;
  gd.ld = ld;
  gd.M = M;
  gd.seg = seg;
  gd.row0 = row0;
  gd.var0 = ctx->n_var;
  gd.has_af = af ? 1 : 0;
  gd.counted = counted ? 1 : 0;
  gd.row0_b = row0;
  gd.Mb = M;
  gd.tiled = tiled ? 1 : 0;
  gd.var0_b = ctx->n_var;
  ctx->genes.push_back(gd);
  ctx->count_slot.push_back(counted ? ctx->n_var : -1);
  ctx->slots.push_back(slots);
  for (int j = 0; j < slots; ++j) {
    ctx->userflags.push_back(flags ? flags[j] : (uint8_t)0xFF);
    ctx->af.push_back(af ? af[j] : 0.0);
  }
  ctx->n_var += slots;
  return RVT_OK;
}

static int push_check(rvt_ctx* ctx, int M, int max_m = kMaxM) {
  if (!ctx->have_null) CTX_FAIL(RVT_E_STATE, "set the null model before pushing genes");
  if (M < 1) CTX_FAIL(RVT_E_BADARG, "gene with no variant (the reference returns -1 / NA: src/Model.h:2637-2640)");
  if (M > max_m) CTX_FAIL(RVT_E_UNSUPPORTED, "M=%d variants; this entry point handles genes of up to %d variants", M, max_m);
  // what the flush could not handle is refused HERE, while the queue is still consistent (a flush that failed half-way
  // used to leave the context stuck: ADVICE r01)
  RVT_CUDA_OK(cudaSetDevice(ctx->device));
  return RVT_OK;
}

// Stage room for a pushed gene: one tiled block if it fits a tensor-core tile, else T = ceil(M/64) blocks of
// consecutive variants, balanced (wide.cuh).
static int stage_tiles(rvt_ctx* ctx, int M, TilePlan* tp) {
  const int T = (M + kTileRows - 1) / kTileRows;
  tp->T = T;
  tp->off.resize(T);
  tp->rows.resize(T);
  tp->r0.resize(T);
  int r0 = 0;
  for (int t = 0; t < T; ++t) {
    const int rows = M / T + (t < M % T ? 1 : 0);
    int rc = stage_alloc(ctx, tiled_bytes(ctx->N, rows), &tp->off[t]);
    if (rc) return rc;
    tp->rows[t] = rows;
    tp->r0[t] = r0;
    r0 += rows;
  }
  return RVT_OK;
}

// append a staged gene: an ordinary descriptor, or (T > 1) a wide gene whose placeholder descriptor is its tile 0
static int push_staged(rvt_ctx* ctx, const TilePlan& tp, int M, const double* af) {
  if (tp.T == 1) return push_common(ctx, ctx->d_stage + tp.off[0], M, 0, af, nullptr, true, kSegStaged, tp.off[0] / 128, true);
  WideGene w;
  w.gene_index = (int)ctx->genes.size();
  w.M = M;
  w.var0 = ctx->n_var;
  w.has_af = af != nullptr;
  for (int t = 0; t < tp.T; ++t) {
    GeneDesc gd;
    memset(&gd, 0, sizeof(gd));
    gd.g = ctx->d_stage + tp.off[t];
    gd.M = gd.Mb = tp.rows[t];
    gd.seg = kSegStaged;
    gd.row0 = gd.row0_b = tp.off[t] / 128;
    gd.var0 = gd.var0_b = ctx->n_var + tp.r0[t];
    gd.counted = 1;
    gd.tiled = 1;
    w.tiles.push_back(gd);
  }
  ctx->wide.push_back(w);
  return push_common(ctx, ctx->d_stage + tp.off[0], tp.rows[0], 0, af, nullptr, true, kSegStaged, tp.off[0] / 128, true, M);
}

static void launch_count(rvt_ctx* ctx, const int8_t* d, int M, int64_t ld, RowCounts* counts) {
  const int64_t per_row = (ctx->N + 16 * 256 - 1) / (16 * 256);
  dim3 grid((unsigned)per_row, (unsigned)M);
  k_count_rows<<<grid, 256, 0, ctx->stream>>>(d, ld, ctx->N, counts);
}

// next landing buffer of the host->device ring (>= need bytes); the copy stream waits until the kernel
// that consumed its previous content has run
static int land_acquire(rvt_ctx* ctx, size_t need, int* slot) {
  if (need > ctx->cap_land) {
    RVT_CUDA_OK(cudaStreamSynchronize(ctx->copy_stream));
    RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    for (auto& p : ctx->d_land) {
      if (p) cudaFree(p);
      p = nullptr;
    }
    for (auto& p : ctx->d_land) RVT_CUDA_OK(cudaMalloc((void**)&p, need));
    ctx->cap_land = need;
  }
  *slot = (int)(ctx->land_seq++ % rvt_ctx::kLandRing);
  RVT_CUDA_OK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_unpacked[*slot], 0));
  return RVT_OK;
}
// the copy is enqueued: make the context stream wait for it
static int land_publish(rvt_ctx* ctx, int slot) {
  RVT_CUDA_OK(cudaEventRecord(ctx->ev_landed[slot], ctx->copy_stream));
  RVT_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->ev_landed[slot], 0));
  return RVT_OK;
}
// the consuming kernel is enqueued: the slot may be overwritten once it has run; with option
// "stream_batch" also enqueue sweep + statistics once enough genes are waiting
static int land_release(rvt_ctx* ctx, int slot) {
  RVT_CUDA_OK(cudaEventRecord(ctx->ev_unpacked[slot], ctx->stream));
  return RVT_OK;
}
static int launch_range(rvt_ctx* ctx, int g0, int g1);
// SKAT-O quadrature of nb genes (jobs[i] -> record d_res[out_index ? out_index[i] : i]).  One launch for as many genes as
// possible: the kernel is latency-bound (serial Davies evaluations), so what it needs is resident warps -- 2 500 genes in
// one launch run at 6 CTAs per SM where two launches of 1 250 ran at 3 (profiles/r02g_qags_ncu.txt).
static size_t qags_scratch_entries(const rvt_ctx* ctx, int nb) {
  if (!ctx->qags_pack) return (size_t)std::min(nb, 2048);
  return (size_t)std::max(1, std::min((nb + kQagsSlots - 1) / kQagsSlots, 6 * ctx->sm_count)) * kQagsSlots;
}
static int launch_qags(rvt_ctx* ctx, const SkatoJob* jobs, int nb, rvt_gene_result* d_res, const int* d_index, cudaStream_t st) {
  if (nb <= 0) return RVT_OK;
  if (!ctx->qags_pack) {
    for (int b0 = 0; b0 < nb; b0 += 2048) {   // one interval list per gene: bounded scratch
      const int n = std::min(2048, nb - b0);
      k_skato_qags<<<n, kQagsThreads, 0, st>>>(jobs + b0, n, ctx->d_qags, d_index ? d_res : d_res + b0, d_index ? d_index + b0 : nullptr,
                                               ctx->wd_cycles);
    }
  } else {
    const int grid = (int)(qags_scratch_entries(ctx, nb) / kQagsSlots);
    RVT_CUDA_OK(cudaMemsetAsync(ctx->d_counter + 1, 0, sizeof(unsigned int), st));
    k_skato_qags_packed<<<grid, kQagsPackThreads, 0, st>>>(jobs, nb, ctx->d_qags, d_res, d_index, ctx->wd_cycles, ctx->d_counter + 1);
  }
  RVT_CUDA_OK(cudaGetLastError());
  return RVT_OK;
}

// Binary trait: the fp64 statistics (dosage.cuh) and the tail of the tile genes of [b0, b0 + nb), enqueued now -- right behind
// their copies -- instead of at flush, where they used to run after the last byte had crossed PCIe (profiles/r02u: copies
// 125 us + statistics 63 us per gene in series = 5.3 k genes/s end to end against 8.1 k for a quantitative trait).  Same
// kernels, same records; genes this does not take (doubles with dosages, blocks without the engine's own counts, wide genes)
// keep the flush path.
static int binary_stream_batch(rvt_ctx* ctx, int b0, int nb, const EngineParams& prm, int* launches) {
  const int64_t N = ctx->N;
  int rc;
  // On a stream of their own: behind the context's stream the 2 ms of statistics of a 32-gene batch held up the unpack
  // kernels of the following genes, the landing ring filled and the copy engine idled (measured: no gain at all).
  if (!ctx->bin_s) {
    RVT_CUDA_OK(cudaStreamCreateWithFlags(&ctx->bin_s, cudaStreamNonBlocking));
    RVT_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_bin_in, cudaEventDisableTiming));
    RVT_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_bin_out, cudaEventDisableTiming));
  }
  RVT_CUDA_OK(cudaEventRecord(ctx->ev_bin_in, ctx->stream));      // tiles unpacked, counts / flags / af of this range in place
  cudaStream_t st = ctx->bin_s;
  RVT_CUDA_OK(cudaStreamWaitEvent(st, ctx->ev_bin_in, 0));
  ctx->bin_streamed.resize(ctx->genes.size(), 0);
  ctx->is_bed.resize(ctx->genes.size(), 0);
  ctx->bs_tg.clear();
  ctx->bs_idx.clear();
  for (int g = b0; g < b0 + nb; ++g) {
    const GeneDesc& gd = ctx->genes[g];
    if (!gd.tiled || !gd.counted || gd.M > kMaxM || ctx->slots[g] > kMaxM) continue;
    bool f64 = false;
    for (auto& dg : ctx->dos) f64 |= dg.gene_index == g;
    for (auto& w : ctx->wide) f64 |= w.gene_index == g;
    if (f64) continue;
    TileGene tg;
    tg.g = gd.g;
    tg.M = gd.M;
    tg.has_af = gd.has_af;
    tg.var0 = gd.var0;
    tg.slot = (int)ctx->bs_tg.size();
    tg.allow_missing = ctx->is_bed[g] ? 1 : 0;
    ctx->bs_tg.push_back(tg);
    ctx->bs_idx.push_back(g);
    ctx->bin_streamed[g] = 1;
  }
  const int nt = (int)ctx->bs_tg.size();
  if (nt == 0) return RVT_OK;
  if ((size_t)nt > ctx->cap_bs_st || (size_t)nt > ctx->cap_bs_tin || (size_t)nt > ctx->cap_bs_idx || (size_t)nt > ctx->cap_bs_tg ||
      (size_t)nt > ctx->cap_jobs || (size_t)nt > ctx->cap_mid || qags_scratch_entries(ctx, nt) > ctx->cap_qags)
    RVT_CUDA_OK(cudaStreamSynchronize(st));   // a buffer is about to move: nothing of the previous batch may still read it
  if ((rc = ensure(ctx, &ctx->d_bs_st, &ctx->cap_bs_st, (size_t)nt, sizeof(DosageStats)))) return rc;
  if ((rc = ensure(ctx, &ctx->d_bs_tin, &ctx->cap_bs_tin, (size_t)nt, sizeof(TailInput)))) return rc;
  if ((rc = ensure(ctx, &ctx->d_bs_idx, &ctx->cap_bs_idx, (size_t)nt, sizeof(int)))) return rc;
  if ((rc = ensure(ctx, &ctx->d_bs_tg, &ctx->cap_bs_tg, (size_t)nt, sizeof(TileGene)))) return rc;
  DosageStats* d_st = (DosageStats*)ctx->d_bs_st;
  TailInput* d_tin = (TailInput*)ctx->d_bs_tin;
  int* d_idx = (int*)ctx->d_bs_idx;
  TileGene* d_tg = (TileGene*)ctx->d_bs_tg;
  RVT_CUDA_OK(cudaMemsetAsync(d_st, 0, sizeof(DosageStats) * nt, st));
  RVT_CUDA_OK(cudaMemcpyAsync(d_tg, ctx->bs_tg.data(), sizeof(TileGene) * nt, cudaMemcpyHostToDevice, st));   // (pageable: staged before the call returns)
  RVT_CUDA_OK(cudaMemcpyAsync(d_idx, ctx->bs_idx.data(), sizeof(int) * nt, cudaMemcpyHostToDevice, st));
  k_tile_cols<<<nt, kTileRows, 0, st>>>(d_tg, nt, N, ctx->d_counts, d_st);
  const int64_t nblk = ((N + 3) / 4 + kSparseThreads - 1) / kSparseThreads;
  const unsigned bx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(nblk, (8 * (int64_t)ctx->sm_count + nt - 1) / nt));
  k_tile_sparse<<<dim3(bx, (unsigned)nt), kSparseThreads, 0, st>>>(d_tg, N, ctx->d_counts, ctx->dX, ctx->C, ctx->dresid, ctx->d_vw, d_st);
  k_tile_prepare<<<nt, 64, 0, st>>>(d_tg, nt, d_st, ctx->d_af, ctx->d_nm, prm, d_tin);
  const bool sk = ctx->skato && ctx->skato_binary;
  const int kld = fin_kld(kTileRows), fsm = fin_smem(kTileRows, ctx->ER, sk), wm_off = kTileRows * kld * 8;
  if (sk) {
    if ((rc = ensure(ctx, (void**)&ctx->d_qags, &ctx->cap_qags, qags_scratch_entries(ctx, nt), sizeof(QagsScratch)))) return rc;
    if ((rc = ensure(ctx, (void**)&ctx->d_jobs, &ctx->cap_jobs, (size_t)nt, sizeof(SkatoJob)))) return rc;
    k_finalize<true><<<nt, kFinThreadsSkato, fsm, st>>>(nullptr, nt, kld, wm_off, fin_uk_off(kTileRows, ctx->ER, sk), ctx->d_flags, ctx->d_af, ctx->d_counts,
                                                         ctx->d_nm, prm, 1, nullptr, ctx->d_res, nullptr, ctx->d_jobs, d_tin, d_idx);
    if ((rc = launch_qags(ctx, ctx->d_jobs, nt, ctx->d_res, d_idx, st))) return rc;
    *launches += 1;
  } else if (ctx->fin_split) {
    if ((rc = ensure(ctx, (void**)&ctx->d_mid, &ctx->cap_mid, (size_t)nt, sizeof(FinMid)))) return rc;
    k_finalize<false, true><<<nt, kFinThreads, fsm, st>>>(nullptr, nt, kld, wm_off, fin_uk_off(kTileRows, ctx->ER, false), ctx->d_flags, ctx->d_af, ctx->d_counts,
                                                           ctx->d_nm, prm, 1, nullptr, ctx->d_res, nullptr, nullptr, d_tin, d_idx, ctx->d_mid);
    k_fin_sturm<<<nt, kFinThreads, 0, st>>>(ctx->d_mid, nt);
    k_fin_tail<<<nt, kFinThreads, 0, st>>>(ctx->d_mid, nt, ctx->d_nm, ctx->d_res, d_idx);
    *launches += 2;
  } else {
    k_finalize<false><<<nt, kFinThreads, fsm, st>>>(nullptr, nt, kld, wm_off, fin_uk_off(kTileRows, ctx->ER, sk), ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm,
                                                     prm, 1, nullptr, ctx->d_res, nullptr, nullptr, d_tin, d_idx);
  }
  RVT_CUDA_OK(cudaGetLastError());
  RVT_CUDA_OK(cudaEventRecord(ctx->ev_bin_out, st));
  ctx->bin_pending = true;     // flush waits for ev_bin_out before it reads the records
  *launches += 4;
  return RVT_OK;
}

static int maybe_stream(rvt_ctx* ctx) {
  const int n = (int)ctx->genes.size();
  // binary trait: the statistics of a batch occupy the SMs for ~60 us per gene; small batches keep them out of the way of the
  // unpack kernels that free the landing ring (8 / 32 / 64 genes: 6.85 / 6.20 / 5.80 k genes/s end to end, profiles/r02u)
  const int sb = (ctx->binary && ctx->bin_stream) ? std::min(ctx->stream_batch, 8) : ctx->stream_batch;
  if (ctx->stream_batch > 0 && n - ctx->launched >= sb) return launch_range(ctx, ctx->launched, n);
  return RVT_OK;
}

int rvt_gene_push_f64(rvt_ctx* ctx, const double* G, int M, const double* af) {
  if (!ctx || !G) return RVT_E_BADARG;
  int rc = push_check(ctx, M, kWideMaxM);
  if (rc) return rc;
  const int64_t N = ctx->N, npad = (N + 127) & ~(int64_t)127;
  size_t need = (size_t)N * M;
  if (need > ctx->cap_stage64) {
    if (ctx->d_stage64) {
      RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->d_stage64);
    }
    RVT_CUDA_OK(cudaMalloc((void**)&ctx->d_stage64, need * sizeof(double)));
    ctx->cap_stage64 = need;
  }
  TilePlan tp;
  const int64_t stage_mark = ctx->stage_used;
  if ((rc = stage_tiles(ctx, M, &tp))) return rc;
  if ((rc = ensure_var(ctx, ctx->n_var + M))) return rc;
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->d_stage64, G, need * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  RVT_CUDA_OK(cudaMemsetAsync(ctx->d_counts + ctx->n_var, 0, sizeof(RowCounts) * M, ctx->stream));
  if (!ctx->d_frac) RVT_CUDA_OK(cudaMalloc((void**)&ctx->d_frac, sizeof(unsigned long long) * 2 * kWideMaxM));
  {
    std::vector<unsigned long long> init(2 * (size_t)M);
    for (int j = 0; j < M; ++j) { init[2 * j] = ~0ull; init[2 * j + 1] = 0ull; }
    RVT_CUDA_OK(cudaMemcpyAsync(ctx->d_frac, init.data(), sizeof(unsigned long long) * 2 * M, cudaMemcpyHostToDevice, ctx->stream));   // (pageable: staged)
  }
  for (int t = 0; t < tp.T; ++t) {
    dim3 grid((unsigned)((npad / 4 + 255) / 256), (unsigned)tp.rows[t]);
    k_pack_f64<<<grid, 256, 0, ctx->stream>>>(ctx->d_stage64 + (size_t)tp.r0[t] * N, N, ctx->d_stage + tp.off[t], tp.rows[t],
                                              ctx->d_counts + ctx->n_var + tp.r0[t], ctx->d_frac + 2 * (size_t)tp.r0[t]);
  }
  RVT_CUDA_OK(cudaGetLastError());
  // the staging buffer is reused by the next push: wait (pageable H2D is synchronous anyway);
  // the row counts also tell whether this gene holds anything but hard calls
  std::vector<RowCounts> rc_host(M);
  RVT_CUDA_OK(cudaMemcpyAsync(rc_host.data(), ctx->d_counts + ctx->n_var, sizeof(RowCounts) * M, cudaMemcpyDeviceToHost, ctx->stream));
  RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  bool dosage = false;
  for (int j = 0; j < M; ++j) dosage |= rc_host[j].bad > 0;
  if (dosage && ctx->f64_imputed) {
    // Is every non-integer entry the mean-imputed value of its column?  Then this is a hard-call gene with missing calls --
    // what DataConsolidator hands to fit() under the default --impute -- and it takes the integer paths of a 2-bit push with
    // code 01 (the tiles hold code 3 there): augmented sweep, wide operand tiles, permutation test included.
    std::vector<unsigned long long> fr(2 * (size_t)M);
    RVT_CUDA_OK(cudaMemcpy(fr.data(), ctx->d_frac, sizeof(unsigned long long) * 2 * M, cudaMemcpyDeviceToHost));
    bool pattern = true;
    for (int j = 0; j < M && pattern; ++j) {
      if (rc_host[j].bad == 0) continue;
      const long long nobs = N - rc_host[j].bad;
      const double ac = (double)((long long)rc_host[j].n1 + 2ll * rc_host[j].n2);
      const double fill = nobs > 0 ? 2.0 * (ac / (double)(2 * nobs)) : 0.0;
      double v;
      memcpy(&v, &fr[2 * j], sizeof(v));
      pattern = fr[2 * j] == fr[2 * j + 1] && fr[2 * j] != ~0ull && fabs(v - fill) <= 1e-12 * std::max(1.0, fill);
    }
    if (pattern) {
      dosage = false;
      ctx->bed_genes.push_back((int)ctx->genes.size());
      ctx->is_bed.resize(ctx->genes.size() + 1, 0);
      ctx->is_bed[ctx->genes.size()] = 1;
    }
  }
  if (dosage && M > kMaxM) {
    // real dosages in a gene wider than a tile: the matrix stays on the device for the dense fp64 statistics of run_wide
    // (k_wide_dos_*); the staged tiles only keep the gene's slot
    (void)stage_mark;
    double* keep = ctx->d_stage64;
    ctx->d_stage64 = nullptr;
    ctx->cap_stage64 = 0;
    rc = push_staged(ctx, tp, M, af);
    if (rc) {
      cudaFree(keep);
      return rc;
    }
    ctx->wide.back().dG = keep;
    return RVT_OK;
  }
  if (dosage) {
    // dosages / mean-imputed values: keep the fp64 matrix for the generic path and hand the
    // staging buffer over to it (the next push allocates a fresh one)
    DosGene dg;
    dg.gene_index = (int)ctx->genes.size();
    dg.M = M;
    dg.dG = ctx->d_stage64;
    dg.has_af = af != nullptr;
    if (af) dg.af.assign(af, af + M);
    ctx->dos.push_back(dg);
    ctx->d_stage64 = nullptr;
    ctx->cap_stage64 = 0;
  }
  return push_staged(ctx, tp, M, af);
}

int rvt_gene_push_i8(rvt_ctx* ctx, const int8_t* G, int M, int64_t ld_in, const double* af) {
  if (!ctx || !G) return RVT_E_BADARG;
  int rc = push_check(ctx, M, kWideMaxM);
  if (rc) return rc;
  const int64_t N = ctx->N, npad = (N + 127) & ~(int64_t)127;
  if (ld_in < N) CTX_FAIL(RVT_E_BADARG, "ld (%lld) < N (%lld)", (long long)ld_in, (long long)N);
  int slot = 0;
  if ((rc = land_acquire(ctx, (size_t)std::max(M, kMaxM) * N, &slot))) return rc;   // one size serves every ordinary gene of this cohort
  TilePlan tp;
  if ((rc = stage_tiles(ctx, M, &tp))) return rc;
  if ((rc = ensure_var(ctx, ctx->n_var + M))) return rc;
  // land the caller's variant-major rows (copy stream), then re-tile and count them on the device
  int8_t* land = ctx->d_land[slot];
  if (ld_in == N)
    RVT_CUDA_OK(cudaMemcpyAsync(land, G, (size_t)M * N, cudaMemcpyHostToDevice, ctx->copy_stream));
  else
    RVT_CUDA_OK(cudaMemcpy2DAsync(land, N, G, ld_in, N, M, cudaMemcpyHostToDevice, ctx->copy_stream));
  if ((rc = land_publish(ctx, slot))) return rc;
  RVT_CUDA_OK(cudaMemsetAsync(ctx->d_counts + ctx->n_var, 0, sizeof(RowCounts) * M, ctx->stream));
  for (int t = 0; t < tp.T; ++t) {
    dim3 grid((unsigned)((npad / 16 + 255) / 256), (unsigned)tp.rows[t]);
    k_tile_rows<<<grid, 256, 0, ctx->stream>>>(land + (size_t)tp.r0[t] * N, N, N, ctx->d_stage + tp.off[t], tp.rows[t],
                                               ctx->d_counts + ctx->n_var + tp.r0[t]);
  }
  RVT_CUDA_OK(cudaGetLastError());
  if ((rc = land_release(ctx, slot))) return rc;
  if ((rc = push_staged(ctx, tp, M, af))) return rc;
  return maybe_stream(ctx);
}

int rvt_gene_push_bed(rvt_ctx* ctx, const uint8_t* bed, int M, int64_t stride, const double* af) {
  if (!ctx || !bed) return RVT_E_BADARG;
  int rc = push_check(ctx, M, kWideMaxM);
  if (rc) return rc;
  const int64_t N = ctx->N, npad = (N + 127) & ~(int64_t)127;
  const int64_t rowb = (N + 3) / 4, pitch = (rowb + 3) & ~(int64_t)3;
  if (stride < rowb) CTX_FAIL(RVT_E_BADARG, "stride (%lld) < ceil(N/4) (%lld)", (long long)stride, (long long)rowb);
  int slot = 0;
  if ((rc = land_acquire(ctx, (size_t)std::max(M, kMaxM) * pitch, &slot))) return rc;
  TilePlan tp;
  if ((rc = stage_tiles(ctx, M, &tp))) return rc;
  if ((rc = ensure_var(ctx, ctx->n_var + M))) return rc;
  // 2 bits per call over PCIe (a quarter of the int8 form) on the copy stream; expanded, re-tiled and
  // counted on the device
  int8_t* land = ctx->d_land[slot];
  if (stride == pitch)   // the last row holds only rowb valid bytes (e.g. the tail of a mapped .bed file)
    RVT_CUDA_OK(cudaMemcpyAsync(land, bed, (size_t)(M - 1) * pitch + rowb, cudaMemcpyHostToDevice, ctx->copy_stream));
  else
    RVT_CUDA_OK(cudaMemcpy2DAsync(land, pitch, bed, stride, rowb, M, cudaMemcpyHostToDevice, ctx->copy_stream));
  if ((rc = land_publish(ctx, slot))) return rc;
  RVT_CUDA_OK(cudaMemsetAsync(ctx->d_counts + ctx->n_var, 0, sizeof(RowCounts) * M, ctx->stream));
  for (int t = 0; t < tp.T; ++t) {
    dim3 grid((unsigned)((npad / 16 + 255) / 256), (unsigned)tp.rows[t]);
    k_unpack_bed<<<grid, 256, 0, ctx->stream>>>(reinterpret_cast<const uint8_t*>(land) + (size_t)tp.r0[t] * pitch, pitch, N,
                                                ctx->d_stage + tp.off[t], tp.rows[t], ctx->d_counts + ctx->n_var + tp.r0[t]);
  }
  RVT_CUDA_OK(cudaGetLastError());
  if ((rc = land_release(ctx, slot))) return rc;
  ctx->bed_genes.push_back((int)ctx->genes.size());
  ctx->is_bed.resize(ctx->genes.size() + 1, 0);
  ctx->is_bed[ctx->genes.size()] = 1;
  if ((rc = push_staged(ctx, tp, M, af))) return rc;
  return maybe_stream(ctx);
}

// genes that arrived as 2-bit rows may hold missing calls (code 01): the counts say which; those are
// mean-imputed on the device (DataConsolidator::imputeGenotypeToMean) and handed to the fp64 path
static int resolve_bed_missing(rvt_ctx* ctx) {
  if (ctx->bed_genes.empty()) return RVT_OK;
  std::vector<RowCounts> hc((size_t)ctx->n_var);
  RVT_CUDA_OK(cudaMemcpyAsync(hc.data(), ctx->d_counts, sizeof(RowCounts) * ctx->n_var, cudaMemcpyDeviceToHost, ctx->stream));
  RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  for (int gi : ctx->bed_genes) {
    const GeneDesc& gd = ctx->genes[gi];
    bool missing = false;
    for (int j = 0; j < ctx->slots[gi]; ++j) missing |= hc[gd.var0 + j].bad > 0;
    if ((size_t)gi < ctx->bin_streamed.size() && ctx->bin_streamed[gi]) {
      ctx->bin_streamed[gi] = missing ? 2 : 1;   // computed behind its copy (imputed on the fly); 2: the permutation test does not cover it
      continue;
    }
    if (!missing) continue;
    if (ctx->slots[gi] > kMaxM) {
      // a wide gene with missing calls: the tensor-core pair sweep on the rows [H ; Mi] of every tile (run_wide, wide.cuh).
      // Without the tensor-core engine (or for a binary trait, refused at push) this ONE gene is reported
      // RVT_GENE_UNSUPPORTED in its record; the rest of the batch is computed
      const bool can = ctx->binary || (ctx->tc.encode && ctx->tc.have_e);   // (binary: k_wide_sparse imputes code 3 on the fly)
      for (size_t w = 0; w < ctx->wide.size(); ++w)
        if (ctx->wide[w].gene_index == gi) {
          if (can) {
            ctx->wide[w].imputed = true;
          } else {
            ctx->wide.erase(ctx->wide.begin() + (long)w);
          }
          break;
        }
      if (!can) ctx->unsupported.push_back(gi);
      continue;
    }
    DosGene dg;   // stays in its int8 tiles: imputed on the fly by the tile statistics kernels (dosage.cuh)
    dg.gene_index = gi;
    dg.M = gd.M;
    dg.dG = nullptr;
    dg.from_bed = true;
    dg.has_af = gd.has_af != 0;
    if (dg.has_af) dg.af.assign(ctx->af.begin() + gd.var0, ctx->af.begin() + gd.var0 + gd.M);
    ctx->dos.push_back(dg);
  }
  ctx->bed_genes.clear();
  return RVT_OK;
}

int rvt_gene_push_dev_i8(rvt_ctx* ctx, const int8_t* dG, int M, int64_t ld, const double* af, const uint8_t* flags) {
  if (!ctx || !dG) return RVT_E_BADARG;
  int rc = push_check(ctx, M);
  if (rc) return rc;
  if (ld < ctx->N || (ld & 15) || ((uintptr_t)dG & 15))
    CTX_FAIL(RVT_E_BADARG, "device block must be 16-byte aligned with ld a multiple of 16 and >= N");
  if (ctx->binary) CTX_FAIL(RVT_E_UNSUPPORTED, "binary trait: caller-owned device blocks are not supported (push host rows instead)");
  if ((rc = ensure_var(ctx, ctx->n_var + M))) return rc;
  if (!flags) {
    RVT_CUDA_OK(cudaMemsetAsync(ctx->d_counts + ctx->n_var, 0, sizeof(RowCounts) * M, ctx->stream));
    launch_count(ctx, dG, M, ld, ctx->d_counts + ctx->n_var);
    RVT_CUDA_OK(cudaGetLastError());
  }
  return push_common(ctx, dG, M, ld, af, flags, flags == nullptr, -1, 0, false);
}

int rvt_pending(const rvt_ctx* ctx) { return ctx ? (int)ctx->genes.size() : 0; }

// user flags override; otherwise derive from counts.  One thread per variant.
__global__ void k_resolve_flags(int64_t n_var, int64_t N, const RowCounts* __restrict__ counts,
                                const uint8_t* __restrict__ user, uint8_t* __restrict__ flags) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_var) return;
  if (user[r] != 0xFF) {
    flags[r] = user[r];
    return;
  }
  const long long n1 = counts[r].n1, n2 = counts[r].n2, n0 = N - n1 - n2, c = n1 + 2 * n2;
  uint8_t f = (c > N) ? kRowFlipped : kRowNormal;
  if (n0 == N || n1 == N || n2 == N) f = kRowSkip;
  flags[r] = f;
}

// splits of the sample axis: S units per gene, `chunk` samples each (a multiple of 512 = one TMA
// stage of the tensor-core kernel = 2 simt tiles).  A unit must stay inside the int32 accumulation
// bounds (2^22 samples for the Gram; 2^18 for the burden rows that ride on the UMMA); beyond that, only
// as many splits as it takes to give every SM ~8 units: each extra split costs one more partial
// (25 KB written by the sweep, read back by the statistics kernel) per gene.
static int split_plan(rvt_ctx* ctx, int n_genes, int* S_out, int64_t* chunk_out) {
  const int64_t N = ctx->N;
  int S = ctx->splits;
  if (S <= 0) {
    const int64_t s_min = (N + 262143) / 262144;
    const int64_t s_fill = (8 * (int64_t)ctx->sm_count + n_genes - 1) / std::max(1, n_genes);
    S = (int)std::min<int64_t>(std::max<int64_t>(s_min, std::min<int64_t>(s_fill, (N + 65535) / 65536)), std::max<int64_t>(16, s_min));
    S = std::max(S, 1);
  }
  int64_t chunk = (((N + S - 1) / S) + 511) & ~(int64_t)511;
  S = (int)((N + chunk - 1) / chunk);
  if (chunk > ((int64_t)1 << 22)) CTX_FAIL(RVT_E_UNSUPPORTED, "split of %lld samples exceeds the int32 accumulation bound; raise 'splits'", (long long)chunk);
  *S_out = S;
  *chunk_out = chunk;
  return RVT_OK;
}

// Enqueue sweep + statistics for the pending genes [g0, g1) on the context stream; records land in
// ctx->d_res[g0..g1).  Nothing here waits for the device: with option "stream_batch" = B the pushes call
// this every B genes, so the kernels of one range run while the next genes are still crossing PCIe on the
// copy stream.  One sweep launch + one statistics launch per batch of <= 2048 genes, back to back -- or, with option
// "overlap" = B, B batches per call with the statistics of batch i on a second stream under the sweep of batch i+1.
// (r01 measured that overlap slower twice -- 17.8 vs 16.9 and 12.9 vs 11.8 ms/step: the sweep's 5-stage ring takes
// 200 KB, so no statistics CTA ever fitted beside a sweep CTA and the two kernels only took turns.  With "tc_stages" 4
// or 3 the ring leaves 65 / 105 KB per SM for them.)
static int launch_range(rvt_ctx* ctx, int g0, int g1) {
  const int n = g1 - g0;
  if (n <= 0) return RVT_OK;
  int rc;
  const int64_t N = ctx->N;
  const int n_total = (int)ctx->genes.size();
  if ((rc = ensure(ctx, (void**)&ctx->d_genes, &ctx->cap_genes, n_total, sizeof(GeneDesc)))) return rc;
  if ((rc = ensure_var(ctx, ctx->n_var))) return rc;
  if (ctx->bin_pending && (size_t)n_total > ctx->cap_res) RVT_CUDA_OK(cudaStreamSynchronize(ctx->bin_s));   // records in flight: the buffer must not move under them
  if ((rc = ensure(ctx, (void**)&ctx->d_res, &ctx->cap_res, n_total, sizeof(rvt_gene_result)))) return rc;
  int S = 0;
  int64_t chunk = 0;
  int nbat = (n + 2047) / 2048;   // equal batches of <= 2048 genes
  if (ctx->overlap > 0 && !ctx->binary) nbat = std::max(nbat, std::min(ctx->overlap, std::max(1, n / ctx->sm_count)));
  const int batch = (n + nbat - 1) / nbat;
  if ((rc = split_plan(ctx, batch, &S, &chunk))) return rc;
  int engine = ctx->engine;
  if (ctx->stage_used > 0) {
    // bind the whole capacity: the maps stay valid while the arena does not move
    rc = tc_bind_segment(&ctx->tc, kSegStaged, ctx->d_stage, ctx->stage_cap, ctx->err, sizeof(ctx->err));
    if (rc) return rc;
  }
  const GeneDesc* hg = ctx->genes.data() + g0;
  bool tc_ok = tc_usable(&ctx->tc, hg, n);
  if (engine == RVT_ENGINE_AUTO) engine = tc_ok ? RVT_ENGINE_TC : RVT_ENGINE_SIMT;
  if (engine == RVT_ENGINE_TC && !tc_ok)
    CTX_FAIL(RVT_E_UNSUPPORTED, "tensor-core engine requested but unavailable for these genes: %s", ctx->tc.why);
  // the wide tensor-core sweep writes two partials per (gene, split): even and odd 128-sample boxes
  const bool wide = engine == RVT_ENGINE_TC && tc_parts_per_unit(&ctx->tc) == 2;
  const int Sp = wide ? 2 * S : S;
  const bool ovl = ctx->overlap > 0 && !ctx->binary;
  if ((rc = ensure(ctx, (void**)&ctx->d_parts, &ctx->cap_parts, (size_t)batch * Sp * (ovl ? 2 : 1), sizeof(SweepPartial)))) return rc;
  if (ctx->skato) {
    if ((rc = ensure(ctx, (void**)&ctx->d_qags, &ctx->cap_qags, qags_scratch_entries(ctx, n), sizeof(QagsScratch)))) return rc;
    if ((rc = ensure(ctx, (void**)&ctx->d_jobs, &ctx->cap_jobs, (size_t)n, sizeof(SkatoJob)))) return rc;
  }
  if (ctx->want_dbg) {
    if ((rc = ensure(ctx, (void**)&ctx->d_dbg, &ctx->cap_dbg, (size_t)n_total * kFinPhases, sizeof(long long)))) return rc;
  } else if (ctx->d_dbg) {
    cudaFree(ctx->d_dbg);
    ctx->d_dbg = nullptr;
    ctx->cap_dbg = 0;
  }
  cudaStream_t st = ctx->stream;
  const int nbatch = (n + batch - 1) / batch;
  // per batch: [0] sweep start, [1] sweep end, [2] finalize start, [3] finalize end
  while ((int)ctx->evpool.size() < 4 * (ctx->pending_timing_batches + nbatch)) {
    cudaEvent_t e;
    RVT_CUDA_OK(cudaEventCreate(&e));
    ctx->evpool.push_back(e);
  }
  if (!ctx->flush_open) {
    RVT_CUDA_OK(cudaEventRecord(ctx->ev[0], st));
    ctx->flush_open = true;
    ctx->n_launch = 0;
  }
  const int64_t v0 = ctx->genes[g0].var0, v1 = ctx->genes[g1 - 1].var0 + ctx->slots[g1 - 1], nv = v1 - v0;
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->d_genes + g0, hg, sizeof(GeneDesc) * n, cudaMemcpyHostToDevice, st));
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->d_userflags + v0, ctx->userflags.data() + v0, nv, cudaMemcpyHostToDevice, st));
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->d_af + v0, ctx->af.data() + v0, sizeof(double) * nv, cudaMemcpyHostToDevice, st));
  k_resolve_flags<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(nv, N, ctx->d_counts + v0, ctx->d_userflags + v0, ctx->d_flags + v0);
  int launches = 1;
  ctx->last_engine = engine;
  ctx->last_S = S;
  EngineParams prm{ctx->beta1, ctx->beta2, ctx->wd_cycles};
  if (engine == RVT_ENGINE_TC && (rc = tc_prepare_maps(&ctx->tc, hg[0].seg, hg, n, st, ctx->err, sizeof(ctx->err))))
    return rc;   // encode every box height up front: no host sync inside the loop
  for (int bi = 0; bi < nbatch; ++bi) {
    const int b0 = g0 + bi * batch;
    const int nb = std::min(batch, g1 - b0);
    // overlap: two partial buffers in turn; buffer k is free again once the statistics kernels that read it have run
    const int pb = ovl ? (int)(ctx->batch_seq++ & 1) : 0;
    SweepPartial* parts = ctx->d_parts + (size_t)pb * batch * Sp;
    if (ovl && ctx->fin_busy[pb]) RVT_CUDA_OK(cudaStreamWaitEvent(st, ctx->ev_fin_done[pb], 0));
    cudaEvent_t* ev = &ctx->evpool[4 * (ctx->pending_timing_batches + bi)];
    RVT_CUDA_OK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned int), st));
    RVT_CUDA_OK(cudaEventRecord(ev[0], st));
    // binary trait: every gene takes the fp64 path at flush (the Gram is weighted by p(1-p), which the integer sweep does
    // not carry) -- the unweighted sweep and its tail would be wasted work on statistics that mean nothing (and SKAT-O's
    // quadrature on such an indefinite K is where the r01 hang sat)
    if (ctx->binary) {
      RVT_CUDA_OK(cudaEventRecord(ev[1], st));
      RVT_CUDA_OK(cudaEventRecord(ev[2], st));
      if (ctx->bin_stream && (rc = binary_stream_batch(ctx, b0, nb, prm, &launches))) return rc;
      RVT_CUDA_OK(cudaEventRecord(ev[3], st));
      continue;
    }
    if (engine == RVT_ENGINE_SIMT) {
      const int grid = std::min(nb * S, ctx->sm_count * 3);
      k_sweep_simt<<<grid, kSimtThreads, kSimtSmem, st>>>(ctx->d_genes + b0, nb, ctx->d_flags, ctx->d_nm, S, chunk, parts, ctx->d_counter);
    } else {
      rc = tc_launch(&ctx->tc, ctx->d_genes + b0, ctx->genes.data() + b0, nb, ctx->d_flags, ctx->d_nm, N, ctx->ER,
                     S, chunk, parts, ctx->d_counter, ctx->sm_count, st, ctx->err, sizeof(ctx->err), false, wide, -1,
                     ctx->aug ? ctx->d_counts : nullptr);
      if (rc) return rc;
    }
    RVT_CUDA_OK(cudaEventRecord(ev[1], st));
    cudaStream_t fs = st;   // the stream the statistics kernels of this batch run on
    if (ovl) {
      fs = ctx->fin_stream;
      RVT_CUDA_OK(cudaEventRecord(ctx->ev_swept[pb], st));
      RVT_CUDA_OK(cudaStreamWaitEvent(fs, ctx->ev_swept[pb], 0));
    }
    RVT_CUDA_OK(cudaEventRecord(ev[2], fs));
    int Mmax = 1;
    for (int i = 0; i < nb; ++i) Mmax = std::max(Mmax, ctx->genes[b0 + i].M);
    const int kld = fin_kld(Mmax), fsm = fin_smem(Mmax, ctx->ER, ctx->skato), wm_off = Mmax * kld * 8;
    if (ctx->skato) {
      k_finalize<true><<<nb, kFinThreadsSkato, fsm, fs>>>(
          ctx->d_genes + b0, nb, kld, wm_off, fin_uk_off(Mmax, ctx->ER, ctx->skato), ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm, prm, Sp, parts, ctx->d_res + b0,
          ctx->d_dbg ? ctx->d_dbg + (size_t)b0 * kFinPhases : nullptr, ctx->d_jobs + (b0 - g0), nullptr, nullptr);
      if (bi == nbatch - 1) {   // the quadrature of the whole range in one launch (see launch_qags)
        if ((rc = launch_qags(ctx, ctx->d_jobs, n, ctx->d_res + g0, nullptr, fs))) return rc;
        launches += 1;
      }
    } else if (ctx->fin_split && !ctx->d_dbg) {
      if ((rc = ensure(ctx, (void**)&ctx->d_mid, &ctx->cap_mid, (size_t)nb, sizeof(FinMid)))) return rc;
      k_finalize<false, true><<<nb, kFinThreads, fsm, fs>>>(
          ctx->d_genes + b0, nb, kld, wm_off, fin_uk_off(Mmax, ctx->ER, false), ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm, prm, Sp, parts, ctx->d_res + b0,
          nullptr, nullptr, nullptr, nullptr, ctx->d_mid);
      k_fin_sturm<<<nb, kFinThreads, 0, fs>>>(ctx->d_mid, nb);
      k_fin_tail<<<nb, kFinThreads, 0, fs>>>(ctx->d_mid, nb, ctx->d_nm, ctx->d_res + b0, nullptr);
      launches += 2;
    } else
      k_finalize<false><<<nb, kFinThreads, fsm, fs>>>(
          ctx->d_genes + b0, nb, kld, wm_off, fin_uk_off(Mmax, ctx->ER, ctx->skato), ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm, prm, Sp, parts, ctx->d_res + b0,
          ctx->d_dbg ? ctx->d_dbg + (size_t)b0 * kFinPhases : nullptr, nullptr, nullptr, nullptr);
    RVT_CUDA_OK(cudaEventRecord(ev[3], fs));
    if (ovl) {
      RVT_CUDA_OK(cudaEventRecord(ctx->ev_fin_done[pb], fs));
      ctx->fin_busy[pb] = true;
    }
    RVT_CUDA_OK(cudaGetLastError());
    launches += 2;
    ctx->last_parts = (int64_t)nb * Sp;
  }
  if (ovl) {
    // everything enqueued on the context stream after this call (the fp64 path, wide genes, the copy of the records, the
    // next pushes' kernels) is ordered after the statistics of these batches
    for (int k = 0; k < 2; ++k)
      if (ctx->fin_busy[k]) RVT_CUDA_OK(cudaStreamWaitEvent(st, ctx->ev_fin_done[k], 0));
  }
  ctx->pending_timing_batches += nbatch;
  ctx->n_launch += launches;
  ctx->launched = g1;
  return RVT_OK;
}

// Genes wider than one tile (wide.cuh): per gene, the T diagonal tile sweeps and the T(T-1)/2 tile-pair sweeps fill the
// M x M integer Gram, one more pass computes the burden collapses over all M variants, then one CTA per gene runs the
// O(M^3) tail on a global-memory workspace.  Their records overwrite what the placeholder descriptors produced.
// Wide genes of a binary-trait run: the p(1-p)-weighted statistics by k_wide_sparse (fp64, from the int8 tiles, missing calls
// imputed on the fly), then the same tail (k_wide_finalize, fp64 mode).
static int run_wide_binary(rvt_ctx* ctx, rvt_gene_result* d_res, int* launches) {
  cudaStream_t st = ctx->stream;
  const int64_t N = ctx->N;
  const int nw = (int)ctx->wide.size();
  int rc;
  const bool sk = ctx->skato && ctx->skato_binary;
  if (sk && (rc = ensure(ctx, (void**)&ctx->d_qags, &ctx->cap_qags, (size_t)nw, sizeof(QagsScratch)))) return rc;
  EngineParams prm{ctx->beta1, ctx->beta2, ctx->wd_cycles};
  std::vector<WideJob> jobs(nw);
  std::vector<void*> to_free;
  auto cleanup = [&]() {
    for (void* p : to_free) cudaFree(p);
  };
  WideJob* d_jobs = nullptr;
  RVT_CUDA_OK(cudaMalloc((void**)&d_jobs, sizeof(WideJob) * nw));
  to_free.push_back(d_jobs);
  for (int wi = 0; wi < nw; ++wi) {
    const WideGene& w = ctx->wide[wi];
    const int T = (int)w.tiles.size();
    uint8_t* ws = nullptr;
    GeneDesc* d_tiles = nullptr;
    cudaError_t e = cudaMalloc((void**)&ws, wide_ws_bytes(w.M));
    if (e == cudaSuccess) to_free.push_back(ws);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_tiles, sizeof(GeneDesc) * T);
    if (e != cudaSuccess) {
      cleanup();
      CTX_FAIL(RVT_E_CUDA, "wide gene (M=%d, binary trait): cudaMalloc: %s", w.M, cudaGetErrorString(e));
    }
    to_free.push_back(d_tiles);
    WideJob jb = wide_job_make(ws, w.M);
    jb.imp = 2;
    jb.out_index = w.gene_index;
    jb.var0 = w.var0;
    jb.has_af = w.has_af ? 1 : 0;
    jb.counted = 1;
    jobs[wi] = jb;
    // accumulators: A (as doubles over A_raw), {S, CW, B} (over De), the burden sums (over coll)
    RVT_CUDA_OK(cudaMemsetAsync(jb.A_raw, 0, sizeof(double) * (size_t)w.M * w.M, st));
    RVT_CUDA_OK(cudaMemsetAsync(jb.De, 0, sizeof(double) * (size_t)w.M * kMaxER, st));
    RVT_CUDA_OK(cudaMemsetAsync(jb.coll, 0, sizeof(long long) * kCollapseN, st));
    RVT_CUDA_OK(cudaMemcpyAsync(d_tiles, w.tiles.data(), sizeof(GeneDesc) * T, cudaMemcpyHostToDevice, st));   // (pageable: staged)
    k_wide_imp_flags<<<(unsigned)((w.M + 255) / 256), 256, 0, st>>>(w.var0, w.M, N, ctx->d_counts, ctx->d_flags);
    const int64_t nchunks = (N + 127) / 128;
    const size_t smem = (size_t)w.M * (sizeof(double) + 3);
    k_wide_sparse<<<(unsigned)std::min<int64_t>(nchunks, (int64_t)ctx->sm_count * 8), kWideCollapseThreads, smem, st>>>(
        d_tiles, T, w.var0, w.M, ctx->d_flags, ctx->d_counts, ctx->d_nm, ctx->dX, ctx->d_vw, reinterpret_cast<double*>(jb.A_raw),
        reinterpret_cast<double*>(jb.De), reinterpret_cast<double*>(jb.coll));
    *launches += 2;
    RVT_CUDA_OK(cudaGetLastError());
  }
  RVT_CUDA_OK(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(WideJob) * nw, cudaMemcpyHostToDevice, st));
  if (sk)
    k_wide_finalize<true><<<nw, kWideThreads, 0, st>>>(d_jobs, nw, ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm, prm, d_res, ctx->d_qags);
  else
    k_wide_finalize<false><<<nw, kWideThreads, 0, st>>>(d_jobs, nw, ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm, prm, d_res, nullptr);
  *launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) CTX_FAIL(RVT_E_CUDA, "wide genes (binary trait): %s", cudaGetErrorString(e));
  return RVT_OK;
}

// Wide genes pushed as doubles that hold REAL dosages: dense fp64 statistics (k_wide_dos_cols / _gram / _burden), then the fp64
// mode of the tail.  Quantitative and binary traits alike (vw = null for the former).
static int run_wide_dosage(rvt_ctx* ctx, rvt_gene_result* d_res, int* launches) {
  cudaStream_t st = ctx->stream;
  const int64_t N = ctx->N;
  int nw = 0;
  for (auto& w : ctx->wide) nw += w.dG != nullptr;
  if (nw == 0) return RVT_OK;
  int rc;
  const bool sk = ctx->skato && (!ctx->binary || ctx->skato_binary);
  if (sk && (rc = ensure(ctx, (void**)&ctx->d_qags, &ctx->cap_qags, (size_t)nw, sizeof(QagsScratch)))) return rc;
  EngineParams prm{ctx->beta1, ctx->beta2, ctx->wd_cycles};
  std::vector<WideJob> jobs;
  std::vector<void*> to_free;
  auto cleanup = [&]() {
    for (void* p : to_free) cudaFree(p);
  };
  const double* vw = ctx->binary ? ctx->d_vw : nullptr;
  for (auto& w : ctx->wide) {
    if (!w.dG) continue;
    uint8_t* ws = nullptr;
    cudaError_t e = cudaMalloc((void**)&ws, wide_ws_bytes(w.M));
    if (e != cudaSuccess) {
      cleanup();
      CTX_FAIL(RVT_E_CUDA, "wide gene (M=%d, dosages): cudaMalloc: %s", w.M, cudaGetErrorString(e));
    }
    to_free.push_back(ws);
    WideJob jb = wide_job_make(ws, w.M);
    jb.imp = 3;
    jb.out_index = w.gene_index;
    jb.var0 = w.var0;
    jb.has_af = w.has_af ? 1 : 0;
    jb.counted = 1;
    jobs.push_back(jb);
    double* A = reinterpret_cast<double*>(jb.A_raw);
    double* SB = reinterpret_cast<double*>(jb.De);
    double* bur = reinterpret_cast<double*>(jb.coll);
    double* csum = reinterpret_cast<double*>(jb.craw);
    RVT_CUDA_OK(cudaMemsetAsync(A, 0, sizeof(double) * (size_t)w.M * w.M, st));
    RVT_CUDA_OK(cudaMemsetAsync(jb.coll, 0, sizeof(long long) * kCollapseN, st));
    k_wide_dos_cols<<<w.M, kWideDosThreads, 0, st>>>(w.dG, N, w.M, w.var0, ctx->d_nm, ctx->dX, vw, SB, csum, ctx->d_flags);
    const int nblk = (w.M + 63) / 64;
    const int npair = nblk * (nblk + 1) / 2;
    const int splits = (int)std::max<int64_t>(1, std::min<int64_t>((N + 4095) / 4096, (4 * (int64_t)ctx->sm_count + npair - 1) / npair));
    const int64_t split_len = (((N + splits - 1) / splits) + 31) / 32 * 32;
    k_wide_dos_gram<<<dim3((unsigned)npair, (unsigned)((N + split_len - 1) / split_len)), 256, 0, st>>>(w.dG, N, w.M, vw, nblk, split_len, A);
    k_wide_dos_burden<<<(unsigned)std::min<int64_t>((N + kWideDosThreads - 1) / kWideDosThreads, (int64_t)ctx->sm_count * 8), kWideDosThreads, (size_t)w.M, st>>>(
        w.dG, N, w.M, w.var0, ctx->d_flags, ctx->d_nm, ctx->dX, vw, bur);
    *launches += 3;
    RVT_CUDA_OK(cudaGetLastError());
  }
  WideJob* d_jobs = nullptr;
  RVT_CUDA_OK(cudaMalloc((void**)&d_jobs, sizeof(WideJob) * nw));
  to_free.push_back(d_jobs);
  RVT_CUDA_OK(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(WideJob) * nw, cudaMemcpyHostToDevice, st));
  if (sk)
    k_wide_finalize<true><<<nw, kWideThreads, 0, st>>>(d_jobs, nw, ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm, prm, d_res, ctx->d_qags);
  else
    k_wide_finalize<false><<<nw, kWideThreads, 0, st>>>(d_jobs, nw, ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm, prm, d_res, nullptr);
  *launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  // the dosage matrices are done with; what is left in ctx->wide are the genes of the integer / sparse paths
  for (size_t k = 0; k < ctx->wide.size();) {
    if (ctx->wide[k].dG) {
      cudaFree(ctx->wide[k].dG);
      ctx->wide.erase(ctx->wide.begin() + (long)k);
    } else {
      ++k;
    }
  }
  if (e != cudaSuccess) CTX_FAIL(RVT_E_CUDA, "wide genes (dosages): %s", cudaGetErrorString(e));
  return RVT_OK;
}

static int run_wide(rvt_ctx* ctx, rvt_gene_result* d_res, int* launches) {
  if (ctx->wide.empty()) return RVT_OK;
  int rc0 = run_wide_dosage(ctx, d_res, launches);
  if (rc0) return rc0;
  if (ctx->wide.empty()) return RVT_OK;
  if (ctx->binary) return run_wide_binary(ctx, d_res, launches);
  int rc;
  if ((rc = tc_bind_segment(&ctx->tc, kSegStaged, ctx->d_stage, ctx->stage_cap, ctx->err, sizeof(ctx->err)))) return rc;
  if (!(ctx->tc.encode && ctx->tc.have_e && ctx->tc.have_seg[kSegStaged]))
    CTX_FAIL(RVT_E_UNSUPPORTED, "genes of more than %d variants need the tensor-core sweep (TMA unavailable: %s)", kMaxM, ctx->tc.why);
  cudaStream_t st = ctx->stream;
  const int64_t N = ctx->N;
  const int nw = (int)ctx->wide.size();
  if ((rc = ensure(ctx, (void**)&ctx->d_zero_flags, &ctx->cap_zero_flags, (size_t)ctx->n_var + kTileRows, 1))) return rc;
  RVT_CUDA_OK(cudaMemsetAsync(ctx->d_zero_flags, 0, (size_t)ctx->n_var + kTileRows, st));
  if (ctx->skato && (rc = ensure(ctx, (void**)&ctx->d_qags, &ctx->cap_qags, (size_t)nw, sizeof(QagsScratch)))) return rc;
  EngineParams prm{ctx->beta1, ctx->beta2, ctx->wd_cycles};
  std::vector<WideJob> jobs(nw);
  std::vector<void*> to_free;
  auto cleanup = [&]() {
    for (void* p : to_free) cudaFree(p);
  };
  WideJob* d_jobs = nullptr;
  RVT_CUDA_OK(cudaMalloc((void**)&d_jobs, sizeof(WideJob) * nw));
  to_free.push_back(d_jobs);
  const int batch = 1024;
  for (int wi = 0; wi < nw; ++wi) {
    const WideGene& w = ctx->wide[wi];
    const int T0 = (int)w.tiles.size();
    // Missing calls (mean-imputed, G = H + Mi diag(delta)): every tile is split into its operand tiles H (hard calls, the
    // fill 0 / 2 at missing entries) and Mi (0/1 indicators) in an auxiliary segment, and the gene is swept as the 2M rows
    // [H_1 .. H_T ; Mi_1 .. Mi_T]: the same diagonal + pair units, exact integers; k_wide_finalize combines them.
    std::vector<GeneDesc> tiles(w.tiles);
    int8_t* d_hm = nullptr;
    if (w.imputed) {
      std::vector<int64_t> off(2 * T0 + 1, 0);
      for (int t = 0; t < 2 * T0; ++t) off[t + 1] = off[t] + tiled_bytes(N, w.tiles[t % T0].M);
      cudaError_t ea = cudaMalloc((void**)&d_hm, (size_t)off[2 * T0]);
      if (ea != cudaSuccess) {
        cleanup();
        CTX_FAIL(RVT_E_CUDA, "wide gene with missing calls (M=%d): cudaMalloc of the operand tiles (%lld bytes): %s", w.M, (long long)off[2 * T0],
                 cudaGetErrorString(ea));
      }
      to_free.push_back(d_hm);
      k_wide_imp_flags<<<(unsigned)((w.M + 255) / 256), 256, 0, st>>>(w.var0, w.M, N, ctx->d_counts, ctx->d_flags);
      const int64_t nwords = ((N + 127) >> 7) * 32;
      tiles.resize(2 * T0);
      for (int t = 0; t < T0; ++t) {
        const GeneDesc& src = w.tiles[t];
        k_split_hm<<<dim3((unsigned)((nwords + 255) / 256), (unsigned)src.M), 256, 0, st>>>(src.g, src.M, N, ctx->d_flags + src.var0, d_hm + off[t],
                                                                                         d_hm + off[T0 + t]);
        GeneDesc h = src, m = src;
        h.g = d_hm + off[t];
        h.seg = kSegAux;
        h.row0 = h.row0_b = off[t] / 128;
        m.g = d_hm + off[T0 + t];
        m.seg = kSegAux;
        m.row0 = m.row0_b = off[T0 + t] / 128;
        m.var0 = m.var0_b = src.var0 + w.M;        // rows M .. 2M-1 of the doubled gene
        tiles[t] = h;
        tiles[T0 + t] = m;
      }
      if ((rc = tc_bind_segment(&ctx->tc, kSegAux, d_hm, (size_t)off[2 * T0], ctx->err, sizeof(ctx->err)))) { cleanup(); return rc; }
      if ((rc = ensure(ctx, (void**)&ctx->d_zero_flags, &ctx->cap_zero_flags, (size_t)ctx->n_var + 2 * (size_t)w.M + kTileRows, 1))) { cleanup(); return rc; }
      RVT_CUDA_OK(cudaMemsetAsync(ctx->d_zero_flags, 0, (size_t)ctx->n_var + 2 * (size_t)w.M + kTileRows, st));
    }
    const int T = (int)tiles.size();
    const int Mw = w.imputed ? 2 * w.M : w.M;      // rows of the swept Gram
    std::vector<GeneDesc> units(tiles);
    for (int t = 0; t < T; ++t)
      for (int u = t + 1; u < T; ++u) {
        GeneDesc g = tiles[t];
        g.row0_b = tiles[u].row0;
        g.Mb = tiles[u].M;
        g.var0_b = tiles[u].var0;
        units.push_back(g);
      }
    const int n_units = (int)units.size();
    uint8_t* ws = nullptr;
    GeneDesc* d_units = nullptr;
    cudaError_t e = cudaMalloc((void**)&ws, wide_ws_bytes(Mw));
    if (e == cudaSuccess) to_free.push_back(ws);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_units, sizeof(GeneDesc) * n_units);
    if (e != cudaSuccess) {
      cleanup();
      CTX_FAIL(RVT_E_CUDA, "wide gene (M=%d): cudaMalloc: %s", w.M, cudaGetErrorString(e));
    }
    to_free.push_back(d_units);
    WideJob jb = wide_job_make(ws, Mw);
    jb.M = w.M;
    jb.imp = w.imputed ? 1 : 0;
    jb.out_index = w.gene_index;
    jb.var0 = w.var0;
    jb.has_af = w.has_af ? 1 : 0;
    jb.counted = 1;
    jobs[wi] = jb;
    RVT_CUDA_OK(cudaMemsetAsync(jb.coll, 0, sizeof(long long) * kCollapseN, st));
    RVT_CUDA_OK(cudaMemcpyAsync(d_units, units.data(), sizeof(GeneDesc) * n_units, cudaMemcpyHostToDevice, st));
    int S = 0;
    int64_t chunk = 0;
    if ((rc = split_plan(ctx, std::min(n_units, batch), &S, &chunk))) { cleanup(); return rc; }
    if ((rc = ensure(ctx, (void**)&ctx->d_parts, &ctx->cap_parts, (size_t)std::min(n_units, batch) * S, sizeof(SweepPartial)))) { cleanup(); return rc; }
    for (int pass = 0; pass < 2; ++pass) {   // 0: diagonal tiles, 1: tile pairs
      const int u0 = pass ? T : 0, u1 = pass ? n_units : T;
      for (int b0 = u0; b0 < u1; b0 += batch) {
        const int nb = std::min(batch, u1 - b0);
        rc = tc_launch(&ctx->tc, d_units + b0, units.data() + b0, nb, ctx->d_zero_flags, ctx->d_nm, N, ctx->ER, S, chunk, ctx->d_parts,
                       ctx->d_counter, ctx->sm_count, st, ctx->err, sizeof(ctx->err), pass == 1, false);
        if (rc) { cleanup(); return rc; }
        k_wide_gather<<<nb, kWideGatherThreads, 0, st>>>(d_units + b0, nb, w.var0, Mw, ctx->ER, S, ctx->d_parts, jb.A_raw, jb.De, pass);
        *launches += 2;
      }
    }
    const int64_t nchunks = (N + 127) / 128;
    k_wide_collapse<<<(unsigned)std::min<int64_t>(nchunks, (int64_t)ctx->sm_count * 8), kWideCollapseThreads, 0, st>>>(
        d_units, T0, w.var0, w.M, ctx->d_flags, ctx->d_nm, jb.coll);   // (imputed: the H tiles -- an imputed value never counts, the fill sees to it)
    *launches += 1;
    RVT_CUDA_OK(cudaGetLastError());
    // `units` is a host temporary of this iteration: the copy above must have left it
    RVT_CUDA_OK(cudaStreamSynchronize(st));
  }
  RVT_CUDA_OK(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(WideJob) * nw, cudaMemcpyHostToDevice, st));
  if (ctx->skato)
    k_wide_finalize<true><<<nw, kWideThreads, 0, st>>>(d_jobs, nw, ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm, prm, d_res, ctx->d_qags);
  else
    k_wide_finalize<false><<<nw, kWideThreads, 0, st>>>(d_jobs, nw, ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm, prm, d_res, nullptr);
  *launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) CTX_FAIL(RVT_E_CUDA, "wide genes: %s", cudaGetErrorString(e));
  return RVT_OK;
}

// jump tables of the rand() recurrence (constant), uploaded once
static int perm_tables(rvt_ctx* ctx) {
  if (ctx->d_lfg) return RVT_OK;
  std::vector<LfgTables> tab(1);
  for (int k = 0; k < 32; ++k) tab[0].zblock[k] = lfg_pow((uint64_t)kLfgBlock << k);
  for (int t = 0; t < kLfgThreads; ++t) tab[0].zthread[t] = lfg_pow((uint64_t)t * kLfgRun);
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->d_lfg, sizeof(LfgTables)));
  RVT_CUDA_OK(cudaMemcpy(ctx->d_lfg, tab.data(), sizeof(LfgTables), cudaMemcpyHostToDevice));
  return RVT_OK;
}

// A6 (perm.cuh): permutation p-values of the SKAT statistic, gene after gene in push order, consuming the glibc
// rand() stream exactly as the reference's serial loop does (src/Model.h:2707-2717).  Runs after the analytic
// results exist (the observed Q is the comparison value).
static int run_perm(rvt_ctx* ctx, const rvt_gene_result* d_res, int n, int* launches) {
  ctx->perm_out.clear();
  ctx->perm_q_log.clear();
  if (ctx->perm_n <= 0) return RVT_OK;
  if (!(ctx->tc.encode && ctx->tc.have_e))
    CTX_FAIL(RVT_E_UNSUPPORTED, "the permutation test needs the tensor-core sweep (TMA unavailable: %s)", ctx->tc.why);
  int rc;
  cudaStream_t st = ctx->stream;
  const int64_t N = ctx->N;
  if (N < 2 || N > 0xFFFFFFF0ll) CTX_FAIL(RVT_E_UNSUPPORTED, "permutation test: N out of range");
  std::vector<rvt_gene_result> hres(n);
  RVT_CUDA_OK(cudaMemcpyAsync(hres.data(), d_res, sizeof(rvt_gene_result) * n, cudaMemcpyDeviceToHost, st));
  RVT_CUDA_OK(cudaStreamSynchronize(st));
  ctx->perm_out.assign(n, rvt_perm_result{});
  if ((rc = perm_tables(ctx))) return rc;
  uint32_t y0[2 * kLfgDeg - 1];
  lfg_seed_window(ctx->perm_seed, y0);
  const int PB = ctx->perm_batch;                 // permutations per batch (multiple of 16)
  const int64_t tile_b = tiled_bytes(N, kTileRows);
  int Mmax = 1, Tmax = 2;
  for (int g = 0; g < n; ++g) Mmax = std::max(Mmax, ctx->slots[g]);
  Mmax = std::max(Mmax, 2 * kTileRows);   // a gene with missing calls: columns for H'r and M'r
  for (auto& w : ctx->wide) Tmax = std::max<int>(Tmax, (int)w.tiles.size());
  bool any_aug = false;
  for (int g = 0; g < n; ++g) any_aug |= ctx->is_dos[g] == 3;
  // scratch layout
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 1023) & ~(size_t)1023; return o; };
  const size_t o_tiles = carve((size_t)(PB / 16) * tile_b);
  const size_t o_draws = carve((size_t)PB * (N - 1) * 4);
  const size_t o_head = carve((size_t)PB * N * 4);
  const size_t o_link = carve((size_t)PB * N * 4);
  const size_t o_root = carve((size_t)PB * N * 4);
  const size_t o_R0 = carve((size_t)N * 4), o_R1 = carve((size_t)N * 4);
  const size_t o_w0 = carve(64 * 4);
  const size_t o_sint = carve((size_t)PB * Mmax * 8);
  const size_t o_w = carve((size_t)2 * Mmax * 8);
  const size_t o_Q = carve((size_t)PB * 8);
  const size_t o_units = carve(sizeof(GeneDesc) * (size_t)Tmax * (PB / 16));
  const size_t o_aux = carve(any_aug ? 2 * (size_t)tile_b : 0);   // H and M tiles of one gene with missing calls
  if (off > ctx->cap_perm) {
    if (ctx->d_perm) cudaFree(ctx->d_perm);
    ctx->d_perm = nullptr;
    ctx->cap_perm = 0;
    RVT_CUDA_OK(cudaMalloc((void**)&ctx->d_perm, off));
    ctx->cap_perm = off;
  }
  uint8_t* base = ctx->d_perm;
  int8_t* d_tiles = (int8_t*)(base + o_tiles);
  uint32_t *d_draws = (uint32_t*)(base + o_draws), *d_head = (uint32_t*)(base + o_head), *d_link = (uint32_t*)(base + o_link),
           *d_root = (uint32_t*)(base + o_root), *d_R[2] = {(uint32_t*)(base + o_R0), (uint32_t*)(base + o_R1)},
           *d_w0 = (uint32_t*)(base + o_w0);
  long long* d_sint = (long long*)(base + o_sint);
  double *d_w = (double*)(base + o_w), *d_Q = (double*)(base + o_Q);
  GeneDesc* d_units = (GeneDesc*)(base + o_units);
  if ((rc = tc_bind_segment(&ctx->tc, kSegPerm, d_tiles, (int64_t)(PB / 16) * tile_b, ctx->err, sizeof(ctx->err)))) return rc;
  int8_t* d_aux = (int8_t*)(base + o_aux);
  if (any_aug && (rc = tc_bind_segment(&ctx->tc, kSegAux, d_aux, 2 * tile_b, ctx->err, sizeof(ctx->err)))) return rc;
  if ((rc = ensure(ctx, (void**)&ctx->d_zero_flags, &ctx->cap_zero_flags, (size_t)ctx->n_var + kTileRows, 1))) return rc;
  RVT_CUDA_OK(cudaMemsetAsync(ctx->d_zero_flags, 0, (size_t)ctx->n_var + kTileRows, st));
  EngineParams prm{ctx->beta1, ctx->beta2, ctx->wd_cycles};
  const int threshold = (int)(1.0 * ctx->perm_n * ctx->perm_alpha * 2);   // Permutation::init, `int threshold`
  std::vector<double> hQ(PB);
  std::vector<GeneDesc> units;
  const bool trace = getenv("RVT_PERM_TRACE") != nullptr;   // diagnostics: synchronise and report after every stage
  auto mark = [&](const char* what) {
    if (!trace) return;
    cudaError_t e = cudaStreamSynchronize(st);
    fprintf(stderr, "[perm] %s: %s\n", what, cudaGetErrorString(e));
    fflush(stderr);
  };
  for (int g = 0; g < n; ++g) {
    rvt_perm_result& rec = ctx->perm_out[g];
    rec.num_perm = ctx->perm_n;
    rec.stat = hres[g].Q;
    rec.stream_pos = (int64_t)ctx->perm_pos;
    rec.p_perm = 1.0;
    const GeneDesc& gd = ctx->genes[g];
    // fit() returned -1 (no polymorphic variant): the reference runs no permutation.  Genes with dosages / missing calls
    // (fp64 path) and caller-owned device blocks without the engine's own counts are not covered: their record says
    // done = 0 and -- the reference WOULD have shuffled for them -- the rand() stream of the genes after them no longer
    // lines up with the reference's (documented in include/rvtests_b200.h).  A binary trait is covered: its permuted
    // statistic is sum_j w_j (g_j' r_pi)^2 with r = y - p, hard calls and digits of r as for a quantitative trait
    // (src/Model.h:2673-2717: one loop for both outcomes).
    rec.stream_ok = ctx->perm_stream_lost ? 0 : 1;
    if (hres[g].status != RVT_GENE_OK || ctx->is_dos[g] == 1 || !gd.tiled || gd.seg < 0 || (!gd.has_af && !gd.counted)) {
      // status OK: the reference's SkatTest::fit would have run its permutation loop here and consumed ActualPerm * (N - 1)
      // draws we cannot know -- every later record of this context says so (stream_ok = 0) until the position is set again
      if (hres[g].status == RVT_GENE_OK) ctx->perm_stream_lost = true;
      continue;
    }
    const bool aug = ctx->is_dos[g] == 3;   // missing calls: two operand tiles, H (fills applied) and M (indicators)
    const std::vector<GeneDesc>* tiles = nullptr;
    std::vector<GeneDesc> one(1, gd);
    for (auto& w : ctx->wide)
      if (w.gene_index == g) tiles = &w.tiles;
    if (!tiles) tiles = &one;
    if (aug) {
      const int Mg = gd.M;
      const int64_t nwords = ((N + 127) >> 7) * 32;
      const int64_t hb = tiled_bytes(N, Mg);
      k_split_hm<<<dim3((unsigned)((nwords + 255) / 256), (unsigned)Mg), 256, 0, st>>>(gd.g, Mg, N, ctx->d_flags + gd.var0, d_aux, d_aux + tile_b);
      one.assign(2, gd);
      one[0].g = d_aux;
      one[0].seg = kSegAux;
      one[0].row0 = one[0].row0_b = 0;
      one[1].g = d_aux + tile_b;
      one[1].seg = kSegAux;
      one[1].row0 = one[1].row0_b = tile_b / 128;
      one[1].var0 = one[1].var0_b = gd.var0 + Mg;   // columns M..2M-1 of sint
      (void)hb;
      tiles = &one;
    }
    const int T = (int)tiles->size(), M = aug ? 2 * gd.M : ctx->slots[g];
    const double obs = hres[g].Q;
    int actual = 0, numX = 0, numEq = 0, cur = 0;
    k_perm_init<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(ctx->d_nm, d_R[0]);
    while (actual < ctx->perm_n && numX + numEq < threshold) {
      const int Pb = std::min(PB, ctx->perm_n - actual), Pb16 = (Pb + 15) & ~15, G16 = Pb16 / 16;
      const uint64_t pos0 = ctx->perm_pos + (uint64_t)actual * (uint64_t)(N - 1);
      uint32_t w0[2 * kLfgDeg - 1];
      lfg_window_at(lfg_pow(pos0 + kLfgWarm), y0, w0);
      RVT_CUDA_OK(cudaMemcpyAsync(d_w0, w0, sizeof(w0), cudaMemcpyHostToDevice, st));   // (pageable: staged before return)
      const uint64_t cnt = (uint64_t)Pb16 * (uint64_t)(N - 1);
      k_lfg_draws<<<(unsigned)((cnt + kLfgBlock - 1) / kLfgBlock), kLfgThreads, 0, st>>>(ctx->d_lfg, d_w0, cnt, d_draws);
      mark("draws");
      RVT_CUDA_OK(cudaMemsetAsync(d_head, 0xFF, (size_t)Pb16 * N * 4, st));
      k_fy_link<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(d_draws, (uint32_t)N, Pb16, d_head, d_link);
      mark("link");
      k_fy_root<<<(unsigned)(((uint64_t)Pb16 * N + 255) / 256), 256, 0, st>>>(d_draws, (uint32_t)N, Pb16, d_head, d_link, d_root);
      mark("root");
      for (int p = 0; p < Pb16; ++p) {
        k_perm_gather<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(d_R[cur], d_root + (size_t)p * N, (uint32_t)N, d_R[cur ^ 1],
                                                                    d_tiles + (size_t)(p / 16) * tile_b, 4 * (p % 16));
        cur ^= 1;
      }
      *launches += 4 + Pb16;
      mark("gather");
      units.clear();
      for (int t = 0; t < T; ++t)
        for (int u = 0; u < G16; ++u) {
          GeneDesc ud = (*tiles)[t];
          ud.row0_b = (int64_t)u * tile_b / 128;
          ud.Mb = kTileRows;
          ud.var0_b = u;
          units.push_back(ud);
        }
      const int n_units = (int)units.size();
      RVT_CUDA_OK(cudaMemcpyAsync(d_units, units.data(), sizeof(GeneDesc) * n_units, cudaMemcpyHostToDevice, st));
      const int batch = 1024;
      int S = 0;
      int64_t chunk = 0;
      if ((rc = split_plan(ctx, std::min(n_units, batch), &S, &chunk))) return rc;
      if ((rc = ensure(ctx, (void**)&ctx->d_parts, &ctx->cap_parts, (size_t)std::min(n_units, batch) * S, sizeof(SweepPartial)))) return rc;
      for (int b0 = 0; b0 < n_units; b0 += batch) {
        const int nb = std::min(batch, n_units - b0);
        rc = tc_launch(&ctx->tc, d_units + b0, units.data() + b0, nb, ctx->d_zero_flags, ctx->d_nm, N, ctx->ER, S, chunk, ctx->d_parts,
                       ctx->d_counter, ctx->sm_count, st, ctx->err, sizeof(ctx->err), true, false, kSegPerm);
        if (rc) return rc;
        mark("sweep");
        k_perm_sint<<<nb, 128, 0, st>>>(d_units + b0, nb, gd.var0, M, S, ctx->d_parts, d_sint);
        mark("sint");
        *launches += 2;
      }
      if (aug)
        k_perm_q_aug<<<1, 256, 0, st>>>(Pb16, gd.M, gd.var0, gd.has_af, d_sint, ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm, prm, d_w, d_Q);
      else
        k_perm_q<<<1, 256, 0, st>>>(Pb16, M, gd.var0, gd.has_af, d_sint, ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm, prm, d_w, d_Q);
      *launches += 1;
      RVT_CUDA_OK(cudaGetLastError());
      RVT_CUDA_OK(cudaMemcpyAsync(hQ.data(), d_Q, sizeof(double) * Pb16, cudaMemcpyDeviceToHost, st));
      RVT_CUDA_OK(cudaStreamSynchronize(st));
      for (int p = 0; p < Pb && actual < ctx->perm_n && numX + numEq < threshold; ++p) {   // Permutation::next / add
        ++actual;
        if (ctx->perm_log) ctx->perm_q_log.push_back(hQ[p]);
        if (hQ[p] > obs) ++numX;
        if (hQ[p] == obs) ++numEq;
      }
    }
    ctx->perm_pos += (uint64_t)actual * (uint64_t)(N - 1);
    rec.actual_perm = actual;
    rec.num_greater = numX;
    rec.num_equal = numEq;
    rec.p_perm = actual ? 1.0 * (numX + 0.5 * numEq) / actual : 1.0;   // Permutation::getPvalue
    rec.done = 1;
  }
  return RVT_OK;
}

static int flush_body(rvt_ctx* ctx, rvt_gene_result* out, int cap, int* n_out, bool to_device);
// A flush that fails must not leave the queue half-consumed (every later flush failed again and the adapters
// printed NA for the rest of the run): whatever the reason, the pending genes are dropped and the fp64-path buffers freed;
// rvt_last_error() keeps the message of the failure.  (A too-small result buffer is the caller's to retry: nothing is dropped.)
static int flush_impl(rvt_ctx* ctx, rvt_gene_result* out, int cap, int* n_out, bool to_device) {
  if (!ctx) return RVT_E_BADARG;
  if (!ctx->genes.empty() && (!out || cap < (int)ctx->genes.size()))
    CTX_FAIL(RVT_E_BADARG, "result buffer too small: %d pending genes, cap %d", (int)ctx->genes.size(), cap);
  const int rc = flush_body(ctx, out, cap, n_out, to_device);
  if (rc != RVT_OK) {
    cudaStreamSynchronize(ctx->stream);
    for (auto& dg : ctx->dos)
      if (dg.dG) cudaFree(dg.dG);
    ctx->dos.clear();
    pending_reset(ctx);
  }
  return rc;
}

static int flush_body(rvt_ctx* ctx, rvt_gene_result* out, int cap, int* n_out, bool to_device) {
  if (!ctx) return RVT_E_BADARG;
  const bool trace = getenv("RVT_FLUSH_TRACE") != nullptr;   // diagnostics: host wall clock of the phases of a flush
  const auto t_begin = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(ctx->stream);
    fprintf(stderr, "[flush] %-24s %8.3f ms\n", what,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
  };
  const int n = (int)ctx->genes.size();
  if (n_out) *n_out = 0;
  if (n == 0) return RVT_OK;
  if (!out || cap < n) CTX_FAIL(RVT_E_BADARG, "result buffer too small: %d pending genes, cap %d", n, cap);
  RVT_CUDA_OK(cudaSetDevice(ctx->device));
  int rc;
  mark("enter (copies landed)");
  if ((rc = launch_range(ctx, ctx->launched, n))) return rc;
  if (ctx->bin_pending) {
    RVT_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ctx->ev_bin_out, 0));
    ctx->bin_pending = false;
  }
  mark("sweep + statistics");
  if ((rc = resolve_bed_missing(ctx))) return rc;
  mark("missing-call check");
  const int64_t N = ctx->N;
  cudaStream_t st = ctx->stream;
  rvt_gene_result* d_res = ctx->d_res;
  EngineParams prm{ctx->beta1, ctx->beta2, ctx->wd_cycles};
  int launches = 0;
  ctx->is_dos.assign(n, 0);
  for (auto& dg : ctx->dos) ctx->is_dos[dg.gene_index] = 1;
  for (auto& w : ctx->wide)
    if (w.imputed || w.dG) ctx->is_dos[w.gene_index] = 1;   // (computed by run_wide; the permutation test does not cover it)
  if (ctx->binary) {
    // binary trait: the Gram is weighted by the per-sample variance p(1-p), which the integer sweep does not carry:
    // every gene takes the fp64 path (its hard-call tiles are expanded on the device)
    for (auto& w : ctx->wide) ctx->is_dos[w.gene_index] = 1;   // wide genes: run_wide_binary (fp64 statistics from the tiles); no permutation test
    for (int g = 0; g < n; ++g) {
      if (ctx->is_dos[g]) continue;
      if ((size_t)g < ctx->bin_streamed.size() && ctx->bin_streamed[g]) {
        ctx->is_dos[g] = ctx->bin_streamed[g] == 2 ? 1 : 2;   // statistics + tail ran in launch_range; 2 = hard calls: the permutation test applies
        continue;
      }
      const GeneDesc& gd = ctx->genes[g];
      if (!gd.tiled) CTX_FAIL(RVT_E_UNSUPPORTED, "binary trait: caller-owned device blocks are not supported");
      DosGene dg;
      dg.gene_index = g;
      dg.M = gd.M;
      dg.dG = nullptr;   // expanded into one shared scratch block inside the loop below (stream-ordered reuse)
      dg.has_af = gd.has_af != 0;
      if (dg.has_af) dg.af.assign(ctx->af.begin() + gd.var0, ctx->af.begin() + gd.var0 + gd.M);
      ctx->dos.push_back(dg);
      ctx->is_dos[g] = 2;   // hard calls, on the fp64 path only because of the weighted Gram: the permutation test still applies
    }
  }
  if (!ctx->dos.empty()) {
    // genes with dosage / imputed values: fp64 statistics, then the same tail (eigen, Davies, SKAT-O, burden);
    // their records overwrite what the hard-call pipeline produced for the same slots
    const int nd = (int)ctx->dos.size();
    if ((rc = ensure(ctx, &ctx->d_dos_st, &ctx->cap_dos_st, (size_t)nd, sizeof(DosageStats)))) return rc;
    if ((rc = ensure(ctx, &ctx->d_dos_tin, &ctx->cap_dos_tin, (size_t)nd, sizeof(TailInput)))) return rc;
    if ((rc = ensure(ctx, &ctx->d_dos_idx, &ctx->cap_dos_idx, (size_t)nd, sizeof(int)))) return rc;
    if ((rc = ensure(ctx, &ctx->d_dos_afd, &ctx->cap_dos_afd, (size_t)nd * kTileRows, sizeof(double)))) return rc;
    DosageStats* d_st = (DosageStats*)ctx->d_dos_st;
    TailInput* d_tin = (TailInput*)ctx->d_dos_tin;
    int* d_idx = (int*)ctx->d_dos_idx;
    double* d_afd = (double*)ctx->d_dos_afd;
    RVT_CUDA_OK(cudaMemsetAsync(d_st, 0, sizeof(DosageStats) * nd, st));
    std::vector<int> idx(nd);
    double* d_expand = nullptr;   // scratch for tile genes without the engine's own counts (loaded cohort, binary trait)
    std::vector<TileGene> tgs;
    for (int i = 0; i < nd; ++i) {
      DosGene& dg = ctx->dos[i];
      idx[i] = dg.gene_index;
      const bool from_tiles = dg.dG == nullptr;
      if (from_tiles && ctx->genes[dg.gene_index].counted) {
        // statistics straight from the int8 tiles, all such genes in one batched launch below
        const GeneDesc& gd = ctx->genes[dg.gene_index];
        TileGene tg;
        tg.g = gd.g;
        tg.M = gd.M;
        tg.has_af = gd.has_af;
        tg.var0 = gd.var0;
        tg.slot = i;
        tg.allow_missing = dg.from_bed ? 1 : 0;
        tgs.push_back(tg);
        continue;
      }
      if (from_tiles) {
        if (!d_expand) RVT_CUDA_OK(cudaMalloc((void**)&d_expand, sizeof(double) * (size_t)N * kMaxM));
        const GeneDesc& gd = ctx->genes[dg.gene_index];
        dim3 grid((unsigned)((N + 255) / 256), (unsigned)gd.M);
        k_impute_tiled_f64<<<grid, 256, 0, st>>>(gd.g, gd.M, N, ctx->d_counts + gd.var0, d_expand);
        dg.dG = d_expand;
      }
      RVT_CUDA_OK(cudaMemsetAsync(d_st[i].cmin, 0xFF, sizeof(unsigned long long) * kTileRows, st));
      if (dg.has_af) RVT_CUDA_OK(cudaMemcpyAsync(d_afd + (size_t)i * kTileRows, dg.af.data(), sizeof(double) * dg.M, cudaMemcpyHostToDevice, st));
      k_dosage_cols<<<ctx->sm_count, kDosThreads, 0, st>>>(dg.dG, N, dg.M, d_st + i);
      k_dosage_stats<<<ctx->sm_count * 2, kDosThreads, 0, st>>>(dg.dG, N, dg.M, ctx->dX, ctx->C, ctx->dresid, ctx->binary ? ctx->d_vw : nullptr, d_st + i);
      k_dosage_prepare<<<1, 64, 0, st>>>(d_st + i, dg.M, dg.has_af ? d_afd + (size_t)i * kTileRows : nullptr, ctx->d_nm, prm, d_tin + i);
      if (from_tiles) dg.dG = nullptr;   // not owned
    }
    TileGene* d_tg = nullptr;
    if (!tgs.empty()) {
      // Genes with missing calls of a quantitative trait ride the AUGMENTED tensor-core sweep (sweep_aug.cuh: the imputed
      // matrix is H + M diag(delta), every sum an exact integer product over the rows [H ; M]); what it does not cover --
      // a binary trait (weighted Gram), tiles of 63 / 64 variants (no spare rows), another segment, no TMA -- keeps the
      // sparse CUDA-core kernel.  The augmented genes are put first so that both sets are contiguous sub-lists.
      const bool aug_ok = ctx->aug && !ctx->binary && ctx->tc.encode && ctx->tc.have_e && (ctx->ER == 16 || ctx->ER == 32);
      auto is_aug = [&](const TileGene& tg) {
        const GeneDesc& gd = ctx->genes[ctx->dos[tg.slot].gene_index];
        return aug_ok && tg.allow_missing && tg.M <= kTileRows - 2 && gd.tiled && gd.seg == kSegStaged;
      };
      std::stable_partition(tgs.begin(), tgs.end(), is_aug);
      int n_aug = 0;
      while (n_aug < (int)tgs.size() && is_aug(tgs[n_aug])) ++n_aug;
      const int ntg = (int)tgs.size();
      if ((rc = ensure(ctx, &ctx->d_dos_tg, &ctx->cap_dos_tg, (size_t)ntg, sizeof(TileGene)))) return rc;
      d_tg = (TileGene*)ctx->d_dos_tg;
      RVT_CUDA_OK(cudaMemcpyAsync(d_tg, tgs.data(), sizeof(TileGene) * ntg, cudaMemcpyHostToDevice, st));
      for (int t0 = 0; t0 < ntg; t0 += 32768) {   // (grid limit)
        const int nt = std::min(32768, ntg - t0);
        k_tile_cols<<<nt, kTileRows, 0, st>>>(d_tg + t0, nt, N, ctx->d_counts, d_st);
        launches += 1;
      }
      std::vector<GeneDesc> aug_desc;   // (host temporaries: the stream is synchronised below)
      std::vector<AugGene> aug_genes;
      if (n_aug > 0) {
        if ((rc = tc_bind_segment(&ctx->tc, kSegStaged, ctx->d_stage, ctx->stage_cap, ctx->err, sizeof(ctx->err)))) return rc;
        aug_desc.resize(n_aug);
        aug_genes.resize(n_aug);
        for (int i = 0; i < n_aug; ++i) {
          ctx->is_dos[ctx->dos[tgs[i].slot].gene_index] = 3;   // (the permutation test covers these: run_perm)
          aug_desc[i] = ctx->genes[ctx->dos[tgs[i].slot].gene_index];
          aug_genes[i].M = tgs[i].M;
          aug_genes[i].slot = tgs[i].slot;
          aug_genes[i].var0 = tgs[i].var0;
        }
        const int abatch = std::min(n_aug, 1024);
        int S = 0;
        int64_t chunk = 0;
        if ((rc = split_plan(ctx, abatch, &S, &chunk))) return rc;
        void *d_adesc = nullptr, *d_agenes = nullptr;
        if ((rc = scratch(ctx, 6, sizeof(GeneDesc) * (size_t)n_aug, &d_adesc))) return rc;
        if ((rc = scratch(ctx, 7, sizeof(AugGene) * (size_t)n_aug, &d_agenes))) return rc;
        if ((rc = ensure(ctx, (void**)&ctx->d_parts, &ctx->cap_parts, (size_t)abatch * S * kAugParts, sizeof(SweepPartial)))) return rc;
        RVT_CUDA_OK(cudaMemcpyAsync(d_adesc, aug_desc.data(), sizeof(GeneDesc) * n_aug, cudaMemcpyHostToDevice, st));
        RVT_CUDA_OK(cudaMemcpyAsync(d_agenes, aug_genes.data(), sizeof(AugGene) * n_aug, cudaMemcpyHostToDevice, st));
        k_aug_flags<<<n_aug, kTileRows, 0, st>>>((const AugGene*)d_agenes, n_aug, N, d_st, ctx->d_flags);
        for (int b0 = 0; b0 < n_aug; b0 += abatch) {
          const int nb = std::min(abatch, n_aug - b0);
          if ((rc = aug_launch(&ctx->tc, kSegStaged, (const GeneDesc*)d_adesc + b0, aug_desc.data() + b0, nb, ctx->d_flags, N, ctx->ER, S, chunk,
                               ctx->d_parts, ctx->sm_count, st, ctx->err, sizeof(ctx->err))))
            return rc;
          k_aug_stats<<<nb, 128, 0, st>>>((const AugGene*)d_agenes + b0, nb, S, ctx->d_parts, ctx->d_flags, ctx->d_counts, ctx->d_nm, d_st);
          launches += 2;
        }
        launches += 1;
      }
      const int nsp = ntg - n_aug;
      if (nsp > 0) {
        const int64_t nblk = ((N + 3) / 4 + kSparseThreads - 1) / kSparseThreads;
        const unsigned bx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(nblk, (8 * (int64_t)ctx->sm_count + nsp - 1) / nsp));
        for (int t0 = n_aug; t0 < ntg; t0 += 32768) {   // (grid.y limit)
          const int nt = std::min(32768, ntg - t0);
          k_tile_sparse<<<dim3(bx, (unsigned)nt), kSparseThreads, 0, st>>>(d_tg + t0, N, ctx->d_counts, ctx->dX, ctx->C, ctx->dresid,
                                                                           ctx->binary ? ctx->d_vw : nullptr, d_st);
          launches += 1;
        }
      }
      for (int t0 = 0; t0 < ntg; t0 += 32768) {
        const int nt = std::min(32768, ntg - t0);
        k_tile_prepare<<<nt, 64, 0, st>>>(d_tg + t0, nt, d_st, ctx->d_af, ctx->d_nm, prm, d_tin);
        launches += 1;
      }
      RVT_CUDA_OK(cudaGetLastError());
      RVT_CUDA_OK(cudaStreamSynchronize(st));   // `tgs`, `aug_desc`, `aug_genes` are host temporaries
      ctx->last_aug = n_aug;
    }
    RVT_CUDA_OK(cudaMemcpyAsync(d_idx, idx.data(), sizeof(int) * nd, cudaMemcpyHostToDevice, st));
    if (ctx->skato && (rc = ensure(ctx, (void**)&ctx->d_qags, &ctx->cap_qags, qags_scratch_entries(ctx, nd), sizeof(QagsScratch)))) return rc;
    if (ctx->skato && (rc = ensure(ctx, (void**)&ctx->d_jobs, &ctx->cap_jobs, (size_t)nd, sizeof(SkatoJob)))) return rc;
    {
      // SKAT-O for a binary trait (SkatO::Fit type "D": the same tail on the p(1-p)-weighted statistics with s2 = 1,
      // finalize.cuh) runs only with "skato_binary" = 1; otherwise skato_ok stays 0
      const bool sk = ctx->skato && (!ctx->binary || ctx->skato_binary);
      const int kld = fin_kld(kTileRows), fsm = fin_smem(kTileRows, ctx->ER, sk), wm_off = kTileRows * kld * 8;
      if (sk) {
        k_finalize<true><<<nd, kFinThreadsSkato, fsm, st>>>(nullptr, nd, kld, wm_off, fin_uk_off(kTileRows, ctx->ER, sk), ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm, prm, 1,
                                                             nullptr, d_res, nullptr, ctx->d_jobs, d_tin, d_idx);
        if ((rc = launch_qags(ctx, ctx->d_jobs, nd, d_res, d_idx, st))) return rc;
        launches += 1;
      } else if (ctx->fin_split) {
        if ((rc = ensure(ctx, (void**)&ctx->d_mid, &ctx->cap_mid, (size_t)nd, sizeof(FinMid)))) return rc;
        k_finalize<false, true><<<nd, kFinThreads, fsm, st>>>(nullptr, nd, kld, wm_off, fin_uk_off(kTileRows, ctx->ER, false), ctx->d_flags, ctx->d_af, ctx->d_counts,
                                                               ctx->d_nm, prm, 1, nullptr, d_res, nullptr, nullptr, d_tin, d_idx, ctx->d_mid);
        k_fin_sturm<<<nd, kFinThreads, 0, st>>>(ctx->d_mid, nd);
        k_fin_tail<<<nd, kFinThreads, 0, st>>>(ctx->d_mid, nd, ctx->d_nm, d_res, d_idx);
        launches += 2;
      } else
        k_finalize<false><<<nd, kFinThreads, fsm, st>>>(nullptr, nd, kld, wm_off, fin_uk_off(kTileRows, ctx->ER, sk), ctx->d_flags, ctx->d_af, ctx->d_counts, ctx->d_nm, prm, 1,
                                                         nullptr, d_res, nullptr, nullptr, d_tin, d_idx);
    }
    RVT_CUDA_OK(cudaGetLastError());
    RVT_CUDA_OK(cudaStreamSynchronize(st));
    launches += 3 * nd + 1;
    if (d_expand) cudaFree(d_expand);
    for (auto& dg : ctx->dos)
      if (dg.dG) cudaFree(dg.dG);
    ctx->dos.clear();
  }
  mark("fp64 path");
  if ((rc = run_wide(ctx, d_res, &launches))) return rc;
  if ((rc = run_perm(ctx, d_res, n, &launches))) return rc;
  if (!ctx->unsupported.empty()) {
    static rvt_gene_result na;   // (static: outlives the asynchronous copies)
    memset(&na, 0, sizeof(na));
    na.status = RVT_GENE_UNSUPPORTED;
    na.p_skat = na.p_liu = 1.0;
    na.p_davies = -1.0;
    for (int gi : ctx->unsupported)
      RVT_CUDA_OK(cudaMemcpyAsync(d_res + gi, &na, sizeof(na), cudaMemcpyHostToDevice, st));
  }
  RVT_CUDA_OK(cudaMemcpyAsync(out, d_res, sizeof(rvt_gene_result) * n, to_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  RVT_CUDA_OK(cudaEventRecord(ctx->ev[1], st));
  RVT_CUDA_OK(cudaStreamSynchronize(st));
  float tot = 0;
  RVT_CUDA_OK(cudaEventElapsedTime(&tot, ctx->ev[0], ctx->ev[1]));
  float ms_sweep = 0.f, ms_fin = 0.f;
  for (int bi = 0; bi < ctx->pending_timing_batches; ++bi) {
    float a = 0, b = 0;
    RVT_CUDA_OK(cudaEventElapsedTime(&a, ctx->evpool[4 * bi], ctx->evpool[4 * bi + 1]));
    RVT_CUDA_OK(cudaEventElapsedTime(&b, ctx->evpool[4 * bi + 2], ctx->evpool[4 * bi + 3]));
    ms_sweep += a;
    ms_fin += b;
  }
  ctx->t_sweep = ms_sweep;
  ctx->t_fin = ms_fin;
  ctx->t_total = tot;
  ctx->n_launch += launches;
  ctx->last_n = n;
  if (n_out) *n_out = n;
  pending_reset(ctx);
  return RVT_OK;
}

int rvt_flush(rvt_ctx* ctx, rvt_gene_result* out, int cap, int* n_out) { return flush_impl(ctx, out, cap, n_out, false); }
int rvt_debug_rand(rvt_ctx* ctx, uint32_t seed, uint64_t pos, int64_t n, int32_t* out) {
  if (!ctx || !out || n < 0) return RVT_E_BADARG;
  RVT_CUDA_OK(cudaSetDevice(ctx->device));
  int rc = perm_tables(ctx);
  if (rc) return rc;
  uint32_t y0[2 * kLfgDeg - 1], w0[2 * kLfgDeg - 1];
  lfg_seed_window(seed, y0);
  lfg_window_at(lfg_pow(pos + kLfgWarm), y0, w0);
  uint32_t *d_w0 = nullptr, *d_out = nullptr;
  RVT_CUDA_OK(cudaMalloc((void**)&d_w0, sizeof(w0)));
  RVT_CUDA_OK(cudaMalloc((void**)&d_out, sizeof(uint32_t) * (size_t)std::max<int64_t>(n, 1)));
  RVT_CUDA_OK(cudaMemcpyAsync(d_w0, w0, sizeof(w0), cudaMemcpyHostToDevice, ctx->stream));
  if (n > 0) k_lfg_draws<<<(unsigned)((n + kLfgBlock - 1) / kLfgBlock), kLfgThreads, 0, ctx->stream>>>(ctx->d_lfg, d_w0, (uint64_t)n, d_out);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_w0);
  cudaFree(d_out);
  if (e != cudaSuccess) CTX_FAIL(RVT_E_CUDA, "rvt_debug_rand: %s", cudaGetErrorString(e));
  return RVT_OK;
}
int rvt_perm_debug_q(rvt_ctx* ctx, double* out, int cap, int* n_out) {
  if (!ctx || !n_out) return RVT_E_BADARG;
  const int n = (int)ctx->perm_q_log.size();
  *n_out = n;
  if (!out) return RVT_OK;
  if (cap < n) CTX_FAIL(RVT_E_BADARG, "permuted statistics: %d available, cap %d", n, cap);
  if (n) memcpy(out, ctx->perm_q_log.data(), sizeof(double) * n);
  return RVT_OK;
}
int rvt_perm_results(rvt_ctx* ctx, rvt_perm_result* out, int cap, int* n_out) {
  if (!ctx || !n_out) return RVT_E_BADARG;
  const int n = (int)ctx->perm_out.size();
  *n_out = n;
  if (!out) return RVT_OK;
  if (cap < n) CTX_FAIL(RVT_E_BADARG, "permutation records: %d available, cap %d", n, cap);
  if (n) memcpy(out, ctx->perm_out.data(), sizeof(rvt_perm_result) * n);
  return RVT_OK;
}
int rvt_flush_dev(rvt_ctx* ctx, rvt_gene_result* d_out, int cap, int* n_out) {
  return flush_impl(ctx, d_out, cap, n_out, true);
}

// last partner of every variant: jmax[i] = max { j >= i : chrom_j == chrom_i, pos_j - pos_i <= window }
static int meta_jmax(const int32_t* pos, const int32_t* chrom, int64_t nv, int64_t window, std::vector<int>* jmax) {
  jmax->resize(nv);
  int64_t j = 0;
  int wmax = 0;
  for (int64_t i = 0; i < nv; ++i) {
    if (j < i) j = i;
    while (j + 1 < nv && chrom[j + 1] == chrom[i] && (int64_t)pos[j + 1] - (int64_t)pos[i] <= window) ++j;
    // the list is sorted: a later variant on the same chromosome never has a smaller position
    (*jmax)[i] = (int)j;
    wmax = std::max<int>(wmax, (int)(j - i));
  }
  return wmax;
}

int rvt_meta_plan(rvt_ctx* ctx, const int32_t* pos, const int32_t* chrom, int64_t nv, int64_t window_bp, int* wmax) {
  if (!ctx || !pos || !chrom || !wmax || nv < 0) return RVT_E_BADARG;
  std::vector<int> jm;
  *wmax = meta_jmax(pos, chrom, nv, window_bp, &jm);
  return RVT_OK;
}

int rvt_meta_flush(rvt_ctx* ctx, const int32_t* pos, const int32_t* chrom, int64_t window_bp,
                   rvt_variant_result* vout, int64_t cap_variants, double* band, int64_t cap_band, int* wmax_out) {
  if (!ctx || !vout) return RVT_E_BADARG;
  const int ngen = (int)ctx->genes.size();
  const int64_t nv = ctx->n_var;
  if (ngen == 0) return RVT_OK;
  if (cap_variants < nv) CTX_FAIL(RVT_E_BADARG, "vout holds %lld records, %lld variants pending", (long long)cap_variants, (long long)nv);
  if (band && (!pos || !chrom)) CTX_FAIL(RVT_E_BADARG, "the covariance band needs pos and chrom");
  if (!ctx->wide.empty()) CTX_FAIL(RVT_E_UNSUPPORTED, "meta: push variant blocks of at most %d variants", kMaxM);
  RVT_CUDA_OK(cudaSetDevice(ctx->device));
  // every pending push is one tile (<= 64 consecutive variants, its own tiled block) of one segment
  const int seg = ctx->genes[0].seg;
  for (const auto& g : ctx->genes)
    if (g.seg != seg || !g.tiled) CTX_FAIL(RVT_E_UNSUPPORTED, "meta: pending variant blocks must live in one engine-owned (tiled) segment");
  int rc;
  if (seg == kSegStaged) {
    if ((rc = tc_bind_segment(&ctx->tc, kSegStaged, ctx->d_stage, ctx->stage_cap, ctx->err, sizeof(ctx->err)))) return rc;
  }
  std::vector<int> jmax;
  int wmax = 0;
  if (band) {
    wmax = meta_jmax(pos, chrom, nv, window_bp, &jmax);
    if (cap_band < nv * (int64_t)(wmax + 1)) CTX_FAIL(RVT_E_BADARG, "band needs %lld doubles", (long long)(nv * (int64_t)(wmax + 1)));
  } else {
    jmax.assign(nv, 0);
    for (int64_t i = 0; i < nv; ++i) jmax[i] = (int)i;
  }
  if (wmax_out) *wmax_out = wmax;
  const int64_t N = ctx->N;
  int S = ctx->splits;
  if (S <= 0) {
    S = (int)std::min<int64_t>(16, std::max<int64_t>(1, (N + 65535) / 65536));
    if (band) {
      // the pair sweeps walk the band split-major (sweep_tc.cuh): one sample chunk of the tiles a window spans -- the
      // partners of a tile plus the tile itself, 64 bytes per sample each -- should sit in L2 (126 MB) with room to spare
      const int64_t wt = (wmax + kTileRows - 1) / kTileRows + 2;
      const int64_t s_l2 = (N * kTileRows * wt + (48ll << 20) - 1) / (48ll << 20);
      S = (int)std::min<int64_t>(64, std::max<int64_t>(S, s_l2));
    }
  }
  // binary trait: an s32 accumulator holds |g d_k e| <= 128 * 127 per sample for 2^31 / 16256 = 132 104 samples
  if (ctx->binary) S = (int)std::max<int64_t>(S, (N + 65535) / 65536);
  int64_t chunk = (((N + S - 1) / S) + 511) & ~(int64_t)511;
  S = (int)((N + chunk - 1) / chunk);
  const int T = ngen;
  std::vector<GeneDesc> tiles(ctx->genes), pairs;
  std::vector<int> tile_of(nv);
  for (int t = 0; t < T; ++t)
    for (int i = 0; i < tiles[t].M; ++i) tile_of[tiles[t].var0 + i] = t;
  if (band) {
    for (int t = 0; t < T; ++t) {
      int jm = 0;
      for (int i = 0; i < tiles[t].M; ++i) jm = std::max(jm, jmax[tiles[t].var0 + i]);
      for (int u = t + 1; u <= tile_of[jm]; ++u) {
        GeneDesc g = tiles[t];
        g.row0_b = tiles[u].row0;
        g.Mb = tiles[u].M;
        g.var0_b = tiles[u].var0;
        pairs.push_back(g);
      }
    }
  }
  cudaStream_t st = ctx->stream;
  // device buffers
  int* d_jmax = nullptr;
  double* d_B = nullptr;
  uint8_t* d_poly = nullptr;
  rvt_variant_result* d_v = nullptr;
  double* d_band = nullptr;
  GeneDesc* d_desc = nullptr;
  uint8_t* d_flags0 = nullptr;
  auto cleanup = [&]() {};   // (the buffers below persist in the context)
  const int batch = 1024;
  const size_t ndesc = ctx->binary ? (size_t)T + pairs.size() : std::max<size_t>((size_t)T, pairs.size());
  {
    const size_t need[7] = {sizeof(int) * (size_t)nv, sizeof(double) * (size_t)nv * kMaxC, (size_t)nv, sizeof(rvt_variant_result) * (size_t)nv,
                            sizeof(GeneDesc) * ndesc, (size_t)nv + kTileRows, band ? sizeof(double) * (size_t)nv * (size_t)(wmax + 1) : 0};
    for (int k = 0; k < 7; ++k)
      if (need[k] > ctx->cap_scr[k]) {   // contents are rewritten every call: no copy on growth
        if (ctx->d_scr[k]) cudaFree(ctx->d_scr[k]);
        ctx->d_scr[k] = nullptr;
        ctx->cap_scr[k] = 0;
        cudaError_t e = cudaMalloc(&ctx->d_scr[k], need[k]);
        if (e != cudaSuccess) CTX_FAIL(RVT_E_CUDA, "meta: cudaMalloc: %s", cudaGetErrorString(e));
        ctx->cap_scr[k] = need[k];
      }
    d_jmax = (int*)ctx->d_scr[0];
    d_B = (double*)ctx->d_scr[1];
    d_poly = (uint8_t*)ctx->d_scr[2];
    d_v = (rvt_variant_result*)ctx->d_scr[3];
    d_desc = (GeneDesc*)ctx->d_scr[4];
    d_flags0 = (uint8_t*)ctx->d_scr[5];
    d_band = band ? (double*)ctx->d_scr[6] : nullptr;
  }
  if ((rc = ensure(ctx, (void**)&ctx->d_parts, &ctx->cap_parts, (size_t)batch * S, sizeof(SweepPartial)))) { cleanup(); return rc; }
  RVT_CUDA_OK(cudaMemcpyAsync(d_jmax, jmax.data(), sizeof(int) * nv, cudaMemcpyHostToDevice, st));
  RVT_CUDA_OK(cudaMemsetAsync(d_flags0, 0, (size_t)nv + kTileRows, st));   // every row "normal": no flip in meta mode
  if (band) {
    // NaN-fill: entries outside a variant's window stay NaN
    RVT_CUDA_OK(cudaMemsetAsync(d_band, 0xFF, sizeof(double) * nv * (size_t)(wmax + 1), st));
  }
  const bool tc_ok = ctx->tc.encode && ctx->tc.have_e && seg >= 0 && ctx->tc.have_seg[seg];
  if (!pairs.empty() && !tc_ok) { cleanup(); CTX_FAIL(RVT_E_UNSUPPORTED, "meta cov needs the tensor-core engine (TMA segment unavailable)"); }
  // binary trait: scratch of the digit passes
  int8_t *d_dig = nullptr, *d_aux = nullptr;
  double *d_acc = nullptr, *d_covxz = nullptr, *d_uraw = nullptr;
  MetabVar* d_mv = nullptr;
  rvt_variant_cc* d_cc = nullptr;
  unsigned long long* d_ncase = nullptr;
  std::vector<int64_t> aux_off(T + 1, 0);
  const int64_t ldd = ((N + 127) >> 7) * 128;
  ctx->metab_nv = 0;
  if (ctx->binary) {
    if (!tc_ok) { cleanup(); CTX_FAIL(RVT_E_UNSUPPORTED, "meta score/cov for a binary trait needs the tensor-core engine"); }
    for (int t = 0; t < T; ++t) aux_off[t + 1] = aux_off[t] + tiled_bytes(N, tiles[t].M);
    const size_t need[6] = {(size_t)(kMetabDigits + 1) * ldd, (size_t)aux_off[T], sizeof(double) * (size_t)nv * (size_t)(wmax + 1),
                            sizeof(MetabVar) * (size_t)nv + sizeof(double) * (size_t)nv + 16, sizeof(rvt_variant_cc) * (size_t)nv,
                            sizeof(double) * (size_t)nv * kMaxC};
    for (int k = 0; k < 6; ++k)
      if (need[k] > ctx->cap_metab[k]) {
        if (ctx->d_metab[k]) cudaFree(ctx->d_metab[k]);
        ctx->d_metab[k] = nullptr;
        ctx->cap_metab[k] = 0;
        cudaError_t e = cudaMalloc(&ctx->d_metab[k], need[k]);
        if (e != cudaSuccess) CTX_FAIL(RVT_E_CUDA, "meta (binary trait): cudaMalloc of %zu bytes: %s", need[k], cudaGetErrorString(e));
        ctx->cap_metab[k] = need[k];
      }
    d_dig = (int8_t*)ctx->d_metab[0];
    d_aux = (int8_t*)ctx->d_metab[1];
    d_acc = (double*)ctx->d_metab[2];
    d_mv = (MetabVar*)ctx->d_metab[3];
    d_uraw = (double*)((char*)ctx->d_metab[3] + sizeof(MetabVar) * (size_t)nv);
    d_ncase = (unsigned long long*)(d_uraw + nv);
    d_cc = (rvt_variant_cc*)ctx->d_metab[4];
    d_covxz = (double*)ctx->d_metab[5];
    if ((rc = tc_bind_segment(&ctx->tc, kSegAux, d_aux, aux_off[T], ctx->err, sizeof(ctx->err)))) { cleanup(); return rc; }
  }
  // BoltLMM::GetCovXX (regression/BoltLMM.cpp:435-460; MetaCovFamQtlBolt::calculateXX, src/Model.cpp:780-805): the entry
  // is g1'(I - ZZ')g2 * xVx_xx_ratio / N -- the projected Gram of this band times a scalar -- where the unrelated-sample
  // model divides by sigma2 N (option "meta_cov_scale" = xVx_xx_ratio of rvt_bolt_fit_null; 0 = unrelated samples)
  const double band_scale = ctx->meta_cov_scale > 0.0 ? ctx->meta_cov_scale / (double)N : 0.0;
  // phase 1: diagonal tiles
  RVT_CUDA_OK(cudaEventRecord(ctx->ev[0], st));
  RVT_CUDA_OK(cudaMemcpyAsync(d_desc, tiles.data(), sizeof(GeneDesc) * T, cudaMemcpyHostToDevice, st));
  for (int b0 = 0; b0 < T; b0 += batch) {
    const int nb = std::min(batch, T - b0);
    if (tc_ok) {
      rc = tc_launch(&ctx->tc, d_desc + b0, tiles.data() + b0, nb, d_flags0, ctx->d_nm, N, ctx->ER, S, chunk, ctx->d_parts,
                     ctx->d_counter, ctx->sm_count, st, ctx->err, sizeof(ctx->err), false);
      if (rc) { cleanup(); return rc; }
    } else {
      RVT_CUDA_OK(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned int), st));
      k_sweep_simt<<<std::min(nb * S, ctx->sm_count * 3), kSimtThreads, kSimtSmem, st>>>(d_desc + b0, nb, d_flags0, ctx->d_nm, S, chunk,
                                                                                       ctx->d_parts, ctx->d_counter);
    }
    k_meta_block<<<nb, kMetaThreads, 0, st>>>(d_desc + b0, nb, ctx->d_nm, S, ctx->d_parts, d_jmax, wmax, d_v, d_B, d_poly,
                                              ctx->binary ? nullptr : d_band, band_scale, d_uraw);
    RVT_CUDA_OK(cudaGetLastError());
  }
  k_meta_hwe<<<(unsigned)nv, kHweThreads, 0, st>>>(nv, d_v);
  RVT_CUDA_OK(cudaGetLastError());
  // phase 2: tile pairs inside the window
  RVT_CUDA_OK(cudaEventRecord(ctx->ev[2], st));
  if (ctx->binary) {
    // digit passes (meta.cuh): A operand = G o d_k in kSegAux, B operand = the plain tiles; units = the diagonal pairs
    // (t, t) followed by the window's pairs; pass kMetabDigits (d = y) needs the diagonal pairs only
    std::vector<GeneDesc> units((size_t)T + pairs.size());
    for (int t = 0; t < T; ++t) {
      GeneDesc g = tiles[t];
      g.row0_b = tiles[t].row0;
      g.Mb = tiles[t].M;
      g.var0_b = tiles[t].var0;
      units[t] = g;
    }
    for (size_t p = 0; p < pairs.size(); ++p) units[(size_t)T + p] = pairs[p];
    for (auto& u : units) {   // the A tile lives in the scaled copy
      const int t = tile_of[u.var0];
      u.g = d_aux + aux_off[t];
      u.seg = kSegAux;
      u.row0 = aux_off[t] / 128;
    }
    RVT_CUDA_OK(cudaStreamSynchronize(st));  // d_desc is rewritten
    RVT_CUDA_OK(cudaMemcpyAsync(d_desc, units.data(), sizeof(GeneDesc) * units.size(), cudaMemcpyHostToDevice, st));
    RVT_CUDA_OK(cudaMemsetAsync(d_ncase, 0, sizeof(unsigned long long), st));
    k_metab_digits<<<(unsigned)((ldd + 255) / 256), 256, 0, st>>>(N, ctx->d_vw, ctx->dresid, d_dig, ldd, d_ncase);
    const int64_t nwords = ((N + 127) >> 7) * 32;
    for (int k = 0; k <= kMetabDigits; ++k) {
      for (int t = 0; t < T; ++t)
        k_metab_scale<<<dim3((unsigned)((nwords + 255) / 256), (unsigned)tiles[t].M), 256, 0, st>>>(tiles[t].g, tiles[t].M, N, d_dig + (size_t)k * ldd,
                                                                                                  d_aux + aux_off[t]);
      RVT_CUDA_OK(cudaGetLastError());
      const size_t nu = (k == kMetabDigits) ? (size_t)T : units.size();
      for (size_t b0 = 0; b0 < nu; b0 += batch) {
        const int nb = (int)std::min<size_t>(batch, nu - b0);
        rc = tc_launch(&ctx->tc, d_desc + b0, units.data() + b0, nb, d_flags0, ctx->d_nm, N, ctx->ER, S, chunk, ctx->d_parts,
                       ctx->d_counter, ctx->sm_count, st, ctx->err, sizeof(ctx->err), true, false, seg);
        if (rc) { cleanup(); return rc; }
        k_metab_acc<<<nb, kMetaThreads, 0, st>>>(d_desc + b0, nb, ctx->d_nm, S, ctx->d_parts, d_jmax, wmax, k, d_acc, d_mv);
        RVT_CUDA_OK(cudaGetLastError());
      }
    }
    k_metab_final<<<(unsigned)((nv + 127) / 128), 128, 0, st>>>(nv, ctx->d_nm, d_acc, wmax, d_mv, d_uraw, d_ncase, d_v, d_cc, d_covxz);
    k_meta_hwe_cc<<<(unsigned)(2 * nv), kHweThreads, 0, st>>>(nv, d_cc);
    if (band) {
      const int64_t ne = nv * (int64_t)(wmax + 1);
      k_metab_band<<<(unsigned)((ne + 127) / 128), 128, 0, st>>>(nv, ctx->d_nm, d_acc, d_covxz, d_poly, d_jmax, wmax, d_band);
    }
    RVT_CUDA_OK(cudaGetLastError());
    ctx->metab_nv = nv;
  } else if (!pairs.empty()) {
    RVT_CUDA_OK(cudaStreamSynchronize(st));  // d_desc is rewritten
    RVT_CUDA_OK(cudaMemcpyAsync(d_desc, pairs.data(), sizeof(GeneDesc) * pairs.size(), cudaMemcpyHostToDevice, st));
    for (size_t b0 = 0; b0 < pairs.size(); b0 += batch) {
      const int nb = (int)std::min<size_t>(batch, pairs.size() - b0);
      rc = tc_launch(&ctx->tc, d_desc + b0, pairs.data() + b0, nb, d_flags0, ctx->d_nm, N, ctx->ER, S, chunk, ctx->d_parts,
                     ctx->d_counter, ctx->sm_count, st, ctx->err, sizeof(ctx->err), true);
      if (rc) { cleanup(); return rc; }
      k_meta_pair<<<nb, kMetaThreads, 0, st>>>(d_desc + b0, nb, ctx->d_nm, S, ctx->d_parts, d_jmax, wmax, d_B, d_poly, d_band, band_scale);
      RVT_CUDA_OK(cudaGetLastError());
    }
  }
  RVT_CUDA_OK(cudaEventRecord(ctx->ev[3], st));
  // (vout / band may be host or device pointers: unified addressing tells)
  RVT_CUDA_OK(cudaMemcpyAsync(vout, d_v, sizeof(rvt_variant_result) * nv, cudaMemcpyDefault, st));
  if (band) RVT_CUDA_OK(cudaMemcpyAsync(band, d_band, sizeof(double) * nv * (size_t)(wmax + 1), cudaMemcpyDefault, st));
  RVT_CUDA_OK(cudaEventRecord(ctx->ev[1], st));
  RVT_CUDA_OK(cudaStreamSynchronize(st));
  {
    // rvt_last_timing of a meta flush: [0] the tile-pair sweeps + band assembly (phase 2), [1] the diagonal tiles + score
    // statistics (phase 1), [2] the whole flush incl. the copies of the results
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&b, ctx->ev[0], ctx->ev[2]);
    cudaEventElapsedTime(&c, ctx->ev[0], ctx->ev[1]);
    ctx->t_sweep = a;
    ctx->t_fin = b;
    ctx->t_total = c;
    ctx->n_launch = (double)pairs.size();
    ctx->last_S = S;
  }
  cleanup();
  pending_reset(ctx);
  return RVT_OK;
}

// ---- A13: FastLMM score step (lmm.cuh) -------------------------------------------------------------
int rvt_lmm_set_null(rvt_ctx* ctx, int64_t N, int C, const float* U, const float* lambda, double delta, double sigma2,
                     const float* uResid, const float* ux) {
  if (!ctx || !U || !lambda || !uResid || !ux) return RVT_E_BADARG;
  if (!(sigma2 > 0.0) || !(delta >= 0.0)) CTX_FAIL(RVT_E_BADARG, "lmm: sigma2 must be > 0 and delta >= 0");
  if (C < 1 || C > kMaxC) CTX_FAIL(RVT_E_UNSUPPORTED, "lmm: C=%d covariate columns; this build supports 1..%d", C, kMaxC);
  // the sweep needs a null-model image for its digit tile: a trivial one (intercept, zero residual)
  int rc = null_model_alloc(ctx, N, 1);
  if (rc) return rc;
  {
    std::vector<double> ones((size_t)N, 1.0), zeros((size_t)N, 0.0);
    RVT_CUDA_OK(cudaMemcpy(ctx->dX, ones.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
    RVT_CUDA_OK(cudaMemcpy(ctx->dy, zeros.data(), sizeof(double) * N, cudaMemcpyHostToDevice));
  }
  if ((rc = null_model_run(ctx, true, 1.0))) return rc;
  if (!(ctx->tc.encode && ctx->tc.have_e))
    CTX_FAIL(RVT_E_UNSUPPORTED, "lmm: the score step needs the tensor-core sweep (TMA unavailable: %s)", ctx->tc.why);
  ctx->have_lmm = false;
  for (void* p : {(void*)ctx->d_lmm, (void*)ctx->d_lmm_tiles, (void*)ctx->d_lmm_vec, (void*)ctx->d_lmm_tsum})
    if (p) cudaFree(p);
  ctx->d_lmm = nullptr; ctx->d_lmm_tiles = nullptr; ctx->d_lmm_vec = nullptr; ctx->d_lmm_tsum = nullptr;
  const int nb = (int)((N + 15) / 16);
  const int64_t tile_b = tiled_bytes(N, kTileRows), nev = 16 * (int64_t)nb;
  cudaStream_t st = ctx->stream;
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->d_lmm, sizeof(LmmNull)));
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->d_lmm_tiles, (size_t)nb * tile_b));
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->d_lmm_vec, sizeof(double) * nev * (3 + kMaxC)));
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->d_lmm_tsum, sizeof(long long) * nev));
  RVT_CUDA_OK(cudaMemsetAsync(ctx->d_lmm_tiles, 0, (size_t)nb * tile_b, st));
  RVT_CUDA_OK(cudaMemsetAsync(ctx->d_lmm_vec, 0, sizeof(double) * nev * (3 + kMaxC), st));
  RVT_CUDA_OK(cudaMemsetAsync(ctx->d_lmm_tsum, 0, sizeof(long long) * nev, st));
  // eigenvectors, a panel of columns at a time
  const int64_t pcols = std::max<int64_t>(1, std::min<int64_t>(N, ((int64_t)256 << 20) / (4 * N)));
  float *d_panel = nullptr, *d_f = nullptr;
  RVT_CUDA_OK(cudaMalloc((void**)&d_panel, sizeof(float) * pcols * N));
  RVT_CUDA_OK(cudaMalloc((void**)&d_f, sizeof(float) * N * (2 + C)));
  for (int64_t c0 = 0; c0 < N; c0 += pcols) {
    const int64_t nc = std::min(pcols, N - c0);
    RVT_CUDA_OK(cudaMemcpyAsync(d_panel, U + (size_t)c0 * N, sizeof(float) * nc * N, cudaMemcpyHostToDevice, st));
    k_lmm_digits<<<(unsigned)nc, 256, 0, st>>>(d_panel, N, c0, ctx->d_lmm_tiles, tile_b, ctx->d_lmm_tsum);
    RVT_CUDA_OK(cudaStreamSynchronize(st));   // the panel buffer is reused
  }
  RVT_CUDA_OK(cudaMemcpyAsync(d_f, lambda, sizeof(float) * N, cudaMemcpyHostToDevice, st));
  RVT_CUDA_OK(cudaMemcpyAsync(d_f + N, uResid, sizeof(float) * N, cudaMemcpyHostToDevice, st));
  RVT_CUDA_OK(cudaMemcpyAsync(d_f + 2 * N, ux, sizeof(float) * N * C, cudaMemcpyHostToDevice, st));
  double *d_a = ctx->d_lmm_vec, *d_d = d_a + nev, *d_t = d_d + nev, *d_w = d_t + nev;
  k_lmm_consts<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, C, d_f, d_f + N, d_f + 2 * N, ctx->d_lmm_tsum, delta, sigma2, d_a, d_d, d_t, d_w);
  LmmNull h;
  memset(&h, 0, sizeof(h));
  h.N = N; h.C = C; h.nb = nb; h.delta = delta; h.sigma2 = sigma2;
  h.a = d_a; h.d = d_d; h.t = d_t; h.w = d_w;
  {   // (ux' D ux)^-1 in double (the reference: float .inverse(), FastLMM.cpp:133-137)
    double A[kMaxC * kMaxC], I[kMaxC * kMaxC];
    for (int l = 0; l < C; ++l)
      for (int m = 0; m < C; ++m) {
        double sacc = 0.0;
        for (int64_t i = 0; i < N; ++i) sacc += (double)ux[(size_t)l * N + i] * (double)ux[(size_t)m * N + i] / ((double)fabsf(lambda[i]) + delta);
        A[l * C + m] = sacc;
        I[l * C + m] = (l == m) ? 1.0 : 0.0;
      }
    for (int k = 0; k < C; ++k) {   // Gauss-Jordan with partial pivoting
      int piv = k;
      for (int r = k + 1; r < C; ++r)
        if (fabs(A[r * C + k]) > fabs(A[piv * C + k])) piv = r;
      if (!(fabs(A[piv * C + k]) > 0.0)) {
        cudaFree(d_panel); cudaFree(d_f);
        CTX_FAIL(RVT_E_NUMERIC, "lmm: ux' D ux is singular");
      }
      for (int c = 0; c < C; ++c) { std::swap(A[k * C + c], A[piv * C + c]); std::swap(I[k * C + c], I[piv * C + c]); }
      const double inv = 1.0 / A[k * C + k];
      for (int c = 0; c < C; ++c) { A[k * C + c] *= inv; I[k * C + c] *= inv; }
      for (int r = 0; r < C; ++r)
        if (r != k) {
          const double f = A[r * C + k];
          for (int c = 0; c < C; ++c) { A[r * C + c] -= f * A[k * C + c]; I[r * C + c] -= f * I[k * C + c]; }
        }
    }
    memcpy(h.xdx_inv, I, sizeof(double) * C * C);
  }
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->d_lmm, &h, sizeof(h), cudaMemcpyHostToDevice, st));
  k_lmm_consts2<<<1, 32, 0, st>>>(ctx->d_lmm);
  RVT_CUDA_OK(cudaGetLastError());
  RVT_CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(d_panel);
  cudaFree(d_f);
  ctx->h_lmm = h;
  if ((rc = tc_bind_segment(&ctx->tc, kSegLmm, ctx->d_lmm_tiles, (int64_t)nb * tile_b, ctx->err, sizeof(ctx->err)))) return rc;
  ctx->have_lmm = true;
  return RVT_OK;
}

int rvt_meta_binary_extras(rvt_ctx* ctx, rvt_variant_cc* cc, double* cov_xz, double* cov_zz, int64_t cap_variants) {
  if (!ctx) return RVT_E_BADARG;
  if (!ctx->binary || ctx->metab_nv == 0) CTX_FAIL(RVT_E_STATE, "no binary-trait rvt_meta_flush to report on");
  const int64_t nv = ctx->metab_nv;
  const int C = ctx->C;
  if ((cc || cov_xz) && cap_variants < nv) CTX_FAIL(RVT_E_BADARG, "%lld variants in the last flush, room for %lld", (long long)nv, (long long)cap_variants);
  RVT_CUDA_OK(cudaSetDevice(ctx->device));
  if (cc) RVT_CUDA_OK(cudaMemcpyAsync(cc, ctx->d_metab[4], sizeof(rvt_variant_cc) * (size_t)nv, cudaMemcpyDefault, ctx->stream));
  if (cov_xz) {
    std::vector<double> h((size_t)nv * kMaxC);
    RVT_CUDA_OK(cudaMemcpyAsync(h.data(), ctx->d_metab[5], sizeof(double) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    for (int64_t v = 0; v < nv; ++v)
      for (int l = 0; l < C; ++l) cov_xz[v * C + l] = h[(size_t)v * kMaxC + l];
  }
  if (cov_zz)
    for (int l = 0; l < C * C; ++l) cov_zz[l] = ctx->h_xvx[l];
  RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return RVT_OK;
}

static int meta_jmax(const int32_t* pos, const int32_t* chrom, int64_t nv, int64_t window, std::vector<int>* jmax);
static int lmm_flush_impl(rvt_ctx* ctx, rvt_lmm_result* out, int64_t cap, const int32_t* pos, const int32_t* chrom, int64_t window_bp,
                          double* band, int64_t cap_band, int* wmax_out);
int rvt_lmm_flush(rvt_ctx* ctx, rvt_lmm_result* out, int64_t cap) { return lmm_flush_impl(ctx, out, cap, nullptr, nullptr, 0, nullptr, 0, nullptr); }
int rvt_lmm_meta_flush(rvt_ctx* ctx, const int32_t* pos, const int32_t* chrom, int64_t window_bp, rvt_lmm_result* out, int64_t cap, double* band,
                       int64_t cap_band, int* wmax) {
  if (!pos || !chrom || !band) return RVT_E_BADARG;
  return lmm_flush_impl(ctx, out, cap, pos, chrom, window_bp, band, cap_band, wmax);
}

static int lmm_flush_impl(rvt_ctx* ctx, rvt_lmm_result* out, int64_t cap, const int32_t* pos, const int32_t* chrom, int64_t window_bp,
                          double* band, int64_t cap_band, int* wmax_out) {
  if (!ctx || !out) return RVT_E_BADARG;
  if (!ctx->have_lmm || ctx->h_lmm.N != ctx->N) CTX_FAIL(RVT_E_STATE, "lmm: call rvt_lmm_set_null first");
  const int ngen = (int)ctx->genes.size();
  const int64_t nv = ctx->n_var;
  if (ngen == 0) return RVT_OK;
  if (cap < nv) CTX_FAIL(RVT_E_BADARG, "lmm: out holds %lld records, %lld variants pending", (long long)cap, (long long)nv);
  if (!ctx->wide.empty()) CTX_FAIL(RVT_E_UNSUPPORTED, "lmm: push variant blocks of at most %d variants", kMaxM);
  for (const auto& g : ctx->genes)
    if (g.seg != kSegStaged || !g.tiled) CTX_FAIL(RVT_E_UNSUPPORTED, "lmm: pending variant blocks must be host pushes (tiled, staged)");
  RVT_CUDA_OK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = tc_bind_segment(&ctx->tc, kSegStaged, ctx->d_stage, ctx->stage_cap, ctx->err, sizeof(ctx->err)))) return rc;
  cudaStream_t st = ctx->stream;
  const int64_t N = ctx->N;
  const int nb = ctx->h_lmm.nb;
  const int64_t tile_b = tiled_bytes(N, kTileRows);
  if ((rc = ensure(ctx, (void**)&ctx->d_zero_flags, &ctx->cap_zero_flags, (size_t)nv + kTileRows, 1))) return rc;
  RVT_CUDA_OK(cudaMemsetAsync(ctx->d_zero_flags, 0, (size_t)nv + kTileRows, st));
  const int batch = 1024;
  int S = 0;
  int64_t chunk = 0;
  if ((rc = split_plan(ctx, std::min(nb, batch), &S, &chunk))) return rc;
  if ((rc = ensure(ctx, (void**)&ctx->d_parts, &ctx->cap_parts, (size_t)std::min(nb, batch) * S, sizeof(SweepPartial)))) return rc;
  GeneDesc* d_units = nullptr;
  double* d_acc = nullptr;
  rvt_lmm_result* d_out = nullptr;
  if ((rc = scratch(ctx, 0, sizeof(GeneDesc) * nb, (void**)&d_units))) return rc;
  if ((rc = scratch(ctx, 1, sizeof(double) * (size_t)nb * kTileRows * kLmmAcc, (void**)&d_acc))) return rc;
  if ((rc = scratch(ctx, 2, sizeof(rvt_lmm_result) * nv, (void**)&d_out))) return rc;
  auto cleanup = [&]() {};   // (scratch slots persist in the context)
  // covariance band (MetaCovFamQtl): keep the rotated rows, plan the tile pairs of the window
  std::vector<int> jmax;
  int wmax = 0;
  double *d_Y = nullptr, *d_band = nullptr;
  LmmVar* d_vars = nullptr;
  const int64_t ldY = 16 * (int64_t)nb;
  if (band) {
    wmax = meta_jmax(pos, chrom, nv, window_bp, &jmax);
    if (wmax_out) *wmax_out = wmax;
    if (cap_band < nv * (int64_t)(wmax + 1)) CTX_FAIL(RVT_E_BADARG, "lmm: band needs %lld doubles", (long long)(nv * (int64_t)(wmax + 1)));
    if ((double)nv * (double)ldY * 8.0 > 48e9) CTX_FAIL(RVT_E_UNSUPPORTED, "lmm: %lld variants x %lld eigenvectors of rotated genotypes exceed the 48 GB kept for them; flush in segments", (long long)nv, (long long)ldY);
    if ((rc = scratch(ctx, 3, sizeof(double) * (size_t)nv * ldY, (void**)&d_Y))) return rc;
    if ((rc = scratch(ctx, 4, sizeof(LmmVar) * (size_t)nv, (void**)&d_vars))) return rc;
    if ((rc = scratch(ctx, 5, sizeof(double) * (size_t)nv * (wmax + 1), (void**)&d_band))) return rc;
    RVT_CUDA_OK(cudaMemsetAsync(d_band, 0xFF, sizeof(double) * (size_t)nv * (wmax + 1), st));   // NaN: outside the window
  }
  std::vector<GeneDesc> units(nb);
  for (int g = 0; g < ngen; ++g) {
    const GeneDesc& gd = ctx->genes[g];
    for (int b = 0; b < nb; ++b) {
      units[b] = gd;
      units[b].row0_b = (int64_t)b * tile_b / 128;
      units[b].Mb = kTileRows;
      units[b].var0_b = b;
    }
    cudaError_t e = cudaMemcpyAsync(d_units, units.data(), sizeof(GeneDesc) * nb, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { cleanup(); CTX_FAIL(RVT_E_CUDA, "lmm: %s", cudaGetErrorString(e)); }
    for (int b0 = 0; b0 < nb; b0 += batch) {
      const int n = std::min(batch, nb - b0);
      rc = tc_launch(&ctx->tc, d_units + b0, units.data() + b0, n, ctx->d_zero_flags, ctx->d_nm, N, ctx->ER, S, chunk, ctx->d_parts,
                     ctx->d_counter, ctx->sm_count, st, ctx->err, sizeof(ctx->err), true, false, kSegLmm);
      if (rc) { cleanup(); return rc; }
      k_lmm_reduce<<<n, 64, 0, st>>>(d_units + b0, n, S, ctx->d_parts, ctx->d_lmm, d_acc + (size_t)b0 * kTileRows * kLmmAcc, d_Y, ldY, gd.var0);
    }
    k_lmm_final<<<1, 64, 0, st>>>(gd.M, nb, d_acc, ctx->d_counts + gd.var0, ctx->d_lmm, d_out + gd.var0, d_vars ? d_vars + gd.var0 : nullptr);
  }
  if (band) {
    // tile pairs inside the window (diagonal pairs included), fp64 Gram of the kept rows, band assembly
    std::vector<LmmPair> pairs;
    std::vector<int> tile_of(nv);
    for (int t = 0; t < ngen; ++t)
      for (int i = 0; i < ctx->genes[t].M; ++i) tile_of[ctx->genes[t].var0 + i] = t;
    for (int t = 0; t < ngen; ++t) {
      int jm = 0;
      for (int i = 0; i < ctx->genes[t].M; ++i) jm = std::max(jm, jmax[ctx->genes[t].var0 + i]);
      for (int u = t; u <= tile_of[jm]; ++u) pairs.push_back(LmmPair{ctx->genes[t].var0, ctx->genes[u].var0, ctx->genes[t].M, ctx->genes[u].M});
    }
    LmmPair* d_pairs = nullptr;
    int* d_jmax = nullptr;
    double* d_G = nullptr;
    const int pbatch = 2048;
    if ((rc = scratch(ctx, 6, sizeof(LmmPair) * pairs.size(), (void**)&d_pairs))) return rc;
    if ((rc = scratch(ctx, 7, sizeof(int) * (size_t)nv, (void**)&d_jmax))) return rc;
    if ((rc = scratch(ctx, 1, sizeof(double) * (size_t)std::min<size_t>(pairs.size(), pbatch) * kTileRows * kTileRows, (void**)&d_G))) return rc;   // (d_acc is dead)
    RVT_CUDA_OK(cudaMemcpyAsync(d_pairs, pairs.data(), sizeof(LmmPair) * pairs.size(), cudaMemcpyHostToDevice, st));
    RVT_CUDA_OK(cudaMemcpyAsync(d_jmax, jmax.data(), sizeof(int) * (size_t)nv, cudaMemcpyHostToDevice, st));
    for (size_t p0 = 0; p0 < pairs.size(); p0 += pbatch) {
      const int np = (int)std::min<size_t>(pbatch, pairs.size() - p0);
      k_lmm_gram<<<np, 256, 0, st>>>(d_pairs + p0, np, d_Y, ldY, N, d_G);
      k_lmm_band<<<np, 128, 0, st>>>(d_pairs + p0, np, d_G, d_vars, ctx->d_lmm, d_jmax, wmax, d_band);
    }
    RVT_CUDA_OK(cudaGetLastError());
    RVT_CUDA_OK(cudaMemcpyAsync(band, d_band, sizeof(double) * (size_t)nv * (wmax + 1), cudaMemcpyDefault, st));
    RVT_CUDA_OK(cudaStreamSynchronize(st));   // `pairs`, `jmax` are host temporaries
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(rvt_lmm_result) * nv, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cleanup();
  if (e != cudaSuccess) CTX_FAIL(RVT_E_CUDA, "lmm: %s", cudaGetErrorString(e));
  pending_reset(ctx);
  return RVT_OK;
}

// ---- A15: BoltLMM null-model fit (bolt.cuh) ---------------------------------------------------------
}  // extern "C"
namespace {
struct BoltDev {
  int64_t N = 0, stride = 0, split_len = 0;
  int M = 0, C = 0, splits = 1;
  uint8_t* bed = nullptr;
  double *Z = nullptr, *tab = nullptr, *zg = nullptr, *gnorm2 = nullptr, *part = nullptr, *Xy = nullptr, *dotp = nullptr, *coef = nullptr;
  cudaStream_t st = nullptr;
  std::vector<void*> owned;
  cudaError_t err = cudaSuccess;
  int cg_total = 0;
  // variant sharding: this rank holds SNP rows [m_off, m_off + M) of M_total; every product that sums over
  // SNPs is a local partial followed by ONE sum over ranks through the caller's collective (ncclAllReduce on `st`)
  int64_t M_total = 0, m_off = 0;
  rvt_allreduce_fn ar = nullptr;
  void* ar_user = nullptr;
  int ar_calls = 0;
  int ar_rc = 0;
  void allreduce(double* buf, int64_t n) {
    if (!ar || err != cudaSuccess || ar_rc) return;
    ++ar_calls;
    ar_rc = ar(ar_user, buf, n, (void*)st);
  }
  template <class T>
  T* alloc(size_t n) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e != cudaSuccess) {
      if (err == cudaSuccess) err = e;
      return nullptr;
    }
    owned.push_back(p);
    return (T*)p;
  }
  void note() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && err == cudaSuccess) err = e;
  }
  ~BoltDev() {
    for (void* p : owned) cudaFree(p);
  }
  size_t rows() const { return (size_t)(N + C); }
  double* vec(int R) { return alloc<double>(rows() * R); }
  int gen = 3;                        // option "bolt_kernels": 1 = first-generation product kernels, 2 = k_bolt_xtv2 / k_bolt_xw2, 3 = the cp.async-pipelined k_bolt_xtv3 / k_bolt_xw3
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_xtv, ev_xw;   // per-launch events of the two panel products
  double ms_xtv = 0.0, ms_xw = 0.0;
  int n_hx = 0;
  cudaEvent_t tick() {
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    return e;
  }
  void collect_times() {             // after a stream synchronize
    for (int w = 0; w < 2; ++w) {
      auto& ev = w ? ev_xw : ev_xtv;
      for (auto& pr : ev) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) (w ? ms_xw : ms_xtv) += ms;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
      }
      ev.clear();
    }
  }
  void launch_xtv(const double* v, int R) {
    cudaEvent_t e0 = tick();
    launch_xtv_(v, R);
    ev_xtv.push_back({e0, tick()});
  }
  template <int RMAX, int CHUNK>
  void xtv3(dim3 g, const double* v, int R, int r0) {
    typedef BoltXtv3Smem<RMAX, CHUNK> Smem;
    static bool attr_set = false;   // per instantiation
    if (!attr_set) {
      cudaFuncSetAttribute(k_bolt_xtv3<RMAX, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
      attr_set = true;
    }
    k_bolt_xtv3<RMAX, CHUNK><<<g, kBoltSnpBlock, sizeof(Smem), st>>>(bed, stride, N, M, tab, v, R, r0, split_len, part);
  }
  template <int RMAX>
  void xw3(const double* W, int R, int r0, double alpha, double beta, const double* add, double* out) {
    const unsigned g = (unsigned)((N + kBoltXw3Threads * 4 - 1) / (kBoltXw3Threads * 4));
    k_bolt_xw3<RMAX><<<g, kBoltXw3Threads, sizeof(BoltXw3Smem<RMAX>), st>>>(bed, stride, N, M, tab, W, R, r0, alpha, beta, add, out);
  }
  void launch_xtv_(const double* v, int R) {
    if (gen >= 3) {
      dim3 g2((unsigned)((M + kBoltXtv2Block - 1) / kBoltXtv2Block), (unsigned)splits);
      for (int r0 = 0; r0 < R; r0 += 16) {
        // 128-sample chunks: 34 / 42 / 58 KB of shared memory per CTA, i.e. 6 / 5 / 3 CTAs per SM (512-sample chunks left
        // one warp per scheduler: 27 % of the fp64 pipe, "wait" stalls -- profiles/r02t_bolt_ncu.txt)
        if (R - r0 <= 4) xtv3<4, 128>(g2, v, R, r0);
        else if (R - r0 <= 8) xtv3<8, 128>(g2, v, R, r0);
        else xtv3<16, 128>(g2, v, R, r0);
      }
      return;
    }
    if (gen >= 2) {
      dim3 g2((unsigned)((M + kBoltXtv2Block - 1) / kBoltXtv2Block), (unsigned)splits);
      for (int r0 = 0; r0 < R; r0 += 16) {
        if (R - r0 <= 4) k_bolt_xtv2<4><<<g2, kBoltSnpBlock, 0, st>>>(bed, stride, N, M, tab, v, R, r0, split_len, part);
        else if (R - r0 <= 8) k_bolt_xtv2<8><<<g2, kBoltSnpBlock, 0, st>>>(bed, stride, N, M, tab, v, R, r0, split_len, part);
        else k_bolt_xtv2<16><<<g2, kBoltSnpBlock, 0, st>>>(bed, stride, N, M, tab, v, R, r0, split_len, part);
      }
      return;
    }
    dim3 g1((unsigned)((M + kBoltXtvBlock - 1) / kBoltXtvBlock), (unsigned)splits);
    if (R <= 4) k_bolt_xtv<4><<<g1, kBoltSnpBlock, 0, st>>>(bed, stride, N, M, tab, v, R, split_len, part);
    else if (R <= 8) k_bolt_xtv<8><<<g1, kBoltSnpBlock, 0, st>>>(bed, stride, N, M, tab, v, R, split_len, part);
    else if (R <= 16) k_bolt_xtv<16><<<g1, kBoltSnpBlock, 0, st>>>(bed, stride, N, M, tab, v, R, split_len, part);
    else k_bolt_xtv<32><<<g1, kBoltSnpBlock, 0, st>>>(bed, stride, N, M, tab, v, R, split_len, part);
  }
  // out = (X X'/M + delta I) v on [v ; Z'v]   (computeHx, BoltLMM.cpp:931-993)
  void Hx(double delta, const double* v, double* out, int R) {
    ++n_hx;
    launch_xtv(v, R);
    k_bolt_xtv_finish<<<(unsigned)(((int64_t)M * R + 255) / 256), 256, 0, st>>>(M, R, C, splits, part, zg, v + (size_t)N * R, 1.0, Xy);
    XW(Xy, R, 1.0 / (double)M_total, delta, v, out);
  }
  // out = alpha [X ; Z'X] W + beta add
  void XW(const double* W, int R, double alpha, double beta, const double* add, double* out) {
    const unsigned g1 = (unsigned)((N + 255) / 256), g4 = (unsigned)((N + 1023) / 1024);
    if (ar) {   // sharded: local partial over this rank's SNPs, one sum over ranks, then the + beta add term
      const double* add0 = add;
      const double beta0 = beta;
      add = nullptr;
      beta = 0.0;
      launch_xw(g1, g4, W, R, alpha, beta, add, out);
      allreduce(out, (int64_t)rows() * R);
      if (add0) k_bolt_axpy<<<(unsigned)((rows() * R + 255) / 256), 256, 0, st>>>((int64_t)rows() * R, beta0, add0, out);
      note();
      return;
    }
    launch_xw(g1, g4, W, R, alpha, beta, add, out);
    note();
  }
  void launch_xw(unsigned g1, unsigned g4, const double* W, int R, double alpha, double beta, const double* add, double* out) {
    cudaEvent_t e0 = tick();
    launch_xw_(g1, g4, W, R, alpha, beta, add, out);
    ev_xw.push_back({e0, tick()});
  }
  void launch_xw_(unsigned g1, unsigned g4, const double* W, int R, double alpha, double beta, const double* add, double* out) {
    if (gen >= 3) {
      for (int r0 = 0; r0 < R; r0 += 16) {
        if (R - r0 <= 4) xw3<4>(W, R, r0, alpha, beta, add, out);
        else if (R - r0 <= 8) xw3<8>(W, R, r0, alpha, beta, add, out);
        else xw3<16>(W, R, r0, alpha, beta, add, out);
      }
      k_bolt_bot<<<(C * R + 63) / 64, 64, 0, st>>>(M, R, C, zg, W, alpha, beta, add ? add + (size_t)N * R : nullptr, out + (size_t)N * R);
      return;
    }
    if (gen >= 2) {
      const unsigned g = (unsigned)((N + 255) / 256);
      for (int r0 = 0; r0 < R; r0 += 16) {
        if (R - r0 <= 4) k_bolt_xw2<4><<<g, 64, 0, st>>>(bed, stride, N, M, tab, W, R, r0, alpha, beta, add, out);
        else if (R - r0 <= 8) k_bolt_xw2<8><<<g, 64, 0, st>>>(bed, stride, N, M, tab, W, R, r0, alpha, beta, add, out);
        else k_bolt_xw2<16><<<g, 64, 0, st>>>(bed, stride, N, M, tab, W, R, r0, alpha, beta, add, out);
      }
      k_bolt_bot<<<(C * R + 63) / 64, 64, 0, st>>>(M, R, C, zg, W, alpha, beta, add ? add + (size_t)N * R : nullptr, out + (size_t)N * R);
      return;
    }
    if (R <= 4) k_bolt_xw<4, 4><<<g4, 256, 0, st>>>(bed, stride, N, M, tab, W, R, alpha, beta, add, out);
    else if (R <= 8) k_bolt_xw<8, 4><<<g4, 256, 0, st>>>(bed, stride, N, M, tab, W, R, alpha, beta, add, out);
    else if (R <= 16) k_bolt_xw<16, 1><<<g1, 256, 0, st>>>(bed, stride, N, M, tab, W, R, alpha, beta, add, out);
    else k_bolt_xw<32, 1><<<g1, 256, 0, st>>>(bed, stride, N, M, tab, W, R, alpha, beta, add, out);
    k_bolt_bot<<<(C * R + 63) / 64, 64, 0, st>>>(M, R, C, zg, W, alpha, beta, add ? add + (size_t)N * R : nullptr, out + (size_t)N * R);
  }
  // X_minus' v / scale -> Xy (host copy optional)
  void XtV(const double* v, int R, double scale) {
    launch_xtv(v, R);
    k_bolt_xtv_finish<<<(unsigned)(((int64_t)M * R + 255) / 256), 256, 0, st>>>(M, R, C, splits, part, zg, v + (size_t)N * R, scale, Xy);
    note();
  }
  void project(double* v, int R) {   // bottom rows = Z' top rows
    k_bolt_project<<<C * R, 256, 0, st>>>(N, R, C, Z, v, v + (size_t)N * R);
    note();
  }
  // projDot / projNorm2 per column (BoltLMM.cpp:1064-1138): two fixed-order stages on the device, R scalars to the host
  void pdot(const double* a, const double* b, int R, double* out) {
    k_bolt_dot<<<kBoltDotCtas, 256, 0, st>>>(N, R, a, b, dotp);
    k_bolt_dot_finish<<<1, 64, 0, st>>>(kBoltDotCtas, R, C, dotp, a + (size_t)N * R, b + (size_t)N * R, coef + 2 * kBoltMaxR);
    cudaMemcpyAsync(out, coef + 2 * kBoltMaxR, sizeof(double) * R, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess && err == cudaSuccess) err = e;
    note();
  }
  // column sums of squares of an M x R matrix (|beta_hat|^2 per right-hand side)
  void colnorm2(const double* a, int64_t rows_, int R, double* out, bool over_snps = true) {
    k_bolt_dot<<<kBoltDotCtas, 256, 0, st>>>(rows_, R, a, a, dotp);
    k_bolt_dot_finish<<<1, 64, 0, st>>>(kBoltDotCtas, R, 0, dotp, nullptr, nullptr, coef + 2 * kBoltMaxR);
    if (over_snps) allreduce(coef + 2 * kBoltMaxR, R);
    cudaMemcpyAsync(out, coef + 2 * kBoltMaxR, sizeof(double) * R, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess && err == cudaSuccess) err = e;
    note();
  }
  // y = ca .* a + cb .* b (columnwise coefficients), all N + C rows
  void axpby(const double* ca, const double* a, const double* cb, const double* b, double* y, int R) {
    cudaMemcpyAsync(coef, ca, sizeof(double) * R, cudaMemcpyHostToDevice, st);
    if (b) cudaMemcpyAsync(coef + kBoltMaxR, cb, sizeof(double) * R, cudaMemcpyHostToDevice, st);
    k_bolt_axpby<<<(unsigned)((rows() * R + 255) / 256), 256, 0, st>>>((int64_t)rows(), R, coef, a, coef + kBoltMaxR, b, y);
    cudaStreamSynchronize(st);   // ca / cb are host temporaries
    note();
  }
  // x = H^-1 y by conjugate gradients (solve, BoltLMM.cpp:745-859); r, p, ap: scratch vectors of the same shape
  void solve(const double* y, double delta, double* x, double* r, double* p, double* ap, int R) {
    std::vector<double> one(R, 1.0), c1(R), c2(R), rsold(R), rsnew(R), pap(R), alpha(R), ratio(R);
    for (int k = 0; k < R; ++k) c1[k] = 1.0 / delta;
    axpby(c1.data(), y, nullptr, nullptr, x, R);            // x = y / delta
    Hx(delta, x, ap, R);
    for (int k = 0; k < R; ++k) c2[k] = -1.0;
    axpby(one.data(), y, c2.data(), ap, r, R);              // r = y - H x
    axpby(one.data(), r, nullptr, nullptr, p, R);           // p = r
    pdot(r, r, R, rsold.data());
    const double tol = 5e-4;
    const int maxIter = (int)std::min<int64_t>(N, 250);
    for (int it = 0; it < maxIter && err == cudaSuccess; ++it) {
      ++cg_total;
      Hx(delta, p, ap, R);
      pdot(p, ap, R, pap.data());
      for (int k = 0; k < R; ++k) {
        alpha[k] = rsold[k] / pap[k];
        if (!std::isfinite(alpha[k])) alpha[k] = 0.0;
      }
      axpby(one.data(), x, alpha.data(), p, x, R);          // x += p alpha
      for (int k = 0; k < R; ++k) c2[k] = -alpha[k];
      axpby(one.data(), r, c2.data(), ap, r, R);            // r -= ap alpha
      pdot(r, r, R, rsnew.data());
      bool all_small = true;
      double maxdiff = 0.0;
      for (int k = 0; k < R; ++k) {
        all_small = all_small && (rsnew[k] < tol);
        maxdiff = std::max(maxdiff, fabs(rsnew[k] - rsold[k]));
      }
      if (all_small) break;          // "reach tolerence"
      if (maxdiff < tol) break;      // "no improvement"
      for (int k = 0; k < R; ++k) {
        ratio[k] = rsnew[k] / rsold[k];
        if (rsnew[k] < tol || !std::isfinite(ratio[k])) ratio[k] = 0.0;
      }
      axpby(one.data(), r, ratio.data(), p, p, R);          // p = r + p ratio
      rsold = rsnew;
    }
  }
};
}  // namespace
extern "C" {

int rvt_bolt_fit_null(rvt_ctx* ctx, const uint8_t* bed, int64_t M, int64_t stride, int64_t N, const double* y, const double* covar,
                      int C, int mc_trials, rvt_bolt_null* out, double* h_inv_y, double* Zout) {
  return rvt_bolt_fit_null_sharded(ctx, bed, M, stride, N, y, covar, C, mc_trials, M, 0, nullptr, nullptr, out, h_inv_y, Zout);
}

int rvt_bolt_fit_null_sharded(rvt_ctx* ctx, const uint8_t* bed, int64_t M, int64_t stride, int64_t N, const double* y, const double* covar,
                              int C, int mc_trials, int64_t M_total, int64_t m_offset, rvt_allreduce_fn allreduce, void* user,
                              rvt_bolt_null* out, double* h_inv_y, double* Zout) {
  if (!ctx || !bed || !y || !covar || !out) return RVT_E_BADARG;
  if (N < 2 || M < 1 || M > 0x7FFFFFF0ll || C < 1 || C > kMaxC) CTX_FAIL(RVT_E_BADARG, "bolt: N, M or C out of range");
  if (m_offset < 0 || m_offset + M > M_total || M_total > 0x7FFFFFF0ll) CTX_FAIL(RVT_E_BADARG, "bolt: shard [%lld, %lld) outside 0..%lld", (long long)m_offset, (long long)(m_offset + M), (long long)M_total);
  if (M != M_total && !allreduce) CTX_FAIL(RVT_E_BADARG, "bolt: a shard of the panel needs the sum-over-ranks callback");
  if (stride < (N + 3) / 4) CTX_FAIL(RVT_E_BADARG, "bolt: stride (%lld) < ceil(N/4)", (long long)stride);
  RVT_CUDA_OK(cudaSetDevice(ctx->device));
  memset(out, 0, sizeof(*out));
  // orthonormal covariate basis (BoltPlinkLoader::extractCovariateBasis keeps the left singular vectors above
  // 1e-8 of the largest singular value; a modified Gram-Schmidt basis spans the same space, and only Z Z' matters)
  std::vector<double> Zh;   // column-major N x Ck
  int Ck = 0;
  for (int c = 0; c < C; ++c) {
    std::vector<double> v(covar + (size_t)c * N, covar + (size_t)(c + 1) * N);
    double n0 = 0.0;
    for (int64_t i = 0; i < N; ++i) n0 += v[i] * v[i];
    for (int pass = 0; pass < 2; ++pass)
      for (int k = 0; k < Ck; ++k) {
        double d = 0.0;
        const double* zk = Zh.data() + (size_t)k * N;
        for (int64_t i = 0; i < N; ++i) d += zk[i] * v[i];
        for (int64_t i = 0; i < N; ++i) v[i] -= d * zk[i];
      }
    double n1 = 0.0;
    for (int64_t i = 0; i < N; ++i) n1 += v[i] * v[i];
    if (!(n1 > 1e-16 * n0) || n0 == 0.0) continue;
    const double inv = 1.0 / sqrt(n1);
    for (int64_t i = 0; i < N; ++i) v[i] *= inv;
    Zh.insert(Zh.end(), v.begin(), v.end());
    ++Ck;
  }
  if (Ck < 1) CTX_FAIL(RVT_E_NUMERIC, "bolt: the covariate matrix has no usable column");
  BoltDev B;
  B.N = N; B.M = (int)M; B.C = Ck; B.stride = stride; B.st = ctx->stream;
  B.M_total = M_total; B.m_off = m_offset; B.ar = allreduce; B.ar_user = user;
  const int mc = mc_trials > 0 ? std::min(mc_trials, 15) : std::max(std::min((int)(4e9 / (double)N / (double)N), 15), 3);   // BoltLMM.cpp:465
  const int R1 = mc + 1;
  const int nSnp = (int)std::min<int64_t>(30, M_total), Rmax = std::max(std::max(R1, nSnp), Ck);
  // sample splits of the X'v product: enough CTAs to fill the device, each a multiple of the staged chunk
  B.gen = ctx->bolt_kernels;
  const int64_t xtv_block = B.gen >= 2 ? kBoltXtv2Block : kBoltXtvBlock;
  const int64_t nblk = (M + xtv_block - 1) / xtv_block;
  int splits = (int)std::max<int64_t>(1, std::min<int64_t>((4 * (int64_t)ctx->sm_count + nblk - 1) / nblk, (N + kBoltChunk - 1) / kBoltChunk));
  B.split_len = (((N + splits - 1) / splits) + kBoltChunk - 1) / kBoltChunk * kBoltChunk;
  B.splits = (int)((N + B.split_len - 1) / B.split_len);
  // the panel: copied when it lives on the host, used in place when the caller already holds it in device memory
  cudaPointerAttributes pattr;
  const bool bed_on_device = cudaPointerGetAttributes(&pattr, bed) == cudaSuccess && pattr.type == cudaMemoryTypeDevice;
  cudaGetLastError();
  const int64_t pitch = bed_on_device ? stride : ((stride + 15) / 16) * 16;   // the engine's own copy: 16-byte row pitch
  B.bed = bed_on_device ? const_cast<uint8_t*>(bed) : B.alloc<uint8_t>((size_t)M * pitch);
  B.stride = pitch;
  if (B.gen >= 3 && (pitch % 4 != 0 || ((uintptr_t)B.bed & 3) != 0)) B.gen = 2;   // cp.async needs 4-byte aligned row segments
  B.Z = B.alloc<double>((size_t)N * Ck);
  B.tab = B.alloc<double>((size_t)M * 4);
  B.zg = B.alloc<double>((size_t)M * Ck);
  B.gnorm2 = B.alloc<double>((size_t)M);
  B.part = B.alloc<double>((size_t)B.splits * M * Rmax);
  B.Xy = B.alloc<double>((size_t)M * Rmax);
  B.dotp = B.alloc<double>((size_t)kBoltDotCtas * Rmax);
  B.coef = B.alloc<double>(3 * kBoltMaxR);
  double *vy = B.vec(Rmax), *vx = B.vec(Rmax), *vr = B.vec(Rmax), *vp = B.vec(Rmax), *vap = B.vec(Rmax), *vxb = B.vec(mc), *ve = B.vec(mc);
  if (B.err != cudaSuccess) CTX_FAIL(RVT_E_CUDA, "bolt: cudaMalloc: %s", cudaGetErrorString(B.err));
  cudaStream_t st = B.st;
  if (!bed_on_device) {
    if (pitch != (N + 3) / 4) RVT_CUDA_OK(cudaMemsetAsync(B.bed, 0, (size_t)M * pitch, st));   // the bytes between a row's end and its pitch are read as words
    RVT_CUDA_OK(cudaMemcpy2DAsync(B.bed, (size_t)pitch, bed, (size_t)stride, (size_t)((N + 3) / 4), (size_t)M, cudaMemcpyHostToDevice, st));
  }
  {   // Z row-major [N][Ck] on the device
    std::vector<double> zr((size_t)N * Ck);
    for (int c = 0; c < Ck; ++c)
      for (int64_t i = 0; i < N; ++i) zr[(size_t)i * Ck + c] = Zh[(size_t)c * N + i];
    RVT_CUDA_OK(cudaMemcpy(B.Z, zr.data(), sizeof(double) * zr.size(), cudaMemcpyHostToDevice));
  }
  if (B.gen >= 3) {
    // counts -> tables by popcount; Z'X as one pass of the X'v product with v = Z (row-major [N][Ck] is the layout of a vector)
    k_bolt_count<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(B.bed, B.stride, N, (int)M, B.tab, B.gnorm2);
    B.launch_xtv_(B.Z, Ck);
    k_bolt_xtv_finish<<<(unsigned)(((int64_t)M * Ck + 255) / 256), 256, 0, st>>>((int)M, Ck, 0, B.splits, B.part, nullptr, nullptr, 1.0, B.Xy);
    k_bolt_gnorm<<<(unsigned)((M + 255) / 256), 256, 0, st>>>((int)M, Ck, B.Xy, B.zg, B.gnorm2);
  } else {
    k_bolt_snp<<<(unsigned)M, 256, 0, st>>>(B.bed, B.stride, N, Ck, B.Z, B.tab, B.zg, B.gnorm2);
  }
  B.note();
  // phenotype: centred (quantitative mode), bottom rows Z'y   (preparePhenotype, BoltPlinkLoader.cpp:143-163)
  BoltRandom rng(12345);
  const size_t rows = (size_t)(N + Ck);
  std::vector<double> hy(rows, 0.0);
  {
    double mean = 0.0;
    for (int64_t i = 0; i < N; ++i) mean += y[i];
    mean /= (double)N;
    if (ctx->bolt_binary) mean = 0.0;   // BoltLMM::enableBinaryMode: "no need to center phenotype for binary trait" (BoltPlinkLoader.cpp:155-158)
    for (int64_t i = 0; i < N; ++i) hy[i] = y[i] - mean;
    for (int c = 0; c < Ck; ++c) {
      double d = 0.0;
      for (int64_t i = 0; i < N; ++i) d += Zh[(size_t)c * N + i] * hy[i];
      hy[N + c] = d;
    }
  }
  // WorkingData::init (BoltLMM.cpp:88-123): beta_rand ~ N(0, 1/M) row by row, x_beta = [X ; Z'X] beta_rand,
  // e_rand ~ N(0, 1) row by row with its covariate rows
  {
    std::vector<double> hb((size_t)M * mc), he(rows * mc, 0.0);
    const double sq = 1.0 / sqrt((double)M_total);
    for (int64_t i = 0; i < M_total; ++i)           // every rank draws the whole stream and keeps its own SNP rows
      for (int j = 0; j < mc; ++j) {
        const double v = rng.normal() * sq;
        if (i >= m_offset && i < m_offset + M) hb[(size_t)(i - m_offset) * mc + j] = v;
      }
    for (int64_t i = 0; i < N; ++i)
      for (int j = 0; j < mc; ++j) he[(size_t)i * mc + j] = rng.normal();
    RVT_CUDA_OK(cudaMemcpy(B.Xy, hb.data(), sizeof(double) * hb.size(), cudaMemcpyHostToDevice));
    B.XW(B.Xy, mc, 1.0, 0.0, nullptr, vxb);
    RVT_CUDA_OK(cudaMemcpy(ve, he.data(), sizeof(double) * he.size(), cudaMemcpyHostToDevice));
    B.project(ve, mc);
  }
  std::vector<double> hxb(rows * mc), he(rows * mc), hY(rows * R1);
  RVT_CUDA_OK(cudaMemcpyAsync(hxb.data(), vxb, sizeof(double) * hxb.size(), cudaMemcpyDeviceToHost, st));
  RVT_CUDA_OK(cudaMemcpyAsync(he.data(), ve, sizeof(double) * he.size(), cudaMemcpyDeviceToHost, st));
  RVT_CUDA_OK(cudaStreamSynchronize(st));
  int evals = 0;
  double last_hh = 0.0;   // |H^-1 y|^2_proj of the data column at the last evaluation
  auto evalREML = [&](double logDelta) -> double {   // BoltLMM.cpp:669-724
    const double delta = exp(logDelta), sd = sqrt(delta);
    for (size_t i = 0; i < rows; ++i) {
      hY[i * R1] = hy[i];
      for (int j = 0; j < mc; ++j) hY[i * R1 + 1 + j] = hxb[i * mc + j] + sd * he[i * mc + j];   // computeY :1046-1060
    }
    cudaMemcpy(vy, hY.data(), sizeof(double) * hY.size(), cudaMemcpyHostToDevice);
    B.solve(vy, delta, vx, vr, vp, vap, R1);
    B.XtV(vx, R1, 1.0 / (double)M_total);            // beta_hat = [X ; Z'X]_minus' H^-1 y / M   (:734-742)
    ++evals;
    double bn[kBoltMaxR], en[kBoltMaxR];
    B.colnorm2(B.Xy, M, R1, bn);                     // |beta_hat|^2 per right-hand side
    B.pdot(vx, vx, R1, en);                          // e_hat = delta H^-1 y: |e_hat|^2_proj = delta^2 |H^-1 y|^2_proj
    for (int r = 0; r < R1; ++r) en[r] *= delta * delta;
    last_hh = en[0] / (delta * delta);
    double rb = 0.0, re = 0.0;
    for (int r = 1; r < R1; ++r) {
      rb += bn[r];
      re += en[r];
    }
    return log((bn[0] / en[0]) / (rb / re));
  };
  // EstimateHeritabilityBolt (BoltLMM.cpp:575-668): secant iteration on log(delta)
  double h2[7] = {0}, ld[7] = {0}, f[7] = {0};
  int i = 0;
  h2[0] = 0.25;
  ld[0] = log((1.0 - h2[0]) / h2[0]);
  f[0] = evalREML(ld[0]);
  i = 1;
  h2[1] = (f[0] < 0) ? 0.25 / 2 : std::min(0.25 * 2, 0.5 * 0.25 + 0.5);
  ld[1] = log((1.0 - h2[1]) / h2[1]);
  f[1] = evalREML(ld[1]);
  for (i = 2; i < 7; ++i) {
    ld[i] = (ld[i - 2] * f[i - 1] - ld[i - 1] * f[i - 2]) / (f[i - 1] - f[i - 2]);
    if (!std::isfinite(ld[i])) {
      --i;
      break;
    }
    if (ld[i] > 5) ld[i] = 5;
    if (ld[i] < -10) ld[i] = -10;
    h2[i] = 1.0 / (1.0 + exp(ld[i]));
    if (fabs(ld[i] - ld[i - 1]) < 0.01) break;
    f[i] = evalREML(ld[i]);
  }
  if (i == 7) --i;
  if (B.err != cudaSuccess) CTX_FAIL(RVT_E_CUDA, "bolt: %s", cudaGetErrorString(B.err));
  const double delta = exp(ld[i]);
  // projDot(y, H^-1 y) of the LAST evaluation (the reference does not re-solve at the final delta): vy / vx still hold them
  double yhv[kBoltMaxR];
  B.pdot(vy, vx, R1, yhv);
  const double sigma2_g = yhv[0] / (double)(N - Ck);
  if (!(sigma2_g > 0.0)) CTX_FAIL(RVT_E_NUMERIC, "bolt: sigma2_g = %g is not positive", sigma2_g);
  const double sigma2_e = delta * sigma2_g;
  const double hn2 = last_hh / (sigma2_g * sigma2_g);       // projNorm2(H_inv_y_), H_inv_y_ = H^-1 y / sigma2_g
  const size_t rows_h = rows;
  std::vector<double> hh(rows_h);
  double* vh = B.vec(1);
  if (!vh) CTX_FAIL(RVT_E_CUDA, "bolt: cudaMalloc");
  k_bolt_bcast<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>((int64_t)rows, R1, 0, 1, 1.0 / sigma2_g, vx, vh);
  RVT_CUDA_OK(cudaMemcpyAsync(hh.data(), vh, sizeof(double) * rows, cudaMemcpyDeviceToHost, st));
  RVT_CUDA_OK(cudaStreamSynchronize(st));
  // EstimateInfStatCalibration (BoltLMM.cpp:1141-1214)
  std::vector<int> idx(nSnp);
  for (int k = 0; k < nSnp; ++k) {
    const int64_t g = (int64_t)(size_t)(rng.next() * (double)M_total);
    idx[k] = (g >= m_offset && g < m_offset + M) ? (int)(g - m_offset) : -1;   // columns of other ranks' SNPs arrive with the sum
  }
  int* d_idx = B.alloc<int>(nSnp);
  if (!d_idx) CTX_FAIL(RVT_E_CUDA, "bolt: cudaMalloc");
  RVT_CUDA_OK(cudaMemcpy(d_idx, idx.data(), sizeof(int) * nSnp, cudaMemcpyHostToDevice));
  k_bolt_columns<<<(unsigned)(((int64_t)N * nSnp + 255) / 256), 256, 0, st>>>(B.bed, B.stride, N, d_idx, nSnp, B.tab, vy);
  B.allreduce(vy, (int64_t)N * nSnp);
  B.project(vy, nSnp);
  const int hx_before_cal = B.n_hx;
  B.solve(vy, delta, vx, vr, vp, vap, nSnp);         // V^-1 x = H^-1 x / sigma2_g
  std::vector<double> xVx(nSnp), xx(nSnp), xVy(nSnp);
  B.pdot(vy, vx, nSnp, xVx.data());
  B.pdot(vy, vy, nSnp, xx.data());
  k_bolt_bcast<<<(unsigned)((rows * nSnp + 255) / 256), 256, 0, st>>>((int64_t)rows, 1, 0, nSnp, 1.0, vh, vr);   // H_inv_y_ in every column
  B.pdot(vy, vr, nSnp, xVy.data());
  if (B.err != cudaSuccess) CTX_FAIL(RVT_E_CUDA, "bolt: %s", cudaGetErrorString(B.err));
  double r0 = 0.0, r1 = 0.0, sxVx = 0.0, sxx = 0.0;
  for (int k = 0; k < nSnp; ++k) {   // 30 scalars: the calibration ratio (BoltLMM.cpp:1166-1181)
    xVx[k] /= sigma2_g;
    const double d = xVy[k];
    const double prosp = d * d / xVx[k], retro = (double)N * d * d / (xx[k] * hn2);
    if (prosp < 5.0) {
      r0 += retro;
      r1 += prosp;
    }
    sxVx += xVx[k];
    sxx += xx[k];
  }
  out->delta = delta;
  out->sigma2_g = sigma2_g;
  out->sigma2_e = sigma2_e;
  out->h2 = h2[i];
  out->h_inv_y_norm2 = hn2;
  out->inf_stat_calibration = (r1 != 0.0) ? r0 / r1 : 1.0;
  out->xvx_xx_ratio = std::isfinite(sxVx / sxx) ? sxVx / sxx : 1.0;
  out->mc_trials = mc;
  out->reml_evals = evals;
  out->cg_iterations = B.cg_total;
  out->n_covariates_kept = Ck;
  cudaStreamSynchronize(st);
  B.collect_times();
  out->ms_xtv = B.ms_xtv;
  out->ms_xw = B.ms_xw;
  out->h_products = B.n_hx;
  out->h_products_calibration = B.n_hx - hx_before_cal;
  out->allreduce_calls = B.ar_calls;
  if (B.ar_rc) CTX_FAIL(RVT_E_CUDA, "bolt: the sum-over-ranks callback returned %d", B.ar_rc);
  for (int k = 0; k < 7; ++k) {
    out->log_delta[k] = (k <= i) ? ld[k] : 0.0;
    out->f[k] = (k < evals) ? f[k] : 0.0;
  }
  if (h_inv_y) memcpy(h_inv_y, hh.data(), sizeof(double) * rows);
  if (Zout) memcpy(Zout, Zh.data(), sizeof(double) * (size_t)N * Ck);
  return RVT_OK;
}

int rvt_synth_load(rvt_ctx* ctx, int n_genes, int M, const uint64_t* keys, const uint32_t* t0, const uint32_t* t1) {
  if (!ctx || !keys || !t0 || !t1 || n_genes < 1) return RVT_E_BADARG;
  int rc = push_check(ctx, M);
  if (rc) return rc;
  const int64_t N = ctx->N, npad = (N + 127) & ~(int64_t)127;
  const int64_t rows = (int64_t)n_genes * M;
  const int64_t gene_bytes = tiled_bytes(N, M);
  if (ctx->d_loaded) {
    cudaFree(ctx->d_loaded);
    ctx->d_loaded = nullptr;
  }
  RVT_CUDA_OK(cudaMalloc((void**)&ctx->d_loaded, (size_t)n_genes * gene_bytes));
  unsigned long long* dk = nullptr;
  uint32_t *d0 = nullptr, *d1 = nullptr;
  RowCounts* dc = nullptr;
  uint8_t* dfl = nullptr;
  double* daf = nullptr;
  RVT_CUDA_OK(cudaMalloc((void**)&dk, rows * 8));
  RVT_CUDA_OK(cudaMalloc((void**)&d0, rows * 4));
  RVT_CUDA_OK(cudaMalloc((void**)&d1, rows * 4));
  RVT_CUDA_OK(cudaMalloc((void**)&dc, rows * sizeof(RowCounts)));
  RVT_CUDA_OK(cudaMalloc((void**)&dfl, rows));
  RVT_CUDA_OK(cudaMalloc((void**)&daf, rows * 8));
  cudaStream_t st = ctx->stream;
  RVT_CUDA_OK(cudaMemcpyAsync(dk, keys, rows * 8, cudaMemcpyHostToDevice, st));
  RVT_CUDA_OK(cudaMemcpyAsync(d0, t0, rows * 4, cudaMemcpyHostToDevice, st));
  RVT_CUDA_OK(cudaMemcpyAsync(d1, t1, rows * 4, cudaMemcpyHostToDevice, st));
  RVT_CUDA_OK(cudaMemsetAsync(dc, 0, rows * sizeof(RowCounts), st));
  const unsigned per_row = (unsigned)((npad / 16 + 255) / 256);
  const int64_t genes_per_launch = std::max<int64_t>(1, 32768 / M);
  for (int64_t g0 = 0; g0 < n_genes; g0 += genes_per_launch) {
    const int64_t ng = std::min<int64_t>(genes_per_launch, n_genes - g0);
    const int64_t r0 = g0 * M;
    k_synth_rows<<<dim3(per_row, (unsigned)(ng * M)), 256, 0, st>>>(ctx->d_loaded + (size_t)g0 * gene_bytes, M, gene_bytes, N, dk + r0,
                                                                    d0 + r0, d1 + r0, dc + r0);
  }
  k_flags_from_counts<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(rows, N, dc, dfl, daf);
  RVT_CUDA_OK(cudaGetLastError());
  ctx->loaded_flags.resize(rows);
  ctx->loaded_af.resize(rows);
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->loaded_flags.data(), dfl, rows, cudaMemcpyDeviceToHost, st));
  RVT_CUDA_OK(cudaMemcpyAsync(ctx->loaded_af.data(), daf, rows * 8, cudaMemcpyDeviceToHost, st));
  RVT_CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(dk); cudaFree(d0); cudaFree(d1); cudaFree(dc); cudaFree(dfl); cudaFree(daf);
  ctx->loaded_rows = rows;
  ctx->loaded_ld = gene_bytes;   // bytes per gene block
  ctx->loaded_genes = n_genes;
  ctx->loaded_M = M;
  rc = tc_bind_segment(&ctx->tc, kSegLoaded, ctx->d_loaded, (int64_t)n_genes * gene_bytes, ctx->err, sizeof(ctx->err));
  if (rc) return rc;
  return RVT_OK;
}

int rvt_loaded_genes(const rvt_ctx* ctx) { return ctx ? ctx->loaded_genes : 0; }

int rvt_push_loaded(rvt_ctx* ctx) {
  if (!ctx) return RVT_E_BADARG;
  if (!ctx->d_loaded) CTX_FAIL(RVT_E_STATE, "no cohort loaded (rvt_synth_load)");
  if (!ctx->genes.empty()) CTX_FAIL(RVT_E_STATE, "flush pending genes first");
  const int M = ctx->loaded_M;
  int rc = push_check(ctx, M);
  if (rc) return rc;
  if ((rc = ensure_var(ctx, (size_t)ctx->loaded_rows))) return rc;
  ctx->genes.reserve(ctx->loaded_genes);
  for (int g = 0; g < ctx->loaded_genes; ++g) {
    const int64_t row0 = (int64_t)g * M;
    // AF: the loader's counts, i.e. what GenotypeCounter::getAF hands to the fitters
    push_common(ctx, ctx->d_loaded + (size_t)g * ctx->loaded_ld, M, 0, ctx->loaded_af.data() + row0,
                ctx->loaded_flags.data() + row0, false, kSegLoaded, ((int64_t)g * ctx->loaded_ld) / 128, true);
  }
  return RVT_OK;
}

int rvt_run_loaded(rvt_ctx* ctx, rvt_gene_result* out, int cap, int* n_out, int results_on_device) {
  int rc = rvt_push_loaded(ctx);
  if (rc) return rc;
  return flush_impl(ctx, out, cap, n_out, results_on_device != 0);
}

int rvt_loaded_read(rvt_ctx* ctx, int64_t row0, int rows, int8_t* out) {
  if (!ctx || !out) return RVT_E_BADARG;
  if (!ctx->d_loaded || row0 < 0 || row0 + rows > ctx->loaded_rows) CTX_FAIL(RVT_E_BADARG, "rows out of range");
  int8_t* tmp = nullptr;
  RVT_CUDA_OK(cudaMalloc((void**)&tmp, (size_t)rows * ctx->N));
  dim3 grid((unsigned)((ctx->N / 16 + 256) / 256), (unsigned)rows);
  k_untile<<<grid, 256, 0, ctx->stream>>>(ctx->d_loaded, ctx->loaded_M, ctx->loaded_ld, row0, ctx->N, tmp);
  RVT_CUDA_OK(cudaGetLastError());
  RVT_CUDA_OK(cudaMemcpyAsync(out, tmp, (size_t)rows * ctx->N, cudaMemcpyDeviceToHost, ctx->stream));
  RVT_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  cudaFree(tmp);
  return RVT_OK;
}

int rvt_debug_partials(rvt_ctx* ctx, void* out, int64_t cap_bytes, int64_t* bytes) {
  if (!ctx || !bytes) return RVT_E_BADARG;
  const int64_t need = (int64_t)ctx->last_parts * (int64_t)sizeof(SweepPartial);
  *bytes = need;
  if (!out) return RVT_OK;
  if (cap_bytes < need) CTX_FAIL(RVT_E_BADARG, "partials need %lld bytes", (long long)need);
  RVT_CUDA_OK(cudaMemcpy(out, ctx->d_parts, (size_t)need, cudaMemcpyDeviceToHost));
  return RVT_OK;
}

int rvt_debug_phases(rvt_ctx* ctx, long long* out, int cap_genes) {
  if (!ctx || !out) return RVT_E_BADARG;
  if (!ctx->d_dbg || cap_genes < ctx->last_n) CTX_FAIL(RVT_E_STATE, "phase counters not enabled (option debug_phases) or buffer too small");
  RVT_CUDA_OK(cudaMemcpy(out, ctx->d_dbg, sizeof(long long) * kFinPhases * ctx->last_n, cudaMemcpyDeviceToHost));
  return RVT_OK;
}

int rvt_last_timing(const rvt_ctx* ctx, double out[4]) {
  if (!ctx || !out) return RVT_E_BADARG;
  out[0] = ctx->t_sweep;
  out[1] = ctx->t_fin;
  out[2] = ctx->t_total;
  out[3] = ctx->n_launch;
  return RVT_OK;
}

}  // extern "C"
