/*
 * rvtests_b200.h -- C ABI of the B200-native rvtests gene engine (librvtests_b200.so).
 *
 * The reference (zhanxw/rvtests @ 8defd6f) has no FFI on this path: its plugin surface is the C++
 * abstract class ModelFitter (src/ModelFitter.h:17-75) whose fit(DataConsolidator*) pulls
 * Matrix/Vector views (base/MathMatrix.h:33-111) and calls the pimpl statistics classes
 *   Skat::Fit                      regression/Skat.h:26-38      / Skat.cpp:29-105
 *   SkatO::Fit                     regression/SkatO.h:28-39     / SkatO.cpp:101-281
 *   LinearRegression::FitLinearModel            regression/LinearRegression.cpp:20-69
 *   LinearRegressionScoreTest::TestCovariate    regression/LinearRegressionScoreTest.cpp:173-263
 *   cmcCollapse / zegginiCollapse               src/Model.cpp:73-89, 115-130
 *   getFlippedToMinorPolymorphicGenotype        src/DataConsolidator.h:128-132, .cpp:46-142
 * The entry points below are what a C++ adapter with the ModelFitter signature binds instead of
 * those classes (rvtests_b200/host/rvt_fitters.h holds such adapters; INTEGRATION.md shows the registration a
 * maintainer adds to src/ModelManager.cpp).  Plain pointers and sizes only; no C++ / torch types.
 *
 * Threading: one host thread per context (the reference's model loop is single-threaded,
 * src/Main.cpp:1249-1253).  All functions return RVT_OK (0) or a negative RVT_E_* code;
 * rvt_last_error() gives the message.  There is NO CPU fallback: without a CUDA device every
 * compute entry point fails with RVT_E_CUDA.
 */
#ifndef RVTESTS_B200_H_
#define RVTESTS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RVT_OK 0
#define RVT_E_BADARG (-1)
#define RVT_E_CUDA (-2)
#define RVT_E_STATE (-3)
#define RVT_E_UNSUPPORTED (-4)
#define RVT_E_NUMERIC (-5)

/* per-gene status (rvt_gene_result.status) */
#define RVT_GENE_OK 0
#define RVT_GENE_NA 2           /* no polymorphic variant: fit() == -1, output "NA" (src/Model.h:2637-2640) */
#define RVT_GENE_BADFLAGS 4     /* caller-supplied flip/skip flags contradict the data */
#define RVT_GENE_BADVALUE 5     /* a genotype outside {0,1,2} reached the hard-call path */
#define RVT_GENE_UNSUPPORTED 7  /* this gene is outside what the engine computes (more than 64 variants AND missing calls when the
                                   tensor-core engine is unavailable); the other genes of the flush are unaffected */
#define RVT_GENE_TIMEOUT 6      /* device watchdog: the SKAT-O quadrature of this gene exceeded its cycle budget
                                   (option "watchdog_ms", default 4000); skato_ok = 0, the other columns are valid */

/* engine selection for rvt_set_option("engine") */
#define RVT_ENGINE_AUTO 0
#define RVT_ENGINE_SIMT 1       /* dp4a CUDA-core sweep */
#define RVT_ENGINE_TC 2         /* tcgen05 (kind::i8) tensor-core sweep, TMA-fed */

typedef struct rvt_ctx rvt_ctx;

/* One record per gene, in push order.  Field <- reference source of the value:
 *   Q, p_skat            SkatTest::writeOutput "%g\t%g" (src/Model.h:2743): Skat::GetQ, pValue
 *   p_davies/fault/p_liu MixtureChiSquare::getPvalue / getLiuPvalue (regression/MixtureChiSquare.cpp)
 *   cmc_nonref, cmc_p    CMCTest::writeOutput NonRefSite, Pvalue (src/Model.h:865-900)
 *   zeg_p                ZegginiTest::writeOutput Pvalue (src/Model.h:1223-1234)
 *   *_U, *_V, *_stat     LinearRegressionScoreTest U, V, stat (LinearRegressionScoreTest.cpp:213-258)
 *   skato_*              SkatOTest::writeOutput Q, rho, Pvalue (src/Model.h:2866-2875)
 */
typedef struct rvt_gene_result {
  double Q;
  double p_skat;
  double p_davies;
  double p_liu;
  int32_t davies_fault;
  int32_t n_lambda;
  int32_t m_poly;
  int32_t status;
  int32_t cmc_nonref;
  int32_t cmc_ok;
  double cmc_U, cmc_V, cmc_stat, cmc_p;
  int32_t zeg_ok;
  int32_t skato_ok;
  double zeg_U, zeg_V, zeg_stat, zeg_p;
  double skato_Q, skato_rho, skato_p;
  double lambda_max;   /* largest kept eigenvalue (diagnostic) */
} rvt_gene_result;

/* Permutation test of `--kernel skat[nPerm=..,alpha=..]` (the reference's default): one record per gene of the last
 * flush, the columns Permutation::writeOutput prints (src/Permutation.h:118-139).  Enabled by
 * rvt_set_option("perm", nPerm) [+ "perm_alpha"]; the shuffles replay glibc's default rand() stream exactly as the
 * reference's serial gene loop consumes it (src/LinearAlgebra.h:8-21, no srand anywhere), starting at
 * "perm_stream_pos" draws (0 in a fresh process) and advancing by ActualPerm * (N-1) per gene.  Quantitative and binary
 * traits alike (src/Model.h:2673-2717), genes with missing calls (2-bit pushes, mean-imputed) included.  NOT covered: a gene
 * with real dosages pushed as doubles, a gene with missing calls of a binary-trait run, of 63-64 or of more than 64 variants, any gene
 * of more than 64 variants of a binary-trait run (done = 0, NA columns);
 * the reference would have shuffled for it, so from such a gene on the stream position -- hence NumGreater / NumEqual of the
 * LATER genes -- no longer replays the reference's: those records carry stream_ok = 0.  rvt_perm_result.stream_pos tells where
 * each gene started. */
typedef struct rvt_perm_result {
  int32_t num_perm;      /* NumPerm */
  int32_t actual_perm;   /* ActualPerm */
  int32_t num_greater;   /* NumGreater */
  int32_t num_equal;     /* NumEqual */
  double stat;           /* Stat: the observed Q */
  double p_perm;         /* PermPvalue */
  int64_t stream_pos;    /* rand() values consumed before this gene's first shuffle */
  int32_t done;          /* 1: permutations ran; 0: gene NA (fit() == -1) or on a path the test does not cover */
  int32_t stream_ok;     /* 1: stream_pos is where the reference's serial loop stands at this gene; 0: an EARLIER gene of this
                          * context (since "perm_stream_pos" / "perm_seed" was last set) was testable in the reference but not
                          * covered here (done = 0 with status OK), so this gene's shuffles -- NumGreater, NumEqual, PermPvalue --
                          * are valid permutation statistics but not the reference's own draws */
} rvt_perm_result;

/* One record per variant from rvt_lmm_flush: FastLMM::TestCovariate, score branch (regression/FastLMM.cpp:215-249) */
typedef struct rvt_lmm_result {
  double af;       /* allele frequency of the pushed hard calls */
  double U, V;     /* Ustat, Vstat (FastLMM::GetUStat / GetVStat) */
  double stat;     /* U^2 / V, or 0 when V <= 0 */
  double pvalue;   /* gsl_cdf_chisq_Q(stat, 1), or 1 when V <= 0 */
  int32_t ok;      /* 0: the block held values outside {0,1,2} */
  int32_t pad;
} rvt_lmm_result;

/* BoltLMM null-model fit (regression/BoltLMM.cpp:169-299): what FitNullModel leaves behind for TestCovariate */
typedef struct rvt_bolt_null {
  double delta;                  /* sigma2_e / sigma2_g at the end of the secant iteration */
  double sigma2_g, sigma2_e, h2;
  double h_inv_y_norm2;          /* projNorm2(H_inv_y_) */
  double inf_stat_calibration;   /* infStatCalibration_ */
  double xvx_xx_ratio;           /* xVx_xx_ratio_ (scales BoltLMM::GetCovXX) */
  double log_delta[7], f[7];     /* the secant path: log(delta) tried and the MC-REML function there */
  int32_t mc_trials, reml_evals, cg_iterations, n_covariates_kept;
  /* measurement: device time (CUDA events on the context stream) of the two panel products over the whole fit -- X'v
   * (BoltLMM.cpp:942-952) and X w (:956-966) --, the number of H-products (computeHx calls) and of sum-over-ranks calls */
  double ms_xtv, ms_xw;
  int32_t h_products, allreduce_calls;
  int32_t h_products_calibration;   /* of h_products, those of the calibration solve (30 right-hand sides instead of MCtrial + 1) */
  int32_t pad;
} rvt_bolt_null;

/* ---- lifetime ------------------------------------------------------------------------------- */
int rvt_ctx_create(int device, rvt_ctx** out);
void rvt_ctx_destroy(rvt_ctx* ctx);
const char* rvt_last_error(const rvt_ctx* ctx);
/* keys: "beta1","beta2" (Beta weight, src/ModelManager.cpp:169-175), "engine", "splits" (0=auto),
 * "skato" (0/1), "skato_binary" (0/1, default 1: with a binary null model SKAT-O = SkatO::Fit type "D",
 * src/Model.h:2854-2858, regression/SkatO.cpp:72-91,133-134,150-158; 0 leaves skato_ok = 0 for a binary trait), "watchdog_ms"
 * (device watchdog of the per-gene SKAT-O quadrature, default 4000; 0 = off), "stream_batch" (B > 0: every B host pushes the engine enqueues sweep + statistics for
 * them right away, so the kernels run while the next genes are still crossing PCIe; rvt_flush then only
 * waits for the tail.  0 = everything at flush.  Options apply to genes pushed after the call.),
 * "binary_stream" (default 1: with a binary null model and "stream_batch" > 0 the fp64 statistics and the tail of the tile genes
 * are enqueued -- on a stream of their own -- right behind their copies instead of at flush; 0 = everything at flush),
 * "bolt_binary" (1: rvt_bolt_fit_null in BoltLMM::enableBinaryMode -- the phenotype is not centred, BoltPlinkLoader.cpp:155-158;
 * the saddle-point calibration of that mode is compiled out upstream: useSaddlePoint = false, BoltLMM.cpp:160),
 * "bolt_kernels" (2 = second-generation panel products of rvt_bolt_fit_null, 1 = the first; same results to rounding),
 * "perm" (nPerm, 0 = analytic p-value only), "perm_alpha" (0.05), "perm_stream_pos", "perm_seed", "perm_batch". */
int rvt_set_option(rvt_ctx* ctx, const char* key, double value);
double rvt_get_info(const rvt_ctx* ctx, const char* key);
/* run on a caller-owned CUDA stream (cudaStream_t), e.g. the framework's current stream, so that
 * the caller's events and collectives order with the engine's kernels; NULL restores the
 * context's own stream */
int rvt_set_stream(rvt_ctx* ctx, void* cuda_stream);

/* ---- null model: replaces LinearRegression::FitLinearModel(cov, phenoVec) -------------------
 * X: N x C column-major doubles INCLUDING the intercept as column 0 (the matrix produced by
 * copyCovariateAndIntercept, src/ModelUtil.h:102-130); y: N.  Host pointers.
 * binary != 0: y in {0,1}; LogisticRegression::FitLogisticModel(cov, phenoVec, 100) (regression/LogisticRegression.cpp:279-339)
 *   on the device, then SKAT / CMC / Zeggini with r = y - p and the per-sample variance v = p(1-p) (src/Model.h:2673-2681,
 *   LogisticRegressionScoreTest.cpp:219-302).  Such genes take the engine's fp64 paths (any width up to 2048); SKAT-O is type "D"
 *   (option "skato_binary", default on); the permutation test runs as for a quantitative trait.  rvt_get_null_model then returns
 *   r, sigma2 = 1 and (X'VX)^-1. */
int rvt_set_null_model(rvt_ctx* ctx, int64_t N, int C, const double* X, const double* y, int binary);
/* "bring your own null": the caller supplies the score vector r (length N) and the variance scale
 * sigma2, the engine only builds (X'X)^-1 and the device images.  This is the score step of the
 * mixed models: BoltLMM::TestCovariate (regression/BoltLMM.cpp:315-338) is U = g.r with
 * r = (I - ZZ')H^-1 y and V = (|g|^2 - |Z'g|^2) * kappa, kappa = |H^-1 y|^2_proj * calibration / N --
 * i.e. rvt_meta_flush after rvt_set_null_residual(ctx, N, C, X, r, kappa).  The null fit that
 * produces H^-1 y (BoltLMM::FitNullModel) is outside this build. */
int rvt_set_null_residual(rvt_ctx* ctx, int64_t N, int C, const double* X, const double* resid, double sigma2);
/* same, X and y already in device memory */
int rvt_set_null_model_dev(rvt_ctx* ctx, int64_t N, int C, const double* dX, const double* dy);
/* resid (N doubles, host, may be NULL), sigma2, xtx_inv (C*C row-major, may be NULL) */
int rvt_get_null_model(rvt_ctx* ctx, double* resid, double* sigma2, double* xtx_inv);
/* beta (C doubles): LinearRegression::GetCovEst, printed in the ##NullModelEstimates block of MetaScore.assoc */
int rvt_get_null_beta(rvt_ctx* ctx, double* beta);

/* ---- genes ---------------------------------------------------------------------------------
 * rvt_gene_push_f64: the reference boundary -- G is dc->getGenotype(): N x M column-major doubles
 *   on the host, imputed, NOT yet flipped (the engine performs convertToMinorAlleleCount +
 *   removeMonomorphicMarker itself).  af: M allele frequencies in the caller's column order
 *   (GenotypeCounter::getAF) or NULL (then AF = column mean / 2 of the kept columns).
 *   Mean-imputed missing calls (DataConsolidator::imputeGenotypeToMean, src/DataConsolidator.cpp:217-245: hard calls plus, per
 *   column, ONE fractional value 2 p^ of its observed calls) are recognised on the device and such a gene is computed exactly like
 *   a 2-bit push with code 01 -- augmented tensor-core sweep, wide operand tiles, permutation test -- with records equal to that
 *   form bit for bit (option "f64_imputed", default 1); any other non-integer value makes it a dosage gene (fp64 paths).
 * rvt_gene_push_i8: same, hard calls as int8 [M][ld] variant-major on the host.
 * rvt_gene_push_dev_i8: block already in device memory (zero-copy; must stay valid until flush).
 *   flags: NULL (engine counts the rows itself) or M bytes 0 normal / 1 flip-to-minor / 2 skip.
 * Width: the host entry points (f64, i8, bed) take genes of 1..2048 variants -- the reference has no limit
 *   (Skat::Fit / MixtureChiSquare size themselves to the gene); a gene of more than 64 variants is cut into
 *   64-variant tiles and its Gram assembled from tile pairs (csrc/wide.cuh).  RVT_E_UNSUPPORTED beyond 2048,
 *   for rvt_gene_push_dev_i8 beyond 64.  A wide gene pushed as doubles that holds real dosages keeps its matrix on the device
 *   for dense fp64 statistics (csrc/wide.cuh: k_wide_dos_*).  Missing calls of a wide gene pushed as 2-bit rows are mean-imputed like those of any gene
 *   (src/DataConsolidator.cpp:217-245): its tiles are split into hard-call and indicator tiles and swept as 2M rows.  With a
 *   binary null model a wide gene takes fp64 statistics computed from its tiles (csrc/wide.cuh: k_wide_sparse), missing calls
 *   included; the permutation test does not cover it.
 * Each push appends one gene; results come back from rvt_flush in push order.  Host buffers are copied
 * asynchronously when they are page-locked: keep them valid and unchanged until rvt_flush returns. */
int rvt_gene_push_f64(rvt_ctx* ctx, const double* G, int M, const double* af);
int rvt_gene_push_i8(rvt_ctx* ctx, const int8_t* G, int M, int64_t ld, const double* af);
int rvt_gene_push_dev_i8(rvt_ctx* ctx, const int8_t* dG, int M, int64_t ld, const double* af,
                         const uint8_t* flags);
/* rvt_gene_push_bed: M variants as PLINK .bed SNP-major rows on the host -- the 2-bit form the
 *   reference itself keeps in RAM for large cohorts (regression/BoltPlinkLoader.cpp:164-264) and
 *   decodes in libVcf/PlinkInputFile.cpp:23-47: sample p = bits 2(p&3)..2(p&3)+1 of byte p>>2 of
 *   the variant's row, 00 -> 0, 10 -> 1, 11 -> 2, 01 -> missing; `stride` bytes between rows
 *   (>= ceil(N/4)).  A quarter of the int8 bytes cross PCIe.  Missing calls are mean-imputed
 *   exactly as DataConsolidator::imputeGenotypeToMean does (src/DataConsolidator.cpp:217-245);
 *   such a gene then takes the fp64 path. */
int rvt_gene_push_bed(rvt_ctx* ctx, const uint8_t* bed, int M, int64_t stride, const double* af);
/* permutation records of the genes of the LAST rvt_flush / rvt_run_loaded, in the same order (empty unless option
 * "perm" > 0).  out == NULL: only *n_out is set. */
int rvt_perm_results(rvt_ctx* ctx, rvt_perm_result* out, int cap, int* n_out);
/* diagnostics: with option "debug_perm_q" = 1, every permuted statistic of the last flush, in the order they ran */
int rvt_perm_debug_q(rvt_ctx* ctx, double* out, int cap, int* n_out);
/* diagnostics: the n values glibc's rand() returns from stream position `pos` after srand(seed), generated by the
 * device kernel the permutation test uses (host array out[n]) */
int rvt_debug_rand(rvt_ctx* ctx, uint32_t seed, uint64_t pos, int64_t n, int32_t* out);
/* ---- mixed-model (FastLMM) score step ---------------------------------------------------------
 * rvt_lmm_set_null: what FastLMM::Impl::FitNullModel leaves behind (regression/FastLMM.cpp:27-140), computed by the caller:
 *   U      N x N floats, column-major, column i = eigenvector i of the kinship (kinshipU.mat)
 *   lambda N eigenvalues (kinshipS; absolute values are taken as the reference does), delta = sigma2_e / sigma2_g,
 *   sigma2 = sigma2_g, uResid = U'y - U'X beta (N), ux = U'X (N x C column-major, intercept included).
 * Then push blocks of <= 64 variants (rvt_gene_push_i8 / _bed, hard calls) and call rvt_lmm_flush: one record per
 * pushed variant, in push order.  Replaces any linear null model set before. */
int rvt_lmm_set_null(rvt_ctx* ctx, int64_t N, int C, const float* U, const float* lambda, double delta, double sigma2,
                     const float* uResid, const float* ux);
int rvt_lmm_flush(rvt_ctx* ctx, rvt_lmm_result* out, int64_t cap);
/* The same flush plus the covariance band of `--meta cov` for related samples: MetaCovFamQtl (src/Model.cpp:437-498) =
 * FastLMM::TransformCentered / GetCovXX / GetCovXZ / GetCovZZ (regression/FastLMM.cpp:538-625) through
 * MetaCovTest::printCovariance (src/Model.cpp:934-1004).  pos / chrom / window_bp / band / wmax as in rvt_meta_flush:
 * band[v*(wmax+1)+d] = (x~_v' D x~_w / sigma2 - covXZ_v covZZ^-1 covXZ_w') / N for w = v + d inside v's window, NaN where
 * either variant is monomorphic or w is outside the window.  band may be a host or a device pointer. */
int rvt_lmm_meta_flush(rvt_ctx* ctx, const int32_t* pos, const int32_t* chrom, int64_t window_bp, rvt_lmm_result* out, int64_t cap,
                       double* band, int64_t cap_band, int* wmax);
/* ---- BoltLMM null model -------------------------------------------------------------------------
 * bed: the PLINK panel, M SNP-major 2-bit rows of `stride` >= ceil(N/4) bytes on the host (samples in phenotype order);
 * y: N phenotypes; covar: N x C column-major, intercept first (the .covar file of BoltPlinkLoader); mc_trials: 0 = the
 * reference's rule max(min(4e9/N^2, 15), 3).  h_inv_y (may be NULL) receives the N + n_covariates_kept values
 * [H^-1 y / sigma2_g ; Z' of it]; Zout (may be NULL) the N x n_covariates_kept orthonormal covariate basis, column-major.
 * The score step then is rvt_set_null_residual(ctx, N, C, X, h - Z Z'h, h_inv_y_norm2 * inf_stat_calibration / N) +
 * rvt_meta_flush.  Random numbers: the reference's generator and seed (libsrc/Random.cpp, 12345). */
int rvt_bolt_fit_null(rvt_ctx* ctx, const uint8_t* bed, int64_t M, int64_t stride, int64_t N, const double* y, const double* covar,
                      int C, int mc_trials, rvt_bolt_null* out, double* h_inv_y, double* Zout);
/* The same fit with the panel's SNPs sharded over ranks (BASELINE configs[4] on 8 GPUs; the reference has no
 * multi-device path, its OpenMP loops over the 64 SNPs of a batch are what this replaces, BoltPlinkLoader.cpp:305-440).
 * This rank holds rows [m_offset, m_offset + M) of the M_total SNP rows (bed, host or DEVICE pointer -- a device panel is
 * used in place); y / covar are replicated.  Every product that sums over SNPs -- the X X'v half of computeHx
 * (BoltLMM.cpp:956-966), |beta_hat|^2 (:689-692), the calibration columns -- is a local partial followed by ONE call of
 * `allreduce`: an in-place SUM over ranks of `count` doubles at DEVICE pointer `buf`, ordered on `cuda_stream` (the context's
 * stream, a cudaStream_t) like ncclAllReduce(buf, buf, count, ncclDouble, ncclSum, comm, stream); return 0 on success.  One
 * call per H-product of the conjugate gradients.  Every rank draws the same random stream and returns the same record /
 * h_inv_y.  allreduce may be NULL only when M == M_total. */
typedef int (*rvt_allreduce_fn)(void* user, double* buf, int64_t count, void* cuda_stream);
int rvt_bolt_fit_null_sharded(rvt_ctx* ctx, const uint8_t* bed, int64_t M, int64_t stride, int64_t N, const double* y, const double* covar,
                              int C, int mc_trials, int64_t M_total, int64_t m_offset, rvt_allreduce_fn allreduce, void* user,
                              rvt_bolt_null* out, double* h_inv_y, double* Zout);
/* number of genes pushed and not yet flushed */
int rvt_pending(const rvt_ctx* ctx);
/* run the sweep + per-gene statistics for every pending gene; out: host array of `cap` records */
int rvt_flush(rvt_ctx* ctx, rvt_gene_result* out, int cap, int* n_out);
/* as rvt_flush but leaves the records in device memory (d_out: device pointer, cap records);
 * asynchronous on the context stream -- used by the multi-GPU gather */
int rvt_flush_dev(rvt_ctx* ctx, rvt_gene_result* d_out, int cap, int* n_out);

/* ---- device-resident synthetic cohort (SURVEY.md 8(d) stream; bench + tests) -----------------
 * Generates n_genes x M variants x N samples of HWE genotypes directly in HBM (packed int8),
 * variant ids first_vid .. first_vid + n_genes*M - 1, with per-variant uint32 thresholds t0,t1 and
 * 64-bit keys supplied by the host (oracle/oracle.py synth_* reproduces the same stream).
 * The genes stay loaded: rvt_run_loaded() pushes all of them (zero-copy) and flushes. */
int rvt_synth_load(rvt_ctx* ctx, int n_genes, int M, const uint64_t* keys, const uint32_t* t0,
                   const uint32_t* t1);
int rvt_loaded_genes(const rvt_ctx* ctx);
/* queue every loaded gene (zero-copy) without flushing: followed by rvt_flush, or -- the genes then being 64-variant tiles of
 * consecutive variants -- by rvt_meta_flush (bench.py --workload meta) */
int rvt_push_loaded(rvt_ctx* ctx);
int rvt_run_loaded(rvt_ctx* ctx, rvt_gene_result* out, int cap, int* n_out, int results_on_device);
/* copy rows [row0,row0+rows) x samples [0,N) of the loaded arena back to the host (tests) */
int rvt_loaded_read(rvt_ctx* ctx, int64_t row0, int rows, int8_t* out /*rows x N*/);

/* ---- single-variant meta-analysis statistics: --meta score,cov --------------------------------
 * Replaces, for unrelated samples and a quantitative trait,
 *   MetaScoreTest::fitWithGivenGenotype / writeOutput   src/Model.h:3188-3260, 3299-3365
 *   MetaCovTest::fitWithGivenGenotype / printCovariance src/Model.cpp:844-1004 (window: src/Model.h:3954-3967)
 * The pending pushes (rvt_gene_push_*; each push = a run of consecutive variants, raw ALT-allele
 * coding, hard calls) are read as ONE ordered variant list v = 0..nv-1.
 *   rvt_meta_plan : host only.  pos/chrom (nv each, sorted by chrom then pos), window in bp
 *                   (ModelManager default 1000000, src/ModelManager.cpp:52,228) -> *wmax = the largest
 *                   number of later variants any variant must be paired with.
 *   rvt_meta_flush: vout[nv] = one rvt_variant_result per variant; band (may be NULL: score only) =
 *                   nv x (wmax+1) doubles, band[v*(wmax+1)+d] = COV entry of variants (v, v+d) divided
 *                   by N as the reference prints it; NaN where either variant is monomorphic (such
 *                   variants are never queued by MetaCovTest) or v+d is outside v's window.
 *                   vout / band may be host or device pointers.
 * Mixed-model band: after rvt_set_null_residual (Bolt score step) set option "meta_cov_scale" = rvt_bolt_null.xvx_xx_ratio:
 *                   the entries are then BoltLMM::GetCovXX / N = g_v'(I - ZZ')g_{v+d} * xVx_xx_ratio / N
 *                   (regression/BoltLMM.cpp:435-460, MetaCovFamQtlBolt, src/Model.cpp:780-805); 0 restores the default. */
typedef struct rvt_variant_result {
  double af;          /* AF                 GenotypeCounter::getAF */
  double ac;          /* INFORMATIVE_ALT_AC GenotypeCounter::getAC */
  double call_rate;   /* CALL_RATE */
  double hwe_p;       /* HWE_PVALUE         SNPHWE */
  int32_t n_ref, n_het, n_alt;   /* N_REF N_HET N_ALT */
  int32_t ok;         /* fitOK: polymorphic and the score test succeeded (else the stats print NA) */
  int32_t polymorphic;
  int32_t pad;
  double U;           /* U_STAT       = U / sigma2 */
  double sqrtV;       /* SQRT_V_STAT  = sqrt(V / sigma2^2) */
  double effect;      /* ALT_EFFSIZE */
  double effect_se;   /* ALT_EFFSIZE_SE (tag "se") */
  double pvalue;      /* PVALUE */
} rvt_variant_result;

int rvt_meta_plan(rvt_ctx* ctx, const int32_t* pos, const int32_t* chrom, int64_t nv, int64_t window_bp, int* wmax);
int rvt_meta_flush(rvt_ctx* ctx, const int32_t* pos, const int32_t* chrom, int64_t window_bp,
                   rvt_variant_result* vout, int64_t cap_variants, double* band, int64_t cap_band, int* wmax);
/* Binary trait (rvt_set_null_model(.., binary = 1)): rvt_meta_flush evaluates MetaUnrelatedBinary (src/Model.h:3669-3784:
 * U = g'(y - p), V = g'Wg - g'WZ (Z'WZ)^-1 Z'Wg with W = diag(p(1 - p)) of the logistic null model, U_STAT = U,
 * SQRT_V_STAT = sqrt(V), ALT_EFFSIZE = U / V, SE = 1 / sqrt(V); LogisticRegressionScoreTest.cpp:219-302 with the intended
 * 1 x 1 solve) and MetaCovUnrelatedBinary (src/Model.cpp:695-778: band entries (g_i'W g_j - covXZ_i covZZ^-1 covXZ_j') / N on
 * the raw genotypes).  The columns MetaScoreTest / MetaCovTest print only for a binary trait come from
 * rvt_meta_binary_extras, for the variants of the LAST rvt_meta_flush:
 *   cc[v]                 genotype counts and exact HWE p among the cases [0] and the controls [1] (the all:case:control
 *                         columns, src/Model.h:3300-3330; AF / AC / call rate follow from the counts)
 *   cov_xz[v*C + l]       covXZ of variant v = g_v'W Z (printCovariance appends covXZ / N of the head variant and
 *   cov_zz[l*C + m]       covZZ / N = Z'WZ / N after the band, src/Model.cpp:985-994)
 * Any of the three pointers may be NULL. */
typedef struct rvt_variant_cc {
  int32_t n[2];                              /* cases, controls */
  int32_t n_ref[2], n_het[2], n_alt[2];
  double hwe_p[2];
} rvt_variant_cc;
int rvt_meta_binary_extras(rvt_ctx* ctx, rvt_variant_cc* cc, double* cov_xz, double* cov_zz, int64_t cap_variants);

/* ---- measurement hooks ----------------------------------------------------------------------- */
/* device milliseconds of the last flush: [0] sweep kernel(s), [1] finalize kernel(s), [2] whole
 * flush on the context stream (CUDA events), [3] number of kernel launches */
int rvt_last_timing(const rvt_ctx* ctx, double out[4]);
/* diagnostic: copy the raw sweep partials of the last flush's final batch (device layout, see
 * rvtests_b200/csrc/common.cuh SweepPartial) to the host; returns the byte count via *bytes */
int rvt_debug_partials(rvt_ctx* ctx, void* out, int64_t cap_bytes, int64_t* bytes);
/* diagnostic: with option "debug_phases"=1 the finalize kernel records SM cycle counts of its
 * phases per gene: [0] split reduction, [1] K build, [2] eigenvalues, [3] Davies, [4] Liu+burden,
 * [5] unused.  out: cap_genes x 6 int64 */
int rvt_debug_phases(rvt_ctx* ctx, long long* out, int cap_genes);

#ifdef __cplusplus
}
#endif
#endif /* RVTESTS_B200_H_ */
