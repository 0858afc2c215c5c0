// rvt_fitters.h -- C++ adapters with the reference's ModelFitter method set, backed by the C ABI
// (include/rvtests_b200.h).  Header-only and templated on the reference's own types so that the
// very same file compiles (a) inside rvtests against src/DataConsolidator.h, base/IO.h (FileWriter)
// and src/Result.h -- see INTEGRATION.md -- and (b) in this repository against the small shims of
// rvtests_b200/host/shim.h used by the adapter test.
//
// Mirrors (names, model names, column headers, NA behaviour):
//   SkatTest     src/Model.h:2612-2772   modelName "Skat"     header "Q\tPvalue"
//   SkatOTest    src/Model.h:2774-2889   modelName "SkatO"    header "Q\trho\tPvalue"
//   CMCTest      src/Model.h:807-907     modelName "CMC"      header "NonRefSite\tPvalue"
//   ZegginiTest  src/Model.h:1170-1242   modelName "Zeggini"  header "Pvalue"
// Behavioural difference, by design: fit() only ENQUEUES the gene; statistics materialise when the
// batch is flushed (every `batch` genes, or in writeFootnote()/the destructor) and the output lines
// are then written in arrival order -- the same deferred-output pattern as the in-tree MetaCovTest
// (src/Model.cpp:828-834, writers outlive models: src/ModelManager.cpp:304-315).  The `siteInfo`
// Result passed to writeOutput() is a reused buffer (src/Main.cpp:1085,1224), so its joined value
// is snapshotted.  Permutation p-values (`skat[nPerm=..,alpha=..]`, the reference's default nPerm = 10000) come
// from the engine's device replay of the reference's rand()-driven shuffles (csrc/perm.cuh).
#ifndef RVT_FITTERS_H_
#define RVT_FITTERS_H_

#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "rvtests_b200.h"

namespace rvtb200 {

// One engine context per process, shared by every adapter so that a gene is uploaded once even
// when several tests (skat, skato, cmc, zeggini) run on it.
template <class DC>
class GeneBatcher {
 public:
  static GeneBatcher& instance() {
    static GeneBatcher b;
    return b;
  }
  ~GeneBatcher() {
    if (ctx_) rvt_ctx_destroy(ctx_);
  }
  void setBatch(int n) { batch_ = n > 0 ? n : 1; }
  void enableSkatO() { skato_ = true; }
  // SkatO::Fit type "D" (src/Model.h:2854-2858) = the engine's "skato_binary" option: on by default (as the reference
  // computes it); enableSkatOBinary(false) prints NA for a binary trait instead
  void enableSkatOBinary(bool on = true) { skato_binary_ = on; }
  // setBinaryOutcome() of any adapter (ModelManager sets every model alike, src/ModelManager.cpp:274-282): logistic null
  void setBinary(bool b) {
    if (b != binary_) have_null_ = false;
    binary_ = b;
  }
  void enablePerm(int nPerm, double alpha) {
    perm_n_ = nPerm;
    perm_alpha_ = alpha;
  }
  // permutation record of a ticket (NULL when the permutation test is off)
  const rvt_perm_result* perm(int ticket) {
    if (ticket < 0 || perm_n_ <= 0) return NULL;
    if (ticket >= (int)results_.size() && !flush()) return NULL;
    return ticket < (int)perms_.size() ? &perms_[ticket] : NULL;
  }
  const char* error() const { return ctx_ ? rvt_last_error(ctx_) : "no context"; }

  // Called from fit(): returns the ticket of the CURRENT gene, uploading it on first sight.
  // A fitter id seen twice means the caller's gene loop has advanced (src/Main.cpp:1249-1253).
  int submit(int fitter_id, DC* dc) {
    if (!ensureContext()) return -1;
    if (current_ < 0 || seen(fitter_id)) {
      if (!pushGene(dc)) return -1;
      seen_.clear();
    }
    seen_.push_back(fitter_id);
    return current_;
  }
  // Statistics of a ticket; flushes the queue when the ticket is still pending.
  const rvt_gene_result* result(int ticket) {
    if (ticket < 0) return NULL;
    if (ticket >= (int)results_.size() && !flush()) return NULL;
    return ticket < (int)results_.size() ? &results_[ticket] : NULL;
  }
  bool shouldFlush() const { return rvt_pending(ctx_) >= batch_; }
  bool flush() {
    int n = ctx_ ? rvt_pending(ctx_) : 0;
    if (n == 0) return true;
    size_t base = results_.size();
    results_.resize(base + n);
    int got = 0;
    if (rvt_flush(ctx_, &results_[base], n, &got) != RVT_OK || got != n) {
      fprintf(stderr, "rvtests_b200: flush failed: %s\n", error());
      results_.resize(base);
      return false;
    }
    perms_.resize(base + n);
    for (int i = 0; i < n; ++i) memset(&perms_[base + i], 0, sizeof(rvt_perm_result));
    if (perm_n_ > 0) {
      int gotp = 0;
      if (rvt_perm_results(ctx_, &perms_[base], n, &gotp) != RVT_OK || gotp != n) {
        fprintf(stderr, "rvtests_b200: permutation records: %s\n", error());
        return false;
      }
    }
    return true;
  }
  int newFitterId() { return next_id_++; }

 private:
  GeneBatcher() : ctx_(NULL), batch_(256), skato_(false), skato_binary_(true), binary_(false), perm_n_(0), perm_alpha_(0.05), current_(-1), next_id_(0), have_null_(false) {}
  bool seen(int id) const {
    for (size_t i = 0; i < seen_.size(); ++i)
      if (seen_[i] == id) return true;
    return false;
  }
  bool ensureContext() {
    if (ctx_) return true;
    if (rvt_ctx_create(0, &ctx_) != RVT_OK) {
      fprintf(stderr, "rvtests_b200: %s\n", error());
      if (ctx_) rvt_ctx_destroy(ctx_);
      ctx_ = NULL;
      return false;  // no CPU fallback: the adapters report fit() == -1 and print NA
    }
    return true;
  }
  // copyCovariateAndIntercept + FitLinearModel (src/ModelUtil.h:102-130, src/Model.h:2672-2699)
  bool ensureNullModel(DC* dc) {
    if (have_null_ && !dc->isPhenotypeUpdated() && !dc->isCovariateUpdated()) return true;
    if (rvt_pending(ctx_) > 0 && !flush()) return false;
    const auto& ph = dc->getPhenotype();
    const auto& cv = dc->getCovariate();
    const int n = ph.rows, c = cv.cols + 1;
    std::vector<double> X((size_t)n * c), y(n);
    for (int i = 0; i < n; ++i) {
      X[i] = 1.0;
      y[i] = ph(i, 0);
    }
    for (int j = 0; j < cv.cols; ++j)
      for (int i = 0; i < n; ++i) X[(size_t)(j + 1) * n + i] = cv(i, j);
    if (skato_) rvt_set_option(ctx_, "skato", 1);
    rvt_set_option(ctx_, "skato_binary", skato_binary_ ? 1 : 0);
    if (perm_n_ > 0) {
      rvt_set_option(ctx_, "perm", perm_n_);
      rvt_set_option(ctx_, "perm_alpha", perm_alpha_);
    }
    if (rvt_set_null_model(ctx_, n, c, X.data(), y.data(), binary_ ? 1 : 0) != RVT_OK) {
      fprintf(stderr, "rvtests_b200: null model: %s\n", error());
      return false;
    }
    have_null_ = true;
    return true;
  }
  bool pushGene(DC* dc) {
    if (!ensureNullModel(dc)) return false;
    const auto& g = dc->getGenotype();  // N x M, column-major doubles, imputed, not flipped
    if (g.cols == 0) {                  // src/Model.h:2637-2640: fit() returns -1, output NA
      rvt_gene_result na;
      memset(&na, 0, sizeof(na));
      na.status = RVT_GENE_NA;
      // keep ticket numbering dense: pending genes first
      if (!flush()) return false;
      results_.push_back(na);
      rvt_perm_result np;
      memset(&np, 0, sizeof(np));
      perms_.push_back(np);
      current_ = (int)results_.size() - 1;
      return true;
    }
    std::vector<double> af(g.cols);
    for (int j = 0; j < g.cols; ++j) af[j] = dc->getMarkerFrequency(j);
    if (rvt_gene_push_f64(ctx_, &g.data[0], g.cols, af.data()) != RVT_OK) {
      fprintf(stderr, "rvtests_b200: push: %s\n", error());
      return false;
    }
    current_ = (int)results_.size() + rvt_pending(ctx_) - 1;
    return true;
  }

  rvt_ctx* ctx_;
  int batch_;
  bool skato_, skato_binary_, binary_;
  int perm_n_;
  double perm_alpha_;
  int current_, next_id_;
  bool have_null_;
  std::vector<int> seen_;
  std::vector<rvt_gene_result> results_;
  std::vector<rvt_perm_result> perms_;
};

// Common machinery: ticket per gene, deferred lines.
template <class DC, class FW, class RES>
class DeferredFitter {
 public:
  DeferredFitter() : ticket_(-1), fp_(NULL), binary_(false) { id_ = GeneBatcher<DC>::instance().newFitterId(); }
  virtual ~DeferredFitter() {}
  const std::string& getModelName() const { return modelName; }
  void setBinaryOutcome() {
    binary_ = true;
    GeneBatcher<DC>::instance().setBinary(true);
  }
  void setQuantitativeOutcome() {
    binary_ = false;
    GeneBatcher<DC>::instance().setBinary(false);
  }
  bool isBinaryOutcome() const { return binary_; }
  bool needToIndexResult() const { return false; }
  void reset() { ticket_ = -1; }
  int fit(DC* dc) {
    ticket_ = GeneBatcher<DC>::instance().submit(id_, dc);
    return ticket_ >= 0 ? 0 : -1;
  }
  void writeOutput(FW* fp, const RES& siteInfo) {
    fp_ = fp;
    Pending p;
    p.ticket = ticket_;
    p.site = siteInfo.joinValue();
    pending_.push_back(p);
    if (GeneBatcher<DC>::instance().shouldFlush()) drain();
  }
  void writeFootnote(FW* fp) {
    if (!fp_) fp_ = fp;
    drain();
  }

 protected:
  virtual void formatLine(const rvt_gene_result* r, std::string* out) const = 0;
  virtual void formatTicket(int ticket, std::string* out) const {
    formatLine(GeneBatcher<DC>::instance().result(ticket), out);
  }
  void drain() {
    if (!fp_) return;
    GeneBatcher<DC>& b = GeneBatcher<DC>::instance();
    for (size_t i = 0; i < pending_.size(); ++i) {
      std::string line = pending_[i].site;
      line += "\t";
      (void)b;
      formatTicket(pending_[i].ticket, &line);
      line += "\n";
      fp_->write(line.c_str());
    }
    pending_.clear();
  }
  static std::string g(double v) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%g", v);
    return buf;
  }
  std::string modelName;

 private:
  struct Pending {
    int ticket;
    std::string site;
  };
  int id_, ticket_;
  FW* fp_;
  bool binary_;
  std::vector<Pending> pending_;
};

template <class DC, class FW, class RES>
class SkatTestB200 : public DeferredFitter<DC, FW, RES> {
 public:
  // SkatTest(int nPerm, double alpha, double beta1, double beta2), src/Model.h:2615-2622 (beta1/beta2: engine options)
  explicit SkatTestB200(int nPerm = 0, double alpha = 0.05) : usePermutation_(nPerm > 0) {
    this->modelName = "Skat";
    if (usePermutation_) GeneBatcher<DC>::instance().enablePerm(nPerm, alpha);
  }
  ~SkatTestB200() { this->drain(); }
  void writeHeader(FW* fp, const RES& siteInfo) {
    siteInfo.writeHeaderTab(fp);
    if (!usePermutation_)
      fp->write("Q\tPvalue\n");
    else   // src/Model.h:2722-2731 + Permutation::writeHeader (src/Permutation.h:56-61)
      fp->write("Q\tPvalue\tNumPerm\tActualPerm\tStat\tNumGreater\tNumEqual\tPermPvalue\n");
  }

 protected:
  void formatLine(const rvt_gene_result* r, std::string* out) const {
    if (!r || r->status != RVT_GENE_OK) {
      *out += "NA\tNA";
      return;
    }
    *out += this->g(r->Q) + "\t" + this->g(r->p_skat);
  }
  void formatTicket(int ticket, std::string* out) const {
    GeneBatcher<DC>& b = GeneBatcher<DC>::instance();
    const rvt_gene_result* r = b.result(ticket);
    formatLine(r, out);
    if (!usePermutation_) return;
    const rvt_perm_result* p = b.perm(ticket);
    if (!r || r->status != RVT_GENE_OK || !p || !p->done) {   // src/Model.h:2736-2740
      *out += "\tNA\tNA\tNA\tNA\tNA\tNA";
      return;
    }
    char buf[160];   // Permutation::updateValue: ints via toString, doubles via floatToString (== %g)
    snprintf(buf, sizeof(buf), "\t%d\t%d\t%g\t%d\t%d\t%g", p->num_perm, p->actual_perm, p->stat, p->num_greater, p->num_equal,
             p->p_perm);
    *out += buf;
  }

 private:
  bool usePermutation_;
};

template <class DC, class FW, class RES>
class SkatOTestB200 : public DeferredFitter<DC, FW, RES> {
 public:
  SkatOTestB200() {
    this->modelName = "SkatO";
    GeneBatcher<DC>::instance().enableSkatO();
  }
  ~SkatOTestB200() { this->drain(); }
  void writeHeader(FW* fp, const RES& siteInfo) {
    siteInfo.writeHeaderTab(fp);
    fp->write("Q\trho\tPvalue\n");
  }

 protected:
  void formatLine(const rvt_gene_result* r, std::string* out) const {
    if (!r || r->status != RVT_GENE_OK || !r->skato_ok) {
      *out += "NA\tNA\tNA";
      return;
    }
    *out += this->g(r->skato_Q) + "\t" + this->g(r->skato_rho) + "\t" + this->g(r->skato_p);
  }
};

template <class DC, class FW, class RES>
class CMCTestB200 : public DeferredFitter<DC, FW, RES> {
 public:
  CMCTestB200() { this->modelName = "CMC"; }
  ~CMCTestB200() { this->drain(); }
  void writeHeader(FW* fp, const RES& siteInfo) {
    siteInfo.writeHeaderTab(fp);
    fp->write("NonRefSite\tPvalue\n");
  }

 protected:
  void formatLine(const rvt_gene_result* r, std::string* out) const {
    if (!r || r->status != RVT_GENE_OK || !r->cmc_ok) {
      *out += "NA\tNA";
      return;
    }
    char buf[32];
    snprintf(buf, sizeof(buf), "%d", r->cmc_nonref);
    *out += std::string(buf) + "\t" + this->g(r->cmc_p);
  }
};

template <class DC, class FW, class RES>
class ZegginiTestB200 : public DeferredFitter<DC, FW, RES> {
 public:
  ZegginiTestB200() { this->modelName = "Zeggini"; }
  ~ZegginiTestB200() { this->drain(); }
  void writeHeader(FW* fp, const RES& siteInfo) {
    siteInfo.writeHeaderTab(fp);
    fp->write("Pvalue\n");
  }

 protected:
  void formatLine(const rvt_gene_result* r, std::string* out) const {
    if (!r || r->status != RVT_GENE_OK || !r->zeg_ok) {
      *out += "NA";
      return;
    }
    *out += this->g(r->zeg_p);
  }
};

}  // namespace rvtb200
#endif  // RVT_FITTERS_H_
