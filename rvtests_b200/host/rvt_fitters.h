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
// (src/Model.cpp:828-834, writers outlive models: src/ModelManager.cpp:304-315).
// Every adapter takes a 4th template parameter BASE: inside rvtests it is the reference's own ModelFitter
// (src/ModelFitter.h:17-75) -- fit / writeHeader / writeOutput / writeFootnote / reset then OVERRIDE its virtuals and the
// objects sit in ModelManager's std::vector<ModelFitter*> like any other model (rvtests_b200/host/ModelB200.h; built
// against the real headers and run next to the reference's fitters by oracle/ref_dropin_shim.cpp).  Stand-alone
// (shim.h, the adapter demos) BASE defaults to a small struct with the same members.  The `siteInfo`
// Result passed to writeOutput() is a reused buffer (src/Main.cpp:1085,1224), so its joined value
// is snapshotted.  Permutation p-values (`skat[nPerm=..,alpha=..]`, the reference's default nPerm = 10000) come
// from the engine's device replay of the reference's rand()-driven shuffles (csrc/perm.cuh).
#ifndef RVT_FITTERS_H_
#define RVT_FITTERS_H_

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "rvtests_b200.h"

namespace rvtb200 {

// what the adapters use of ModelFitter, for builds without the reference tree
struct StandaloneBase {
  StandaloneBase() : modelName("UninitializedModel"), binaryOutcome(false), indexResult(false) {}
  virtual ~StandaloneBase() {}
  const std::string& getModelName() const { return modelName; }
  bool isBinaryOutcome() const { return binaryOutcome; }
  void setBinaryOutcome() { binaryOutcome = true; }
  void setQuantitativeOutcome() { binaryOutcome = false; }
  bool needToIndexResult() const { return indexResult; }
  virtual void reset() {}

 protected:
  std::string modelName;
  bool binaryOutcome, indexResult;
};

// One engine per process, shared by every adapter so that a gene is uploaded once even when several tests (skat, skato,
// cmc, zeggini) run on it.  It drives ONE CONTEXT PER DEVICE: setDevices(n) / $RVTESTS_B200_DEVICES = n (default 1) deals
// the genes round-robin over devices first .. first + n - 1 ($RVTESTS_B200_DEVICE = first, default 0); a flush runs the
// contexts concurrently, one host thread each (the C ABI's rule: one thread per context), and hands the records back in
// arrival order.  The permutation test replays ONE rand() stream in gene order (csrc/perm.cuh), so with nPerm > 0 a
// single device is used.
template <class DC>
class GeneBatcher {
 public:
  static GeneBatcher& instance() {
    static GeneBatcher b;
    return b;
  }
  ~GeneBatcher() {
    for (size_t k = 0; k < ctx_.size(); ++k)
      if (ctx_[k]) rvt_ctx_destroy(ctx_[k]);
  }
  void setBatch(int n) { batch_ = n > 0 ? n : 1; }
  // Beta(MAF; beta1, beta2) weights of --kernel skat[beta1=..,beta2=..] / skato[..] (src/ModelManager.cpp:169-185): one
  // engine serves every adapter, so the pair is process-wide (the reference's default 1, 25 unless a model says otherwise)
  void setBeta(double b1, double b2) {
    if (b1 != beta1_ || b2 != beta2_) have_null_ = false;
    beta1_ = b1;
    beta2_ = b2;
  }
  void setDevice(int first) { device_ = first; }      // before the first fit(); default $RVTESTS_B200_DEVICE or 0
  void setDevices(int n) { n_devices_ = n; }          // before the first fit(); default $RVTESTS_B200_DEVICES or 1
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
  const char* error() const { return last_error_.c_str(); }

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
  bool shouldFlush() const { return (int)owner_.size() >= batch_; }
  bool flush() {
    const int n = (int)owner_.size();
    if (n == 0) return true;
    const int nc = (int)ctx_.size();
    std::vector<std::vector<rvt_gene_result> > res(nc);
    std::vector<std::vector<rvt_perm_result> > prm(nc);
    std::vector<int> ok(nc, 1);
    std::vector<std::thread> th;
    for (int k = 0; k < nc; ++k) {
      if (nc == 1)
        flushOne(k, &res[k], &prm[k], &ok[k]);
      else
        th.push_back(std::thread(&GeneBatcher::flushOne, this, k, &res[k], &prm[k], &ok[k]));
    }
    for (size_t k = 0; k < th.size(); ++k) th[k].join();
    bool all = true;
    for (int k = 0; k < nc; ++k)
      if (!ok[k]) {
        last_error_ = rvt_last_error(ctx_[k]);
        fprintf(stderr, "rvtests_b200: flush failed on device %d: %s\n", device_ + k, last_error_.c_str());
        all = false;
      }
    // hand the records back in arrival order; a failed device leaves NA records (status != OK) for its genes
    std::vector<size_t> at(nc, 0);
    for (int i = 0; i < n; ++i) {
      const int k = owner_[i];
      rvt_gene_result r;
      rvt_perm_result p;
      memset(&r, 0, sizeof(r));
      memset(&p, 0, sizeof(p));
      r.status = RVT_GENE_NA;
      if (ok[k] && at[k] < res[k].size()) {
        r = res[k][at[k]];
        if (at[k] < prm[k].size()) p = prm[k][at[k]];
      }
      ++at[k];
      results_.push_back(r);
      perms_.push_back(p);
    }
    owner_.clear();
    return all;
  }
  int newFitterId() { return next_id_++; }
  // adapters register for their lifetime: when the last one is gone (ModelManager::close deletes the models,
  // src/ModelManager.cpp:304-315) the batcher forgets the run -- tickets, the current gene, the null model -- so that a
  // second ModelManager in the same process starts clean
  void attach() { ++n_attached_; }
  void detach() {
    if (--n_attached_ > 0) return;
    if (!owner_.empty()) flush();
    results_.clear();
    perms_.clear();
    seen_.clear();
    current_ = -1;
    have_null_ = false;
    skato_ = false;
    perm_n_ = 0;
  }

 private:
  GeneBatcher()
      : n_attached_(0), device_(-1), n_devices_(-1), beta1_(1.0), beta2_(25.0), batch_(256), skato_(false), skato_binary_(true),
        binary_(false), perm_n_(0), perm_alpha_(0.05), current_(-1), next_id_(0), next_ctx_(0), have_null_(false) {}
  bool seen(int id) const {
    for (size_t i = 0; i < seen_.size(); ++i)
      if (seen_[i] == id) return true;
    return false;
  }
  void flushOne(int k, std::vector<rvt_gene_result>* res, std::vector<rvt_perm_result>* prm, int* ok) {
    const int n = rvt_pending(ctx_[k]);
    res->resize(n);
    prm->clear();
    if (n == 0) return;
    int got = 0;
    if (rvt_flush(ctx_[k], res->data(), n, &got) != RVT_OK || got != n) {
      *ok = 0;
      return;
    }
    if (perm_n_ > 0) {
      prm->resize(n);
      int gotp = 0;
      if (rvt_perm_results(ctx_[k], prm->data(), n, &gotp) != RVT_OK || gotp != n) *ok = 0;
    }
  }
  bool ensureContext() {
    if (!ctx_.empty()) return true;
    if (device_ < 0) {
      const char* e = getenv("RVTESTS_B200_DEVICE");
      device_ = e ? atoi(e) : 0;
    }
    if (n_devices_ < 1) {
      const char* e = getenv("RVTESTS_B200_DEVICES");
      n_devices_ = e ? atoi(e) : 1;
      if (n_devices_ < 1) n_devices_ = 1;
    }
    if (perm_n_ > 0) n_devices_ = 1;   // one rand() stream in gene order
    for (int k = 0; k < n_devices_; ++k) {
      rvt_ctx* c = NULL;
      if (rvt_ctx_create(device_ + k, &c) != RVT_OK) {
        last_error_ = c ? rvt_last_error(c) : "context allocation failed";
        fprintf(stderr, "rvtests_b200: %s\n", last_error_.c_str());
        if (c) rvt_ctx_destroy(c);
        for (size_t j = 0; j < ctx_.size(); ++j) rvt_ctx_destroy(ctx_[j]);
        ctx_.clear();
        return false;  // no CPU fallback: the adapters report fit() == -1 and print NA
      }
      ctx_.push_back(c);
    }
    return true;
  }
  // copyCovariateAndIntercept + FitLinearModel (src/ModelUtil.h:102-130, src/Model.h:2672-2699), on every device
  bool ensureNullModel(DC* dc) {
    if (have_null_ && !dc->isPhenotypeUpdated() && !dc->isCovariateUpdated()) return true;
    if (!owner_.empty() && !flush()) return false;
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
    for (size_t k = 0; k < ctx_.size(); ++k) {
      rvt_ctx* c_ = ctx_[k];
      rvt_set_option(c_, "skato", skato_ ? 1 : 0);
      rvt_set_option(c_, "skato_binary", skato_binary_ ? 1 : 0);
      rvt_set_option(c_, "beta1", beta1_);
      rvt_set_option(c_, "beta2", beta2_);
      rvt_set_option(c_, "perm", perm_n_ > 0 ? perm_n_ : 0);
      if (perm_n_ > 0) rvt_set_option(c_, "perm_alpha", perm_alpha_);
      if (rvt_set_null_model(c_, n, c, X.data(), y.data(), binary_ ? 1 : 0) != RVT_OK) {
        last_error_ = rvt_last_error(c_);
        fprintf(stderr, "rvtests_b200: null model: %s\n", last_error_.c_str());
        return false;
      }
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
    const int k = next_ctx_;
    if (rvt_gene_push_f64(ctx_[k], &g.data[0], g.cols, af.data()) != RVT_OK) {
      last_error_ = rvt_last_error(ctx_[k]);
      fprintf(stderr, "rvtests_b200: push: %s\n", last_error_.c_str());
      return false;
    }
    next_ctx_ = (next_ctx_ + 1) % (int)ctx_.size();
    owner_.push_back(k);
    current_ = (int)results_.size() + (int)owner_.size() - 1;
    return true;
  }

  int n_attached_;
  std::vector<rvt_ctx*> ctx_;
  int device_, n_devices_;
  double beta1_, beta2_;
  int batch_;
  bool skato_, skato_binary_, binary_;
  int perm_n_;
  double perm_alpha_;
  int current_, next_id_, next_ctx_;
  bool have_null_;
  std::string last_error_;
  std::vector<int> seen_;
  std::vector<int> owner_;   // pending genes in arrival order: the context each one was pushed to
  std::vector<rvt_gene_result> results_;
  std::vector<rvt_perm_result> perms_;
};

// Common machinery: ticket per gene, deferred lines.
template <class DC, class FW, class RES, class BASE = StandaloneBase>
class DeferredFitter : public BASE {
 public:
  DeferredFitter() : ticket_(-1), fp_(NULL) {
    id_ = GeneBatcher<DC>::instance().newFitterId();
    GeneBatcher<DC>::instance().attach();
  }
  virtual ~DeferredFitter() { GeneBatcher<DC>::instance().detach(); }
  // ModelFitter::reset() clears the model's own Result (src/ModelFitter.h:46); the adapters keep none
  virtual void reset() {
    BASE::reset();
    ticket_ = -1;
  }
  virtual int fit(DC* dc) {
    // ModelManager flips the (non-virtual) outcome flag of every model alike (src/ModelManager.cpp:274-282): read it here
    GeneBatcher<DC>::instance().setBinary(this->isBinaryOutcome());
    ticket_ = GeneBatcher<DC>::instance().submit(id_, dc);
    return ticket_ >= 0 ? 0 : -1;
  }
  virtual void writeHeader(FW* fp, const RES& siteInfo) = 0;
  virtual void writeOutput(FW* fp, const RES& siteInfo) {
    fp_ = fp;
    Pending p;
    p.ticket = ticket_;
    p.site = siteInfo.joinValue();
    pending_.push_back(p);
    if (GeneBatcher<DC>::instance().shouldFlush()) drain();
  }
  virtual void writeFootnote(FW* fp) {
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
    for (size_t i = 0; i < pending_.size(); ++i) {
      std::string line = pending_[i].site;
      line += "\t";
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

 private:
  struct Pending {
    int ticket;
    std::string site;
  };
  int id_, ticket_;
  FW* fp_;
  std::vector<Pending> pending_;
};

template <class DC, class FW, class RES, class BASE = StandaloneBase>
class SkatTestB200 : public DeferredFitter<DC, FW, RES, BASE> {
 public:
  // SkatTest(int nPerm, double alpha, double beta1, double beta2), src/Model.h:2615-2622 (beta1/beta2: engine options)
  explicit SkatTestB200(int nPerm = 0, double alpha = 0.05, double beta1 = 1.0, double beta2 = 25.0) : usePermutation_(nPerm > 0) {
    this->modelName = "Skat";
    GeneBatcher<DC>::instance().setBeta(beta1, beta2);
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

template <class DC, class FW, class RES, class BASE = StandaloneBase>
class SkatOTestB200 : public DeferredFitter<DC, FW, RES, BASE> {
 public:
  explicit SkatOTestB200(double beta1 = 1.0, double beta2 = 25.0) {   // SkatOTest(beta1, beta2), src/Model.h:2776-2783
    this->modelName = "SkatO";
    GeneBatcher<DC>::instance().setBeta(beta1, beta2);
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

template <class DC, class FW, class RES, class BASE = StandaloneBase>
class CMCTestB200 : public DeferredFitter<DC, FW, RES, BASE> {
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

template <class DC, class FW, class RES, class BASE = StandaloneBase>
class ZegginiTestB200 : public DeferredFitter<DC, FW, RES, BASE> {
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
