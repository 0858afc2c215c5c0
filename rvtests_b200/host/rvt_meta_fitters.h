// rvt_meta_fitters.h -- ModelFitter-shaped adapters for `--meta score,cov`, backed by rvt_meta_flush (include/rvtests_b200.h).
// Same conventions as rvt_fitters.h (templated on the reference's DataConsolidator / FileWriter / Result so that the file
// compiles inside rvtests and against host/shim.h here).
//
// Mirrors:
//   MetaScoreTest  src/Model.h:3154-3365   modelName "MetaScore"; columns AF INFORMATIVE_ALT_AC CALL_RATE HWE_PVALUE N_REF
//                  N_HET N_ALT U_STAT SQRT_V_STAT ALT_EFFSIZE [ALT_EFFSIZE_SE] PVALUE (:3263-3279); statistics of a
//                  monomorphic or failed site print NA, its counts are still printed (:3290-3345); the
//                  "##NullModelEstimates" block of MetaUnrelatedQtl::PrintNullModel (:3525-3541) precedes the header
//   MetaCovTest    src/Model.cpp:807-1004, src/Model.h:3954-3967: one line per polymorphic variant, CHROM START_POS END_POS
//                  NUM_MARKER MARKER_POS COV, listing every later polymorphic variant within windowSize bp on the same
//                  chromosome (the loci queued when the head is popped), COV entries divided by N and printed as floats
// Behavioural difference, by design (as in rvt_fitters.h): fit() only records the variant; variants are packed 64 to a block,
// blocks are evaluated in segments, and the lines are written in arrival order when a segment is flushed (every
// `segment` variants, at a chromosome change, in writeFootnote / the destructor).  A segment keeps the trailing variants
// whose window is still open and re-submits them with the next segment, so covariance windows are never cut.
// Binary traits (the fitter's isBinaryOutcome(), read at fit time): MetaUnrelatedBinary / MetaCovUnrelatedBinary -- the site
// columns become all:case:control triples (src/Model.h:3300-3330), the null-model block prints Sigma2 NA NA (:3733-3747),
// and every MetaCov line carries ":covXZ/N:covZZ/N" of its head variant after the band (src/Model.cpp:985-994).
// Not covered: kinship (family) models, hemizygous regions, dosages (a site with a value outside {0,1,2}
// prints NA statistics).
#ifndef RVT_META_FITTERS_H_
#define RVT_META_FITTERS_H_

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <deque>
#include <string>
#include <vector>

#include "rvtests_b200.h"
#include "rvt_fitters.h"   // StandaloneBase
#include "rvt_summary.h"   // SummaryHook

namespace rvtb200 {

template <class DC>
class MetaBatcher {
 public:
  static MetaBatcher& instance() {
    static MetaBatcher b;
    return b;
  }
  ~MetaBatcher() {
    if (ctx_) rvt_ctx_destroy(ctx_);
  }
  void setSegment(int nVariants) { segment_ = nVariants > 64 ? nVariants : 64; }
  void setWindow(int bp) { window_ = bp; }
  void enableCov() { want_cov_ = true; }
  const char* error() const { return ctx_ ? rvt_last_error(ctx_) : "no context"; }
  int newFitterId() { return next_id_++; }
  // every adapter attaches in its constructor and detaches in its destructor (after draining its own lines): when the last
  // one of a run is gone the batcher forgets the run, so that a second ModelManager in the same process starts clean
  void attach() { ++n_attached_; }
  void detach() {
    if (--n_attached_ > 0) return;
    sites_.clear();
    blocks_.clear();
    open_ = Block{0, 0, std::vector<int8_t>()};
    seen_.clear();
    chroms_.clear();
    current_ = -1;
    pending_new_ = 0;
    have_null_ = false;
    want_cov_ = false;
    binary_ = false;
    if (ctx_ && rvt_pending(ctx_) > 0) {   // (a failed flush left pushes behind)
      rvt_ctx_destroy(ctx_);
      ctx_ = NULL;
    }
  }

  struct Site {
    int chrom, pos;
    bool hard;           // every value in {0,1,2}
    bool score_done, cov_done;
    rvt_variant_result r;
    rvt_variant_cc cc;      // binary trait: cases [0] / controls [1]
    std::string cov_line;   // empty: nothing to print (monomorphic)
  };

  // Called from fit(): ticket of the CURRENT variant, recorded on first sight (a fitter id seen twice = next variant)
  int submit(int fitter_id, DC* dc, bool binary = false) {
    if (!ensureContext()) return -1;
    if (binary != binary_) {   // the outcome type changed: a new null model
      if (!flush(true)) return -1;
      binary_ = binary;
      have_null_ = false;
    }
    if (current_ < 0 || seen(fitter_id)) {
      if (!record(dc)) return -1;
      seen_.clear();
    }
    seen_.push_back(fitter_id);
    return current_;
  }
  const Site* site(int ticket, bool need_cov) {
    if (ticket < 0 || ticket >= (int)sites_.size()) return NULL;
    if (!sites_[ticket].score_done || (need_cov && !sites_[ticket].cov_done)) return NULL;
    return &sites_[ticket];
  }
  // evaluate what is pending; final = no later variant will extend any window (end of input / destructor)
  bool flush(bool final) {
    if (!ctx_) return true;
    // nothing new since the last evaluation, and no open window that `final` would close
    if (pending_new_ == 0 && !(final && want_cov_ && !blocks_.empty())) return true;
    closeBlock();
    if (blocks_.empty()) return true;
    std::vector<int> tix;
    for (size_t b = 0; b < blocks_.size(); ++b) {
      const Block& k = blocks_[b];
      if (rvt_gene_push_i8(ctx_, k.g.data(), k.M, N_, NULL) != RVT_OK) return fail("push");
      for (int j = 0; j < k.M; ++j) tix.push_back(k.first + j);
    }
    const int64_t nv = (int64_t)tix.size();
    std::vector<int32_t> pos(nv), chrom(nv);
    for (int64_t v = 0; v < nv; ++v) {
      pos[v] = sites_[tix[v]].pos;
      chrom[v] = sites_[tix[v]].chrom;
    }
    int wmax = 0;
    std::vector<double> band;
    if (want_cov_) {
      if (rvt_meta_plan(ctx_, pos.data(), chrom.data(), nv, window_, &wmax) != RVT_OK) return fail("plan");
      band.resize((size_t)nv * (wmax + 1));
    }
    std::vector<rvt_variant_result> vr(nv);
    if (rvt_meta_flush(ctx_, pos.data(), chrom.data(), window_, vr.data(), nv, want_cov_ ? band.data() : NULL,
                       (int64_t)band.size(), &wmax) != RVT_OK)
      return fail("flush");
    std::vector<rvt_variant_cc> cc;
    std::vector<double> xz, zz;
    if (binary_) {
      cc.resize(nv);
      xz.resize((size_t)nv * C_);
      zz.resize((size_t)C_ * C_);
      if (rvt_meta_binary_extras(ctx_, cc.data(), xz.data(), zz.data(), nv) != RVT_OK) return fail("binary extras");
    }
    const int last_pos = pos[nv - 1], last_chrom = chrom[nv - 1];
    int64_t first_open = nv;   // first variant whose window may still grow
    for (int64_t v = 0; v < nv; ++v) {
      Site& s = sites_[tix[v]];
      if (!s.score_done) {
        s.r = vr[v];
        if (binary_) s.cc = cc[v];
        if (!s.hard) s.r.ok = 0;
        s.score_done = true;
      }
      if (s.cov_done || !want_cov_) continue;
      // the reference pops a head once a locus farther than windowSize (or on another chromosome) arrives
      const bool closed = final || chrom[v] != last_chrom || abs(last_pos - pos[v]) > window_;
      if (!closed) {
        if (first_open == nv) first_open = v;
        continue;
      }
      s.cov_done = true;
      if (!vr[v].polymorphic || !s.hard) continue;   // never queued: no line
      std::string mp, cv;
      int n = 0;
      char buf[64];
      for (int d = 0; d <= wmax && v + d < nv; ++d) {
        const double c = band[(size_t)v * (wmax + 1) + d];
        if (c != c) continue;   // outside the window, or the partner is monomorphic
        snprintf(buf, sizeof(buf), "%s%d", n ? "," : "", pos[v + d]);
        mp += buf;
        snprintf(buf, sizeof(buf), "%s%g", n ? "," : "", (double)(float)c);   // toString(float) (src/Model.h:4034-4041)
        cv += buf;
        ++n;
        end_pos_ = pos[v + d];
      }
      if (binary_) {   // printCovariance: s += ':' covXZ / n  ':' lower triangle of covZZ / n (src/Model.cpp:985-994)
        const float scale = 1.0f / (float)N_;
        cv += ':';
        for (int l = 0; l < C_; ++l) {
          snprintf(buf, sizeof(buf), "%s%g", l ? "," : "", (double)((float)xz[(size_t)v * C_ + l] * scale));
          cv += buf;
        }
        cv += ':';
        for (int i = 0; i < C_; ++i)
          for (int j = 0; j <= i; ++j) {
            snprintf(buf, sizeof(buf), "%s%g", (i || j) ? "," : "", zz[(size_t)i * C_ + j] * (double)scale);
            cv += buf;
          }
      }
      snprintf(buf, sizeof(buf), "%d", n);
      s.cov_line = chromName(chrom[v]) + "\t" + itoa(pos[v]) + "\t" + itoa(end_pos_) + "\t" + buf + "\t" + mp + "\t" + cv;
    }
    if (!want_cov_) first_open = nv;
    // keep the blocks that still hold an open variant (a suffix), drop the rest
    size_t keep_from = blocks_.size();
    int64_t v0 = 0;
    for (size_t b = 0; b < blocks_.size(); ++b) {
      if (first_open < v0 + blocks_[b].M) {
        keep_from = b;
        break;
      }
      v0 += blocks_[b].M;
    }
    blocks_.erase(blocks_.begin(), blocks_.begin() + keep_from);
    pending_new_ = 0;
    return true;
  }
  bool shouldFlush() const { return pending_new_ >= segment_; }
  // null model for the ##NullModelEstimates block (valid after the first submit)
  bool nullModel(std::vector<double>* beta, std::vector<double>* var, double* sigma2) {
    if (!ctx_ || C_ <= 0) return false;
    beta->resize(C_);
    std::vector<double> xtx((size_t)C_ * C_);
    if (rvt_get_null_beta(ctx_, beta->data()) != RVT_OK) return false;
    if (rvt_get_null_model(ctx_, NULL, sigma2, xtx.data()) != RVT_OK) return false;
    var->resize(C_);
    for (int i = 0; i < C_; ++i) (*var)[i] = xtx[(size_t)i * C_ + i] * *sigma2;   // covB = (X'X)^-1 sigma2, LinearRegression.cpp:62-66
    return true;                                                                   // (binary: (X'VX)^-1 of the logistic fit, sigma2 = 1)
  }
  bool binary() const { return binary_; }

 private:
  struct Block {
    int first, M;             // ticket of row 0, rows
    std::vector<int8_t> g;    // [M][N] variant-major hard calls
  };
  MetaBatcher()
      : ctx_(NULL), N_(0), C_(0), segment_(4096), window_(1000000), want_cov_(false), current_(-1), next_id_(0), have_null_(false),
        pending_new_(0), end_pos_(0) {}
  bool fail(const char* what) {
    fprintf(stderr, "rvtests_b200: meta %s failed: %s\n", what, error());
    return false;
  }
  static std::string itoa(int v) {
    char buf[32];
    snprintf(buf, sizeof(buf), "%d", v);
    return buf;
  }
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
      return false;   // no CPU fallback
    }
    return true;
  }
  bool ensureNullModel(DC* dc) {
    if (have_null_ && !dc->isPhenotypeUpdated() && !dc->isCovariateUpdated()) return true;
    if (!flush(true)) return false;
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
    if (rvt_set_null_model(ctx_, n, c, X.data(), y.data(), binary_ ? 1 : 0) != RVT_OK) return fail("null model");
    N_ = n;
    C_ = c;
    have_null_ = true;
    return true;
  }
  int chromId(const std::string& name) {
    for (size_t i = 0; i < chroms_.size(); ++i)
      if (chroms_[i] == name) return (int)i;
    chroms_.push_back(name);
    return (int)chroms_.size() - 1;
  }
  const std::string& chromName(int id) const { return chroms_[id]; }
  void closeBlock() {
    if (open_.M > 0) {
      open_.g.resize((size_t)open_.M * N_);
      blocks_.push_back(open_);
      open_ = Block();
      open_.M = 0;
    }
  }
  bool record(DC* dc) {
    if (!ensureNullModel(dc)) return false;
    const auto& g = dc->getGenotype();   // N x 1
    Site s;
    memset(&s.r, 0, sizeof(s.r));
    memset(&s.cc, 0, sizeof(s.cc));
    s.score_done = s.cov_done = false;
    s.hard = g.cols == 1 && g.rows == N_;
    const auto& info = dc->getResult();
    s.chrom = chromId(info["CHROM"]);
    s.pos = atoi(info["POS"].c_str());
    // a chromosome change closes every window: evaluate what is pending first
    if (!sites_.empty() && sites_.back().chrom != s.chrom && !flush(true)) return false;
    if (open_.M == 0) {
      open_.first = (int)sites_.size();
      open_.g.assign((size_t)64 * N_, 0);
    }
    int8_t* row = open_.g.data() + (size_t)open_.M * N_;
    if (s.hard)
      for (int i = 0; i < N_; ++i) {
        const double v = g(i, 0);
        if (v == 0.0 || v == 1.0 || v == 2.0)
          row[i] = (int8_t)v;
        else {
          s.hard = false;
          break;
        }
      }
    if (!s.hard) memset(row, 0, (size_t)N_);   // placeholder row (monomorphic): statistics print NA, never queued
    sites_.push_back(s);
    current_ = (int)sites_.size() - 1;
    if (++open_.M == 64) closeBlock();
    ++pending_new_;
    return true;
  }

  rvt_ctx* ctx_;
  int N_, C_, segment_, window_;
  bool want_cov_, binary_ = false;
  int n_attached_ = 0;
  int current_, next_id_;
  bool have_null_;
  int pending_new_, end_pos_;
  std::vector<int> seen_;
  std::vector<std::string> chroms_;
  std::vector<Site> sites_;
  std::deque<Block> blocks_;
  Block open_ = Block{0, 0, std::vector<int8_t>()};
};

// BASE: the reference's ModelFitter inside rvtests (ModelB200.h), StandaloneBase otherwise -- see rvt_fitters.h
template <class DC, class FW, class RES, class BASE = StandaloneBase>
class MetaScoreTestB200 : public BASE {
 public:
  explicit MetaScoreTestB200(bool outputSE = false) : outputSE_(outputSE), ticket_(-1), fp_(NULL), header_(false) {
    this->modelName = "MetaScore";
    this->indexResult = true;   // src/Model.h:3163: ModelManager opens a bgzipped, tabix-indexed writer for this model
    id_ = MetaBatcher<DC>::instance().newFitterId();
    MetaBatcher<DC>::instance().attach();
  }
  ~MetaScoreTestB200() {
    drain(true);
    MetaBatcher<DC>::instance().detach();
  }
  virtual void reset() {
    BASE::reset();
    ticket_ = -1;
  }
  virtual int fit(DC* dc) {
    ticket_ = MetaBatcher<DC>::instance().submit(id_, dc, this->isBinaryOutcome());
    return ticket_ >= 0 ? 0 : -1;
  }
  virtual void writeHeader(FW*, const RES&) {}   // deferred: the header follows the null-model block (src/Model.h:3261-3281)
  virtual void writeOutput(FW* fp, const RES& siteInfo) {
    fp_ = fp;
    if (!header_) {
      site_header_ = siteInfo.joinHeader();
      header_ = true;
    }
    Pending p;
    p.ticket = ticket_;
    p.site = siteInfo.joinValue();
    pending_.push_back(p);
    if (MetaBatcher<DC>::instance().shouldFlush()) drain(false);
  }
  virtual void writeFootnote(FW* fp) {
    if (!fp_) fp_ = fp;
    drain(true);
  }
  // g_SummaryHeader of the reference (src/Main.cpp:775-780): its block opens the file and its covariate labels name the rows of
  // the null-model estimates (MetaScoreTest::writeSummaryAndHeader, src/Model.h:3283-3297).  Not owned.
  void setSummaryHeader(SummaryHook<FW>* h) { summary_ = h; }

 private:
  static std::string g(double v) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%g", v);
    return buf;
  }
  void printHeader() {
    MetaBatcher<DC>& b = MetaBatcher<DC>::instance();
    std::vector<double> beta, var;
    double sigma2 = 0;
    if (summary_) summary_->outputHeader(fp_);
    const std::vector<std::string>* labels = summary_ ? &summary_->getCovLabel() : NULL;
    if (b.nullModel(&beta, &var, &sigma2)) {   // MetaUnrelatedQtl::PrintNullModel, src/Model.h:3525-3541
      fp_->write("##NullModelEstimates\n");
      fp_->write("## - Name\tBeta\tSD\n");
      fp_->write(("## - Intercept\t" + g(beta[0]) + "\t" + g(var[0]) + "\n").c_str());
      for (size_t i = 1; i < beta.size(); ++i) {
        char nm[32];
        snprintf(nm, sizeof(nm), "Cov%d", (int)i);   // (no summary header recorded: placeholder names)
        if (labels && i - 1 >= labels->size()) break;   // PrintNullModel stops at the last label
        fp_->write((std::string("## - ") + (labels ? (*labels)[i - 1] : std::string(nm)) + "\t" + g(beta[i]) + "\t" + g(var[i]) + "\n").c_str());
      }
      fp_->write(b.binary() ? "## - Sigma2\tNA\tNA\n" : ("## - Sigma2\t" + g(sigma2) + "\tNA\n").c_str());
    }
    std::string h = site_header_ + "\tAF\tINFORMATIVE_ALT_AC\tCALL_RATE\tHWE_PVALUE\tN_REF\tN_HET\tN_ALT\tU_STAT\tSQRT_V_STAT\tALT_EFFSIZE";
    if (outputSE_) h += "\tALT_EFFSIZE_SE";
    h += "\tPVALUE\n";
    fp_->write(h.c_str());
  }
  void drain(bool final) {
    if (!fp_ || pending_.empty()) return;
    MetaBatcher<DC>& b = MetaBatcher<DC>::instance();
    if (!b.flush(final)) return;
    if (!printed_header_) {
      printHeader();
      printed_header_ = true;
    }
    size_t i = 0;
    for (; i < pending_.size(); ++i) {
      const typename MetaBatcher<DC>::Site* s = b.site(pending_[i].ticket, false);
      if (!s && pending_[i].ticket >= 0) break;
      std::string line = pending_[i].site + "\t";
      if (!s) {
        line += "NA\tNA\tNA\tNA\tNA\tNA\tNA\tNA\tNA\tNA";
        if (outputSE_) line += "\tNA";
        line += "\tNA";
      } else {
        const rvt_variant_result& r = s->r;
        char buf[160];
        if (!b.binary()) {
          snprintf(buf, sizeof(buf), "%d\t%d\t%d", r.n_ref, r.n_het, r.n_alt);
          line += g(r.af) + "\t" + g(r.ac) + "\t" + g(r.call_rate) + "\t" + g(r.hwe_p) + "\t" + buf + "\t";
        } else {   // all:case:control (src/Model.h:3300-3330); complete hard calls, so GenotypeCounter reduces to the counts
          const rvt_variant_cc& q = s->cc;
          double ac[2], af[2], cr[2];
          for (int w = 0; w < 2; ++w) {
            ac[w] = (double)q.n_het[w] + 2.0 * (double)q.n_alt[w];
            af[w] = q.n[w] ? 0.5 * ac[w] / (double)q.n[w] : -1.0;   // GenotypeCounter::getAF / getCallRate of an empty set
            cr[w] = q.n[w] ? 1.0 : 0.0;
          }
          line += g(r.af) + ":" + g(af[0]) + ":" + g(af[1]) + "\t" + g(r.ac) + ":" + g(ac[0]) + ":" + g(ac[1]) + "\t" + g(r.call_rate) + ":" + g(cr[0]) +
                  ":" + g(cr[1]) + "\t" + g(r.hwe_p) + ":" + g(q.hwe_p[0]) + ":" + g(q.hwe_p[1]) + "\t";
          snprintf(buf, sizeof(buf), "%d:%d:%d\t%d:%d:%d\t%d:%d:%d\t", r.n_ref, q.n_ref[0], q.n_ref[1], r.n_het, q.n_het[0], q.n_het[1], r.n_alt,
                   q.n_alt[0], q.n_alt[1]);
          line += buf;
        }
        if (r.ok) {
          line += g(r.U) + "\t" + g(r.sqrtV) + "\t" + g(r.effect);
          if (outputSE_) line += "\t" + ((r.sqrtV > 0.0) ? g(r.effect_se) : std::string("NA"));
          line += "\t" + g(r.pvalue);
        } else {
          line += "NA\tNA\tNA";
          if (outputSE_) line += "\tNA";
          line += "\tNA";
        }
      }
      line += "\n";
      fp_->write(line.c_str());
    }
    pending_.erase(pending_.begin(), pending_.begin() + i);
  }
  struct Pending {
    int ticket;
    std::string site;
  };
  std::string site_header_;
  SummaryHook<FW>* summary_ = NULL;
  bool outputSE_;
  int id_, ticket_;
  FW* fp_;
  bool header_, printed_header_ = false;
  std::vector<Pending> pending_;
};

template <class DC, class FW, class RES, class BASE = StandaloneBase>
class MetaCovTestB200 : public BASE {
 public:
  explicit MetaCovTestB200(int windowSize = 1000000) : ticket_(-1), fp_(NULL) {
    this->modelName = "MetaCov";
    this->indexResult = true;
    MetaBatcher<DC>& b = MetaBatcher<DC>::instance();
    id_ = b.newFitterId();
    b.attach();
    b.enableCov();
    b.setWindow(windowSize);
  }
  ~MetaCovTestB200() {   // MetaCovTest::~MetaCovTest prints what is still queued (src/Model.cpp:828-834)
    drain(true);
    MetaBatcher<DC>::instance().detach();
  }
  virtual void reset() {
    BASE::reset();
    ticket_ = -1;
  }
  virtual int fit(DC* dc) {
    ticket_ = MetaBatcher<DC>::instance().submit(id_, dc, this->isBinaryOutcome());
    return ticket_ >= 0 ? 0 : -1;
  }
  // MetaCovTest::writeHeader: the summary block, then the column names (src/Model.h:3944-3953)
  virtual void writeHeader(FW* fp, const RES&) {
    if (summary_) summary_->outputHeader(fp);
    fp->write("CHROM\tSTART_POS\tEND_POS\tNUM_MARKER\tMARKER_POS\tCOV\n");
  }
  void setSummaryHeader(SummaryHook<FW>* h) { summary_ = h; }
  virtual void writeOutput(FW* fp, const RES&) {
    fp_ = fp;
    pending_.push_back(ticket_);
    if (MetaBatcher<DC>::instance().shouldFlush()) drain(false);
  }
  virtual void writeFootnote(FW* fp) {
    if (!fp_) fp_ = fp;
    drain(true);
  }

 private:
  void drain(bool final) {
    if (!fp_ || pending_.empty()) return;
    MetaBatcher<DC>& b = MetaBatcher<DC>::instance();
    if (!b.flush(final)) return;
    size_t i = 0;
    for (; i < pending_.size(); ++i) {
      if (pending_[i] < 0) continue;
      const typename MetaBatcher<DC>::Site* s = b.site(pending_[i], true);
      if (!s) break;   // its window is still open
      if (!s->cov_line.empty()) fp_->write((s->cov_line + "\n").c_str());
    }
    pending_.erase(pending_.begin(), pending_.begin() + i);
  }
  int id_, ticket_;
  FW* fp_;
  SummaryHook<FW>* summary_ = NULL;
  std::vector<int> pending_;
};

}  // namespace rvtb200
#endif  // RVT_META_FITTERS_H_
