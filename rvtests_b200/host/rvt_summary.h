// rvt_summary.h -- the "##" block a `--meta score` file starts with: SummaryHeader (src/Summary.h:24-185), printed by
// MetaScoreTest::writeSummaryAndHeader before the null-model estimates (src/Model.h:3283-3297).
//
//   Summary::add          min, v[(int)(n * 0.25)], v[(int)(n * 0.5)], v[(int)(n * 0.75)], max of the SORTED values, the mean, and
//                         the sample standard deviation (n - 1), printed squared under "variance" (src/Summary.h:27-42, 135-139)
//   outputHeader          ##ProgramName .. ##InverseNormal, ##TraitSummary + one line per trait, ##Covariates= + ##CovariateSummary
//                         when there are covariates, ##ResidualModelEstimates when --useResidualAsPhenotype recorded one
// Inside rvtests the adapters print the reference's own g_SummaryHeader (ModelB200.h installs that hook); a host without the
// reference's base/ records its trait and covariates here.  SummaryHook<FW> is what MetaScoreTestB200 consumes.
#ifndef RVT_SUMMARY_H_
#define RVT_SUMMARY_H_

#include <math.h>
#include <stdio.h>

#include <algorithm>
#include <string>
#include <vector>

namespace rvtb200 {

template <class FW>
struct SummaryHook {
  virtual ~SummaryHook() {}
  virtual void outputHeader(FW* fp) = 0;
  virtual const std::vector<std::string>& getCovLabel() const = 0;
};

struct ColumnSummary {
  double min, q1, median, q3, max, mean, sd;
  int n;
  ColumnSummary() : min(0), q1(0), median(0), q3(0), max(0), mean(0), sd(0), n(0) {}
  void add(const std::vector<double>& v) {
    n = (int)v.size();
    if (n == 0) return;
    std::vector<double> t = v;
    std::sort(t.begin(), t.end());
    min = t[0];
    q1 = t[(size_t)(n * 0.25)];
    median = t[(size_t)(n * 0.5)];
    q3 = t[(size_t)(n * 0.75)];
    max = t[n - 1];
    double s = 0.0;   // calculateMean / calculateSampleSD, base/CommonFunction.h
    for (int i = 0; i < n; ++i) s += v[i];
    mean = s / n;
    double ss = 0.0;
    for (int i = 0; i < n; ++i) ss += (v[i] - mean) * (v[i] - mean);
    sd = n > 1 ? sqrt(ss / (n - 1)) : 0.0;
  }
};

template <class FW>
class SummaryHeaderB200 : public SummaryHook<FW> {
 public:
  explicit SummaryHeaderB200(const char* version = "rvtests_b200") : version_(version), inverse_normal_(false) {}
  void recordPhenotype(const char* label, const std::vector<double>& pheno) {
    pheno_label_.push_back(label);
    ColumnSummary s;
    s.add(pheno);
    pheno_.push_back(s);
  }
  void setInverseNormalize(bool b) { inverse_normal_ = b; }
  void recordCovariate(const std::vector<std::string>& labels, const std::vector<std::vector<double> >& columns) {
    cov_label_ = labels;
    cov_.clear();
    for (size_t j = 0; j < columns.size(); ++j) {
      ColumnSummary s;
      s.add(columns[j]);
      cov_.push_back(s);
    }
  }
  // rows "name beta sd" of --useResidualAsPhenotype (non-finite values print NA)
  void recordEstimation(const std::vector<std::string>& names, const std::vector<double>& beta, const std::vector<double>& sd) {
    est_name_ = names;
    est_beta_ = beta;
    est_sd_ = sd;
  }
  virtual void outputHeader(FW* fp) {
    const int n = pheno_.empty() ? 0 : pheno_[0].n;
    std::string s = "##ProgramName=Rvtests\n##Version=" + version_ + "\n";
    static const char* keys[6] = {"Samples", "AnalyzedSamples", "Families", "AnalyzedFamilies", "Founders", "AnalyzedFounders"};
    char buf[512];
    for (int k = 0; k < 6; ++k) {
      snprintf(buf, sizeof(buf), "##%s=%d\n", keys[k], n);
      s += buf;
    }
    s += std::string("##InverseNormal=") + (inverse_normal_ ? "ON" : "OFF") + "\n";
    s += "##TraitSummary\tmin\t25th\tmedian\t75th\tmax\tmean\tvariance\n";
    for (size_t i = 0; i < pheno_.size(); ++i) s += line(pheno_label_[i], pheno_[i]);
    if (!cov_.empty()) {
      s += "##Covariates=";
      for (size_t i = 0; i < cov_.size(); ++i) s += (i ? "," : "") + cov_label_[i];
      s += "\n##CovariateSummary\tmin\t25th\tmedian\t75th\tmax\tmean\tvariance\n";
      for (size_t i = 0; i < cov_.size(); ++i) s += line(cov_label_[i], cov_[i]);
    }
    if (!est_name_.empty()) {
      s += "##ResidualModelEstimates\n## - Name\tBeta\tSD\n";
      for (size_t i = 0; i < est_name_.size(); ++i) s += "## - " + est_name_[i] + "\t" + num(est_beta_[i]) + "\t" + num(est_sd_[i]) + "\n";
    }
    fp->write(s.c_str());
  }
  virtual const std::vector<std::string>& getCovLabel() const { return cov_label_; }

 private:
  static std::string num(double v) {
    if (!std::isfinite(v)) return "NA";
    char buf[64];
    snprintf(buf, sizeof(buf), "%g", v);
    return buf;
  }
  static std::string line(const std::string& label, const ColumnSummary& c) {
    char buf[512];
    snprintf(buf, sizeof(buf), "##%s\t%g\t%g\t%g\t%g\t%g\t%g\t%g\n", label.c_str(), c.min, c.q1, c.median, c.q3, c.max, c.mean, c.sd * c.sd);
    return buf;
  }
  std::string version_;
  bool inverse_normal_;
  std::vector<std::string> pheno_label_, cov_label_, est_name_;
  std::vector<ColumnSummary> pheno_, cov_;
  std::vector<double> est_beta_, est_sd_;
};

}  // namespace rvtb200
#endif  // RVT_SUMMARY_H_
