// shim.h -- minimal stand-ins for the reference types the adapters touch, used ONLY by the adapter
// test of this repository (in rvtests itself the real headers are used):
//   Matrix            base/MathMatrix.h:33-111   (rows, cols, std::vector<double> data, column-major)
//   DataConsolidator  src/DataConsolidator.h:126-137,185-186,223-224  (the accessors fit() pulls)
//   FileWriter        base/IO.h  (write(const char*))
//   Result            src/Result.h:20-249 (writeHeaderTab / joinValue of the site columns)
#ifndef RVT_SHIM_H_
#define RVT_SHIM_H_
#include <string.h>

#include <string>
#include <vector>

namespace shim {

struct Matrix {
  int rows, cols;
  std::vector<double> data;  // column-major
  Matrix() : rows(0), cols(0) {}
  void Dimension(int r, int c) {
    rows = r;
    cols = c;
    data.assign((size_t)r * c, 0.0);
  }
  double& operator()(int i, int j) { return data[(size_t)j * rows + i]; }
  const double& operator()(int i, int j) const { return data[(size_t)j * rows + i]; }
};

struct FileWriter {
  std::string out;
  int write(const char* s) {
    out += s;
    return (int)strlen(s);
  }
};

struct Result {
  std::vector<std::string> keys, values;
  void writeHeaderTab(FileWriter* fp) const {
    for (size_t i = 0; i < keys.size(); ++i) fp->write((keys[i] + "\t").c_str());
  }
  // siteInfo["CHROM"], siteInfo["POS"] (src/Result.h:218-238)
  const std::string& operator[](const std::string& key) const {
    static const std::string na = "NA";
    for (size_t i = 0; i < keys.size(); ++i)
      if (keys[i] == key) return values[i];
    return na;
  }
  std::string joinHeader() const {
    std::string s;
    for (size_t i = 0; i < keys.size(); ++i) s += (i ? "\t" : "") + keys[i];
    return s;
  }
  std::string joinValue() const {
    std::string s;
    for (size_t i = 0; i < values.size(); ++i) s += (i ? "\t" : "") + values[i];
    return s;
  }
};

struct DataConsolidator {
  Matrix pheno, cov, geno;
  std::vector<double> af;
  bool phenoUpdated, covUpdated;
  DataConsolidator() : phenoUpdated(true), covUpdated(true) {}
  const Matrix& getPhenotype() const { return pheno; }
  const Matrix& getCovariate() const { return cov; }
  const Matrix& getGenotype() const { return geno; }
  double getMarkerFrequency(int col) const { return af[col]; }
  bool isPhenotypeUpdated() const { return phenoUpdated; }
  bool isCovariateUpdated() const { return covUpdated; }
  Result site;                                   // dc->getResult(): the site columns of the current variant
  Result& getResult() { return site; }
};

}  // namespace shim
#endif
