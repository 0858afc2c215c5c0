// oracle/ref_model_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// Drives the REFERENCE's own model layer the way src/Main.cpp does: DataConsolidator::consolidate(), then
// ModelFitter::reset() / fit(&dc) / writeOutput() of SkatTest, SkatOTest, CMCTest, ZegginiTest (gene loop,
// Main.cpp:1221-1254) and of MetaScoreTest, MetaCovTest (single-variant loop, Main.cpp:1092-1147), writing the
// reference's own `.assoc` text.  src/Model.cpp (+ src/Model.h), src/DataConsolidator.cpp and the files they need
// are compiled UNMODIFIED from /root/reference into oracle/_ref/libmodel_ref.so (oracle/Makefile) against
// oracle/eigen_standin and the vendored GSL.  Nothing from the reference is copied here: this file calls its public
// classes and DEFINES the few external symbols whose own translation units cannot be built in this image:
//   * FileWriter's constructors (base/IO.cpp needs samtools' bgzf + bzip2): a plain stdio text writer behind the
//     reference's AbstractFileWriter interface (base/IO.h:172-180) -- the formatting code is the reference's;
//   * BoltLMM, KinshipHolder, Plink{Input,Output}File members, BufferedReader: paths not exercised here
//     (no kinship, no --boltPlink); they abort if ever reached;
//   * the globals of src/Main.cpp that the model layer reads (logger, VERSION, g_SummaryHeader, one FLAG).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "third/eigen/Eigen/Core"

#include "base/IO.h"
#include "base/Logger.h"
#include "base/KinshipHolder.h"
#include "base/ParRegion.h"
#include "libVcf/PlinkInputFile.h"
#include "libVcf/PlinkOutputFile.h"
#include "regression/BoltLMM.h"
#include "src/DataConsolidator.h"
#include "src/Model.h"
#include "src/ModelFitter.h"
#include "src/ModelParser.h"
#include "src/Result.h"
#include "src/Summary.h"
#include "base/SimpleMatrix.h"

// ---------------------------------------------------------------- globals of src/Main.cpp
Logger* logger = NULL;
const char* VERSION = "reference-build-under-test";
SummaryHeader* g_SummaryHeader = NULL;
namespace parameter {
bool FLAG_boltPlinkNoCheck = false;
}

// ---------------------------------------------------------------- FileWriter on stdio
namespace {
class StdioWriter : public AbstractFileWriter {
 public:
  StdioWriter() : f_(NULL) {}
  int open(const char* fn, bool append = false) {
    f_ = fopen(fn, append ? "a" : "w");
    return f_ ? 0 : -1;
  }
  void close() {
    if (f_) fclose(f_);
    f_ = NULL;
  }
  int write(const char* s) { return fputs(s, f_) >= 0 ? (int)strlen(s) : -1; }
  int writeLine(const char* s) {
    int r = write(s);
    fputc('\n', f_);
    return r + 1;
  }
  ~StdioWriter() { close(); }

 private:
  FILE* f_;
};
[[noreturn]] void unreachable(const char* what) {
  fprintf(stderr, "ref_model_shim: %s is not available in this build\n", what);
  abort();
}
}  // namespace
AbstractFileWriter::~AbstractFileWriter() {}
FileWriter::FileWriter(const std::string& fileName, bool append) {
  StdioWriter* w = new StdioWriter;
  if (w->open(fileName.c_str(), append)) unreachable("output file");
  this->fp = w;
  this->fpRaw = NULL;
  this->createBuffer();
}
FileWriter::FileWriter(const std::string& fileName, FileType) : FileWriter(fileName, false) {}
bool fileExists(std::string fn) {
  FILE* f = fopen(fn.c_str(), "r");
  if (f) fclose(f);
  return f != NULL;
}
BufferedReader::BufferedReader(const char*, int) { unreachable("BufferedReader"); }
int BufferedReader::readLineBySep(std::vector<std::string>*, const char*) { unreachable("BufferedReader"); }
bool BufferedReader::isEof() { return true; }
void BufferedReader::close() {}
int BufferedReader::getc() { return EOF; }
int BufferedReader::read(void*, int) { return 0; }

// ---------------------------------------------------------------- unreachable subsystems
BoltLMM::BoltLMM() : impl_(NULL) {}
BoltLMM::~BoltLMM() {}
int BoltLMM::FitNullModel(const std::string&, const Matrix*) { unreachable("BoltLMM"); }
void BoltLMM::GetCovXX(const FloatMatrixRef&, const FloatMatrixRef&, float*) { unreachable("BoltLMM"); }
int BoltLMM::TestCovariate(const Matrix&) { unreachable("BoltLMM"); }
double BoltLMM::GetAF() { unreachable("BoltLMM"); }
double BoltLMM::GetU() { unreachable("BoltLMM"); }
double BoltLMM::GetV() { unreachable("BoltLMM"); }
double BoltLMM::GetEffect() { unreachable("BoltLMM"); }
double BoltLMM::GetPvalue() { unreachable("BoltLMM"); }
void BoltLMM::enableBinaryMode() { unreachable("BoltLMM"); }
KinshipHolder::KinshipHolder() { this->matK = this->matS = this->matU = NULL; this->pSample = NULL; this->loaded = false; }
KinshipHolder::~KinshipHolder() {}
int KinshipHolder::setSample(const std::vector<std::string>&) { unreachable("KinshipHolder"); }
int KinshipHolder::setFile(const std::string&) { unreachable("KinshipHolder"); }
int KinshipHolder::setEigenFile(const std::string&) { unreachable("KinshipHolder"); }
int KinshipHolder::load() { unreachable("KinshipHolder"); }
int PlinkInputFile::calculateMAF(std::vector<double>*) { unreachable("PlinkInputFile"); }
int PlinkInputFile::calculateMissing(std::vector<double>*, std::vector<double>*) { unreachable("PlinkInputFile"); }
void PlinkOutputFile::init(const char*) { unreachable("PlinkOutputFile"); }
int PlinkOutputFile::extractFAMWithPhenotype(PlinkInputFile&, const std::vector<int>&, const SimpleMatrix&) { unreachable("PlinkOutputFile"); }
int PlinkOutputFile::extractBIM(PlinkInputFile&, const std::vector<int>&) { unreachable("PlinkOutputFile"); }
int PlinkOutputFile::extractBED(PlinkInputFile&, const std::vector<int>&, const std::vector<int>&) { unreachable("PlinkOutputFile"); }

// ---------------------------------------------------------------- the driver
namespace {
void fill(const double* p, int r, int c, Matrix* m) {
  m->Dimension(r, c);
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i) (*m)(i, j) = p[(size_t)j * r + i];
}
void ensure_logger(const char* prefix) {
  if (!logger) logger = new Logger((std::string(prefix) + ".log").c_str());
}
// g_SummaryHeader as src/Main.cpp:775-780 sets it up (trait and covariate summaries of the output header)
void record_summary(int N, int n_cov, const double* cov, const double* pheno) {
  delete g_SummaryHeader;
  g_SummaryHeader = new SummaryHeader;
  SimpleMatrix m(N, n_cov);
  std::vector<std::string> names;
  for (int j = 0; j < n_cov; ++j) {
    for (int i = 0; i < N; ++i) m[i][j] = cov[(size_t)j * N + i];
    char b[32];
    snprintf(b, sizeof b, "cov%d", j);
    names.push_back(b);
  }
  m.setColName(names);
  g_SummaryHeader->recordCovariate(m);
  g_SummaryHeader->recordPhenotype("Trait", std::vector<double>(pheno, pheno + N));
}
}  // namespace

extern "C" {
// Gene loop of Main.cpp:1221-1254 over `n_genes` genes that share N samples.  G holds the genes back to back, gene k
// being N x M[k] column-major raw genotypes (missing < 0) as GenotypeExtractor hands them over; cov is N x n_cov WITHOUT
// the intercept (the fitters add it, src/ModelUtil.h:102-130).  Writes <prefix>.{Skat,SkatO,CMC,Zeggini}.assoc.
int ref_run_gene_models(int N, int n_genes, const int* M, const double* G, int n_cov, const double* cov,
                        const double* pheno, int n_perm, double alpha, int binary, const char* prefix) {
  ensure_logger(prefix);
  record_summary(N, n_cov, cov, pheno);
  Matrix phenotypeMatrix, covariate;
  fill(pheno, N, 1, &phenotypeMatrix);
  fill(cov, N, n_cov, &covariate);
  for (int j = 0; j < n_cov; ++j) {
    char b[32];
    snprintf(b, sizeof b, "cov%d", j);
    covariate.SetColumnLabel(j, b);
  }
  ParRegion par;
  DataConsolidator dc;
  dc.setStrategy(DataConsolidator::IMPUTE_MEAN);
  dc.setParRegion(&par);
  std::vector<ModelFitter*> model;
  model.push_back(new SkatTest(n_perm, alpha, 1.0, 25.0));
  model.push_back(new SkatOTest(1.0, 25.0));
  model.push_back(new CMCTest);
  model.push_back(new ZegginiTest);
  std::vector<FileWriter*> fOuts;
  static const char* kSpec[] = {"skat", "skato", "cmc", "zeggini"};
  for (size_t m = 0; m < model.size(); ++m) {
    ModelParser parser;  // src/ModelManager.cpp:35-40, :275 -- every model gets setParameter(parser)
    parser.parse(kSpec[m]);
    model[m]->setParameter(parser);
    if (binary) model[m]->setBinaryOutcome(); else model[m]->setQuantitativeOutcome();
    fOuts.push_back(new FileWriter(std::string(prefix) + "." + model[m]->getModelName() + ".assoc"));
  }
  Result& buf = dc.getResult();
  buf.addHeader("Gene");
  buf.addHeader("RANGE");
  buf.addHeader("N_INFORMATIVE");
  buf.addHeader("NumVar");
  buf.addHeader("NumPolyVar");
  for (size_t m = 0; m < model.size(); ++m) model[m]->writeHeader(fOuts[m], buf);
  Matrix& genotype = dc.getOriginalGenotype();
  size_t off = 0;
  for (int k = 0; k < n_genes; ++k) {
    fill(G + off, N, M[k], &genotype);
    off += (size_t)N * M[k];
    std::vector<GenotypeCounter> counter(M[k]);  // GenotypeExtractor fills these while reading (src/GenotypeExtractor.cpp)
    for (int j = 0; j < M[k]; ++j) {
      char b[32];
      snprintf(b, sizeof b, "1:%d", 1000 * k + j + 1);
      genotype.SetColumnLabel(j, b);
      for (int i = 0; i < N; ++i) counter[j].add(genotype(i, j));
    }
    dc.setGenotypeCounter(counter);
    buf.clearValue();
    dc.consolidate(phenotypeMatrix, covariate, genotype);
    char name[32];
    snprintf(name, sizeof name, "GENE%d", k);
    buf.updateValue("Gene", name);
    buf.updateValue("RANGE", "1:1-2");
    buf.updateValue("N_INFORMATIVE", genotype.rows);
    buf.updateValue("NumVar", genotype.cols);
    buf.updateValue("NumPolyVar", dc.getFlippedToMinorPolymorphicGenotype().cols);
    for (size_t m = 0; m < model.size(); ++m) {
      model[m]->reset();
      model[m]->fit(&dc);
      model[m]->writeOutput(fOuts[m], buf);
    }
  }
  for (size_t m = 0; m < model.size(); ++m) {
    model[m]->writeFootnote(fOuts[m]);
    delete model[m];
    delete fOuts[m];
  }
  return 0;
}

// Single-variant loop of Main.cpp:1092-1147 with MetaScoreTest + MetaCovTest(window): variant j is column j of G
// (N x n_var column-major raw genotypes) at 1:pos[j].  Writes <prefix>.MetaScore.assoc and <prefix>.MetaCov.assoc
// (the covariance lines are flushed by MetaCovTest's destructor, src/Model.cpp:828-834).
int ref_run_meta_models(int N, int n_var, const double* G, const int* pos, int n_cov, const double* cov,
                        const double* pheno, int window, const char* prefix) {
  ensure_logger(prefix);
  record_summary(N, n_cov, cov, pheno);
  Matrix phenotypeMatrix, covariate;
  fill(pheno, N, 1, &phenotypeMatrix);
  fill(cov, N, n_cov, &covariate);
  for (int j = 0; j < n_cov; ++j) {
    char b[32];
    snprintf(b, sizeof b, "cov%d", j);
    covariate.SetColumnLabel(j, b);
  }
  ParRegion par;
  DataConsolidator dc;
  dc.setStrategy(DataConsolidator::IMPUTE_MEAN);
  dc.setParRegion(&par);
  std::vector<ModelFitter*> model;
  model.push_back(new MetaScoreTest);
  model.push_back(new MetaCovTest(window));
  std::vector<FileWriter*> fOuts;
  static const char* kSpec[] = {"score", "cov"};
  for (size_t m = 0; m < model.size(); ++m) {
    ModelParser parser;
    parser.parse(kSpec[m]);
    model[m]->setParameter(parser);
    model[m]->setQuantitativeOutcome();
    fOuts.push_back(new FileWriter(std::string(prefix) + "." + model[m]->getModelName() + ".assoc"));
  }
  Result& buf = dc.getResult();
  buf.addHeader("CHROM");
  buf.addHeader("POS");
  buf.addHeader("REF");
  buf.addHeader("ALT");
  buf.addHeader("N_INFORMATIVE");
  for (size_t m = 0; m < model.size(); ++m) model[m]->writeHeader(fOuts[m], buf);
  Matrix& genotype = dc.getOriginalGenotype();
  for (int j = 0; j < n_var; ++j) {
    fill(G + (size_t)j * N, N, 1, &genotype);
    char b[32];
    snprintf(b, sizeof b, "1:%d", pos[j]);
    genotype.SetColumnLabel(0, b);
    std::vector<GenotypeCounter> counter(1);
    for (int i = 0; i < N; ++i) counter[0].add(genotype(i, 0));
    dc.setGenotypeCounter(counter);
    buf.clearValue();
    buf.updateValue("CHROM", "1");
    buf.updateValue("POS", pos[j]);
    buf.updateValue("REF", "A");
    buf.updateValue("ALT", "C");
    dc.consolidate(phenotypeMatrix, covariate, genotype);
    buf.updateValue("N_INFORMATIVE", toString(genotype.rows));
    for (size_t m = 0; m < model.size(); ++m) {
      model[m]->reset();
      model[m]->fit(&dc);
      model[m]->writeOutput(fOuts[m], buf);
    }
  }
  for (size_t m = 0; m < model.size(); ++m) {
    model[m]->writeFootnote(fOuts[m]);
    delete model[m];  // MetaCovTest flushes its queue here, so the writer must still be open
  }
  for (size_t m = 0; m < model.size(); ++m) delete fOuts[m];
  return 0;
}
}
