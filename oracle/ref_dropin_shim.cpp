// oracle/ref_dropin_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// The drop-in, literally: the REFERENCE's own ModelManager (src/ModelManager.cpp with the two registration lines of
// rvtests_b200/host/ModelB200.h applied to a scratch copy by oracle/patch_model_manager.py) creates the models by name
// -- `--kernel skat[..],skato --burden cmc,zeggini` -- and the gene loop of src/Main.cpp:1221-1254 drives whatever it
// created through the ModelFitter interface on a real DataConsolidator.  With RVTESTS_B200 unset those are the
// reference's SkatTest / SkatOTest / CMCTest / ZegginiTest; with RVTESTS_B200=1 they are the B200 adapters
// (true ModelFitter subclasses) calling librvtests_b200.so.  Same loop, same DataConsolidator, same writers: the test
// diffs the two sets of `.assoc` files.  Links against oracle/_ref/libmodel_ref.so (the reference's model layer and the
// stubs of what cannot be built in this image) and rvtests_b200/librvtests_b200.so.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "base/IO.h"
#include "base/Logger.h"
#include "base/ParRegion.h"
#include "base/SimpleMatrix.h"
#include "src/DataConsolidator.h"
#include "src/GenotypeCounter.h"
#include "src/ModelFitter.h"
#include "src/ModelManager.h"
#include "src/Result.h"
#include "src/Summary.h"
#include "src/TabixUtil.h"

#include "ModelB200.h"

extern Logger* logger;
extern SummaryHeader* g_SummaryHeader;
// a global of src/Main.cpp that src/Model.h reads (single-variant Wald output); not defined by libmodel_ref.so
namespace parameter {
bool FLAG_hideCovar = false;
}

// src/ModelManager.cpp:318-327 indexes bgzipped outputs with tabix (src/TabixUtil.cpp needs htslib): not reached here,
// none of the four models asks for an indexed result
int tabixIndexFile(const std::string&, int, char, int, int, int) { return 0; }

namespace {
void fill(const double* p, int r, int c, Matrix* m) {
  m->Dimension(r, c);
  for (int j = 0; j < c; ++j)
    for (int i = 0; i < r; ++i) (*m)(i, j) = p[(size_t)j * r + i];
}
}  // namespace

extern "C" {
// Same inputs as ref_run_gene_models (ref_model_shim.cpp).  kernel / burden: the reference's own model lists, e.g.
// "skat[nPerm=0],skato" and "cmc,zeggini".  use_b200 sets / clears RVTESTS_B200 for the duration of the call.
// Writes <prefix>.<ModelName>.assoc through ModelManager's own writers.
int dropin_run_gene_models(int N, int n_genes, const int* M, const double* G, int n_cov, const double* cov, const double* pheno,
                           const char* kernel, const char* burden, int binary, int use_b200, int batch, const char* prefix) {
  if (!logger) logger = new Logger((std::string(prefix) + ".log").c_str());
  if (!g_SummaryHeader) g_SummaryHeader = new SummaryHeader;
  if (use_b200)
    setenv("RVTESTS_B200", "1", 1);
  else
    unsetenv("RVTESTS_B200");
  if (use_b200 && batch > 0) rvtb200::GeneBatcher<DataConsolidator>::instance().setBatch(batch);
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
  {
    ModelManager modelManager(prefix);   // src/Main.cpp:782-803
    if (binary)
      modelManager.setBinaryOutcome();
    else
      modelManager.setQuantitativeOutcome();
    modelManager.create("burden", burden);
    modelManager.create("kernel", kernel);
    const std::vector<ModelFitter*>& model = modelManager.getModel();
    const std::vector<FileWriter*>& fOuts = modelManager.getResultFile();
    const size_t numModel = model.size();
    Result& buf = dc.getResult();
    buf.addHeader("Gene");
    buf.addHeader("RANGE");
    buf.addHeader("N_INFORMATIVE");
    buf.addHeader("NumVar");
    buf.addHeader("NumPolyVar");
    for (size_t m = 0; m < numModel; m++) model[m]->writeHeader(fOuts[m], buf);
    Matrix& genotype = dc.getOriginalGenotype();
    size_t off = 0;
    for (int k = 0; k < n_genes; ++k) {
      fill(G + off, N, M[k], &genotype);
      off += (size_t)N * M[k];
      std::vector<GenotypeCounter> counter(M[k]);
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
      for (size_t m = 0; m != numModel; m++) {
        model[m]->reset();
        model[m]->fit(&dc);
        model[m]->writeOutput(fOuts[m], buf);
      }
    }
  }   // ModelManager::close: writeFootnote + delete models, then the writers
  return 0;
}

// Single-variant loop of src/Main.cpp:1092-1147 through ModelManager::create("meta", "score[se],cov[windowSize=..]"): variant j
// is column j of G (N x n_var column-major raw genotypes) at 1:pos[j].  Writes <prefix>.MetaScore.assoc.gz and
// <prefix>.MetaCov.assoc.gz -- plain text here: the writer behind FileWriter(.., BGZIP) is the stdio stand-in of
// ref_model_shim.cpp, and the tabix step of ModelManager::close is stubbed.
int dropin_run_meta_models(int N, int n_var, const double* G, const int* pos, int n_cov, const double* cov, const double* pheno,
                           const char* meta, int use_b200, int segment, const char* prefix, int binary) {
  if (!logger) logger = new Logger((std::string(prefix) + ".log").c_str());
  {
    delete g_SummaryHeader;   // the trait / covariate summaries MetaScoreTest prints in its header (src/Main.cpp:775-780)
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
  if (use_b200)
    setenv("RVTESTS_B200", "1", 1);
  else
    unsetenv("RVTESTS_B200");
  if (use_b200 && segment > 0) rvtb200::MetaBatcher<DataConsolidator>::instance().setSegment(segment);
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
  {
    ModelManager modelManager(prefix);
    if (binary)
      modelManager.setBinaryOutcome();   // MetaUnrelatedBinary / MetaCovUnrelatedBinary (phenotype already 0 / 1)
    else
      modelManager.setQuantitativeOutcome();
    modelManager.create("meta", meta);
    const std::vector<ModelFitter*>& model = modelManager.getModel();
    const std::vector<FileWriter*>& fOuts = modelManager.getResultFile();
    const size_t numModel = model.size();
    Result& buf = dc.getResult();
    buf.addHeader("CHROM");
    buf.addHeader("POS");
    buf.addHeader("REF");
    buf.addHeader("ALT");
    buf.addHeader("N_INFORMATIVE");
    for (size_t m = 0; m < numModel; m++) model[m]->writeHeader(fOuts[m], buf);
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
      for (size_t m = 0; m != numModel; m++) {
        model[m]->reset();
        model[m]->fit(&dc);
        model[m]->writeOutput(fOuts[m], buf);
      }
    }
  }
  return 0;
}
}
