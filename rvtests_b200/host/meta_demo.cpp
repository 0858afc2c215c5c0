// meta_demo.cpp -- drives the `--meta score,cov` adapters (rvt_meta_fitters.h) like the reference's single-variant loop
// (src/Main.cpp:1084-1117: for each variant { consolidate; for each model { reset; fit; writeOutput } }; then
// ModelManager::close -> writeFootnote, models deleted before the writers) and prints both .assoc tables.
//   input: int32 N, C1, nVar; double y[N]; double cov[N*C1] (col-major); per variant: int32 chrom, pos; double g[N]
//   usage: meta_demo problem.bin segment window [prefix [binary]]
// With a prefix the tables are ALSO written the way ModelManager writes a model with needToIndexResult()
// (src/ModelManager.cpp:285-327): "<prefix>.MetaScore.assoc.gz" / "<prefix>.MetaCov.assoc.gz", bgzipped, and tabix-indexed
// when the writers close (rvt_bgzf.h).  binary = 1: setBinaryOutcome() on both models (case / control phenotype, 0 / 1).
#include <stdio.h>
#include <stdlib.h>

#include "rvt_bgzf.h"
#include "rvt_meta_fitters.h"
#include "shim.h"

struct DemoWriter : shim::FileWriter {   // FileWriter(fn, BGZIP) of base/IO.h next to the in-memory copy the tests read
  rvtb200::IndexedAssocWriter gz;
  bool on = false;
  int write(const char* s) {
    if (on) gz.write(s);
    return shim::FileWriter::write(s);
  }
};
typedef rvtb200::MetaScoreTestB200<shim::DataConsolidator, DemoWriter, shim::Result> MetaScoreTest;
typedef rvtb200::MetaCovTestB200<shim::DataConsolidator, DemoWriter, shim::Result> MetaCovTest;

template <class T>
static void rd(FILE* f, T* p, size_t n) {
  if (fread(p, sizeof(T), n, f) != n) {
    fprintf(stderr, "short read\n");
    exit(2);
  }
}

int main(int argc, char** argv) {
  if (argc < 4) return 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  const int segment = atoi(argv[2]), window = atoi(argv[3]);
  int N, C1, nVar;
  rd(f, &N, 1);
  rd(f, &C1, 1);
  rd(f, &nVar, 1);
  shim::DataConsolidator dc;
  dc.pheno.Dimension(N, 1);
  rd(f, &dc.pheno.data[0], N);
  dc.cov.Dimension(N, C1);
  if (C1) rd(f, &dc.cov.data[0], (size_t)N * C1);
  rvtb200::MetaBatcher<shim::DataConsolidator>::instance().setSegment(segment);
  DemoWriter fw[2];
  if (argc > 4) {
    const std::string prefix = argv[4];
    fw[0].on = fw[0].gz.open((prefix + ".MetaScore.assoc.gz").c_str());
    fw[1].on = fw[1].gz.open((prefix + ".MetaCov.assoc.gz").c_str());
    if (!fw[0].on || !fw[1].on) return 3;
  }
  int rc_close = 0;
  {
    MetaScoreTest score(true);   // --meta score[se]
    MetaCovTest cov(window);
    if (argc > 5 && atoi(argv[5])) {
      score.setBinaryOutcome();
      cov.setBinaryOutcome();
    }
    const char* keys[5] = {"CHROM", "POS", "REF", "ALT", "N_INFORMATIVE"};
    for (int k = 0; k < 5; ++k) dc.site.keys.push_back(keys[k]);
    dc.site.values.resize(5);
    score.writeHeader(&fw[0], dc.site);
    cov.writeHeader(&fw[1], dc.site);
    dc.geno.Dimension(N, 1);
    for (int v = 0; v < nVar; ++v) {
      int chrom, pos;
      rd(f, &chrom, 1);
      rd(f, &pos, 1);
      rd(f, &dc.geno.data[0], N);
      char buf[32];
      snprintf(buf, sizeof(buf), "%d", chrom);
      dc.site.values[0] = buf;
      snprintf(buf, sizeof(buf), "%d", pos);
      dc.site.values[1] = buf;
      dc.site.values[2] = "A";
      dc.site.values[3] = "G";
      snprintf(buf, sizeof(buf), "%d", N);
      dc.site.values[4] = buf;
      score.reset(); score.fit(&dc); score.writeOutput(&fw[0], dc.site);
      cov.reset(); cov.fit(&dc); cov.writeOutput(&fw[1], dc.site);
      dc.phenoUpdated = dc.covUpdated = false;
    }
    score.writeFootnote(&fw[0]);
    cov.writeFootnote(&fw[1]);
  }
  // ModelManager::close: the writers close after the models are gone, then createIndex()
  for (int k = 0; k < 2; ++k)
    if (fw[k].on && fw[k].gz.close() != 0) rc_close = 4;
  printf("#MetaScore\n%s#MetaCov\n%s", fw[0].out.c_str(), fw[1].out.c_str());
  fclose(f);
  return rc_close;
}
