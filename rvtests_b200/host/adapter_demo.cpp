// adapter_demo.cpp -- drives the ModelFitter-shaped adapters exactly like the reference's gene loop
// (src/Main.cpp:1221-1254: for each gene { consolidate; for each model { reset; fit; writeOutput } };
// ModelManager::close -> writeFootnote) on a problem read from a small binary file, and prints the
// four .assoc tables.  Built and run by tests/test_gpu_adapters.py on the GPU box.
//   input: int32 N, C1 (covariates w/o intercept), nGenes; double y[N]; double cov[N*C1] (col-major);
//          per gene: int32 M; double G[N*M] (col-major); double af[M]
#include <stdio.h>
#include <stdlib.h>

#include "rvt_fitters.h"
#include "shim.h"

typedef rvtb200::SkatTestB200<shim::DataConsolidator, shim::FileWriter, shim::Result> SkatTest;
typedef rvtb200::SkatOTestB200<shim::DataConsolidator, shim::FileWriter, shim::Result> SkatOTest;
typedef rvtb200::CMCTestB200<shim::DataConsolidator, shim::FileWriter, shim::Result> CMCTest;
typedef rvtb200::ZegginiTestB200<shim::DataConsolidator, shim::FileWriter, shim::Result> ZegginiTest;

template <class T>
static void rd(FILE* f, T* p, size_t n) {
  if (fread(p, sizeof(T), n, f) != n) {
    fprintf(stderr, "short read\n");
    exit(2);
  }
}

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  int batch = atoi(argv[2]);
  int N, C1, nGenes;
  rd(f, &N, 1);
  rd(f, &C1, 1);
  rd(f, &nGenes, 1);
  shim::DataConsolidator dc;
  dc.pheno.Dimension(N, 1);
  rd(f, &dc.pheno.data[0], N);
  dc.cov.Dimension(N, C1);
  if (C1) rd(f, &dc.cov.data[0], (size_t)N * C1);
  rvtb200::GeneBatcher<shim::DataConsolidator>::instance().setBatch(batch);

  const int nPerm = argc > 3 ? atoi(argv[3]) : 0;        // skat[nPerm=..,alpha=..]
  const double alpha = argc > 4 ? atof(argv[4]) : 0.05;
  const bool binary = argc > 5 && atoi(argv[5]) != 0;    // ModelManager: setBinaryOutcome() on every model
  if (argc > 6) rvtb200::GeneBatcher<shim::DataConsolidator>::instance().enableSkatOBinary(atoi(argv[6]) != 0);   // default: on
  SkatTest skat(nPerm, alpha);
  SkatOTest skato;
  CMCTest cmc;
  ZegginiTest zeg;
  if (binary) {
    skat.setBinaryOutcome();
    skato.setBinaryOutcome();
    cmc.setBinaryOutcome();
    zeg.setBinaryOutcome();
  }
  shim::FileWriter fw[4];
  shim::Result site;
  site.keys.push_back("Range");
  site.keys.push_back("N_INFORMATIVE");
  site.keys.push_back("NumVar");
  site.values.resize(3);
  skat.writeHeader(&fw[0], site);
  skato.writeHeader(&fw[1], site);
  cmc.writeHeader(&fw[2], site);
  zeg.writeHeader(&fw[3], site);
  for (int g = 0; g < nGenes; ++g) {
    int M;
    rd(f, &M, 1);
    dc.geno.Dimension(N, M);
    if (M) rd(f, &dc.geno.data[0], (size_t)N * M);
    dc.af.resize(M);
    if (M) rd(f, &dc.af[0], M);
    char buf[64];
    snprintf(buf, sizeof(buf), "gene%d", g);
    site.values[0] = buf;  // the reused siteInfo buffer
    snprintf(buf, sizeof(buf), "%d", N);
    site.values[1] = buf;
    snprintf(buf, sizeof(buf), "%d", M);
    site.values[2] = buf;
    skat.reset(); skat.fit(&dc); skat.writeOutput(&fw[0], site);
    skato.reset(); skato.fit(&dc); skato.writeOutput(&fw[1], site);
    cmc.reset(); cmc.fit(&dc); cmc.writeOutput(&fw[2], site);
    zeg.reset(); zeg.fit(&dc); zeg.writeOutput(&fw[3], site);
    dc.phenoUpdated = dc.covUpdated = false;
  }
  skat.writeFootnote(&fw[0]);
  skato.writeFootnote(&fw[1]);
  cmc.writeFootnote(&fw[2]);
  zeg.writeFootnote(&fw[3]);
  const char* names[4] = {"Skat", "SkatO", "CMC", "Zeggini"};
  for (int m = 0; m < 4; ++m) printf("#%s\n%s", names[m], fw[m].out.c_str());
  fclose(f);
  return 0;
}
