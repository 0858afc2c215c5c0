// ingest_demo.cpp -- the ingestion headers against the C ABI, end to end on the GPU box:
//   ingest_demo vcf  <file.vcf> <setFile> <pheno.txt>     plain-text VCF; pheno.txt: one "sample value" pair per line
//   ingest_demo bed  <plink-prefix> <setFile>            phenotype = column 6 of the .fam
//   ingest_demo vcfgz <file.vcf.gz> <setFile> <pheno.txt>  bgzipped VCF + .tbi: every range of a set is a tabix query (rvt_bgzf.h),
//                                                         as VCFInputFile's range mode reads it (libVcf/VCFInputFile.cpp)
//   ingest_demo bgen <file.bgen> <setFile> <pheno.txt>    BGEN dosages (rvt_bgen.h), mean-imputed, pushed as doubles
// For every set of the setFile (src/Main.cpp:138-173) the variants in its ranges are packed (rvt_vcf_pack.h) or taken
// straight out of the mapped .bed (rvt_bed_file.h), pushed with rvt_gene_push_bed, and the SKAT / CMC / Zeggini records of
// one rvt_flush are printed.  Intercept-only null model; samples in file order (the caller's DataLoader job otherwise).
#include <stdio.h>
#include <stdlib.h>

#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "rvt_bed_file.h"
#include "rvt_bgen.h"
#include "rvt_bgzf.h"
#include "rvt_vcf_pack.h"

static void readPheno(const char* path, std::map<std::string, double>* ph) {
  std::ifstream pf(path);
  std::string id;
  double v;
  while (pf >> id >> v) (*ph)[id] = v;
}

static int die(rvt_ctx* ctx, const char* what) {
  fprintf(stderr, "%s: %s\n", what, ctx ? rvt_last_error(ctx) : "");
  return 1;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const std::string mode = argv[1];
  rvtb200::GeneRangeMap sets;
  if (sets.loadRangeFile(argv[3]) <= 0) return die(NULL, "no sets");
  rvt_ctx* ctx = NULL;
  if (rvt_ctx_create(0, &ctx) != RVT_OK) return die(ctx, "rvt_ctx_create");
  std::vector<std::string> names;
  int pushed = 0;
  if (mode == "bed") {
    rvtb200::BedFile bf;
    if (bf.open(argv[2])) {
      fprintf(stderr, "%s\n", bf.error().c_str());
      return 1;
    }
    const int64_t n = bf.numSample();
    std::vector<double> X((size_t)n, 1.0);
    if (rvt_set_null_model(ctx, n, 1, X.data(), bf.phenotype().data(), 0) != RVT_OK) return die(ctx, "null model");
    std::vector<int> rows;
    for (size_t g = 0; g < sets.size(); ++g) {
      bf.rowsIn(sets.ranges(g), &rows);
      if (rows.empty()) continue;
      if (bf.push(ctx, rows) != RVT_OK) return die(ctx, "push");
      names.push_back(sets.name(g));
      ++pushed;
    }
  } else if (mode == "vcfgz") {
    if (argc < 5) return 2;
    rvtb200::TabixReader tr;
    if (!tr.open(argv[2])) {
      fprintf(stderr, "%s: %s\n", argv[2], tr.error().c_str());
      return 1;
    }
    std::vector<std::string> hdr;
    tr.readHeader(&hdr);
    std::string header;
    for (size_t i = 0; i < hdr.size(); ++i)
      if (hdr[i].compare(0, 6, "#CHROM") == 0) header = hdr[i];
    std::map<std::string, double> ph;
    readPheno(argv[4], &ph);
    std::vector<std::string> keep;
    for (std::map<std::string, double>::const_iterator it = ph.begin(); it != ph.end(); ++it) keep.push_back(it->first);
    rvtb200::VcfGenePacker pk;
    const int n = pk.setHeader(header.data(), header.size(), &keep);
    if (n <= 0) return die(NULL, "VCF header / phenotype samples");
    std::vector<double> X((size_t)n, 1.0), y((size_t)n);
    for (int i = 0; i < n; ++i) y[i] = ph[pk.sampleNames()[i]];
    if (rvt_set_null_model(ctx, n, 1, X.data(), y.data(), 0) != RVT_OK) return die(ctx, "null model");
    std::string line;
    for (size_t g = 0; g < sets.size(); ++g) {
      pk.clear();
      pk.ranges() = sets.ranges(g);
      const rvtb200::VcfRangeSet& rs = sets.ranges(g);
      for (size_t k = 0; k < rs.size(); ++k) {
        if (!tr.query(rs.chrom(k), rs.begin(k), rs.end(k))) continue;   // a sequence the file does not hold
        while (tr.next(&line))
          if (pk.addRecord(line.data(), line.size()) < 0) return die(NULL, "malformed VCF record");
      }
      if (pk.numVariant() == 0) continue;
      if (pk.push(ctx) != RVT_OK) return die(ctx, "push");
      names.push_back(sets.name(g));
      ++pushed;
    }
  } else if (mode == "bgen") {
    if (argc < 5) return 2;
    rvtb200::BgenReader br;
    if (!br.open(argv[2])) {
      fprintf(stderr, "%s: %s\n", argv[2], br.error().c_str());
      return 1;
    }
    std::map<std::string, double> ph;
    readPheno(argv[4], &ph);
    const int n = (int)br.numSample();
    std::vector<double> X((size_t)n, 1.0), y((size_t)n);
    for (int i = 0; i < n; ++i) {
      if (!ph.count(br.sampleIdentifier()[i])) return die(NULL, "a BGEN sample without a phenotype");
      y[i] = ph[br.sampleIdentifier()[i]];
    }
    if (rvt_set_null_model(ctx, n, 1, X.data(), y.data(), 0) != RVT_OK) return die(ctx, "null model");
    for (size_t g = 0; g < sets.size(); ++g) {
      const rvtb200::VcfRangeSet& rs = sets.ranges(g);
      std::vector<double> block, af;
      for (size_t k = 0; k < rs.size(); ++k) {
        if (!br.open(argv[2])) return 1;   // (no .bgi: every range is a scan that skips the genotype blocks outside it)
        br.setRange(rs.chrom(k), (uint32_t)rs.begin(k), (uint32_t)rs.end(k));
        while (br.readRecord()) {
          const size_t at = block.size();
          br.appendDosages(&block);
          // GenotypeCounter::getAF (src/GenotypeCounter.h:43-49), then DataConsolidator::imputeGenotypeToMean with its integer
          // accumulator (src/DataConsolidator.cpp:217-245)
          double sum = 0.0;
          int ac = 0, an = 0;
          bool any = false;
          for (int i = 0; i < n; ++i) {
            const double v = block[at + i];
            if (v >= 0) {
              sum += v;
              ac = (int)(ac + v);
              an += 2;
            } else
              any = true;
          }
          af.push_back(0.5 * sum / n);
          if (any) {
            const double fill = 2.0 * (an == 0 ? 0.0 : 1.0 * ac / an);
            for (int i = 0; i < n; ++i)
              if (block[at + i] < 0) block[at + i] = fill;
          }
        }
        if (!br.error().empty()) {
          fprintf(stderr, "%s: %s\n", argv[2], br.error().c_str());
          return 1;
        }
      }
      if (af.empty()) continue;
      if (rvt_gene_push_f64(ctx, block.data(), (int)af.size(), af.data()) != RVT_OK) return die(ctx, "push");
      names.push_back(sets.name(g));
      ++pushed;
    }
  } else {
    if (argc < 5) return 2;
    std::ifstream vcf(argv[2]);
    std::vector<std::string> lines;
    std::string line, header;
    while (std::getline(vcf, line)) {
      if (line.compare(0, 6, "#CHROM") == 0) header = line;
      else if (!line.empty() && line[0] != '#') lines.push_back(line);
    }
    std::map<std::string, double> ph;
    readPheno(argv[4], &ph);
    std::vector<std::string> keep;
    for (std::map<std::string, double>::const_iterator it = ph.begin(); it != ph.end(); ++it) keep.push_back(it->first);
    rvtb200::VcfGenePacker pk;
    const int n = pk.setHeader(header.data(), header.size(), &keep);
    if (n <= 0) return die(NULL, "VCF header / phenotype samples");
    std::vector<double> X((size_t)n, 1.0), y((size_t)n);
    for (int i = 0; i < n; ++i) y[i] = ph[pk.sampleNames()[i]];   // phenotype in VCF column order
    if (rvt_set_null_model(ctx, n, 1, X.data(), y.data(), 0) != RVT_OK) return die(ctx, "null model");
    for (size_t g = 0; g < sets.size(); ++g) {
      pk.clear();
      pk.ranges() = sets.ranges(g);
      for (size_t k = 0; k < lines.size(); ++k)
        if (pk.addRecord(lines[k].data(), lines[k].size()) < 0) return die(NULL, "malformed VCF record");
      if (pk.numVariant() == 0) continue;
      if (pk.push(ctx) != RVT_OK) return die(ctx, "push");
      names.push_back(sets.name(g));
      ++pushed;
    }
  }
  std::vector<rvt_gene_result> res((size_t)(pushed ? pushed : 1));
  int got = 0;
  if (rvt_flush(ctx, res.data(), (int)res.size(), &got) != RVT_OK) return die(ctx, "rvt_flush");
  printf("Set\tNumPolyVar\tQ\tPvalue\tNonRefSite\tCMC_P\tZeggini_P\n");
  for (int g = 0; g < got; ++g)
    printf("%s\t%d\t%g\t%g\t%d\t%g\t%g\n", names[g].c_str(), res[g].m_poly, res[g].Q, res[g].p_skat, res[g].cmc_nonref,
           res[g].cmc_p, res[g].zeg_p);
  rvt_ctx_destroy(ctx);
  return 0;
}
