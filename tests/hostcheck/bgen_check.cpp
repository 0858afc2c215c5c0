// ctypes shim over rvtests_b200/host/rvt_bgen.h for tests/test_bgen_reader.py (test infrastructure)
#include "rvt_bgen.h"

#include <string>

static std::string g(float v) {
  char b[64];
  snprintf(b, sizeof b, "%g", v);
  return b;
}

extern "C" {
// text dump: line 1 = sample identifiers; then per variant
//   chrom pos rsid varid alleles(,) P|U  and per sample "probabilities(,)|dosage" -- '.' entries for a missing sample
// -> length; -1 open error, -2 read error, -3 buffer too small
long bg_dump(const char* path, const char* chrom, unsigned beg, unsigned end, char* out, long cap, int* layout, int* compression,
             unsigned* n_sample, unsigned* n_marker) {
  rvtb200::BgenReader r;
  if (!r.open(path)) return -1;
  *layout = r.layout();
  *compression = r.compression();
  *n_sample = r.numSample();
  *n_marker = r.numMarker();
  if (chrom && chrom[0]) r.setRange(chrom, beg, end);
  std::string s;
  for (size_t i = 0; i < r.sampleIdentifier().size(); ++i) s += (i ? "\t" : "") + r.sampleIdentifier()[i];
  s += "\n";
  while (r.readRecord()) {
    char b[64];
    snprintf(b, sizeof b, "%u", r.pos);
    s += r.chrom + "\t" + b + "\t" + r.rsid + "\t" + r.varid + "\t";
    for (size_t a = 0; a < r.alleles.size(); ++a) s += (a ? "," : "") + r.alleles[a];
    s += r.phased ? "\tP" : "\tU";
    for (unsigned i = 0; i < r.numSample(); ++i) {
      s += "\t";
      for (int k = r.index[i]; k < r.index[i + 1]; ++k) s += (k > r.index[i] ? "," : "") + (r.missing[i] ? std::string(".") : g(r.prob[k]));
      snprintf(b, sizeof b, "|%.17g", r.dosage((int)i));
      s += b;
    }
    s += "\n";
  }
  if (!r.error().empty()) {
    snprintf(out, (size_t)cap, "%s", r.error().c_str());
    return -2;
  }
  if ((long)s.size() + 1 > cap) return -3;
  memcpy(out, s.c_str(), s.size() + 1);
  return (long)s.size();
}
}
