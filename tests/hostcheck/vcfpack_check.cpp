// tests/hostcheck/vcfpack_check.cpp -- C wrapper around the product's host-side VCF packer (rvtests_b200/host/rvt_vcf_pack.h)
// so that the CPU tests can drive it through ctypes.  rvt_gene_push_bed is referenced by VcfGenePacker::push only; the
// wrapper never calls it, and the test library is linked without the engine (the symbol stays an unused inline reference).
#include <string>
#include <vector>

#include "../../rvtests_b200/host/rvt_vcf_pack.h"

extern "C" int rvt_gene_push_bed(rvt_ctx*, const uint8_t*, int, int64_t, const double*) { return RVT_E_UNSUPPORTED; }
extern "C" int rvt_gene_push_f64(rvt_ctx*, const double*, int, const double*) { return RVT_E_UNSUPPORTED; }

static rvtb200::VcfGenePacker g_p;

extern "C" {
int vp_gt(const char* s, int len) { return rvtb200::vcfGenotype(s, len); }
// keep: '\n'-separated names or NULL
int vp_header(const char* line, const char* keep) {
  std::vector<std::string> k;
  if (keep) {
    std::string s(keep);
    size_t b = 0;
    while (b < s.size()) {
      size_t e = s.find('\n', b);
      if (e == std::string::npos) e = s.size();
      if (e > b) k.push_back(s.substr(b, e - b));
      b = e + 1;
    }
  }
  return g_p.setHeader(line, strlen(line), keep ? &k : NULL);
}
int vp_set_range(const char* spec) {
  g_p.ranges().clear();
  return spec && spec[0] ? g_p.ranges().add(spec) : 0;
}
void vp_clear() { g_p.clear(); }
int vp_add(const char* line, int len) { return g_p.addRecord(line, (size_t)len); }
int vp_num_variant() { return g_p.numVariant(); }
long long vp_stride() { return g_p.stride(); }
long long vp_num_sample() { return g_p.numSample(); }
void vp_get(unsigned char* rows, double* af, int* counts) {
  const int m = g_p.numVariant();
  if (m == 0) return;
  memcpy(rows, g_p.rows(), (size_t)m * g_p.stride());
  memcpy(af, g_p.af(), sizeof(double) * m);
  memcpy(counts, g_p.counts(), sizeof(int) * 4 * m);
}
int vp_gt_male02(const char* s, int len) { return rvtb200::vcfGenotypeMale02(s, len); }
int vp_par_is_hemi(const char* xLabel, const char* parRegion, const char* chrom, int pos) {
  rvtb200::VcfParRegion p;
  p.init(xLabel, parRegion);
  return p.isHemiRegion(chrom, pos) ? 1 : 0;
}
// sex == NULL: X handling off
int vp_set_sex(const int* sex, int n, const char* xLabel, const char* parRegion) {
  g_p.parRegion().init(xLabel ? xLabel : "", parRegion ? parRegion : "");
  return g_p.setSex(sex ? std::vector<int>(sex, sex + n) : std::vector<int>());
}
int vp_count_alt(const char* s, int len, int alt) { return rvtb200::vcfCountAltAllele(s, len, alt); }
int vp_count_male_alt2(const char* s, int len, int alt) { return rvtb200::vcfCountMaleAltAllele2(s, len, alt); }
void vp_set_multi(int on) { g_p.setMultiAllelic(on != 0); }
void vp_set_freq(double lo, double hi) { g_p.setFreqRange(lo, hi); }
int vp_parse_range(const char* s, char* chrom, int* beg, int* end) {
  std::string c;
  const bool ok = rvtb200::vcfParseRange(s, &c, beg, end);
  strncpy(chrom, c.c_str(), 63);
  chrom[63] = 0;
  return ok ? 0 : -1;
}
static rvtb200::GeneRangeMap g_map;
int vp_load_gene_file(const char* path, const char* only) {
  g_map.clear();
  return g_map.loadGeneFile(path, only ? only : "");
}
int vp_load_range_file(const char* path, const char* only) {
  g_map.clear();
  return g_map.loadRangeFile(path, only ? only : "");
}
const char* vp_map_name(int i) { return g_map.name(i).c_str(); }
int vp_map_contains(int i, const char* chrom, int pos) { return g_map.ranges(i).contains(chrom, strlen(chrom), pos) ? 1 : 0; }
int vp_map_nranges(int i) { return (int)g_map.ranges(i).size(); }
void vp_set_filters(int gd_min, int gd_max, int gq_min, int gq_max) {
  g_p.setDepthFilter(gd_min, gd_max);
  g_p.setQualFilter(gq_min, gq_max);
}
void vp_set_dosage_tag(const char* tag) { g_p.setDosageTag(tag ? tag : ""); }
// raw != 0: as read; else after imputeDosagesToMean()
void vp_get_dosages(double* out, double* af, int* counts, int raw) {
  const int m = g_p.numVariant();
  if (m == 0) return;
  if (!raw) g_p.imputeDosagesToMean();
  memcpy(out, g_p.dosages(), sizeof(double) * (size_t)m * (size_t)g_p.numSample());
  memcpy(af, g_p.af(), sizeof(double) * m);
  memcpy(counts, g_p.counts(), sizeof(int) * 4 * m);
}
const char* vp_variant_name(int j) { return g_p.variantName(j).c_str(); }
const char* vp_sample_name(int i) { return g_p.sampleNames()[i].c_str(); }
}
