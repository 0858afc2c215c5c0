// tests/hostcheck/bedfile_check.cpp -- C wrapper around rvtests_b200/host/rvt_bed_file.h for the CPU tests (ctypes).
// rvt_gene_push_bed is replaced by a recorder so that BedFile::push can be checked without the engine.
#include <string>
#include <vector>

#include "../../rvtests_b200/host/rvt_bed_file.h"

static std::vector<uint8_t> g_rows;
static std::vector<double> g_af;
static int g_m = 0;
static long long g_stride = 0;
extern "C" int rvt_gene_push_bed(rvt_ctx*, const uint8_t* bed, int M, int64_t stride, const double* af) {
  g_m = M;
  g_stride = stride;
  g_rows.assign(bed, bed + (size_t)M * stride);
  g_af.assign(af, af + M);
  return RVT_OK;
}
extern "C" int rvt_gene_push_f64(rvt_ctx*, const double*, int, const double*) { return RVT_E_UNSUPPORTED; }

static rvtb200::BedFile g_f;
static std::vector<int> g_sel;

extern "C" {
int bf_open(const char* prefix) { return g_f.open(prefix); }
const char* bf_error() { return g_f.error().c_str(); }
long long bf_num_sample() { return g_f.numSample(); }
int bf_num_marker() { return g_f.numMarker(); }
long long bf_stride() { return g_f.stride(); }
const char* bf_sample(int i) { return g_f.sampleNames()[i].c_str(); }
int bf_sex(int i) { return g_f.sex()[i]; }
double bf_pheno(int i) { return g_f.phenotype()[i]; }
int bf_pos(int j) { return g_f.pos(j); }
const char* bf_chrom(int j) { return g_f.chrom(j).c_str(); }
int bf_marker_index(const char* key) { return g_f.markerIndex(key); }
double bf_af(int j, int* counts) { return g_f.alleleFrequency(j, counts); }
void bf_row(int j, unsigned char* out) { memcpy(out, g_f.row(j), (size_t)g_f.stride()); }
int bf_select(const char* ranges, int* out, int cap) {
  rvtb200::VcfRangeSet rs;
  if (rs.add(ranges) < 0) return -1;
  g_f.rowsIn(rs, &g_sel);
  for (size_t k = 0; k < g_sel.size() && (int)k < cap; ++k) out[k] = g_sel[k];
  return (int)g_sel.size();
}
int bf_push(const int* rows, int n) { return g_f.push(NULL, std::vector<int>(rows, rows + n)); }
int bf_pushed(unsigned char* rows, double* af) {
  if (rows) memcpy(rows, g_rows.data(), g_rows.size());
  if (af) memcpy(af, g_af.data(), sizeof(double) * g_af.size());
  return g_m;
}
}
