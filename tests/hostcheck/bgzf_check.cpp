// ctypes shim over rvtests_b200/host/rvt_bgzf.h for tests/test_bgzf_tabix.py (test infrastructure)
#include "rvt_bgzf.h"

extern "C" {
// write `n` bytes of text in pieces of `piece` bytes through IndexedAssocWriter -> its close() code
int bz_write_indexed(const char* path, const char* text, long n, int piece) {
  rvtb200::IndexedAssocWriter w;
  if (!w.open(path)) return -9;
  std::string buf;
  for (long o = 0; o < n; o += piece) {
    buf.assign(text + o, (size_t)((n - o < piece) ? n - o : piece));
    w.write(buf.c_str());
  }
  return w.close();
}
int bz_printf_check(const char* path) {
  rvtb200::IndexedAssocWriter w(path);
  w.printf("#%s\n", "CHROM\tPOS\tX");
  for (int i = 1; i <= 50; ++i) w.printf("%d\t%d\t%g\n", 1 + i / 30, 100 * i, 0.5 * i);
  std::string big(10000, 'x');
  w.printf("3\t77\t%s\n", big.c_str());
  return w.close();
}
int bz_reg2bin(unsigned beg, unsigned end) { return rvtb200::TabixIndex::reg2bin(beg, end); }
}

// ---- rvt_summary.h --------------------------------------------------------------------------------------------------
#include "rvt_summary.h"
namespace {
struct StrWriter {
  std::string out;
  int write(const char* s) {
    out += s;
    return (int)strlen(s);
  }
};
}  // namespace
extern "C" int sh_render(int N, int n_cov, const double* y, const double* cov /* col-major */, const char* version, char* out, int cap) {
  rvtb200::SummaryHeaderB200<StrWriter> sh(version);
  sh.recordPhenotype("Trait", std::vector<double>(y, y + N));
  std::vector<std::string> labels;
  std::vector<std::vector<double> > cols;
  for (int j = 0; j < n_cov; ++j) {
    char b[32];
    snprintf(b, sizeof b, "cov%d", j);
    labels.push_back(b);
    cols.push_back(std::vector<double>(cov + (size_t)j * N, cov + (size_t)(j + 1) * N));
  }
  sh.recordCovariate(labels, cols);
  StrWriter w;
  sh.outputHeader(&w);
  if ((int)w.out.size() + 1 > cap) return -1;
  memcpy(out, w.out.c_str(), w.out.size() + 1);
  return (int)w.out.size();
}

// ---- readers --------------------------------------------------------------------------------------------------------
// lines of <path> overlapping chrom:beg-end (1-based inclusive) joined by '\n' -> length, -1 unknown sequence, -2 open error
extern "C" long bz_query(const char* path, const char* chrom, int beg, int end, char* out, long cap) {
  rvtb200::TabixReader r;
  if (!r.open(path)) return -2;
  if (!r.query(chrom, beg, end)) return -1;
  std::string all, line;
  while (r.next(&line)) all += line + "\n";
  if ((long)all.size() + 1 > cap) return -3;
  memcpy(out, all.c_str(), all.size() + 1);
  return (long)all.size();
}
extern "C" long bz_header(const char* path, char* out, long cap) {
  rvtb200::TabixReader r;
  if (!r.open(path)) return -2;
  std::vector<std::string> h;
  r.readHeader(&h);
  std::string all;
  for (size_t i = 0; i < h.size(); ++i) all += h[i] + "\n";
  if ((long)all.size() + 1 > cap) return -3;
  memcpy(out, all.c_str(), all.size() + 1);
  return (long)all.size();
}
// sequential read of a BGZF file through BgzfReader::getline -> number of lines, bytes summed into *nbytes
extern "C" long bz_count_lines(const char* path, long* nbytes) {
  rvtb200::BgzfReader r;
  if (!r.open(path)) return -2;
  std::string l;
  long n = 0;
  *nbytes = 0;
  while (r.getline(&l)) {
    ++n;
    *nbytes += (long)l.size() + 1;
  }
  return r.error().empty() ? n : -4;
}
