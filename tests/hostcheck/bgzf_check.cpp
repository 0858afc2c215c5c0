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
