// tools/vcf_pack_time.cpp -- single-thread throughput of the host-side VCF packer (rvtests_b200/host/rvt_vcf_pack.h) on one
// 500 000-sample record:  g++ -O2 -std=c++11 -I include -I rvtests_b200/host tools/vcf_pack_time.cpp -o /tmp/vcf_pack_time && /tmp/vcf_pack_time
#include <chrono>
#include <stdio.h>
#include <string>
#include "rvt_vcf_pack.h"
extern "C" int rvt_gene_push_bed(rvt_ctx*, const uint8_t*, int, int64_t, const double*) { return 0; }
int main() {
  const int N = 500000;
  std::string hdr = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT";
  for (int i = 0; i < N; ++i) hdr += "\tS" + std::to_string(i);
  std::string rec = "1\t100\t.\tA\tG\t50\tPASS\t.\tGT:GQ";
  unsigned s = 1;
  for (int i = 0; i < N; ++i) { s = s * 1664525u + 1013904223u; unsigned r = s >> 24; rec += r < 240 ? "\t0/0:9" : r < 252 ? "\t0/1:9" : r < 254 ? "\t1/1:9" : "\t./.:9"; }
  rvtb200::VcfGenePacker pk;
  pk.setHeader(hdr.data(), hdr.size());
  auto t0 = std::chrono::steady_clock::now();
  const int R = 20;
  for (int k = 0; k < R; ++k) pk.addRecord(rec.data(), rec.size());
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("%d records of %zu bytes: %.3f s -> %.1f MB/s, %.2f M genotypes/s; af=%g\n", R, rec.size(), sec, R * rec.size() / sec / 1e6, R * (double)N / sec / 1e6, pk.af()[0]);
}
