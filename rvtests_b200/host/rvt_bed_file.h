// rvt_bed_file.h -- genotype ingestion for the engine: a binary PLINK fileset
// (prefix.bed / .bim / .fam) as the source of rvt_gene_push_bed() rows.  SNP-major .bed rows ARE the engine's host format,
// so a gene is "the rows whose .bim position falls into its ranges", pushed straight out of the file mapping: no decode,
// no N x M Matrix.
//
// Follows the reference's reader, libVcf/PlinkInputFile.h:14-139:
//   .bed  three header bytes 0x6c 0x1b mode; mode 0x01 = SNP-major (the only layout the engine takes; 0x00 = individual-
//         major is refused here with an error, the reference transposes it while decoding, PlinkInputFile.cpp:57-96);
//         ceil(N/4) bytes per marker, sample p in bits 2(p&3).. of byte p>>2 (PlinkInputFile.cpp:23-47)
//   .bim  six whitespace-separated columns chrom, id, cM, pos, a1, a2; the key of a marker is its id, or "chrom:pos" when the
//         id is "."; a duplicated key is an error
//   .fam  six columns, sample id = column 2, sex = column 5 (atoi), phenotype = column 6 (atof); duplicated id = error
// AF per row as GenotypeCounter::getAF would count the decoded calls (missing stays in the denominator).
// Header-only, C++11; POSIX mmap for the .bed.
#ifndef RVT_BED_FILE_H_
#define RVT_BED_FILE_H_

#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <string>
#include <vector>

#include "rvt_vcf_pack.h"   // VcfRangeSet
#include "rvtests_b200.h"

namespace rvtb200 {

class BedFile {
 public:
  BedFile() : map_(NULL), map_len_(0), fd_(-1), n_(0), stride_(0) {}
  ~BedFile() { close(); }

  // 0 on success; otherwise error() says why (the reference aborts / exits in the same situations)
  int open(const std::string& prefix) {
    close();
    if (readFam(prefix + ".fam") || readBim(prefix + ".bim")) return -1;
    n_ = (int64_t)indv_.size();
    stride_ = (n_ + 3) / 4;
    fd_ = ::open((prefix + ".bed").c_str(), O_RDONLY);
    if (fd_ < 0) return fail("Cannot open binary PLINK file!");
    struct stat st;
    if (fstat(fd_, &st) != 0 || st.st_size < 3) return fail("Encounter error when reading plink BED files.");
    map_len_ = (size_t)st.st_size;
    map_ = (const uint8_t*)mmap(NULL, map_len_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (map_ == (const uint8_t*)MAP_FAILED) {
      map_ = NULL;
      return fail("mmap of the .bed file failed");
    }
    if (map_[0] != 0x6c || map_[1] != 0x1b) return fail("Magic number of binary PLINK file does not match!");
    if (map_[2] == 0x00) return fail("individual-major .bed: convert to SNP-major (plink --make-bed) for the engine");
    if (map_[2] != 0x01) return fail("Unrecognized major mode in binary PLINK file.");
    if (map_len_ < 3 + (size_t)stride_ * pos_.size()) return fail(".bed is shorter than .bim x .fam imply");
    return 0;
  }
  void close() {
    if (map_) munmap((void*)map_, map_len_);
    if (fd_ >= 0) ::close(fd_);
    map_ = NULL;
    fd_ = -1;
    map_len_ = 0;
  }
  const std::string& error() const { return err_; }

  int64_t numSample() const { return n_; }
  int numMarker() const { return (int)pos_.size(); }
  int64_t stride() const { return stride_; }
  const std::vector<std::string>& sampleNames() const { return indv_; }
  const std::vector<int>& sex() const { return sex_; }
  const std::vector<double>& phenotype() const { return pheno_; }
  const std::string& chrom(int j) const { return chrom_[j]; }
  int pos(int j) const { return pos_[j]; }
  const std::string& markerName(int j) const { return snp_[j]; }
  // index of the marker with this key (id, or "chrom:pos" for id "."), -1 if absent
  int markerIndex(const std::string& key) const {
    std::map<std::string, int>::const_iterator it = snp2idx_.find(key);
    return it == snp2idx_.end() ? -1 : it->second;
  }
  const uint8_t* row(int j) const { return map_ + 3 + (size_t)j * (size_t)stride_; }

  // rows whose (chrom, pos) lies in the range set, in file order
  void rowsIn(const VcfRangeSet& ranges, std::vector<int>* out) const {
    out->clear();
    for (int j = 0; j < numMarker(); ++j)
      if (ranges.contains(chrom_[j].data(), chrom_[j].size(), pos_[j])) out->push_back(j);
  }

  // GenotypeCounter::getAF of row j: 0.5 * (#het + 2 #hom-alt) / N; counts[4] (optional): hom-ref, het, hom-alt, missing
  double alleleFrequency(int j, int* counts = NULL) const {
    static const CountTable tab;
    const uint8_t* r = row(j);
    int64_t c[4] = {0, 0, 0, 0};
    const int64_t full = n_ / 4;
    for (int64_t b = 0; b < full; ++b)
      for (int k = 0; k < 4; ++k) c[k] += tab.t[r[b]][k];
    for (int64_t p = full * 4; p < n_; ++p) ++c[(r[p >> 2] >> (2 * (p & 3))) & 3];
    // codes: 00 hom-ref, 01 missing, 10 het, 11 hom-alt (libVcf/PlinkInputFile.h:206-209)
    if (counts) {
      counts[0] = (int)c[0];
      counts[1] = (int)c[2];
      counts[2] = (int)c[3];
      counts[3] = (int)c[1];
    }
    return n_ ? 0.5 * (double)(c[2] + 2 * c[3]) / (double)n_ : -1.0;
  }

  // push the given rows as one gene.  Consecutive rows go out of the mapping directly; otherwise they are gathered first.
  int push(rvt_ctx* ctx, const std::vector<int>& rows) {
    if (rows.empty()) return RVT_E_BADARG;
    std::vector<double> af(rows.size());
    bool consecutive = true;
    for (size_t k = 0; k < rows.size(); ++k) {
      af[k] = alleleFrequency(rows[k]);
      if (k && rows[k] != rows[k - 1] + 1) consecutive = false;
    }
    if (consecutive) return rvt_gene_push_bed(ctx, row(rows[0]), (int)rows.size(), stride_, &af[0]);
    gather_.resize(rows.size() * (size_t)stride_);
    for (size_t k = 0; k < rows.size(); ++k) memcpy(&gather_[k * (size_t)stride_], row(rows[k]), (size_t)stride_);
    return rvt_gene_push_bed(ctx, &gather_[0], (int)rows.size(), stride_, &af[0]);
  }

 private:
  struct CountTable {
    uint8_t t[256][4];
    CountTable() {
      for (int b = 0; b < 256; ++b) {
        t[b][0] = t[b][1] = t[b][2] = t[b][3] = 0;
        for (int k = 0; k < 4; ++k) ++t[b][(b >> (2 * k)) & 3];
      }
    }
  };
  int fail(const char* why) {
    err_ = why;
    close();
    return -1;
  }
  // LineReader::readLineBySep(&fd, " \t"): split on every blank or tab (empty fields are kept)
  static void split(const std::string& line, std::vector<std::string>* fd) {
    fd->clear();
    size_t b = 0;
    while (true) {
      const size_t e = line.find_first_of(" \t", b);
      if (e == std::string::npos) {
        fd->push_back(line.substr(b));
        return;
      }
      fd->push_back(line.substr(b, e - b));
      b = e + 1;
    }
  }
  template <class F>
  int eachLine(const std::string& path, F f) {
    FILE* fp = fopen(path.c_str(), "rt");
    if (!fp) return fail("Cannot open binary PLINK file!");
    std::string line;
    char buf[65536];
    int rc = 0;
    while (rc == 0 && fgets(buf, sizeof(buf), fp)) {
      line += buf;
      if (line.empty() || line[line.size() - 1] != '\n') continue;   // long line: keep reading
      while (!line.empty() && (line[line.size() - 1] == '\n' || line[line.size() - 1] == '\r')) line.erase(line.size() - 1);
      rc = f(line);
      line.clear();
    }
    if (rc == 0 && !line.empty()) rc = f(line);
    fclose(fp);
    return rc;
  }
  int readBim(const std::string& path) {
    chrom_.clear();
    snp_.clear();
    pos_.clear();
    snp2idx_.clear();
    std::vector<std::string> fd;
    return eachLine(path, [&](const std::string& line) -> int {
      split(line, &fd);
      if (fd.size() != 6) return fail("Wrong format in bim file.");
      const std::string key = fd[1] == "." ? fd[0] + ":" + fd[3] : fd[1];
      if (snp2idx_.count(key)) return fail("Error found: duplicated marker name or chromosomal position");
      snp2idx_[key] = (int)pos_.size();
      chrom_.push_back(fd[0]);
      snp_.push_back(fd[1]);
      pos_.push_back(atoi(fd[3].c_str()));
      return 0;
    });
  }
  int readFam(const std::string& path) {
    indv_.clear();
    sex_.clear();
    pheno_.clear();
    std::map<std::string, int> seen;
    std::vector<std::string> fd;
    return eachLine(path, [&](const std::string& line) -> int {
      split(line, &fd);
      if (fd.size() != 6) return fail("Wrong format in fam file.");
      if (seen.count(fd[1])) return fail("duplicated person id");
      seen[fd[1]] = 1;
      indv_.push_back(fd[1]);
      sex_.push_back(atoi(fd[4].c_str()));
      pheno_.push_back(atof(fd[5].c_str()));
      return 0;
    });
  }

  const uint8_t* map_;
  size_t map_len_;
  int fd_;
  int64_t n_, stride_;
  std::vector<std::string> indv_, chrom_, snp_;
  std::vector<int> sex_, pos_;
  std::vector<double> pheno_;
  std::map<std::string, int> snp2idx_;
  std::vector<uint8_t> gather_;
  std::string err_;
};

}  // namespace rvtb200

#endif  // RVT_BED_FILE_H_
