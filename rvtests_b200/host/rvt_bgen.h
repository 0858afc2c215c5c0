// rvt_bgen.h -- a BGEN reader that yields what the reference's --inBgen path feeds DataConsolidator: per variant one dosage
// per sample (or MISSING_GENOTYPE = -9), N x M column-major doubles for rvt_gene_push_f64.  Header-only C++11 + zlib
// (zstd-compressed files need libzstd.so.1 at run time; it is looked up with dlopen, there is no link dependency).
//
// Follows, file:line in the reference:
//   BGenFile::BGenFile            libBgen/BGenFile.cpp:8-131   offset, header block (LH, M, N, "bgen", free data, flags:
//                                                               bits 0-1 compression, 2-5 layout, 31 sample identifiers),
//                                                               sample identifier block, then fseek(offset + 4)
//   parseLayout1                  :163-245   v1.1: N, ids, chrom, pos, two alleles; zlib block of N x 3 uint16 / 32768;
//                                            three zeros = missing
//   parseLayout2                  :247-389   v1.2: K alleles, C [D], payload (none / zlib / zstd): N, K, min / max ploidy,
//                                            one ploidy+missing byte per sample, phased flag, B bits; then per sample
//                                            (C(Z+K-1, K-1) - 1) probabilities (unphased) or Z (K-1) (phased) of B bits each,
//                                            the last one of every group restored as 1 - sum
//   BitReader                     libBgen/BitReader.h:15-79    little-endian bit stream, value * (1 / (2^B - 1)) in FLOAT
//   BGenGenotypeExtractor::getGenotype   src/BGenGenotypeExtractor.cpp:413-472: missing -> -9; two alleles: p1 + 2 p2 (read at
//                                            index + 1, index + 2 whatever the ploidy); one allele: 2; more: (p1 + 2 p2) / (p0 + p1 + p2)
//                                            of the first three entries, -9 when that total is 0; ploidy other than 1 or 2: -9
// Arithmetic is the reference's (float probabilities, the float remainder, a double dosage), so the doubles are the ones
// the reference's Matrix holds.  Pinned on the golden outputs of the reference's own reader tests
// (libBgen/test/*.bgen -> *.vcf.correct: every probability of every sample, %g) -- tests/test_bgen_reader.py.
// Range mode: the reference consults the .bgi sidecar (sqlite3, libBgen/BGenIndex.cpp); this reader filters while scanning
// (setRange), which needs no index and reads each variant's identifying block only for the skipped ones.
#ifndef RVT_BGEN_H_
#define RVT_BGEN_H_

#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <string>
#include <vector>

namespace rvtb200 {

class BgenReader {
 public:
  enum { kNoCompression = 0, kZlib = 1, kZstd = 2 };
  static constexpr double kMissingGenotype = -9.0;   // MISSING_GENOTYPE, libBgen/BGenFile.h:15

  BgenReader() : K(0), pos(0), phased(false), bits(0), fp_(NULL), file_size_(0), offset_(0), M_(0), N_(0), flags_(0), use_range_(false),
                 rbeg_(0), rend_(0) {}
  ~BgenReader() { close(); }
  void close() {
    if (fp_) fclose(fp_);
    fp_ = NULL;
  }
  bool open(const char* path) {
    close();
    error_.clear();
    fp_ = fopen(path, "rb");
    if (!fp_) return fail("cannot open the file");
    fseeko(fp_, 0, SEEK_END);
    file_size_ = (uint64_t)ftello(fp_);
    fseeko(fp_, 0, SEEK_SET);
    uint32_t LH;
    char magic[4];
    if (!rd(&offset_, 4) || !rd(&LH, 4) || !rd(&M_, 4) || !rd(&N_, 4) || !rd(magic, 4)) return fail("short header");
    if (memcmp(magic, "bgen", 4) != 0 && memcmp(magic, "\0\0\0\0", 4) != 0) return fail("bgen magic number does not match");
    if (LH < 20) return fail("header block too short");
    free_data_.resize(LH - 20);
    if (!free_data_.empty() && !rd(&free_data_[0], free_data_.size())) return fail("short header");
    if (!rd(&flags_, 4)) return fail("short header");
    if (layout() != 1 && layout() != 2) return fail("unsupported layout");
    if (compression() > 2) return fail("unsupported compression");
    if (layout() == 1 && compression() != kZlib) return fail("layout 1 is read with zlib compression only (as the reference)");
    sample_.clear();
    if (flags_ >> 31) {
      uint32_t LSI, N2;
      if (!rd(&LSI, 4) || !rd(&N2, 4)) return fail("short sample identifier block");
      if (N2 != N_ || (uint64_t)LSI + LH > offset_) return fail("inconsistent sample identifier block");
      sample_.resize(N_);
      for (uint32_t i = 0; i < N_; ++i)
        if (!rdString(2, &sample_[i])) return fail("short sample identifier block");
    }
    if (fseeko(fp_, (off_t)offset_ + 4, SEEK_SET) != 0) return fail("seek");
    return true;
  }
  uint32_t numSample() const { return N_; }
  uint32_t numMarker() const { return M_; }
  int layout() const { return (int)((flags_ >> 2) & 0xf); }
  int compression() const { return (int)(flags_ & 3); }
  const std::vector<std::string>& sampleIdentifier() const { return sample_; }
  // keep only variants of `chrom` with begin <= pos <= end (1-based inclusive); clearRange() reads everything
  void setRange(const std::string& chrom, uint32_t begin, uint32_t end) {
    use_range_ = true;
    rchrom_ = chrom;
    rbeg_ = begin;
    rend_ = end;
  }
  void clearRange() { use_range_ = false; }

  // the next variant (inside the range, if one is set); false at the end of the file or on error (see error())
  bool readRecord() {
    while (true) {
      if (!fp_ || (uint64_t)ftello(fp_) >= file_size_) return false;
      uint32_t n_row = N_;
      if (layout() == 1 && !rd(&n_row, 4)) return fail("short variant block");
      if (!rdString(2, &varid) || !rdString(2, &rsid) || !rdString(2, &chrom) || !rd(&pos, 4)) return fail("short variant block");
      K = 2;
      if (layout() == 2 && !rd(&K, 2)) return fail("short variant block");
      alleles.resize(K);
      for (int a = 0; a < K; ++a)
        if (!rdString(4, &alleles[a])) return fail("short variant block");
      uint32_t C = 0, D = 0;
      if (!rd(&C, 4)) return fail("short variant block");
      size_t payload = C;
      if (layout() == 1) {
        D = n_row * 6;
      } else if (compression() == kNoCompression) {
        D = C;
      } else {
        if (C < 4 || !rd(&D, 4)) return fail("short variant block");
        payload = C - 4;
      }
      const bool wanted = !use_range_ || (chrom == rchrom_ && pos >= rbeg_ && pos <= rend_);
      if (!wanted) {
        if (fseeko(fp_, (off_t)payload, SEEK_CUR) != 0) return fail("seek");
        continue;
      }
      cbuf_.resize(payload);
      if (payload && !rd(&cbuf_[0], payload)) return fail("short genotype block");
      buf_.resize(D);
      const int comp = layout() == 1 ? kZlib : compression();
      if (comp == kNoCompression) {
        buf_ = cbuf_;
      } else if (comp == kZlib) {
        unsigned long n = D;
        if (uncompress(buf_.data(), &n, cbuf_.data(), (unsigned long)payload) != Z_OK || n != D) return fail("zlib: corrupt genotype block");
      } else {
        if (!zstd(buf_.data(), D, cbuf_.data(), payload)) return false;
      }
      return layout() == 1 ? parse1(n_row) : parse2();
    }
  }

  // current variant
  std::string varid, rsid, chrom;
  uint16_t K;
  uint32_t pos;
  std::vector<std::string> alleles;
  bool phased;
  int bits;
  std::vector<uint8_t> missing, ploidy;
  std::vector<float> prob;
  std::vector<int> index;   // sample i owns prob[index[i] .. index[i + 1])

  // BGenGenotypeExtractor::getGenotype outside hemizygous regions
  double dosage(int i) const {
    if (missing[i]) return kMissingGenotype;
    if (ploidy[i] != 1 && ploidy[i] != 2) return kMissingGenotype;
    if (alleles.size() == 1) return 2.0;
    const size_t b = (size_t)index[i];
    const float p0 = at(b), p1 = at(b + 1), p2 = at(b + 2);   // (a haploid sample owns two entries: the third one read is its
    if (alleles.size() == 2) return p1 + p2 * 2.0;             //  neighbour's first, as in the reference; 0 past the end)
    const double total = p0 + p1 + p2;
    return total > 0.0 ? (p1 + p2 * 2.0) / total : kMissingGenotype;
  }
  // append the current variant as one column of N doubles (all samples, or keep[] in that order)
  void appendDosages(std::vector<double>* out, const std::vector<int>* keep = NULL) const {
    if (keep) {
      for (size_t k = 0; k < keep->size(); ++k) out->push_back(dosage((*keep)[k]));
    } else {
      for (uint32_t i = 0; i < N_; ++i) out->push_back(dosage((int)i));
    }
  }
  const std::string& error() const { return error_; }

 private:
  float at(size_t k) const { return k < prob.size() ? prob[k] : 0.0f; }
  bool fail(const char* what) {
    error_ = what;
    return false;
  }
  bool rd(void* p, size_t n) { return fread(p, 1, n, fp_) == n; }
  bool rdString(int len_bytes, std::string* out) {
    uint32_t n = 0;
    if (len_bytes == 2) {
      uint16_t n16;
      if (!rd(&n16, 2)) return false;
      n = n16;
    } else if (!rd(&n, 4))
      return false;
    if ((uint64_t)n > file_size_) return false;
    out->resize(n);
    return n == 0 || rd(&(*out)[0], n);
  }
  bool zstd(void* dst, size_t dst_len, const void* src, size_t src_len) {
    typedef size_t (*decompress_t)(void*, size_t, const void*, size_t);
    static decompress_t fn = NULL;
    static bool tried = false;
    if (!tried) {
      tried = true;
      void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL);
      if (h) fn = (decompress_t)dlsym(h, "ZSTD_decompress");
    }
    if (!fn) return fail("zstd-compressed BGEN: libzstd.so.1 is not available");
    if (fn(dst, dst_len, src, src_len) != dst_len) return fail("zstd: corrupt genotype block");
    return true;
  }
  bool parse1(uint32_t n_row) {
    if (n_row != N_) return fail("layout 1: the variant block has a different number of samples");
    missing.assign(N_, 0);
    ploidy.assign(N_, 2);
    phased = false;
    bits = 16;
    prob.resize((size_t)N_ * 3);
    index.resize(N_ + 1);
    for (uint32_t i = 0; i < N_; ++i) {
      uint16_t v[3];
      memcpy(v, buf_.data() + (size_t)i * 6, 6);
      index[i] = 3 * (int)i;
      for (int k = 0; k < 3; ++k) prob[(size_t)i * 3 + k] = (float)v[k] / 32768;
      missing[i] = v[0] == 0 && v[1] == 0 && v[2] == 0;
    }
    index[N_] = 3 * (int)N_;
    return true;
  }
  static int choose(int n, int m) {   // BGenFile::choose (int arithmetic, as there)
    if (m == 1) return n;
    if (n == 1) return 1;
    int r = 1;
    for (int i = 0; i < m; ++i) r *= (n - i);
    for (int i = 0; i < m; ++i) r /= (i + 1);
    return r;
  }
  bool parse2() {
    const size_t D = buf_.size();
    if (D < 10 + (size_t)N_) return fail("layout 2: short genotype block");
    uint32_t n_indv;
    memcpy(&n_indv, buf_.data(), 4);
    if (n_indv != N_) return fail("layout 2: the variant block has a different number of samples");
    const uint8_t* pm = buf_.data() + 8;
    phased = buf_[8 + N_] != 0;
    bits = buf_[8 + N_ + 1];
    if (bits < 1 || bits > 32) return fail("layout 2: bits per probability out of range");
    // BitReader
    const uint8_t* data = buf_.data() + 8 + N_ + 2;
    const size_t len = D - 8 - N_ - 2;
    size_t off = 0;
    unsigned avail = 0;
    uint64_t value = 0;
    const uint64_t mask = bits == 64 ? ~0ull : ((1ull << bits) - 1);
    float scale = 1.0;
    for (int i = 0; i < bits; ++i) scale *= 2;
    scale -= 1;
    scale = 1.0 / scale;
    const int B = bits;
    struct Next {
      const uint8_t* data;
      size_t len;
      size_t* off;
      unsigned* avail;
      uint64_t* value;
      uint64_t mask;
      float scale;
      int B;
      float operator()() const {
        if (B == 8 && *off < len) return (float)data[(*off)++] * scale;
        if (B == 16 && *off + 2 <= len) {
          uint16_t v;
          memcpy(&v, data + *off, 2);
          *off += 2;
          return (float)v * scale;
        }
        if (B == 32 && *off + 4 <= len) {
          uint32_t v;
          memcpy(&v, data + *off, 4);
          *off += 4;
          return (float)v * scale;
        }
        if (B == 8 || B == 16 || B == 32) return 0.0f;   // truncated block
        while (*avail < (unsigned)B && *off < len) {
          *value |= ((uint64_t)data[*off]) << *avail;
          ++*off;
          *avail += 8;
        }
        const float res = (float)(*value & mask);
        *avail -= B;
        *value >>= B;
        return res * scale;
      }
    } next = {data, len, &off, &avail, &value, mask, scale, B};
    missing.resize(N_);
    ploidy.resize(N_);
    index.clear();
    index.reserve(N_ + 1);
    prob.clear();
    for (uint32_t i = 0; i < N_; ++i) {
      index.push_back((int)prob.size());
      const int Z = pm[i] & 0x3f;
      ploidy[i] = (uint8_t)Z;
      missing[i] = (pm[i] & 0x80) != 0;
      if (phased) {
        for (int j = 0; j < Z; ++j) {
          float remain = 1.0;
          for (int k = 0; k < K - 1; ++k) {
            const float p = next();
            prob.push_back(p);
            remain -= p;
          }
          prob.push_back(remain);
        }
      } else {
        const int nc = choose(Z + K - 1, K - 1);
        float remain = 1.0;
        for (int j = 0; j < nc - 1; ++j) {
          const float p = next();
          prob.push_back(p);
          remain -= p;
        }
        prob.push_back(remain);
      }
    }
    index.push_back((int)prob.size());
    return true;
  }

  FILE* fp_;
  uint64_t file_size_;
  uint32_t offset_, M_, N_, flags_;
  std::vector<uint8_t> free_data_, cbuf_, buf_;
  std::vector<std::string> sample_;
  bool use_range_;
  std::string rchrom_;
  uint32_t rbeg_, rend_;
  std::string error_;
};

}  // namespace rvtb200
#endif  // RVT_BGEN_H_
