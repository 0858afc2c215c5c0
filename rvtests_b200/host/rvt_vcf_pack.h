// rvt_vcf_pack.h -- genotype ingestion for the engine: VCF text records -> PLINK 2-bit
// SNP-major rows + the per-variant AF side table, i.e. exactly what rvt_gene_push_bed() takes, without ever building
// the reference's N x M `Matrix` of doubles (24-200 MB per gene at the BASELINE sizes).
//
// What it replaces on the reference side (hard calls, autosomal / non-hemizygous sites, no GD/GQ filter, no dosage tag,
// no multi-allelic expansion -- the default `--inVcf` path of a gene-based run):
//   VCFRecord::parse / parseSite / parseIndividual        libVcf/VCFRecord.h:30-201     (tab-separated columns)
//   VCFRecord::getFormatIndex("GT")                       libVcf/VCFRecord.h:280-306    (PREFIX match, first hit)
//   VCFIndividual::parse + justGet(idx)                   libVcf/VCFIndividual.h:27-59, 95-100 (':'-separated subfields;
//                                                         a column with too few subfields yields an empty value = missing)
//   VCFValue::getGenotype                                 libVcf/VCFValue.h:74-116      (the GT grammar, quirks included)
//   VCFGenotypeExtractor::extractMultipleGenotype         src/VCFGenotypeExtractor.cpp:29-140, getGenotype :397-439
//   GenotypeCounter::add / getAF                          src/GenotypeCounter.h:14-52   (AF = 0.5 * sumAC / nSample, missing
//                                                         calls stay in the denominator)
//   RangeList "chr:beg-end[,chr:beg-end...]" sets          src/Main.cpp --setFile / --rangeList (1-based, inclusive ends)
// Dosage mode (`--dosage TAG`, VCFGenotypeExtractor.cpp:70-76, 404-406): setDosageTag("DS") makes every record contribute the
// TAG subfield read with VCFValue::toDouble (= atof: "." and a truncated column read as 0.0, NOT missing) as one column of an
// N x M column-major double block for rvt_gene_push_f64; negative entries are mean-imputed with the reference's rule first.
// Samples come out in VCF column order restricted to the kept names (VCFRecord::includePeople semantics); the caller
// orders the phenotype accordingly, as DataLoader does.  Missing calls become the .bed code 01 and are mean-imputed on the
// device (DataConsolidator::imputeGenotypeToMean).  Header-only, C++11, no dependency but the C ABI header.
#ifndef RVT_VCF_PACK_H_
#define RVT_VCF_PACK_H_

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <ctype.h>

#include <set>
#include <string>
#include <utility>
#include <vector>

#include "rvtests_b200.h"

namespace rvtb200 {

enum { kVcfMissing = -9 };  // MISSING_GENOTYPE, libVcf/VCFConstant.h:4

// VCFValue::getGenotype (libVcf/VCFValue.h:74-116) on the GT subfield s[0, len).  Reading one past the end yields '\0'
// like the reference's NUL-terminated in-place buffer.
inline int vcfGenotype(const char* s, int len) {
  int p = 0;
  const char c0 = len > 0 ? s[0] : '\0';
  if (c0 == '.') return kVcfMissing;
  if (c0 < '0') return kVcfMissing;   // "Wrong genotype detected. [1]"
  int g = c0 - '0';
  if (g > 1) return kVcfMissing;      // multi-allelic (or any byte above '1')
  p++;
  if (p >= len) return g;             // haploid call
  if (s[p] != '|' && s[p] != '/') return kVcfMissing;
  p++;
  if (p >= len) return kVcfMissing;   // "Wrong genotype length = 2"
  if (s[p] == '.') return kVcfMissing;
  if (s[p] < '0') {
    // "Wrong genotype detected. [2]": the reference reports and carries on WITHOUT adding a second allele
  } else {
    const int a2 = s[p] - '0';
    if (a2 > 1) return kVcfMissing;
    g += a2;
  }
  p++;
  if (p != len) return kVcfMissing;
  return g;
}

// VCFValue::getMaleNonParGenotype02 (libVcf/VCFValue.h:125-142, getAllele1 / getAllele2 :159-179, isHaploid :242): a male's call
// outside the pseudo-autosomal regions of X -- "0" / "1" or a homozygous diploid call -> 0 / 2, everything else missing.
// (getAllele1/2 read a byte below '0' as allele 0 after a REPORT, and look at byte 2 without checking the separator.)
inline int vcfGenotypeMale02(const char* s, int len) {
  const char c0 = len > 0 ? s[0] : '\0';
  if (c0 == '.') return kVcfMissing;
  const int g = c0 < '0' ? 0 : c0 - '0';
  if (len == 1) return g == 0 ? 0 : g == 1 ? 2 : kVcfMissing;
  if (2 >= len) return kVcfMissing;
  const char c2 = s[2];
  if (c2 == '.') return kVcfMissing;
  const int g2 = c2 < '0' ? 0 : c2 - '0';
  if (g == g2) {
    if (g == 0) return 0;
    if (g == 1) return 2;
  }
  return kVcfMissing;
}

// --multipleAllele: VCFValue::countAltAllele(alt) (libVcf/VCFValue.h:180-213): copies of alt allele `alt` (1-based) in the call.
// Other alleles count 0 (so 0/2 is 0 for alt 1 and 1 for alt 2); '.', a wrong separator, a missing second allele or trailing
// bytes -> missing; a non-digit allele is reported and counts 0.  An EMPTY value (truncated sample column) is read as
// missing here -- the reference walks past the end of its one-byte default buffer in that case.
inline int vcfCountAltAllele(const char* s, int len, int alt) {
  if (len <= 0) return kVcfMissing;
  if (s[0] == '.') return kVcfMissing;
  int g = (s[0] - '0' == alt) ? 1 : 0;
  if (len == 1) return g;
  if (s[1] != '|' && s[1] != '/') return kVcfMissing;
  if (len == 2) return kVcfMissing;
  if (s[2] == '.') return kVcfMissing;
  if (!(s[2] < '0' || s[2] > '9')) g += (s[2] - '0' == alt) ? 1 : 0;
  if (len != 3) return kVcfMissing;
  return g;
}

// VCFValue::countMaleNonParAltAllele2(alt) (libVcf/VCFValue.h:214-234): haploid call -> 2 if it is `alt` else 0; diploid
// call with two equal alleles -> 2 if they are `alt` else 0; unequal alleles or a missing allele -> missing
inline int vcfCountMaleAltAllele2(const char* s, int len, int alt) {
  const char c0 = len > 0 ? s[0] : '\0';
  if (c0 == '.') return kVcfMissing;
  const int g = c0 < '0' ? 0 : c0 - '0';
  if (len == 1) return g == alt ? 2 : 0;
  if (2 >= len) return kVcfMissing;
  if (s[2] == '.') return kVcfMissing;
  const int g2 = s[2] < '0' ? 0 : s[2] - '0';
  if (g == g2) return (g == alt ? 1 : 0) + (g2 == alt ? 1 : 0);
  return kVcfMissing;
}

// ParRegion (base/ParRegion.h:18-147): X labels + pseudo-autosomal intervals; hemizygous = on an X label and outside them
class VcfParRegion {
 public:
  VcfParRegion() { init("", ""); }   // Main.cpp passes --xLabel / --xParRegion, both empty by default: X,23 and hg19
  void init(const std::string& xLabel, const std::string& parRegion) {
    label_.clear();
    region_.clear();
    if (xLabel.empty()) {
      label_.insert("X");
      label_.insert("23");
    } else {
      size_t b = 0;
      while (b <= xLabel.size()) {
        size_t e = xLabel.find(',', b);
        if (e == std::string::npos) e = xLabel.size();
        label_.insert(xLabel.substr(b, e - b));
        b = e + 1;
      }
    }
    std::string r;
    for (size_t i = 0; i < parRegion.size(); ++i) r += (char)tolower((unsigned char)parRegion[i]);
    if (r.empty()) r = "hg19";
    if (r == "hg19" || r == "b37" || r == "grch37") {
      add(60001, 2699520);
      add(154931044, 155260560);
    } else if (r == "hg18" || r == "b36" || r == "grch36") {
      add(1, 2709520);
      add(154584238, 154913754);
    } else if (r == "hg38" || r == "b38" || r == "grch38") {
      add(10001, 2781479);
      add(155701383, 156030895);
    } else {
      size_t b = 0;
      while (b <= parRegion.size()) {
        size_t e = parRegion.find(',', b);
        if (e == std::string::npos) e = parRegion.size();
        const std::string piece = parRegion.substr(b, e - b);
        b = e + 1;
        // stringTokenize(piece, "-"): exactly two fields, the first one not empty; an empty second = to the end
        const size_t d = piece.find('-');
        if (d == std::string::npos || piece.find('-', d + 1) != std::string::npos || d == 0) continue;
        const std::string hi = piece.substr(d + 1);
        add(atoi(piece.substr(0, d).c_str()), hi.empty() ? INT32_MAX : atoi(hi.c_str()));
      }
    }
  }
  bool isHemiRegion(const std::string& chrom, int pos) const {
    if (!label_.count(chrom)) return false;
    for (size_t i = 0; i < region_.size(); ++i)
      if (pos >= region_[i].first && pos <= region_[i].second) return false;
    return true;
  }

 private:
  void add(int b, int e) { region_.push_back(std::make_pair(b, e)); }
  std::set<std::string> label_;
  std::vector<std::pair<int, int> > region_;
};

// str2int (base/TypeConversion.cpp:55-135): blanks, an optional '-', blanks, then 1..10 digits (anything after them is
// ignored); false when there is no digit or the value leaves the 32-bit range
inline bool vcfStr2Int(const char* in, int* out) {
  while (*in == ' ') ++in;
  bool neg = false;
  if (*in == '-') {
    neg = true;
    ++in;
  }
  if (*in == '\0') return false;
  while (*in == ' ') ++in;
  size_t len = 0;
  while (in[len] >= '0' && in[len] <= '9') ++len;
  if (len < 1 || len > 10) return false;
  unsigned long v = 0;
  for (size_t i = 0; i < len; ++i) v = v * 10 + (unsigned long)(in[i] - '0');
  if (neg ? v > (unsigned long)INT32_MAX + 1 : v > (unsigned long)INT32_MAX) return false;
  *out = neg ? (int)(-(long)v) : (int)v;
  return true;
}

// parseRangeFormat (base/RangeList.cpp:78-125): "chr:beg-end" (1-based, inclusive); "chr:beg" and "chr:beg-" run to
// 1 << 29 (tabix's constant); a piece without a usable begin, with a negative bound or with beg > end does not conform
inline bool vcfParseRange(const std::string& s, std::string* chr, int* beg, int* end) {
  size_t i = 0;
  chr->clear();
  while (i < s.size() && s[i] != ':') chr->push_back(s[i++]);
  ++i;
  std::string t;
  while (i < s.size() && s[i] != '-') t.push_back(s[i++]);
  int b = 0;
  if (!vcfStr2Int(t.c_str(), &b) || b < 0) return false;
  *beg = b;
  if (i >= s.size()) {
    *end = 1 << 29;
    return true;
  }
  ++i;
  if (i >= s.size()) {
    *end = 1 << 29;
    return true;
  }
  int e = 0;
  if (!vcfStr2Int(s.c_str() + i, &e) || e < 0 || b > e) return false;
  *end = e;
  return true;
}

// a set of ranges as --rangeList / a --setFile line give them: "chr:beg-end[,chr:beg-end...]" (RangeList::addRangeList,
// base/RangeList.cpp:134-150: pieces that do not conform are skipped) or chromosome + bounds (RangeList::addRange)
class VcfRangeSet {
 public:
  void clear() { r_.clear(); }
  bool empty() const { return r_.empty(); }
  size_t size() const { return r_.size(); }
  // returns the number of ranges added (malformed pieces are skipped, as in the reference)
  int add(const std::string& spec) {
    int added = 0;
    size_t b = 0;
    while (b <= spec.size()) {
      size_t e = spec.find(',', b);
      if (e == std::string::npos) e = spec.size();
      Range r;
      if (vcfParseRange(spec.substr(b, e - b), &r.chrom, &r.beg, &r.end)) {
        r_.push_back(r);
        ++added;
      }
      b = e + 1;
    }
    return added;
  }
  void addRange(const std::string& chrom, int beg, int end) {
    Range r;
    r.chrom = chrom;
    r.beg = beg;
    r.end = end;
    r_.push_back(r);
  }
  // the i-th range (for index / region queries: TabixReader::query, BgenReader::setRange)
  const std::string& chrom(size_t i) const { return r_[i].chrom; }
  int begin(size_t i) const { return r_[i].beg; }
  int end(size_t i) const { return r_[i].end; }
  bool contains(const char* chrom, size_t chrom_len, int pos) const {
    for (size_t i = 0; i < r_.size(); ++i)
      if (r_[i].chrom.size() == chrom_len && memcmp(r_[i].chrom.data(), chrom, chrom_len) == 0 && pos >= r_[i].beg && pos <= r_[i].end)
        return true;
    return false;
  }

 private:
  struct Range {
    std::string chrom;
    int beg, end;
  };
  std::vector<Range> r_;
};

// --geneFile (refFlat) and --setFile: gene / set name -> ranges, names in order of first appearance
// (loadGeneFile / loadRangeFile, src/Main.cpp:91-122, 138-173; OrderedMap keeps insertion order)
class GeneRangeMap {
 public:
  size_t size() const { return names_.size(); }
  const std::string& name(size_t i) const { return names_[i]; }
  const VcfRangeSet& ranges(size_t i) const { return sets_[i]; }
  void clear() {
    names_.clear();
    sets_.clear();
  }
  // refFlat lines "gene transcript chrom strand txStart txEnd ...", blank- or tab-separated: the gene's range is
  // chopChr(chrom):txStart-txEnd, several lines of one gene add up.  only: comma-separated gene names to keep ("" = all).
  // A line with fewer than 6 columns stops the reading (the reference logs an error and breaks).  Returns the number of
  // genes, < 0 when the file cannot be opened.
  int loadGeneFile(const std::string& path, const std::string& only = "") {
    const std::set<std::string> keep = makeSet(only);
    std::vector<std::string> fd;
    return eachLine(path, [&](const std::string& line) -> bool {
      split(line, &fd);
      if (fd.size() < 6) return false;
      if (!keep.empty() && !keep.count(fd[0])) return true;
      std::string chr = fd[2];
      if (chr.size() > 3 && (chr[0] == 'c' || chr[0] == 'C') && (chr[1] == 'h' || chr[1] == 'H') && (chr[2] == 'r' || chr[2] == 'R'))
        chr = chr.substr(3);
      slot(fd[0]).addRange(chr, atoi(fd[4].c_str()), atoi(fd[5].c_str()));
      return true;
    });
  }
  // "setName range[,range...]" lines; lines with fewer than 2 columns or an empty column are skipped
  int loadRangeFile(const std::string& path, const std::string& only = "") {
    const std::set<std::string> keep = makeSet(only);
    std::vector<std::string> fd;
    return eachLine(path, [&](const std::string& line) -> bool {
      split(line, &fd);
      if (fd.size() < 2) return true;
      if (!keep.empty() && !keep.count(fd[0])) return true;
      if (fd[0].empty() || fd[1].empty()) return true;
      slot(fd[0]).add(fd[1]);
      return true;
    });
  }

 private:
  VcfRangeSet& slot(const std::string& name) {
    for (size_t i = 0; i < names_.size(); ++i)
      if (names_[i] == name) return sets_[i];
    names_.push_back(name);
    sets_.push_back(VcfRangeSet());
    return sets_.back();
  }
  static std::set<std::string> makeSet(const std::string& csv) {
    std::set<std::string> out;
    size_t b = 0;
    while (b < csv.size()) {
      size_t e = csv.find(',', b);
      if (e == std::string::npos) e = csv.size();
      if (e > b) out.insert(csv.substr(b, e - b));
      b = e + 1;
    }
    return out;
  }
  static void split(const std::string& line, std::vector<std::string>* fd) {   // readLineBySep(&fd, "\t ")
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
    if (!fp) return -1;
    std::string line;
    char buf[65536];
    bool go = true;
    while (go && fgets(buf, sizeof(buf), fp)) {
      line += buf;
      if (line.empty() || line[line.size() - 1] != '\n') continue;
      while (!line.empty() && (line[line.size() - 1] == '\n' || line[line.size() - 1] == '\r')) line.erase(line.size() - 1);
      go = f(line);
      line.clear();
    }
    if (go && !line.empty()) f(line);
    fclose(fp);
    return (int)names_.size();
  }
  std::vector<std::string> names_;
  std::vector<VcfRangeSet> sets_;
};

class VcfGenePacker {
 public:
  VcfGenePacker() : ncol_(0), n_(0), stride_(0), m_(0), multi_(false), freq_min_(0.0), freq_max_(0.0), need_gd_(false), need_gq_(false), gd_min_(0), gd_max_(0), gq_min_(0), gq_max_(0) {}

  // the "#CHROM\tPOS\t...\tFORMAT\tS1\tS2..." line.  keep: names to include (NULL or empty: everyone); samples are
  // emitted in VCF column order (VCFRecord::createIndividual + includePeople, libVcf/VCFRecord.h:203-231).
  // Returns the number of samples kept, < 0 when the line has no sample columns or repeats/misses a kept name.
  int setHeader(const char* line, size_t len, const std::vector<std::string>* keep = NULL) {
    while (len && (line[len - 1] == '\n' || line[len - 1] == '\r')) --len;
    std::set<std::string> want;
    if (keep) want.insert(keep->begin(), keep->end());
    col_to_out_.clear();
    names_.clear();
    size_t b = 0;
    int col = 0;
    while (b <= len) {
      const char* t = (const char*)memchr(line + b, '\t', len - b);
      const size_t e = t ? (size_t)(t - line) : len;
      if (col >= 9) {
        const std::string name(line + b, e - b);
        if (name.empty()) return -2;   // the reference exits on an empty column header
        if (want.empty() || want.count(name)) {
          col_to_out_.push_back((int)names_.size());
          names_.push_back(name);
        } else {
          col_to_out_.push_back(-1);
        }
      }
      ++col;
      b = e + 1;
    }
    ncol_ = (int)col_to_out_.size();
    n_ = (int64_t)names_.size();
    stride_ = (n_ + 3) / 4;
    if ((int64_t)sex_.size() != n_) sex_.clear();   // a sex vector belongs to one kept-sample set: call setSex again after setHeader
    clear();
    if (ncol_ == 0) return -1;
    if (!want.empty() && names_.size() != want.size()) return -3;
    return (int)n_;
  }

  // "" (default): hard calls from GT.  Otherwise the FORMAT key whose value is taken as the dosage.
  void setDosageTag(const std::string& tag) {
    dosage_tag_ = tag;
    clear();
  }
  bool dosageMode() const { return !dosage_tag_.empty(); }
  // --multipleAllele (VCFGenotypeExtractor::extractMultipleGenotype, src/VCFGenotypeExtractor.cpp:44-50, 90-97, 113-122): a
  // record with K comma-separated ALT alleles becomes K variant rows, row a counting the copies of alt allele a; the
  // variant is then named "chrom:posREF/ALTa".  addRecord returns K.  (With a dosage tag the reference warns and keeps one
  // row named after the LAST alt allele; so does this.)
  void setMultiAllelic(bool on) { multi_ = on; }
  // --freqLower / --freqUpper (src/VCFGenotypeExtractor.cpp:98-110): a row whose MAF = min(AF, 1 - AF) (GenotypeCounter::
  // getMAF, AF over ALL kept samples) lies below freq_min or above freq_max is dropped again; a bound <= 0 is off.
  // addRecord then returns the number of rows that stayed.
  void setFreqRange(double freq_min, double freq_max) {
    freq_min_ = freq_min;
    freq_max_ = freq_max;
  }
  // Sex of the KEPT samples in output order (PLINK coding: 1 male, 2 female, anything else unknown) switches on the
  // reference's X handling (VCFGenotypeExtractor::getGenotype, src/VCFGenotypeExtractor.cpp:416-428, 404-415): at a site
  // in a hemizygous region (parRegion()) a male is coded 0 / 2 (vcfGenotypeMale02; a dosage is doubled), a female as
  // usual, unknown sex as missing (dosage: as read).  An empty vector (default) switches it off.
  int setSex(const std::vector<int>& sex) {
    if (!sex.empty() && (int64_t)sex.size() != n_) return -1;
    sex_ = sex;
    return 0;
  }
  VcfParRegion& parRegion() { return par_; }
  // --indvDepthMin/Max, --indvQualMin/Max (VCFGenotypeExtractor::checkGD / checkGQ, src/VCFGenotypeExtractor.cpp:304-333,
  // 429-431): a call whose GD (GQ) subfield, read with atoi, lies below min or above max becomes missing; 0 = no bound on
  // that side, and a negative pair switches the filter off.  With the filter on, a record WITHOUT the key reads 0 for
  // everyone (justGet of index -1 is the empty value), so any min > 0 blanks the whole variant, as in the reference.
  void setDepthFilter(int gd_min, int gd_max) {
    need_gd_ = gd_min >= 0 && gd_max >= 0;
    gd_min_ = gd_min;
    gd_max_ = gd_max;
  }
  void setQualFilter(int gq_min, int gq_max) {
    need_gq_ = gq_min >= 0 && gq_max >= 0;
    gq_min_ = gq_min;
    gq_max_ = gq_max;
  }
  int64_t numSample() const { return n_; }
  int64_t stride() const { return stride_; }
  const std::vector<std::string>& sampleNames() const { return names_; }
  VcfRangeSet& ranges() { return ranges_; }

  // start the next gene: drops the rows collected so far (the range set stays; change it through ranges())
  void clear() {
    m_ = 0;
    rows_.clear();
    af_.clear();
    counts_.clear();
    names_var_.clear();
    dos_.clear();
  }

  // One VCF line.  Returns 1 when the record became the next variant row of the current gene, 0 when it was skipped
  // (meta/header line, or outside the range set when one is given), < 0 on a malformed record (fewer than 9 site
  // columns, or a sample count different from the header's: VCFRecord::parseIndividual returns -1 for both).
  int addRecord(const char* line, size_t len) {
    while (len && (line[len - 1] == '\n' || line[len - 1] == '\r')) --len;
    if (len == 0 || line[0] == '#') return 0;
    if (ncol_ == 0) return -10;
    // the nine site columns
    size_t fb[9], fe[9];
    size_t b = 0;
    for (int c = 0; c < 9; ++c) {
      if (b > len) return -1;
      const char* t = (const char*)memchr(line + b, '\t', len - b);
      if (!t && c < 8) return -1;
      fb[c] = b;
      fe[c] = t ? (size_t)(t - line) : len;
      b = fe[c] + 1;
    }
    if (fe[8] >= len) return -1;   // no sample columns at all
    const int pos = atoi(std::string(line + fb[1], fe[1] - fb[1]).c_str());
    if (!ranges_.empty() && !ranges_.contains(line + fb[0], fe[0] - fb[0], pos)) return 0;
    const bool dosage = dosageMode();
    const int gt = formatIndex(line + fb[8], fe[8] - fb[8], dosage ? dosage_tag_.c_str() : "GT");

    const int gd_idx = need_gd_ ? formatIndex(line + fb[8], fe[8] - fb[8], "GD") : -1;
    const int gq_idx = need_gq_ ? formatIndex(line + fb[8], fe[8] - fb[8], "GQ") : -1;
    const bool filtered = need_gd_ || need_gq_;
    const bool hemi = (int64_t)sex_.size() == n_ && n_ > 0 && par_.isHemiRegion(std::string(line + fb[0], fe[0] - fb[0]), pos);

    // alt alleles to expand (multi-allelic mode), else one pass with alt = 0 = the plain GT grammar
    std::vector<std::string> alts;
    if (multi_) {
      size_t ab = fb[4];
      while (ab <= fe[4]) {
        const char* c = (const char*)memchr(line + ab, ',', fe[4] - ab);
        const size_t ae = c ? (size_t)(c - line) : fe[4];
        alts.push_back(std::string(line + ab, ae - ab));   // empty tokens kept, as stringTokenize keeps them: "A,,T" is three alleles
        ab = ae + 1;
      }
      if (alts.empty()) alts.push_back(std::string());
    }
    const int n_pass = multi_ && !dosage ? (int)alts.size() : 1;
    const size_t b_samples = b, row_first = rows_.size(), dos_first = dos_.size();
    const int m_first = m_;
    int dropped = 0;
    for (int pass = 0; pass < n_pass; ++pass) {
      const int alt = multi_ && !dosage ? pass + 1 : 0;
      b = b_samples;
      const size_t row0 = rows_.size(), dos0 = dos_.size();
      if (dosage)
        dos_.resize(dos0 + (size_t)n_, (double)kVcfMissing);
      else
        rows_.resize(row0 + (size_t)stride_, 0);
      uint8_t* row = dosage ? NULL : &rows_[row0];
      double* drow = dosage ? &dos_[dos0] : NULL;
      double sum_ac = 0.0;
      int cnt[4] = {0, 0, 0, 0};   // hom-ref, het, hom-alt, missing
      const bool plain = gt == 0 && !hemi && !filtered && alt == 0;   // GT first, no X / filter / allele expansion
      int col = 0;
      while (true) {
        if (col >= ncol_) {   // "VCF header have LESS people than VCF content!"
          rollback(row_first, dos_first, m_first);
          return -2;
        }
        size_t e = b;   // sample columns are a few bytes long: a plain scan beats a library call here
        while (e < len && line[e] != '\t') ++e;
        const bool t = e < len;
        const int o = col_to_out_[col];
        if (o >= 0 && dosage) {
          // VCFIndividual::justGet(idx).toDouble(): atof of the subfield; an absent subfield is the empty string = 0.0;
          // no such FORMAT key at all = MISSING_GENOTYPE (VCFGenotypeExtractor.cpp:434-438)
          double g = (double)kVcfMissing;
          if (gt >= 0) {
            size_t sb, se;
            g = subfield(line, b, e, gt, &sb, &se) ? atof(std::string(line + sb, se - sb).c_str()) : 0.0;
            if (hemi && sex_[o] == 1) g *= 2.0;   // imputed male dosages on X lie in [0, 1]
            if (filtered && !passFilters(line, b, e, gd_idx, gq_idx)) g = (double)kVcfMissing;
          }
          drow[o] = g;
          // GenotypeCounter::add (src/GenotypeCounter.h:14-33)
          if (g < 0) {
            ++cnt[3];
          } else if (g < 2.0 / 3) {
            ++cnt[0];
            sum_ac += g;
          } else if (g < 4.0 / 3) {
            ++cnt[1];
            sum_ac += g;
          } else if (g <= 2.0) {
            ++cnt[2];
            sum_ac += g;
          } else {
            ++cnt[3];
          }
        } else if (o >= 0) {
          int g = kVcfMissing;
          if (plain && e - b >= 3 && (b + 3 == e || line[b + 3] == ':') && (line[b] == '0' || line[b] == '1') &&
              (line[b + 1] == '/' || line[b + 1] == '|') && (line[b + 2] == '0' || line[b + 2] == '1')) {
            g = (line[b] - '0') + (line[b + 2] - '0');   // the common case, same value as the general grammar below
          } else if (gt >= 0) {
            // the gt-th ':'-separated subfield; a column with fewer subfields reads as the empty value = missing
            size_t sb, se;
            const bool have = subfield(line, b, e, gt, &sb, &se);
            const char* v = have ? line + sb : "";
            const int vl = have ? (int)(se - sb) : 0;
            if (!hemi || sex_[o] == 2)
              g = alt ? vcfCountAltAllele(v, vl, alt) : vcfGenotype(v, vl);
            else if (sex_[o] == 1)
              g = alt ? vcfCountMaleAltAllele2(v, vl, alt) : vcfGenotypeMale02(v, vl);
            else
              g = kVcfMissing;
            if (filtered && !passFilters(line, b, e, gd_idx, gq_idx)) g = kVcfMissing;
          }
          // .bed codes, sample 0 in the low bits (libVcf/PlinkInputFile.h:206-209): 00 hom-ref, 10 het, 11 hom-alt, 01 missing
          const unsigned code = g == 0 ? 0u : g == 1 ? 2u : g == 2 ? 3u : 1u;
          row[o >> 2] |= (uint8_t)(code << ((o & 3) * 2));
          ++cnt[g == 0 ? 0 : g == 1 ? 1 : g == 2 ? 2 : 3];
        }
        ++col;
        if (!t) break;
        b = e + 1;
      }
      if (col != ncol_) {   // "VCF header have MORE people than VCF content!"
        rollback(row_first, dos_first, m_first);
        return -3;
      }
      // GenotypeCounter::getAF: 0.5 * sumAC / nSample, nSample counting the missing calls too
      if (!dosage) sum_ac = (double)(cnt[1] + 2 * cnt[2]);
      const double af_row = n_ ? 0.5 * sum_ac / (double)n_ : -1.0;
      const double maf = af_row > 0.5 ? 1.0 - af_row : af_row;
      if ((freq_min_ > 0. && freq_min_ > maf) || (freq_max_ > 0. && freq_max_ < maf)) {   // "undo loaded contents"
        rows_.resize(row0);
        dos_.resize(dos0);
        ++dropped;
        continue;
      }
      af_.push_back(af_row);
      for (int k = 0; k < 4; ++k) counts_.push_back(cnt[k]);
      std::string name = std::string(line + fb[0], fe[0] - fb[0]) + ":" + std::string(line + fb[1], fe[1] - fb[1]);
      if (multi_) name += std::string(line + fb[3], fe[3] - fb[3]) + "/" + (dosage ? alts.back() : alts[pass]);
      names_var_.push_back(name);
      ++m_;
    }  // pass
    return n_pass - dropped;
  }

  int numVariant() const { return m_; }
  const uint8_t* rows() const { return rows_.empty() ? NULL : &rows_[0]; }
  const double* af() const { return af_.empty() ? NULL : &af_[0]; }
  // [4 * j + k]: hom-ref, het, hom-alt, missing of variant j
  const int* counts() const { return counts_.empty() ? NULL : &counts_[0]; }
  const std::string& variantName(int j) const { return names_var_[j]; }   // "chrom:pos" (VCFGenotypeExtractor.cpp:113-116)

  // dosage mode: the N x M column-major block (variant j = entries [j * N, (j + 1) * N)), raw as read
  const double* dosages() const { return dos_.empty() ? NULL : &dos_[0]; }

  // DataConsolidator::imputeGenotypeToMean (src/DataConsolidator.cpp:217-245) on the dosage block, in place: in a column
  // holding a negative entry those become 2 * ac / an, with the reference's INTEGER accumulator `int ac; ac += g`
  // (truncation after every addition) over the non-negative entries.  Hard-call rows need none of this: code 01 is
  // imputed on the device.
  void imputeDosagesToMean() {
    for (int j = 0; j < m_; ++j) {
      double* c = &dos_[(size_t)j * (size_t)n_];
      bool any = false;
      int ac = 0, an = 0;
      for (int64_t i = 0; i < n_; ++i) {
        if (c[i] >= 0) {
          ac = (int)(ac + c[i]);
          an += 2;
        } else {
          any = true;
        }
      }
      if (!any) continue;
      const double g = 2.0 * (an == 0 ? 0.0 : 1.0 * ac / an);
      for (int64_t i = 0; i < n_; ++i)
        if (c[i] < 0) c[i] = g;
    }
  }

  // hand the collected gene to the engine (the entry points copy; the packer can be cleared right after)
  int push(rvt_ctx* ctx) {
    if (m_ == 0) return RVT_E_BADARG;
    if (dosageMode()) {
      imputeDosagesToMean();
      return rvt_gene_push_f64(ctx, dosages(), m_, af());
    }
    return rvt_gene_push_bed(ctx, rows(), m_, stride_, af());
  }

 private:
  void rollback(size_t rows_size, size_t dos_size, int m) {
    rows_.resize(rows_size);
    dos_.resize(dos_size);
    af_.resize((size_t)m);
    counts_.resize((size_t)m * 4);
    names_var_.resize((size_t)m);
    m_ = m;
  }
  // the idx-th ':'-separated subfield of the sample column line[b, e) as [*sb, *se); false when the column has fewer
  // subfields or idx < 0 (VCFIndividual::justGet then hands out the empty default value)
  static bool subfield(const char* line, size_t b, size_t e, int idx, size_t* sb_out, size_t* se_out) {
    if (idx < 0) return false;
    size_t sb = b;
    int k = 0;
    while (k < idx) {
      const char* cpos = (const char*)memchr(line + sb, ':', e - sb);
      if (!cpos) return false;
      sb = (size_t)(cpos - line) + 1;
      ++k;
    }
    const char* cpos = (const char*)memchr(line + sb, ':', e - sb);
    *sb_out = sb;
    *se_out = cpos ? (size_t)(cpos - line) : e;
    return true;
  }
  static int subfieldInt(const char* line, size_t b, size_t e, int idx) {
    size_t sb, se;
    if (!subfield(line, b, e, idx, &sb, &se)) return 0;   // atoi("")
    return atoi(std::string(line + sb, se - sb).c_str());
  }
  bool passFilters(const char* line, size_t b, size_t e, int gd_idx, int gq_idx) const {
    if (need_gd_) {
      const int gd = subfieldInt(line, b, e, gd_idx);
      if (gd_min_ > 0 && gd < gd_min_) return false;
      if (gd_max_ > 0 && gd > gd_max_) return false;
    }
    if (need_gq_) {
      const int gq = subfieldInt(line, b, e, gq_idx);
      if (gq_min_ > 0 && gq < gq_min_) return false;
      if (gq_max_ > 0 && gq > gq_max_) return false;
    }
    return true;
  }

  // VCFRecord::getFormatIndex (libVcf/VCFRecord.h:280-306): index of the first FORMAT key that STARTS WITH `key`
  static int formatIndex(const char* f, size_t len, const char* key) {
    const size_t kl = strlen(key);
    size_t b = 0;
    int idx = 0;
    while (b < len) {
      if (kl <= len - b && memcmp(f + b, key, kl) == 0) return idx;
      ++idx;
      const char* c = (const char*)memchr(f + b, ':', len - b);
      if (!c) return -1;
      b = (size_t)(c - f) + 1;
    }
    return -1;
  }

  int ncol_;                      // sample columns in the header
  int64_t n_, stride_;
  int m_;
  std::vector<int> col_to_out_;   // VCF sample column -> output sample index, -1 = not kept
  std::vector<std::string> names_, names_var_;
  std::vector<uint8_t> rows_;
  std::vector<double> dos_;       // dosage mode: M columns of N doubles
  std::string dosage_tag_;
  std::vector<int> sex_;
  VcfParRegion par_;
  bool multi_;
  double freq_min_, freq_max_;
  bool need_gd_, need_gq_;
  int gd_min_, gd_max_, gq_min_, gq_max_;
  std::vector<double> af_;
  std::vector<int> counts_;
  VcfRangeSet ranges_;
};

}  // namespace rvtb200

#endif  // RVT_VCF_PACK_H_
