// oracle/ref_vcf_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
// Drives the REFERENCE's own VCF record parser, compiled unmodified from /root/reference/libVcf
// (VCFRecord.cpp, VCFIndividual.cpp, VCFInfo.cpp, VCFHeader.cpp, VCFBuffer.cpp, VCFValue.cpp, PeopleSet.cpp), the way
// VCFGenotypeExtractor::extractMultipleGenotype does for hard calls outside hemizygous regions with no GD/GQ filter
// (src/VCFGenotypeExtractor.cpp:29-140, 397-439): createIndividual(header) -> parse(line) -> getFormatIndex("GT") ->
// people[i]->justGet(idx).getGenotype().  Output goes to oracle/_ref/libvcf_ref.so only.
#include <string.h>

#include <string>

#include "base/ParRegion.h"
#include "base/RangeList.h"
#include "libVcf/VCFRecord.h"

// BufferedReader (base/IO.cpp needs bzip2/zlib trees that are not built here) is only reached from
// include/excludePeopleFromFile, which this shim never calls
#include <stdlib.h>
static void unreachable(const char* w) {
  fprintf(stderr, "ref_vcf_shim: %s reached\n", w);
  abort();
}
BufferedReader::BufferedReader(const char*, int) { unreachable("BufferedReader"); }
int BufferedReader::readLineBySep(std::vector<std::string>*, const char*) { unreachable("BufferedReader"); return 0; }
bool BufferedReader::isEof() { return true; }
void BufferedReader::close() {}
int BufferedReader::getc() { return EOF; }
int BufferedReader::read(void*, int) { return 0; }
int BufferedReader::readLine(std::string*) { unreachable("BufferedReader"); return 0; }

extern "C" {
// header: the "#CHROM\tPOS..." line; record: one data line (no newline).  out[cap] receives the genotype per sample
// (0/1/2, MISSING_GENOTYPE = -9); chrom (>= 64 bytes) and pos the site.  Returns the number of samples, < 0 on a parse error.
int ref_vcf_genotypes(const char* header, const char* record, int* out, int cap, char* chrom, int* pos) {
  VCFRecord r;
  r.createIndividual(std::string(header));
  r.includeAllPeople();
  std::string line(record);
  if (r.parse(&line)) {
    r.deleteIndividual();
    return -1;
  }
  const int idx = r.getFormatIndex("GT");
  VCFPeople& people = r.getPeople();
  const int n = (int)people.size();
  for (int i = 0; i < n && i < cap; ++i) out[i] = idx >= 0 ? people[i]->justGet(idx).getGenotype() : MISSING_GENOTYPE;
  strncpy(chrom, r.getChrom(), 63);
  chrom[63] = 0;
  *pos = r.getPos();
  r.deleteIndividual();
  return n;
}

// hard calls with the per-individual depth / quality filters of VCFGenotypeExtractor::getGenotype (:429-431).  checkGD /
// checkGQ (:304-317) live in a file that cannot be built here (it pulls in the tabix / BCF readers), so their eight lines
// are restated below ON TOP OF the reference's own parser values (justGet(idx).toInt()); negative min AND max = filter off.
int ref_vcf_genotypes_filtered(const char* header, const char* record, int gd_min, int gd_max, int gq_min, int gq_max, int* out,
                               int cap) {
  VCFRecord r;
  r.createIndividual(std::string(header));
  r.includeAllPeople();
  std::string line(record);
  if (r.parse(&line)) {
    r.deleteIndividual();
    return -1;
  }
  const int idx = r.getFormatIndex("GT");
  const int GDidx = r.getFormatIndex("GD");
  const int GQidx = r.getFormatIndex("GQ");
  const bool needGD = gd_min >= 0 && gd_max >= 0, needGQ = gq_min >= 0 && gq_max >= 0;
  VCFPeople& people = r.getPeople();
  const int n = (int)people.size();
  for (int i = 0; i < n && i < cap; ++i) {
    int g = idx >= 0 ? people[i]->justGet(idx).getGenotype() : MISSING_GENOTYPE;
    if (idx >= 0 && needGD) {
      const int gd = people[i]->justGet(GDidx).toInt();
      if ((gd_min > 0 && gd < gd_min) || (gd_max > 0 && gd > gd_max)) g = MISSING_GENOTYPE;
    }
    if (idx >= 0 && needGQ) {
      const int gq = people[i]->justGet(GQidx).toInt();
      if ((gq_min > 0 && gq < gq_min) || (gq_max > 0 && gq > gq_max)) g = MISSING_GENOTYPE;
    }
    out[i] = g;
  }
  r.deleteIndividual();
  return n;
}

// parseRangeFormat (base/RangeList.cpp:78-125): 0 = parsed
int ref_parse_range(const char* s, char* chrom, unsigned int* beg, unsigned int* end) {
  std::string c;
  const int rc = parseRangeFormat(std::string(s), &c, beg, end);
  strncpy(chrom, c.c_str(), 63);
  chrom[63] = 0;
  return rc;
}

// VCFValue::getMaleNonParGenotype02 on one GT string (libVcf/VCFValue.h:125-142)
int ref_vcf_gt_male02(const char* s, int len) {
  char buf[64];
  if (len > 63) len = 63;
  memcpy(buf, s, len);
  buf[len] = 0;
  VCFValue v(buf, 0, len);
  return v.getMaleNonParGenotype02();
}

// ParRegion::isHemiRegion (base/ParRegion.h) for the --xLabel / --xParRegion strings
int ref_par_is_hemi(const char* xLabel, const char* parRegion, const char* chrom, int pos) {
  ParRegion p(xLabel, parRegion);
  return p.isHemiRegion(chrom, pos) ? 1 : 0;
}

// the sex / hemizygous-region branches of VCFGenotypeExtractor::getGenotype (src/VCFGenotypeExtractor.cpp:404-428), restated
// over the reference's own VCFValue and ParRegion: dosage != 0 -> toDouble (male x 2 in a hemizygous region), else hard
// calls (male: getMaleNonParGenotype02, female: getGenotype, unknown sex: missing).  out: doubles.
int ref_vcf_genotypes_sex(const char* header, const char* record, const char* xLabel, const char* parRegion, const int* sex,
                          const char* dosage_tag, double* out, int cap) {
  VCFRecord r;
  r.createIndividual(std::string(header));
  r.includeAllPeople();
  std::string line(record);
  if (r.parse(&line)) {
    r.deleteIndividual();
    return -1;
  }
  const bool useDosage = dosage_tag && dosage_tag[0];
  const int idx = r.getFormatIndex(useDosage ? dosage_tag : "GT");
  ParRegion par(xLabel, parRegion);
  const bool hemi = par.isHemiRegion(r.getChrom(), r.getPos());
  VCFPeople& people = r.getPeople();
  const int n = (int)people.size();
  for (int i = 0; i < n && i < cap; ++i) {
    double g = MISSING_GENOTYPE;
    if (idx >= 0) {
      VCFIndividual& indv = *people[i];
      if (useDosage) {
        g = indv.justGet(idx).toDouble();
        if (hemi && sex[i] == PLINK_MALE) g *= 2.0;
      } else if (!hemi) {
        g = indv.justGet(idx).getGenotype();
      } else if (sex[i] == PLINK_MALE) {
        g = indv.justGet(idx).getMaleNonParGenotype02();
      } else if (sex[i] == PLINK_FEMALE) {
        g = indv.justGet(idx).getGenotype();
      }
    }
    out[i] = g;
  }
  r.deleteIndividual();
  return n;
}

// --multipleAllele: VCFValue::countAltAllele / countMaleNonParAltAllele2 on one GT string (libVcf/VCFValue.h:180-234)
int ref_vcf_count_alt(const char* s, int len, int alt) {
  char buf[64];
  if (len > 63) len = 63;
  memcpy(buf, s, len);
  buf[len] = 0;
  VCFValue v(buf, 0, len);
  return v.countAltAllele(alt);
}
int ref_vcf_count_male_alt2(const char* s, int len, int alt) {
  char buf[64];
  if (len > 63) len = 63;
  memcpy(buf, s, len);
  buf[len] = 0;
  VCFValue v(buf, 0, len);
  return v.countMaleNonParAltAllele2(alt);
}
// hard-call branches of VCFGenotypeExtractor::getGenotypeForAltAllele (src/VCFGenotypeExtractor.cpp:459-473) for alt allele
// `alt` of one record, over the reference's VCFValue and ParRegion; sex == NULL: nobody is in a hemizygous region.
// n_alt receives the number of comma-separated ALT alleles (parseAltAllele, :393-395).
int ref_vcf_genotypes_alt(const char* header, const char* record, const char* xLabel, const char* parRegion, const int* sex,
                          int alt, int* out, int cap, int* n_alt) {
  VCFRecord r;
  r.createIndividual(std::string(header));
  r.includeAllPeople();
  std::string line(record);
  if (r.parse(&line)) {
    r.deleteIndividual();
    return -1;
  }
  std::vector<std::string> alts;
  stringTokenize(r.getAlt(), ",", &alts);
  *n_alt = (int)alts.size();
  const int idx = r.getFormatIndex("GT");
  ParRegion par(xLabel, parRegion);
  const bool hemi = sex && par.isHemiRegion(r.getChrom(), r.getPos());
  VCFPeople& people = r.getPeople();
  const int n = (int)people.size();
  for (int i = 0; i < n && i < cap; ++i) {
    int g = MISSING_GENOTYPE;
    if (idx >= 0) {
      VCFIndividual& indv = *people[i];
      if (!hemi)
        g = indv.justGet(idx).countAltAllele(alt);
      else if (sex[i] == PLINK_MALE)
        g = indv.justGet(idx).countMaleNonParAltAllele2(alt);
      else if (sex[i] == PLINK_FEMALE)
        g = indv.justGet(idx).countAltAllele(alt);
    }
    out[i] = g;
  }
  r.deleteIndividual();
  return n;
}

// dosage mode (src/VCFGenotypeExtractor.cpp:70-76, 404-406): justGet(getFormatIndex(tag)).toDouble() per sample; no such key
// -> MISSING_GENOTYPE (:434-438)
int ref_vcf_dosages(const char* header, const char* record, const char* tag, double* out, int cap) {
  VCFRecord r;
  r.createIndividual(std::string(header));
  r.includeAllPeople();
  std::string line(record);
  if (r.parse(&line)) {
    r.deleteIndividual();
    return -1;
  }
  const int idx = r.getFormatIndex(tag);
  VCFPeople& people = r.getPeople();
  const int n = (int)people.size();
  for (int i = 0; i < n && i < cap; ++i) out[i] = idx >= 0 ? people[i]->justGet(idx).toDouble() : (double)MISSING_GENOTYPE;
  r.deleteIndividual();
  return n;
}

// VCFValue::getGenotype on one GT string (libVcf/VCFValue.h:74-116)
int ref_vcf_gt(const char* s, int len) {
  char buf[64];
  if (len > 63) len = 63;
  memcpy(buf, s, len);
  buf[len] = 0;
  VCFValue v(buf, 0, len);
  return v.getGenotype();
}
}
