// rvt_bgzf.h -- BGZF (blocked gzip) + tabix, header-only C++11 + zlib: the writer and index builder of the `--meta` outputs, and
// (second half) the reader of bgzipped, tabix-indexed input text.
//
// Replaces, for a host that does not link the reference's base/ and third/tabix:
//   FileWriter(fn, BGZIP) = BGZipFileWriter       base/IO.h:645-676 (bgzf_open(fn, "w"), bgzf_write, bgzf_close)
//   ModelManager::create / close / createIndex    src/ModelManager.cpp:285-327: a model with needToIndexResult() (MetaScore,
//                                                 MetaCov, ...) writes "<prefix>.<Model>.assoc.gz" and, after the footnotes,
//   tabixIndexFile(fn)                            src/TabixUtil.cpp:5-15: ti_index_build with {preset 0 (generic), seq col 1,
//                                                 begin col 2, end col 0, meta '#', skip 0} -> "<fn>.tbi"
// File formats (SAM/tabix specifications; third/tabix-0.2.6 inside the reference is the pinned implementation):
//   BGZF  a series of gzip members of <= 64 KiB, each with the extra subfield 'B','C' (len 2) = BSIZE = member size - 1,
//         raw deflate payload, CRC32 and ISIZE; a 28-byte empty member marks the end of the file.  A position is a VIRTUAL
//         OFFSET (file offset of the member << 16 | offset inside its uncompressed payload).
//   TBI   "TBI\1", n_ref, the six configuration words, the NUL-separated names, then per name the binning index
//         (bin -> list of (begin, end) virtual offsets of the chunks holding records of that bin; UCSC bins, 16 kb leaves)
//         and the linear index (per 16 kb window the virtual offset of the first record overlapping it); the file is itself
//         BGZF-compressed.  TabixIndex::build restates ti_index_core (third/tabix-0.2.6/index.c): a chunk is closed when the bin
//         changes, chunks of one bin that touch the same BGZF member are merged, empty windows of the linear index inherit
//         the previous offset, the windows of the first record keep 0.  The reader sees the offset after a record that ends
//         its member as (next member, 0) (bgzf_getline), so offsets are derived from UNCOMPRESSED positions once the member
//         table is known -- the writer need not know, while writing a line, whether the member will grow.
// tests/test_bgzf_tabix.py holds the output against zlib/gzip, and the index against ti_index_build of tabix-0.2.6 compiled
// from the reference's own tarball (oracle/_ref/libtabix_ref.so) on the very file this writer produced.
#ifndef RVT_BGZF_H_
#define RVT_BGZF_H_

#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include <zlib.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

namespace rvtb200 {

class BgzfWriter {
 public:
  struct Member {
    uint64_t ustart;   // uncompressed offset of its first byte
    uint64_t caddr;    // file offset of the member
    uint32_t ulen;
  };
  BgzfWriter() : fp_(NULL), level_(Z_DEFAULT_COMPRESSION), utotal_(0), caddr_(0), ok_(true) {}
  ~BgzfWriter() { close(); }
  bool open(const char* path, int level = -1) {
    close();
    fp_ = fopen(path, "wb");
    level_ = (level < 0 || level > 9) ? Z_DEFAULT_COMPRESSION : level;
    utotal_ = caddr_ = 0;
    members_.clear();
    buf_.clear();
    ok_ = fp_ != NULL;
    return ok_;
  }
  bool write(const void* data, size_t n) {
    if (!fp_) return false;
    const uint8_t* p = (const uint8_t*)data;
    while (n > 0) {
      const size_t room = kPayload - buf_.size(), take = n < room ? n : room;
      buf_.insert(buf_.end(), p, p + take);
      p += take;
      n -= take;
      if (buf_.size() == kPayload) flushMember();
    }
    return ok_;
  }
  uint64_t tell() const { return utotal_ + buf_.size(); }   // UNCOMPRESSED offset of the next byte
  bool close() {
    if (!fp_) return ok_;
    if (!buf_.empty()) flushMember();
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (fwrite(eof, 1, sizeof(eof), fp_) != sizeof(eof)) ok_ = false;
    eof_addr_ = caddr_;
    if (fclose(fp_) != 0) ok_ = false;
    fp_ = NULL;
    return ok_;
  }
  // valid after close(): virtual offset of an uncompressed position, as a BGZF READER reports it
  uint64_t virtualOffset(uint64_t u) const {
    size_t lo = 0, hi = members_.size();
    while (lo < hi) {   // first member with ustart + ulen > u
      const size_t mid = (lo + hi) / 2;
      if (members_[mid].ustart + members_[mid].ulen > u)
        hi = mid;
      else
        lo = mid + 1;
    }
    if (lo == members_.size()) return eof_addr_ << 16;
    return (members_[lo].caddr << 16) | (u - members_[lo].ustart);
  }
  const std::vector<Member>& members() const { return members_; }

 private:
  static const size_t kPayload = 0xff00;   // uncompressed bytes per member: the deflated member always fits 64 KiB
  void flushMember() {
    uint8_t out[0x10000 + 64];
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level_, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
      ok_ = false;
      return;
    }
    zs.next_in = buf_.data();
    zs.avail_in = (uInt)buf_.size();
    zs.next_out = out + 18;
    zs.avail_out = 0x10000 - 18 - 8;
    const int rc = deflate(&zs, Z_FINISH);
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) {   // (incompressible 0xff00 bytes deflate to < 0xff00 + 5 * 2 + ..: cannot happen)
      ok_ = false;
      return;
    }
    const uint32_t clen = (uint32_t)zs.total_out, total = clen + 18 + 8;
    static const uint8_t head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(out, head, 16);
    out[16] = (uint8_t)((total - 1) & 0xff);
    out[17] = (uint8_t)((total - 1) >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), buf_.data(), (uInt)buf_.size()), isize = (uint32_t)buf_.size();
    for (int k = 0; k < 4; ++k) {
      out[18 + clen + k] = (uint8_t)(crc >> (8 * k));
      out[22 + clen + k] = (uint8_t)(isize >> (8 * k));
    }
    if (fwrite(out, 1, total, fp_) != total) ok_ = false;
    Member m = {utotal_, caddr_, isize};
    members_.push_back(m);
    utotal_ += isize;
    caddr_ += total;
    buf_.clear();
  }
  FILE* fp_;
  int level_;
  uint64_t utotal_, caddr_, eof_addr_ = 0;
  bool ok_;
  std::vector<uint8_t> buf_;
  std::vector<Member> members_;
};

struct TabixConf {
  int32_t preset, sc, bc, ec, meta_char, line_skip;
};

class TabixIndex {
 public:
  // tabixIndexFile's defaults (src/TabixUtil.h:6-7)
  explicit TabixIndex(int chromCol = 1, int beginCol = 2, int endCol = 0, char meta = '#', int skip = 0) : lineno_(0) {
    conf_.preset = 0;
    conf_.sc = chromCol;
    conf_.bc = beginCol;
    conf_.ec = endCol;
    conf_.meta_char = meta;
    conf_.line_skip = skip;
  }
  // one complete line (without its '\n') that occupies uncompressed bytes [ubeg, uend) of the data file (uend counts the '\n')
  bool addLine(const char* line, size_t len, uint64_t ubeg, uint64_t uend) {
    Line l;
    l.ubeg = ubeg;
    l.uend = uend;
    l.tid = -1;
    l.beg = l.end = -1;
    ++lineno_;
    if (lineno_ <= (uint64_t)conf_.line_skip || (len > 0 && line[0] == (char)conf_.meta_char)) {
      lines_.push_back(l);
      return true;
    }
    // ti_get_intv, generic preset
    const char *ss = NULL, *se = NULL;
    size_t b = 0;
    int id = 1;
    for (size_t i = 0; i <= len; ++i) {
      if (i == len || line[i] == '\t') {
        if (id == conf_.sc) {
          ss = line + b;
          se = line + i;
        } else if (id == conf_.bc) {
          l.beg = l.end = strtol(std::string(line + b, i - b).c_str(), NULL, 0);
          --l.beg;
          if (l.beg < 0) l.beg = 0;
          if (l.end < 1) l.end = 1;
        } else if (id == conf_.ec) {
          l.end = strtol(std::string(line + b, i - b).c_str(), NULL, 0);
        }
        b = i + 1;
        ++id;
      }
    }
    if (!ss || l.beg < 0 || l.end < 0) {
      error_ = "line " + std::to_string((unsigned long long)lineno_) + " cannot be parsed";
      return false;
    }
    const std::string name(ss, se - ss);
    std::map<std::string, int>::iterator it = tid_.find(name);
    if (it == tid_.end()) {
      l.tid = (int)names_.size();
      tid_[name] = l.tid;
      names_.push_back(name);
    } else
      l.tid = it->second;
    lines_.push_back(l);
    return true;
  }
  // ti_index_core over the recorded lines; `bz` must be closed (its member table is final)
  bool build(const BgzfWriter& bz) {
    const size_t nref = names_.size();
    bins_.assign(nref, std::map<uint32_t, std::vector<Chunk> >());
    lidx_.assign(nref, std::vector<uint64_t>());
    const uint32_t kNone = 0xffffffffu;
    uint32_t last_bin = kNone, save_bin = kNone;
    int last_tid = -1, save_tid = -1;
    long last_coor = -1;
    uint64_t save_off = 0, last_off = 0, end_off = 0;
    bool have0 = false;
    int beg0 = 0, end0 = 0;
    for (size_t k = 0; k < lines_.size(); ++k) {
      const Line& l = lines_[k];
      end_off = bz.virtualOffset(l.uend);
      if (l.tid < 0) {   // skipped / meta line
        last_off = end_off;
        continue;
      }
      if (last_tid != l.tid) {
        if (last_tid > l.tid) return fail("the chromosome blocks are not continuous: is the file sorted?");
        last_tid = l.tid;
        last_bin = kNone;
      } else if (last_coor > l.beg)
        return fail("the file is out of order");
      {   // insert_offset2: linear index
        std::vector<uint64_t>& lx = lidx_[l.tid];
        const int b = (int)(l.beg >> 14), e = (int)((l.end - 1) >> 14);
        if ((int)lx.size() < e + 1) lx.resize(e + 1, 0);
        for (int i = b; i <= e; ++i)
          if (lx[i] == 0) lx[i] = last_off;
        if (last_off == 0) {
          have0 = true;
          beg0 = b;
          end0 = e;
        }
      }
      const uint32_t bin = (uint32_t)reg2bin((uint32_t)l.beg, (uint32_t)l.end);
      if (bin != last_bin) {
        if (save_bin != kNone) addChunk(save_tid, save_bin, save_off, last_off);
        save_off = last_off;
        save_bin = last_bin = bin;
        save_tid = l.tid;
      }
      last_off = end_off;
      last_coor = l.beg;
    }
    if (save_tid >= 0) addChunk(save_tid, save_bin, save_off, bz.virtualOffset(lines_.empty() ? 0 : lines_.back().uend));
    for (size_t t = 0; t < nref; ++t) {
      // merge_chunks: neighbours of one bin that touch the same BGZF member
      for (std::map<uint32_t, std::vector<Chunk> >::iterator it = bins_[t].begin(); it != bins_[t].end(); ++it) {
        std::vector<Chunk>& c = it->second;
        size_t m = 0;
        for (size_t l = 1; l < c.size(); ++l) {
          if ((c[m].v >> 16) == (c[l].u >> 16))
            c[m].v = c[l].v;
          else
            c[++m] = c[l];
        }
        c.resize(m + 1);
      }
      // fill_missing
      for (size_t j = 1; j < lidx_[t].size(); ++j)
        if (lidx_[t][j] == 0) lidx_[t][j] = lidx_[t][j - 1];
    }
    if (have0 && nref > 0 && !lidx_[0].empty())
      for (int i = beg0; i <= end0; ++i) lidx_[0][i] = 0;
    return true;
  }
  // ti_index_save (little endian), BGZF-compressed; bins in ascending order (tabix writes them in hash-table order; readers
  // look them up by number)
  bool save(const char* path) const {
    BgzfWriter out;
    if (!out.open(path)) return false;
    std::vector<uint8_t> b;
    put(&b, "TBI\1", 4);
    put32(&b, (int32_t)names_.size());
    put(&b, &conf_, sizeof(conf_));
    int32_t l = 0;
    for (size_t i = 0; i < names_.size(); ++i) l += (int32_t)names_[i].size() + 1;
    put32(&b, l);
    for (size_t i = 0; i < names_.size(); ++i) put(&b, names_[i].c_str(), names_[i].size() + 1);
    for (size_t t = 0; t < names_.size(); ++t) {
      put32(&b, (int32_t)bins_[t].size());
      for (std::map<uint32_t, std::vector<Chunk> >::const_iterator it = bins_[t].begin(); it != bins_[t].end(); ++it) {
        put32(&b, (int32_t)it->first);
        put32(&b, (int32_t)it->second.size());
        for (size_t c = 0; c < it->second.size(); ++c) {
          put(&b, &it->second[c].u, 8);
          put(&b, &it->second[c].v, 8);
        }
      }
      put32(&b, (int32_t)lidx_[t].size());
      if (!lidx_[t].empty()) put(&b, lidx_[t].data(), 8 * lidx_[t].size());
    }
    out.write(b.data(), b.size());
    return out.close();
  }
  const std::string& error() const { return error_; }

  static int reg2bin(uint32_t beg, uint32_t end) {
    --end;
    if (beg >> 14 == end >> 14) return 4681 + (beg >> 14);
    if (beg >> 17 == end >> 17) return 585 + (beg >> 17);
    if (beg >> 20 == end >> 20) return 73 + (beg >> 20);
    if (beg >> 23 == end >> 23) return 9 + (beg >> 23);
    if (beg >> 26 == end >> 26) return 1 + (beg >> 26);
    return 0;
  }

 private:
  struct Line {
    uint64_t ubeg, uend;
    int tid;
    long beg, end;
  };
  struct Chunk {
    uint64_t u, v;
  };
  bool fail(const char* what) {
    error_ = what;
    return false;
  }
  void addChunk(int tid, uint32_t bin, uint64_t u, uint64_t v) {
    Chunk c = {u, v};
    bins_[tid][bin].push_back(c);
  }
  static void put(std::vector<uint8_t>* b, const void* p, size_t n) { b->insert(b->end(), (const uint8_t*)p, (const uint8_t*)p + n); }
  static void put32(std::vector<uint8_t>* b, int32_t v) { put(b, &v, 4); }
  TabixConf conf_;
  uint64_t lineno_;
  std::vector<Line> lines_;
  std::vector<std::string> names_;
  std::map<std::string, int> tid_;
  std::vector<std::map<uint32_t, std::vector<Chunk> > > bins_;
  std::vector<std::vector<uint64_t> > lidx_;
  std::string error_;
};

// What ModelManager hands a model with needToIndexResult(): FileWriter's write / printf / close, producing
// "<path>" (BGZF) and, at close, "<path>.tbi".  Lines may arrive in pieces.
class IndexedAssocWriter {
 public:
  IndexedAssocWriter() : open_(false), line_start_(0) {}
  explicit IndexedAssocWriter(const char* path) : open_(false), line_start_(0) { open(path); }
  ~IndexedAssocWriter() { close(); }
  bool open(const char* path) {
    path_ = path;
    open_ = bz_.open(path);
    line_.clear();
    line_start_ = 0;
    idx_ = TabixIndex();
    return open_;
  }
  int write(const char* s) {
    const size_t n = strlen(s);
    if (!open_) return -1;
    size_t b = 0;
    for (size_t i = 0; i < n; ++i)
      if (s[i] == '\n') {
        line_.append(s + b, i - b);
        bz_.write(s + b, i - b + 1);
        if (!idx_.addLine(line_.c_str(), line_.size(), line_start_, bz_.tell())) index_error_ = idx_.error();
        line_start_ = bz_.tell();
        line_.clear();
        b = i + 1;
      }
    if (b < n) {
      line_.append(s + b, n - b);
      bz_.write(s + b, n - b);
    }
    return (int)n;
  }
  int printf(const char* fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    const int n = vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (n < 0) return n;
    if ((size_t)n < sizeof(buf)) return write(buf);
    std::vector<char> big((size_t)n + 1);
    va_start(ap, fmt);
    vsnprintf(big.data(), big.size(), fmt, ap);
    va_end(ap);
    return write(big.data());
  }
  // -> 0 ok, -1 data file, -2 index (the data file is complete and readable either way, as after a failed tabixIndexFile)
  int close() {
    if (!open_) return 0;
    open_ = false;
    if (!line_.empty()) {   // a last line without '\n': tabix still indexes it
      if (!idx_.addLine(line_.c_str(), line_.size(), line_start_, bz_.tell())) index_error_ = idx_.error();
      line_.clear();
    }
    if (!bz_.close()) return -1;
    if (!index_error_.empty() || !idx_.build(bz_)) {
      if (index_error_.empty()) index_error_ = idx_.error();
      fprintf(stderr, "rvtests_b200: tabix index failed on file [ %s ]: %s\n", path_.c_str(), index_error_.c_str());
      return -2;
    }
    return idx_.save((path_ + ".tbi").c_str()) ? 0 : -2;
  }
  const std::string& indexError() const { return index_error_; }

 private:
  BgzfWriter bz_;
  TabixIndex idx_;
  bool open_;
  std::string path_, line_, index_error_;
  uint64_t line_start_;
};

// ---- the input side: bgzipped, tabix-indexed text (the reference's --inVcf file.vcf.gz with --rangeList / --setFile) ---------
//   VCFInputFile in range mode      libVcf/VCFInputFile.cpp:20-120 (ti_open, ti_queryi per range, ti_read), base/IO.h BGZipFileReader
//   third/tabix-0.2.6/index.c       ti_index_load, reg2bins, ti_iter_first / ti_iter_read: the bins a region overlaps, their chunks
//                                   with end > the linear-index offset of the region's first 16 kb window, sorted and merged;
//                                   lines are read from each chunk and kept when they overlap the region
// BgzfReader inflates one member at a time (a member is at most 64 KiB either way) and addresses by virtual offset.
class BgzfReader {
 public:
  BgzfReader() : fp_(NULL), addr_(0), next_addr_(0), off_(0), eof_(false) {}
  ~BgzfReader() { close(); }
  bool open(const char* path) {
    close();
    fp_ = fopen(path, "rb");
    addr_ = next_addr_ = 0;
    off_ = 0;
    block_.clear();
    eof_ = false;
    error_.clear();
    return fp_ != NULL;
  }
  void close() {
    if (fp_) fclose(fp_);
    fp_ = NULL;
  }
  bool seek(uint64_t voffset) {
    if (!fp_) return false;
    const uint64_t a = voffset >> 16;
    if (a != addr_ || block_.empty()) {
      next_addr_ = a;
      eof_ = false;
      if (!readMember()) return false;
    }
    off_ = (size_t)(voffset & 0xffff);
    return off_ <= block_.size();
  }
  // as bgzf_tell of a reader that has just consumed up to here: a position at the end of a member is (next member, 0)
  uint64_t tell() const { return off_ >= block_.size() ? next_addr_ << 16 : (addr_ << 16) | off_; }
  // one line without its '\n' (a trailing '\r' is kept, as bgzf_getline does); false at the end of the file
  bool getline(std::string* line) {
    line->clear();
    bool any = false;
    while (true) {
      if (off_ >= block_.size()) {
        if (eof_ || !readMember()) return any;
        if (block_.empty()) continue;   // empty member (the EOF marker, or a flush)
      }
      const char* b = (const char*)block_.data() + off_;
      const char* nl = (const char*)memchr(b, '\n', block_.size() - off_);
      any = true;
      if (nl) {
        line->append(b, nl - b);
        off_ += (size_t)(nl - b) + 1;
        return true;
      }
      line->append(b, block_.size() - off_);
      off_ = block_.size();
    }
  }
  // the whole remaining payload (small files: the .tbi)
  bool readAll(std::vector<uint8_t>* out) {
    out->clear();
    while (true) {
      if (off_ < block_.size()) out->insert(out->end(), block_.begin() + off_, block_.end());
      off_ = block_.size();
      if (eof_ || !readMember()) break;
    }
    return error_.empty();
  }
  const std::string& error() const { return error_; }

 private:
  bool readMember() {
    block_.clear();
    off_ = 0;
    addr_ = next_addr_;
    if (fseeko(fp_, (off_t)addr_, SEEK_SET) != 0) return fail("seek");
    uint8_t h[18];
    const size_t got = fread(h, 1, 18, fp_);
    if (got == 0) {
      eof_ = true;
      return false;
    }
    if (got != 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return fail("not a BGZF member");
    const unsigned xlen = h[10] | (h[11] << 8);
    if (xlen != 6 || h[12] != 'B' || h[13] != 'C') return fail("BGZF member without the BC subfield first");   // as bgzf.c checks
    const unsigned bsize = (h[16] | (h[17] << 8)) + 1u;
    if (bsize < 26) return fail("short BGZF member");
    std::vector<uint8_t> c(bsize - 18);
    if (fread(c.data(), 1, c.size(), fp_) != c.size()) return fail("truncated BGZF member");
    const size_t clen = c.size() - 8;
    const uint32_t isize = c[clen + 4] | (c[clen + 5] << 8) | (c[clen + 6] << 16) | ((uint32_t)c[clen + 7] << 24);
    block_.resize(isize);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return fail("inflateInit2");
    zs.next_in = c.data();
    zs.avail_in = (uInt)clen;
    zs.next_out = block_.data();
    zs.avail_out = isize;
    const int rc = isize ? inflate(&zs, Z_FINISH) : Z_STREAM_END;
    inflateEnd(&zs);
    if (rc != Z_STREAM_END || (isize && zs.total_out != isize)) return fail("inflate");
    next_addr_ = addr_ + bsize;
    return true;
  }
  bool fail(const char* what) {
    error_ = what;
    eof_ = true;
    block_.clear();
    return false;
  }
  FILE* fp_;
  uint64_t addr_, next_addr_;
  size_t off_;
  bool eof_;
  std::vector<uint8_t> block_;
  std::string error_;
};

class TabixReader {
 public:
  // opens <path> and <path>.tbi
  bool open(const char* path) {
    names_.clear();
    refs_.clear();
    if (!data_.open(path)) return fail("cannot open the data file");
    BgzfReader ix;
    std::vector<uint8_t> b;
    if (!ix.open((std::string(path) + ".tbi").c_str()) || !ix.readAll(&b)) return fail("cannot read the .tbi index");
    size_t o = 0;
    if (b.size() < 36 || memcmp(b.data(), "TBI\1", 4) != 0) return fail("wrong magic number in the index");
    const int32_t n_ref = get32(b, 4);
    memcpy(&conf_, b.data() + 8, 24);
    const int32_t l_nm = get32(b, 32);
    o = 36;
    if (o + (size_t)l_nm > b.size()) return fail("truncated index");
    for (size_t i = o, st = o; i < o + (size_t)l_nm; ++i)
      if (b[i] == 0) {
        names_.push_back(std::string((const char*)b.data() + st, i - st));
        st = i + 1;
      }
    o += l_nm;
    refs_.resize(n_ref);
    for (int32_t t = 0; t < n_ref; ++t) {
      if (o + 4 > b.size()) return fail("truncated index");
      const int32_t n_bin = get32(b, o);
      o += 4;
      for (int32_t k = 0; k < n_bin; ++k) {
        if (o + 8 > b.size()) return fail("truncated index");
        const uint32_t bin = (uint32_t)get32(b, o);
        const int32_t n_chunk = get32(b, o + 4);
        o += 8;
        if (o + 16 * (size_t)n_chunk > b.size()) return fail("truncated index");
        std::vector<Chunk>& c = refs_[t].bins[bin];
        c.resize(n_chunk);
        if (n_chunk) memcpy(c.data(), b.data() + o, 16 * (size_t)n_chunk);
        o += 16 * (size_t)n_chunk;
      }
      if (o + 4 > b.size()) return fail("truncated index");
      const int32_t n_intv = get32(b, o);
      o += 4;
      if (o + 8 * (size_t)n_intv > b.size()) return fail("truncated index");
      refs_[t].lidx.resize(n_intv);
      if (n_intv) memcpy(refs_[t].lidx.data(), b.data() + o, 8 * (size_t)n_intv);
      o += 8 * (size_t)n_intv;
    }
    return true;
  }
  const std::vector<std::string>& names() const { return names_; }
  // header = the leading lines that start with the meta character (ti_query before any region: "#..." lines of a VCF)
  bool readHeader(std::vector<std::string>* lines) {
    lines->clear();
    if (!data_.seek(0)) return false;
    std::string l;
    while (true) {
      const uint64_t at = data_.tell();
      if (!data_.getline(&l)) break;
      if (l.empty() || l[0] != (char)conf_.meta_char) {
        data_.seek(at);
        break;
      }
      lines->push_back(l);
    }
    return true;
  }
  // region in the reference's convention: 1-based, inclusive [beg, end] (RangeList); false when the sequence is unknown
  bool query(const std::string& chrom, int beg1, int end1) {
    chunks_.clear();
    cur_ = 0;
    active_ = false;
    tid_ = -1;
    for (size_t i = 0; i < names_.size(); ++i)
      if (names_[i] == chrom) tid_ = (int)i;
    if (tid_ < 0) return false;
    qbeg_ = beg1 > 0 ? beg1 - 1 : 0;   // 0-based half-open
    qend_ = end1 < 1 ? 1 : end1;
    if (qend_ > (1 << 29)) qend_ = 1 << 29;
    if (qbeg_ >= qend_) return true;   // nothing can overlap
    const Ref& r = refs_[tid_];
    uint64_t min_off = 0;
    if (!r.lidx.empty()) {
      const size_t w = (size_t)(qbeg_ >> 14);
      min_off = w >= r.lidx.size() ? r.lidx.back() : r.lidx[w];
    }
    static const int first[6] = {0, 1, 9, 73, 585, 4681}, shift[6] = {29, 26, 23, 20, 17, 14};
    for (int lv = 0; lv < 6; ++lv)
      for (int bn = first[lv] + (qbeg_ >> shift[lv]); bn <= first[lv] + ((qend_ - 1) >> shift[lv]); ++bn) {
        std::map<uint32_t, std::vector<Chunk> >::const_iterator it = r.bins.find((uint32_t)bn);
        if (it == r.bins.end()) continue;
        for (size_t c = 0; c < it->second.size(); ++c)
          if (it->second[c].v > min_off) chunks_.push_back(it->second[c]);
      }
    std::sort(chunks_.begin(), chunks_.end(), chunkLess);
    size_t m = 0;   // merge chunks that overlap or touch
    for (size_t l = 1; l < chunks_.size(); ++l) {
      if (chunks_[l].u <= chunks_[m].v) {
        if (chunks_[l].v > chunks_[m].v) chunks_[m].v = chunks_[l].v;
      } else
        chunks_[++m] = chunks_[l];
    }
    if (!chunks_.empty()) chunks_.resize(m + 1);
    return true;
  }
  // the next line overlapping the queried region, in file order
  bool next(std::string* line) {
    while (cur_ < chunks_.size()) {
      if (!active_) {
        if (!data_.seek(chunks_[cur_].u)) return false;
        active_ = true;
      }
      if (data_.tell() >= chunks_[cur_].v || !data_.getline(line)) {
        ++cur_;
        active_ = false;
        continue;
      }
      if (!line->empty() && (*line)[0] == (char)conf_.meta_char) continue;
      std::string nm;
      long b, e;
      if (!interval(*line, &nm, &b, &e)) continue;
      if (nm != names_[tid_]) continue;
      if (b >= qend_) {   // sorted: nothing further in this chunk can overlap
        ++cur_;
        active_ = false;
        continue;
      }
      if (e > qbeg_) return true;
    }
    return false;
  }
  const std::string& error() const { return error_; }

 private:
  struct Chunk {
    uint64_t u, v;
  };
  struct Ref {
    std::map<uint32_t, std::vector<Chunk> > bins;
    std::vector<uint64_t> lidx;
  };
  static bool chunkLess(const Chunk& a, const Chunk& b) { return a.u < b.u; }
  static int32_t get32(const std::vector<uint8_t>& b, size_t o) {
    int32_t v;
    memcpy(&v, b.data() + o, 4);
    return v;
  }
  bool fail(const char* what) {
    error_ = what;
    return false;
  }
  // ti_get_intv for the generic and VCF presets (0-based half-open)
  bool interval(const std::string& line, std::string* name, long* beg, long* end) const {
    size_t b = 0;
    int id = 1;
    *beg = *end = -1;
    bool have_name = false;
    const int preset = conf_.preset & 0xffff;
    for (size_t i = 0; i <= line.size(); ++i) {
      if (i == line.size() || line[i] == '\t') {
        if (id == conf_.sc) {
          name->assign(line, b, i - b);
          have_name = true;
        } else if (id == conf_.bc) {
          *beg = *end = strtol(line.substr(b, i - b).c_str(), NULL, 0);
          if (!(conf_.preset & 0x10000)) --*beg; else ++*end;
          if (*beg < 0) *beg = 0;
          if (*end < 1) *end = 1;
        } else if (preset == 0 && id == conf_.ec) {
          *end = strtol(line.substr(b, i - b).c_str(), NULL, 0);
        } else if (preset == 2 && id == 4 && b < i) {   // VCF: the REF allele spans the record
          *end = *beg + (long)(i - b);
        }
        b = i + 1;
        ++id;
      }
    }
    return have_name && *beg >= 0 && *end >= 0;
  }
  BgzfReader data_;
  TabixConf conf_;
  std::vector<std::string> names_;
  std::vector<Ref> refs_;
  std::vector<Chunk> chunks_;
  size_t cur_ = 0;
  bool active_ = false;
  int tid_ = -1;
  long qbeg_ = 0, qend_ = 0;
  std::string error_;
};

}  // namespace rvtb200
#endif  // RVT_BGZF_H_
