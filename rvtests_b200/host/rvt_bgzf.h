// rvt_bgzf.h -- the output side of `--meta`: a BGZF (blocked gzip) writer and a tabix index builder, header-only C++11 + zlib.
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

}  // namespace rvtb200
#endif  // RVT_BGZF_H_
