// fastx.cu -- host ingest of FASTA/FASTQ(.gz) read files: the reads of one or more
// chunk files (temp/{sample}/reads/{hap}_{i-of-N}.fq.gz, workflow/rules/tagONT.smk:17) parsed in
// parallel, laid out back to back in ONE (pinned) host buffer + offsets, ready for gvs_reads_set.
//
// Record semantics are those of the `readfq` port the reference's Nim tools use
// (workflow/src/kmerpos_annot3.nim:85, workflow/src/rlen.nim:13; pinned by tests/golden/rlen_b8):
// the name ends at the first white space, multi-line sequences are concatenated, blank lines are
// ignored, trailing CR/LF are stripped, other characters (internal spaces included) are kept, empty
// records are emitted, FASTA and FASTQ records may be mixed, a truncated quality string still
// yields its record.  No CUDA kernel here: decompression and line splitting are host work; pinning
// is what lets gvs_reads_set overlap the PCIe copy with the probe.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

struct FileReads {
  std::vector<u8> seq;
  std::vector<u64> len;
  std::string names;
  std::vector<u64> name_end;
  std::string err;
};

static bool slurp(const char* path, std::vector<u8>& data, std::string& err) {
  gzFile f = gzopen(path, "rb");  // transparent for files that are not gzip
  if (!f) {
    err = std::string("cannot open ") + path;
    return false;
  }
  gzbuffer(f, 1 << 20);
  size_t cap = 1 << 22, n = 0;
  data.resize(cap);
  for (;;) {
    if (n == cap) {
      cap *= 2;
      data.resize(cap);
    }
    size_t want = cap - n;
    if (want > (1u << 30)) want = 1u << 30;
    int r = gzread(f, data.data() + n, (unsigned)want);
    if (r < 0) {
      int en = 0;
      err = std::string(path) + ": " + gzerror(f, &en);
      gzclose(f);
      return false;
    }
    if (r == 0) break;
    n += (size_t)r;
  }
  gzclose(f);
  data.resize(n);
  return true;
}

static inline bool is_ws(u8 c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == 0x0b || c == 0x0c; }

// [b, e) of the next line (without its '\n'); false at end of data
struct Lines {
  const u8* p;
  const u8* end;
  bool next(const u8*& b, const u8*& e) {
    if (p >= end) return false;
    b = p;
    const u8* nl = (const u8*)memchr(p, '\n', (size_t)(end - p));
    if (nl) {
      e = nl;
      p = nl + 1;
    } else {
      e = end;
      p = end;
    }
    return true;
  }
};
static inline const u8* rstrip_crlf(const u8* b, const u8* e) {
  while (e > b && (e[-1] == '\r' || e[-1] == '\n')) e--;
  return e;
}

static void parse(const std::vector<u8>& data, FileReads& out) {
  Lines L{data.data(), data.data() + data.size()};
  const u8 *b = nullptr, *e = nullptr;
  bool have_hdr = false;
  out.seq.reserve(data.size());
  for (;;) {
    if (!have_hdr) {
      bool found = false;
      while (L.next(b, e))
        if (e > b && (*b == '>' || *b == '@')) {
          found = true;
          break;
        }
      if (!found) return;
    }
    have_hdr = false;
    // name = header up to the first white space ("" when the header starts with a blank)
    const u8* hb = b + 1;
    const u8* he = rstrip_crlf(hb, e);
    const u8* ne = hb;
    if (hb < he && *hb != ' ' && *hb != '\t')
      while (ne < he && !is_ws(*ne)) ne++;
    bool all_ws = true;
    for (const u8* q = hb; q < he; q++)
      if (!is_ws(*q)) { all_ws = false; break; }
    if (all_ws) ne = hb;
    out.names.append((const char*)hb, (size_t)(ne - hb));
    out.name_end.push_back(out.names.size());
    const size_t s0 = out.seq.size();
    bool plus = false;
    while (L.next(b, e)) {
      if (e > b && (*b == '@' || *b == '+' || *b == '>')) {
        if (*b == '+') plus = true; else have_hdr = true;
        break;
      }
      const u8* se = rstrip_crlf(b, e);
      out.seq.insert(out.seq.end(), b, se);
    }
    const u64 slen = out.seq.size() - s0;
    out.len.push_back(slen);
    if (plus) {  // FASTQ: skip quality lines until they cover the sequence (at least one line)
      u64 q = 0;
      while (L.next(b, e)) {
        q += (u64)(rstrip_crlf(b, e) - b);
        if (q >= slen) break;
      }
    }
    if (!plus && !have_hdr) return;  // end of data after a FASTA record
  }
}

}  // namespace

// Page-locking host memory costs about as much as parsing into it: the page-locked word arrays of freed batches are
// kept (a few, process-wide) and handed to the next gvs_fastx_read_packed that fits -- a run streams many batches of
// similar size through the same two buffers.
namespace {
struct PinnedCache {
  std::mutex mu;
  std::vector<std::pair<void*, size_t>> free_list;
  static constexpr size_t MAX_KEEP = 3;
  void* take(size_t bytes, size_t* cap) {
    std::lock_guard<std::mutex> l(mu);
    size_t best = free_list.size();
    for (size_t i = 0; i < free_list.size(); i++)
      if (free_list[i].second >= bytes && (best == free_list.size() || free_list[i].second < free_list[best].second)) best = i;
    if (best == free_list.size()) return nullptr;
    void* p = free_list[best].first;
    *cap = free_list[best].second;
    free_list.erase(free_list.begin() + (long)best);
    return p;
  }
  void give(void* p, size_t cap) {
    void* drop = nullptr;
    {
      std::lock_guard<std::mutex> l(mu);
      free_list.emplace_back(p, cap);
      if (free_list.size() > MAX_KEEP) {  // the smallest goes
        size_t w = 0;
        for (size_t i = 1; i < free_list.size(); i++)
          if (free_list[i].second < free_list[w].second) w = i;
        drop = free_list[w].first;
        free_list.erase(free_list.begin() + (long)w);
      }
    }
    if (drop) cudaFreeHost(drop);
  }
};
PinnedCache g_pinned_words;
}  // namespace

struct gvs_fastx_impl {
  u32* words = nullptr;
  bool words_pinned = false;
  size_t words_cap = 0;  // bytes of a page-locked `words` from the cache (0: plain allocation of gvs_fastx_pack)
  std::string names;
  std::vector<u64> name_off;
  std::vector<u64> read_off;
  std::vector<u64> chunk_first;
  u8* seq = nullptr;
  size_t seq_cap = 0;
  bool pinned = false;
};

extern "C" int gvs_fastx_read(const char* const* paths, uint32_t n_files, int threads, int pin, gvs_fastx* out, char* err,
                              uint64_t err_len) {
  auto fail = [&](const std::string& m) {
    if (err && err_len) snprintf(err, (size_t)err_len, "%s", m.c_str());
    return GVS_E_ARG;
  };
  if (!out || (n_files && !paths)) return fail("gvs_fastx_read: null argument");
  memset(out, 0, sizeof(*out));
  std::vector<FileReads> files(n_files);
  if (threads < 1) threads = 1;
  if ((u32)threads > n_files) threads = (int)(n_files ? n_files : 1);
  std::atomic<u32> next{0};
  auto work = [&]() {
    for (;;) {
      u32 i = next.fetch_add(1);
      if (i >= n_files) break;
      std::vector<u8> data;
      if (slurp(paths[i], data, files[i].err)) parse(data, files[i]);
    }
  };
  {
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
  }
  for (u32 i = 0; i < n_files; i++)
    if (!files[i].err.empty()) return fail(files[i].err);
  gvs_fastx_impl* R = new gvs_fastx_impl();
  u64 n_reads = 0, total = 0;
  R->chunk_first.push_back(0);
  for (u32 i = 0; i < n_files; i++) {
    n_reads += files[i].len.size();
    total += files[i].seq.size();
    R->chunk_first.push_back(n_reads);
  }
  R->read_off.reserve(n_reads + 1);
  R->name_off.reserve(n_reads + 1);
  R->read_off.push_back(0);
  R->name_off.push_back(0);
  R->seq_cap = total + 64;  // slack: the probe stages 16-byte vectors
  if (pin && cudaHostAlloc((void**)&R->seq, R->seq_cap, cudaHostAllocDefault) == cudaSuccess) {
    R->pinned = true;
  } else {
    cudaGetLastError();
    R->seq = (u8*)malloc(R->seq_cap);
    if (!R->seq) {
      delete R;
      return fail("gvs_fastx_read: out of host memory");
    }
  }
  std::vector<u64> file_base(n_files + 1, 0);
  for (u32 i = 0; i < n_files; i++) {
    file_base[i + 1] = file_base[i] + files[i].seq.size();
    u64 o = file_base[i];
    for (u64 l : files[i].len) {
      o += l;
      R->read_off.push_back(o);
    }
    u64 nb = R->names.size();
    R->names += files[i].names;
    for (u64 ne : files[i].name_end) R->name_off.push_back(nb + ne);
  }
  {
    std::atomic<u32> nx{0};
    auto copy = [&]() {
      for (;;) {
        u32 i = nx.fetch_add(1);
        if (i >= n_files) break;
        if (!files[i].seq.empty()) memcpy(R->seq + file_base[i], files[i].seq.data(), files[i].seq.size());
        std::vector<u8>().swap(files[i].seq);
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(copy);
    copy();
    for (auto& t : pool) t.join();
  }
  memset(R->seq + total, 0, R->seq_cap - total);
  out->impl = R;
  out->seq = R->seq;
  out->read_off = R->read_off.data();
  out->names = R->names.data();
  out->name_off = R->name_off.data();
  out->chunk_first = R->chunk_first.data();
  out->n_reads = n_reads;
  out->total_bases = total;
  out->n_files = n_files;
  out->pinned = R->pinned ? 1 : 0;
  return 0;
}

// 2-bit codes of kmer.encode (nim-kmer 0.2.6, pinned by tests/golden/kat_bytes): A/a 0, C/c 1, G/g 2,
// T/t/U/u 3, bytes 0x01..0x03 themselves, everything else 0 -- for this path the packing loses nothing.
// The byte map itself lives in hostpack.cpp (AVX2 with a table for blocks that hold anything but ACGT).
extern "C" void gvs_hostpack_range(const uint8_t* ascii, uint64_t n, uint32_t* words, uint64_t w0, uint64_t w1);

extern "C" int gvs_pack_2bit(const uint8_t* ascii, uint64_t n, uint32_t* words, int threads) {
  if ((n && !ascii) || !words) return GVS_E_ARG;
  const u64 nw = (n + 15) / 16;
  if (threads < 1) threads = 1;
  if (threads == 1 || nw < 65536) {
    gvs_hostpack_range(ascii, n, words, 0, nw);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(gvs_hostpack_range, ascii, n, words, nw * t / threads, nw * (t + 1) / threads);
    for (auto& t : pool) t.join();
  }
  return 0;
}

extern "C" int gvs_fastx_pack(gvs_fastx* fx, int threads, int pin) {
  if (!fx || !fx->impl) return GVS_E_ARG;
  gvs_fastx_impl* R = (gvs_fastx_impl*)fx->impl;
  if (!R->words) {
    const u64 nw = (fx->total_bases + 15) / 16;
    const size_t bytes = (nw + 16) * 4;
    if (pin && cudaHostAlloc((void**)&R->words, bytes, cudaHostAllocDefault) == cudaSuccess) {
      R->words_pinned = true;
    } else {
      cudaGetLastError();
      R->words = (u32*)malloc(bytes);
      if (!R->words) return GVS_E_NOMEM;
    }
    memset(R->words + nw, 0, 16 * 4);
    int rc = gvs_pack_2bit(R->seq, fx->total_bases, R->words, threads);
    if (rc) return rc;
  }
  fx->words = R->words;
  fx->n_words = (fx->total_bases + 15) / 16;
  return 0;
}

extern "C" void gvs_fastx_free(gvs_fastx* fx) {
  if (!fx || !fx->impl) return;
  gvs_fastx_impl* R = (gvs_fastx_impl*)fx->impl;
  if (R->words) {
    if (R->words_pinned && R->words_cap) g_pinned_words.give(R->words, R->words_cap);
    else if (R->words_pinned) cudaFreeHost(R->words);
    else free(R->words);
  }
  if (R->seq) {
    if (R->pinned) cudaFreeHost(R->seq); else free(R->seq);
  }
  delete R;
  memset(fx, 0, sizeof(*fx));
}

// ------------------------------------------------------------------------------------------------------------
// Fused parse + 2-bit pack: the ingest's native path.  The ASCII bases never land in memory as a batch: every file
// is streamed block by block (plain files straight out of the page cache through mmap, gzip files through a
// cache-sized inflate window), the readfq state machine above runs as a push parser across block boundaries, and
// the sequence bytes go through kmer.encode's byte map (hostpack.cpp, pinned by tests/golden/kat_bytes) into
// per-file 2-bit words.  The files' word strings are then shifted into ONE page-locked array of big-endian
// 16-base words: the layout gvs_reads_set_packed hands to the PACKED probe variant, a quarter of the ASCII bytes
// on the PCIe link.
// ------------------------------------------------------------------------------------------------------------
namespace {

struct PackedFile {
  std::vector<u32> words;  // 2-bit words of this file's bases, starting at base 0 of the file
  u64 n_bases = 0;
  u8 pend[16];             // bases of the unfinished last word
  u32 n_pend = 0;
  std::vector<u64> len;
  std::string names;
  std::vector<u64> name_end;
  std::string err;

  void append(const u8* b, const u8* e) {
    u64 n = (u64)(e - b);
    if (!n) return;
    n_bases += n;
    if (n_pend) {
      while (n_pend < 16 && b < e) pend[n_pend++] = *b++;
      if (n_pend < 16) return;
      u32 w;
      gvs_hostpack_range(pend, 16, &w, 0, 1);
      words.push_back(w);
      n_pend = 0;
      n = (u64)(e - b);
    }
    const u64 full = n / 16;
    if (full) {
      const size_t at = words.size();
      words.resize(at + full);
      gvs_hostpack_range(b, full * 16, words.data() + at, 0, full);
      b += full * 16;
    }
    while (b < e) pend[n_pend++] = *b++;
  }
  void finish() {
    if (n_pend) {
      u32 w;
      gvs_hostpack_range(pend, n_pend, &w, 0, 1);  // zero-padded partial word
      words.push_back(w);
      n_pend = 0;
    }
  }
};

// readfq as a push parser: feed() takes the file's bytes in arbitrary pieces, end() flushes the last line.
// Lines are classified by their first byte at the moment they start; sequence bytes are packed as they come
// (trailing CRs of a line are held back until the line's end is seen: rstrip), header bytes are collected,
// quality bytes are only counted.
struct FastxPush {
  PackedFile& out;
  enum State { SEEK, SEQ, QUAL } st = SEEK;   // SEEK: before the first header / after a finished FASTQ record
  enum Line { L_NONE, L_SKIP, L_HDR, L_SEQ, L_QUAL, L_PLUS } line = L_NONE;
  bool at_line_start = true;
  std::string hdr;        // bytes of the current header line (after the marker)
  u64 held_cr = 0;        // trailing '\r' of the current sequence / quality line seen so far
  u64 seq_len = 0;        // bases of the current record
  u64 qual = 0;           // quality bytes of the current record (rstripped lines)
  u64 line_len = 0;       // rstripped length of the current quality line so far
  bool in_record = false;

  explicit FastxPush(PackedFile& o) : out(o) {}

  void open_record() {  // header line complete: name = up to the first white space ("" when it starts with a blank)
    size_t he = hdr.size();
    while (he > 0 && (hdr[he - 1] == '\r' || hdr[he - 1] == '\n')) he--;
    size_t ne = 0;
    if (he > 0 && hdr[0] != ' ' && hdr[0] != '\t')
      while (ne < he && !is_ws((u8)hdr[ne])) ne++;
    bool all_ws = true;
    for (size_t q = 0; q < he; q++)
      if (!is_ws((u8)hdr[q])) { all_ws = false; break; }
    if (all_ws) ne = 0;
    out.names.append(hdr.data(), ne);
    out.name_end.push_back(out.names.size());
    hdr.clear();
    seq_len = 0;
    qual = 0;
    in_record = true;
    st = SEQ;
  }
  void close_record() {
    if (in_record) out.len.push_back(seq_len);
    in_record = false;
  }
  void seq_bytes(const u8* b, const u8* e) {
    // hold back the trailing CRs of what we have of the line; CRs followed by other bytes are sequence
    const u8* t = e;
    while (t > b && t[-1] == '\r') t--;
    if (t > b) {
      if (held_cr) {
        static const u8 crs[16] = {'\r', '\r', '\r', '\r', '\r', '\r', '\r', '\r', '\r', '\r', '\r', '\r', '\r', '\r', '\r', '\r'};
        seq_len += held_cr;
        while (held_cr) {
          u64 c = held_cr < 16 ? held_cr : 16;
          out.append(crs, crs + c);
          held_cr -= c;
        }
      }
      out.append(b, t);
      seq_len += (u64)(t - b);
    }
    held_cr += (u64)(e - t);
  }
  void qual_bytes(const u8* b, const u8* e) {
    const u8* t = e;
    while (t > b && t[-1] == '\r') t--;
    if (t > b) {
      line_len += held_cr + (u64)(t - b);
      held_cr = 0;
    }
    held_cr += (u64)(e - t);
  }
  void start_line(u8 c) {
    held_cr = 0;
    line_len = 0;
    switch (st) {
      case SEEK:
        line = (c == '>' || c == '@') ? L_HDR : L_SKIP;
        break;
      case SEQ:
        if (c == '>' || c == '@') {
          close_record();
          line = L_HDR;
        } else if (c == '+') {
          line = L_PLUS;
        } else {
          line = L_SEQ;
        }
        break;
      case QUAL:
        line = L_QUAL;
        break;
    }
  }
  void end_line() {  // the line's '\n' (or the end of the data) has been seen
    switch (line) {
      case L_HDR:
        open_record();
        break;
      case L_PLUS:
        st = QUAL;
        break;
      case L_QUAL:
        qual += line_len;
        if (qual >= seq_len) {  // quality covers the sequence (at least one line was read): record complete
          close_record();
          st = SEEK;
        }
        break;
      default:
        break;
    }
    held_cr = 0;
    line = L_NONE;
    at_line_start = true;
  }
  void feed(const u8* p, const u8* end) {
    while (p < end) {
      if (at_line_start) {
        if (*p == '\n') {  // empty line: never a marker; an empty quality line still counts as a line read
          if (st == QUAL) {
            line = L_QUAL;
            line_len = 0;
            end_line();
          }
          p++;
          continue;
        }
        start_line(*p);
        at_line_start = false;
        if (line == L_HDR) p++;  // the marker itself
        if (p >= end) break;
      }
      const u8* nl = (const u8*)memchr(p, '\n', (size_t)(end - p));
      const u8* e = nl ? nl : end;
      switch (line) {
        case L_HDR: hdr.append((const char*)p, (size_t)(e - p)); break;
        case L_SEQ: seq_bytes(p, e); break;
        case L_QUAL: qual_bytes(p, e); break;
        default: break;
      }
      if (!nl) return;
      end_line();
      p = nl + 1;
    }
  }
  void end() {
    if (!at_line_start) end_line();  // last line without '\n'
    close_record();                  // FASTA record at the end of the data / truncated FASTQ
    out.finish();
  }
};

static bool is_gzip(int fd) {
  u8 m[2] = {0, 0};
  return pread(fd, m, 2, 0) == 2 && m[0] == 0x1f && m[1] == 0x8b;
}

static void ingest_file(const char* path, u64 block, PackedFile& out) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) {
    out.err = std::string("cannot open ") + path;
    return;
  }
  struct stat sb;
  if (fstat(fd, &sb) != 0) {
    out.err = std::string("cannot stat ") + path;
    close(fd);
    return;
  }
  FastxPush P(out);
  if (is_gzip(fd) || !S_ISREG(sb.st_mode)) {
    gzFile f = gzdopen(fd, "rb");  // takes the descriptor over
    if (!f) {
      out.err = std::string("cannot open ") + path;
      close(fd);
      return;
    }
    gzbuffer(f, 1 << 18);
    std::vector<u8> buf(block);  // cache-sized: the inflated bytes are packed while they are still in L2
    for (;;) {
      int r = gzread(f, buf.data(), (unsigned)buf.size());
      if (r < 0) {
        int en = 0;
        out.err = std::string(path) + ": " + gzerror(f, &en);
        gzclose(f);
        return;
      }
      if (r == 0) break;
      P.feed(buf.data(), buf.data() + r);
    }
    gzclose(f);
  } else if (sb.st_size > 0) {
    const u64 size = (u64)sb.st_size;
    void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) {
      out.err = std::string("cannot map ") + path;
      close(fd);
      return;
    }
    madvise(m, size, MADV_SEQUENTIAL);
    out.words.reserve(size / 16 / 2 + 64);
    const u8* p = (const u8*)m;
    for (u64 o = 0; o < size; o += block) P.feed(p + o, p + (o + block < size ? o + block : size));
    munmap(m, size);
    close(fd);
  } else {
    close(fd);
  }
  P.end();
}

}  // namespace

// value of global word g of a file whose base 0 sits at global base B: local words shifted by B % 16 bases
static inline u32 shifted_word(const std::vector<u32>& loc, i64 j, u32 sh) {
  const u32 hi = (j - 1 >= 0 && (u64)(j - 1) < loc.size()) ? loc[(size_t)(j - 1)] : 0u;
  const u32 lo = (j >= 0 && (u64)j < loc.size()) ? loc[(size_t)j] : 0u;
  return sh ? ((hi << (32 - 2 * sh)) | (lo >> (2 * sh))) : lo;
}

extern "C" int gvs_fastx_read_packed(const char* const* paths, uint32_t n_files, int threads, int pin, uint64_t block_bytes,
                                     gvs_fastx* out, char* err, uint64_t err_len) {
  auto fail = [&](const std::string& m) {
    if (err && err_len) snprintf(err, (size_t)err_len, "%s", m.c_str());
    return GVS_E_ARG;
  };
  if (!out || (n_files && !paths)) return fail("gvs_fastx_read_packed: null argument");
  memset(out, 0, sizeof(*out));
  if (block_bytes == 0) block_bytes = 1u << 20;
  auto t_begin = std::chrono::steady_clock::now();
  std::vector<PackedFile> files(n_files);
  if (threads < 1) threads = 1;
  if ((u32)threads > n_files) threads = (int)(n_files ? n_files : 1);
  {
    std::atomic<u32> next{0};
    auto work = [&]() {
      for (;;) {
        u32 i = next.fetch_add(1);
        if (i >= n_files) break;
        ingest_file(paths[i], block_bytes, files[i]);
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
  }
  for (u32 i = 0; i < n_files; i++)
    if (!files[i].err.empty()) return fail(files[i].err);
  const bool timing = getenv("GVS_FASTX_TIMING") != nullptr;
  auto t_parse = std::chrono::steady_clock::now();
  gvs_fastx_impl* R = new gvs_fastx_impl();
  u64 n_reads = 0, total = 0;
  std::vector<u64> base(n_files + 1, 0);
  R->chunk_first.push_back(0);
  for (u32 i = 0; i < n_files; i++) {
    n_reads += files[i].len.size();
    total += files[i].n_bases;
    base[i + 1] = total;
    R->chunk_first.push_back(n_reads);
  }
  R->read_off.reserve(n_reads + 1);
  R->name_off.reserve(n_reads + 1);
  R->read_off.push_back(0);
  R->name_off.push_back(0);
  for (u32 i = 0; i < n_files; i++) {
    u64 o = base[i];
    for (u64 l : files[i].len) {
      o += l;
      R->read_off.push_back(o);
    }
    const u64 nb = R->names.size();
    R->names += files[i].names;
    for (u64 ne : files[i].name_end) R->name_off.push_back(nb + ne);
  }
  const u64 nw = (total + 15) / 16;
  const size_t bytes = (nw + 16) * 4;  // slack: the probe stages whole vectors
  size_t cap = 0;
  if (pin && (R->words = (u32*)g_pinned_words.take(bytes, &cap)) != nullptr) {
    R->words_pinned = true;
    R->words_cap = cap;
  } else if (pin && cudaHostAlloc((void**)&R->words, bytes + bytes / 8, cudaHostAllocDefault) == cudaSuccess) {
    R->words_pinned = true;
    R->words_cap = bytes + bytes / 8;
  } else {
    cudaGetLastError();
    R->words = (u32*)malloc(bytes);
    if (!R->words) {
      delete R;
      return fail("gvs_fastx_read_packed: out of host memory");
    }
  }
  memset(R->words + nw, 0, 16 * 4);
  // every file's words shifted to its place: words that lie entirely inside one file are written in parallel, the
  // words that files share (first / last word of a file) are OR-ed together afterwards
  for (u32 i = 0; i < n_files; i++) {
    if (base[i + 1] == base[i]) continue;
    R->words[base[i] / 16] = 0;
    R->words[(base[i + 1] - 1) / 16] = 0;
  }
  {
    std::atomic<u32> nx{0};
    auto place = [&]() {
      for (;;) {
        u32 i = nx.fetch_add(1);
        if (i >= n_files) break;
        if (base[i + 1] == base[i]) continue;
        const u64 g0 = base[i] / 16, g1 = (base[i + 1] - 1) / 16;  // first / last global word the file touches
        const u32 sh = (u32)(base[i] % 16);
        // interior words: both local words exist, no bounds to check (a vectorisable shift / a plain copy); the few
        // words left over at the file's end take the checked path
        const std::vector<u32>& loc = files[i].words;
        const u64 nl = loc.size();
        const u64 jmax = g1 - g0 < nl ? g1 - g0 : nl;  // j in [1, jmax): loc[j - 1] and loc[j] exist
        u64 g = g0 + 1;
        if (jmax > 1) {
          const u32* L = loc.data();
          u32* W = R->words + g0;
          if (sh == 0) {
            memcpy(W + 1, L + 1, (size_t)(jmax - 1) * 4);
          } else {
            const u32 ls = 32 - 2 * sh, rs = 2 * sh;
            for (u64 j = 1; j < jmax; j++) W[j] = (L[j - 1] << ls) | (L[j] >> rs);
          }
          g = g0 + jmax;
        }
        for (; g < g1; g++) R->words[g] = shifted_word(loc, (i64)(g - g0), sh);
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(place);
    place();
    for (auto& t : pool) t.join();
  }
  for (u32 i = 0; i < n_files; i++) {
    if (base[i + 1] == base[i]) continue;
    const u64 g0 = base[i] / 16, g1 = (base[i + 1] - 1) / 16;
    const u32 sh = (u32)(base[i] % 16);
    R->words[g0] |= shifted_word(files[i].words, 0, sh);
    if (g1 != g0) R->words[g1] |= shifted_word(files[i].words, (i64)(g1 - g0), sh);
    std::vector<u32>().swap(files[i].words);
  }
  if (timing) {
    auto t_end = std::chrono::steady_clock::now();
    fprintf(stderr, "[gvs_fastx_read_packed] parse+pack %.3f s, layout %.3f s (%llu bases, %u files, %d threads)\n",
            std::chrono::duration<double>(t_parse - t_begin).count(), std::chrono::duration<double>(t_end - t_parse).count(),
            (unsigned long long)total, n_files, threads);
  }
  out->impl = R;
  out->seq = nullptr;
  out->read_off = R->read_off.data();
  out->names = R->names.data();
  out->name_off = R->name_off.data();
  out->chunk_first = R->chunk_first.data();
  out->n_reads = n_reads;
  out->total_bases = total;
  out->n_files = n_files;
  out->pinned = R->words_pinned ? 1 : 0;
  out->words = R->words;
  out->n_words = nw;
  return 0;
}

