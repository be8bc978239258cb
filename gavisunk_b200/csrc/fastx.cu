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
#include <zlib.h>

#include <atomic>
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

struct gvs_fastx_impl {
  u32* words = nullptr;
  bool words_pinned = false;
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
    if (R->words_pinned) cudaFreeHost(R->words); else free(R->words);
  }
  if (R->seq) {
    if (R->pinned) cudaFreeHost(R->seq); else free(R->seq);
  }
  delete R;
  memset(fx, 0, sizeof(*fx));
}
