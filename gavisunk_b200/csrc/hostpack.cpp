// hostpack.cpp -- kmer.encode's byte map on the host, 128 bases per AVX2 step (plain C++: g++ compiles this
// file, nvcc never sees the intrinsics).  The map is the one pinned by tests/golden/kat_bytes (nim-kmer 0.2.6,
// call site workflow/src/kmerpos_annot3.nim:88): A/a 0, C/c 1, G/g 2, T/t/U/u 3, bytes 0x01..0x03 themselves,
// everything else 0.  Output: big-endian words of 16 bases (first base in the two top bits), the layout the
// PACKED probe variant reads (probe.cu p_stage_issue).
#include <stdint.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

static uint8_t g_lut[256];
static struct LutInit {
  LutInit() {
    memset(g_lut, 0, sizeof g_lut);
    g_lut['C'] = g_lut['c'] = 1;
    g_lut['G'] = g_lut['g'] = 2;
    g_lut['T'] = g_lut['t'] = g_lut['U'] = g_lut['u'] = 3;
    g_lut[1] = 1; g_lut[2] = 2; g_lut[3] = 3;
  }
} g_lut_init;

static inline uint32_t pack16_lut(const uint8_t* p) {
  uint32_t v = 0;
  for (int i = 0; i < 16; i++) v = (v << 2) | g_lut[p[i]];
  return v;
}

// words [w0, w1) of the batch `ascii[0..n)`; the last word may be partial (zero-padded)
static void pack_scalar(const uint8_t* ascii, uint64_t n, uint32_t* words, uint64_t w0, uint64_t w1) {
  for (uint64_t w = w0; w < w1; w++) {
    const uint64_t b = w * 16;
    if (b + 16 <= n) {
      words[w] = pack16_lut(ascii + b);
    } else {
      uint32_t v = 0;
      for (int i = 0; i < 16; i++) v = (v << 2) | (b + i < n ? g_lut[ascii[b + i]] : 0);
      words[w] = v;
    }
  }
}

#if defined(__x86_64__)
// 128 bases -> eight words per step.  A letter's code is ((x >> 1) ^ (x >> 2)) & 3; the codes are turned back
// into letters with one PSHUFB and compared with the upper-cased input, and a block that holds anything but
// A/C/G/T (N runs, U, control bytes) takes the table instead.
__attribute__((target("avx2"))) static void pack_avx2(const uint8_t* ascii, uint64_t n, uint32_t* words, uint64_t w0,
                                                      uint64_t w1) {
  const __m256i m3 = _mm256_set1_epi8(3);
  const __m256i up = _mm256_set1_epi8((char)0xDF);
  const __m256i acgt = _mm256_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 'A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0,
                                        0, 0, 0, 0, 0, 0);
  const __m256i mul1 = _mm256_set1_epi16(0x0104);      // c[2i] * 4 + c[2i+1]
  const __m256i mul2 = _mm256_set1_epi32(0x00010010);  // n[2i] * 16 + n[2i+1]
  const __m256i bswap = _mm256_setr_epi8(3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12, 3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8,
                                         15, 14, 13, 12);
  const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
  const uint64_t full = n / 16;  // words whose 16 bases all exist
  const uint64_t stop = w1 < full ? w1 : full;
  uint64_t w = w0;
  for (; w + 8 <= stop; w += 8) {
    const uint8_t* p = ascii + w * 16;
    // 2 KiB ahead: +10 % at 16 threads on the bench host (4 KiB pages under nested paging keep the hardware
    // prefetcher short of the next page)
    _mm_prefetch((const char*)(p + 2048), _MM_HINT_T0);
    _mm_prefetch((const char*)(p + 2048 + 64), _MM_HINT_T0);
    __m256i f[4];
    __m256i ok = _mm256_set1_epi8(-1);
    for (int i = 0; i < 4; i++) {
      const __m256i x = _mm256_loadu_si256((const __m256i*)(p + 32 * i));
      f[i] = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(x, 1), _mm256_srli_epi16(x, 2)), m3);
      ok = _mm256_and_si256(ok, _mm256_cmpeq_epi8(_mm256_shuffle_epi8(acgt, f[i]), _mm256_and_si256(x, up)));
    }
    if (__builtin_expect(_mm256_movemask_epi8(ok) != -1, 0)) {
      for (int i = 0; i < 8; i++) words[w + i] = pack16_lut(p + 16 * i);
      continue;
    }
    // 32-bit lanes: four bases in the low byte, first base in its top bits
    for (int i = 0; i < 4; i++) f[i] = _mm256_madd_epi16(_mm256_maddubs_epi16(f[i], mul1), mul2);
    // bytes of both 128-bit halves: [f0 f1 f2 f3] x 4 bytes; byte-swapped they are the big-endian words
    const __m256i b = _mm256_packus_epi16(_mm256_packus_epi32(f[0], f[1]), _mm256_packus_epi32(f[2], f[3]));
    const __m256i v = _mm256_permutevar8x32_epi32(_mm256_shuffle_epi8(b, bswap), order);
    _mm256_storeu_si256((__m256i*)(words + w), v);
  }
  if (w < w1) pack_scalar(ascii, n, words, w, w1);
}
#endif

#if defined(__x86_64__)
// The same map with AVX-512BW: 256 bases -> sixteen words per step (half the instructions per base of the AVX2 path; the
// packers of the host pipeline are compute-bound, one core packs ~6 GB/s with AVX2).
__attribute__((target("avx512f,avx512bw"))) static void pack_avx512(const uint8_t* ascii, uint64_t n, uint32_t* words, uint64_t w0,
                                                                     uint64_t w1) {
  const __m512i m3 = _mm512_set1_epi8(3);
  const __m512i up = _mm512_set1_epi8((char)0xDF);
  const __m512i acgt = _mm512_broadcast_i32x4(_mm_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0));
  const __m512i mul1 = _mm512_set1_epi16(0x0104);      // c[2i] * 4 + c[2i+1]
  const __m512i mul2 = _mm512_set1_epi32(0x00010010);  // n[2i] * 16 + n[2i+1]
  const __m512i bswap = _mm512_broadcast_i32x4(_mm_setr_epi8(3, 2, 1, 0, 7, 6, 5, 4, 11, 10, 9, 8, 15, 14, 13, 12));
  // after the in-lane packs, lane k holds [block 0 | block 1 | block 2 | block 3] x (bases 16k .. 16k+15): word i*4+k sits at 4k+i
  const __m512i order = _mm512_setr_epi32(0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15);
  const uint64_t full = n / 16;
  const uint64_t stop = w1 < full ? w1 : full;
  uint64_t w = w0;
  const bool stream = (((uintptr_t)(words + w0)) & 63) == 0 && stop - w0 >= (1u << 12);  // aligned, large range: streaming stores
  for (; w + 16 <= stop; w += 16) {
    const uint8_t* p = ascii + w * 16;
    _mm_prefetch((const char*)(p + 4096), _MM_HINT_T0);
    _mm_prefetch((const char*)(p + 4096 + 64), _MM_HINT_T0);
    _mm_prefetch((const char*)(p + 4096 + 128), _MM_HINT_T0);
    _mm_prefetch((const char*)(p + 4096 + 192), _MM_HINT_T0);
    __m512i f[4];
    __mmask64 ok = ~(__mmask64)0;
    for (int i = 0; i < 4; i++) {
      const __m512i x = _mm512_loadu_si512((const void*)(p + 64 * i));
      f[i] = _mm512_and_si512(_mm512_xor_si512(_mm512_srli_epi16(x, 1), _mm512_srli_epi16(x, 2)), m3);
      ok &= _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(acgt, f[i]), _mm512_and_si512(x, up));
    }
    if (__builtin_expect(ok != ~(__mmask64)0, 0)) {
      for (int i = 0; i < 16; i++) words[w + i] = pack16_lut(p + 16 * i);
      continue;
    }
    for (int i = 0; i < 4; i++) f[i] = _mm512_madd_epi16(_mm512_maddubs_epi16(f[i], mul1), mul2);
    const __m512i b = _mm512_packus_epi16(_mm512_packus_epi32(f[0], f[1]), _mm512_packus_epi32(f[2], f[3]));
    const __m512i v = _mm512_permutexvar_epi32(order, _mm512_shuffle_epi8(b, bswap));
    // the words are read next by the DMA engine, not by this core: a line-sized streaming store skips the read for ownership
    if (stream) _mm512_stream_si512((__m512i*)(words + w), v);
    else _mm512_storeu_si512((void*)(words + w), v);
  }
  if (stream) _mm_sfence();
  if (w < w1) pack_avx2(ascii, n, words, w, w1);
}
#endif

// exported to fastx.cu / ctx.cu (not part of the C ABI)
extern "C" void gvs_hostpack_range(const uint8_t* ascii, uint64_t n, uint32_t* words, uint64_t w0, uint64_t w1) {
#if defined(__x86_64__)
  static const bool have_avx512 = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx2");
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx512 && w1 - w0 >= 16) {
    pack_avx512(ascii, n, words, w0, w1);
    return;
  }
  if (have_avx2) {
    pack_avx2(ascii, n, words, w0, w1);
    return;
  }
#endif
  pack_scalar(ascii, n, words, w0, w1);
}
extern "C" int gvs_hostpack_simd(void) {
#if defined(__x86_64__)
  if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx2")) return 2;
  return __builtin_cpu_supports("avx2") ? 1 : 0;
#else
  return 0;
#endif
}
