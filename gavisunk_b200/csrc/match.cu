// match.cu -- SUNK matching of reads: the GPU replacement of kmerpos_annot3's main loop
// (workflow/src/kmerpos_annot3.nim:81-97).
//
//   k_tile_index   first read boundary after each 4096-base tile (binary search, tiny)
//   k_probe        ★ hot kernel: 128-bit coalesced loads of ASCII bases -> 2-bit packed tile in
//                  shared memory -> rolling forward / reverse-complement k-mer per thread strip ->
//                  canonical min -> L2-resident blocked-Bloom word -> (rare) exact 32-byte bucket
//                  probe in HBM -> hits appended per tile, in position order inside the tile
//   k_gather_hits  tiles' hit runs copied into global position order (scan of tile counts)
//   k_emit_flags / scans / k_emit_rows
//                  consecutive-(contig,group) suppression across the whole chunk (prevLoc is
//                  never reset between reads, Q4) and the position drift it causes (Q3)
#include "table.cuh"

#define PT 256              // threads per probe block
#define WPT 16              // windows per thread
#define TILE (PT * WPT)     // 4096 window start positions per tile
#define HALO_WORDS 2        // k-1 <= 31 extra bases = 2 packed words
#define BWORDS ((TILE + 64) / 32 + 2)
#define MAXB 32             // boundaries remembered per tile for hit -> read mapping

// error flag bits (ctx->counters[1])
#define FLAG_KEYERROR 1u
#define FLAG_OVERFLOW 2u

// ASCII -> 2-bit codes for 4 bytes at once.  Mapping pinned by probing the reference ELF with every
// byte value (tests/golden/kat_bytes): A/a 0, C/c 1, G/g 2, T/t/U/u 3, bytes 0x01..0x03 map to
// themselves, everything else 0 (nim-kmer 0.2.6 encode, called at kmerpos_annot3.nim:88).
__device__ __forceinline__ u32 codes4(u32 x) {
  u32 f = ((x >> 1) ^ (x >> 2)) & 0x03030303u;  // ACGT(U) -> 0123 for letters
  u32 l = x | 0x20202020u;
  u32 letter = __vcmpeq4(l, 0x63636363u) | __vcmpeq4(l, 0x67676767u) | __vcmpeq4(l, 0x74747474u) |
               __vcmpeq4(l, 0x75757575u);
  u32 low = __vcmpeq4(x & 0xFCFCFCFCu, 0u);      // bytes 0..3
  return (f & letter) | (x & low & 0x03030303u);
}
// 4 bytes of 2-bit codes -> 8 bits (first base in the low bits)
__device__ __forceinline__ u32 squeeze4(u32 c) {
  c = (c | (c >> 6)) & 0x000F000Fu;
  c = (c | (c >> 12)) & 0xFFu;
  return c;
}
__device__ __forceinline__ u32 pack16(uint4 v) {
  return squeeze4(codes4(v.x)) | (squeeze4(codes4(v.y)) << 8) | (squeeze4(codes4(v.z)) << 16) |
         (squeeze4(codes4(v.w)) << 24);
}
__device__ __forceinline__ u32 load_pack16(const u8* __restrict__ seq, u64 g, u64 total) {
  uint4 v = make_uint4(0, 0, 0, 0);
  if (g + 16 <= total) {
    v = __ldg((const uint4*)(seq + g));
  } else if (g < total) {
    u32 w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16 && g + i < total; i++) w[i >> 2] |= (u32)seq[g + i] << (8 * (i & 3));
    v = make_uint4(w[0], w[1], w[2], w[3]);
  }
  return pack16(v);
}

__global__ void __launch_bounds__(256) k_tile_index(const u64* __restrict__ read_off, u64 n_reads, u64 n_tiles,
                                                    u32* __restrict__ tile_first) {
  u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  u64 ts = t * TILE;
  // smallest j in [1, n_reads] with read_off[j] > ts (exists: read_off[n_reads] = total > ts)
  u64 lo = 1, hi = n_reads;
  while (lo < hi) {
    u64 mid = (lo + hi) >> 1;
    if (__ldg(read_off + mid) > ts) hi = mid; else lo = mid + 1;
  }
  tile_first[t] = (u32)lo;
}

struct ProbeParams {
  const u8* __restrict__ seq;
  u64 total;
  const u64* __restrict__ read_off;
  u64 n_reads;
  const u32* __restrict__ tile_first;
  u64 n_tiles;
  TabView tab;
  int k;
  u32* hit_read;
  u32* hit_w;
  u32* hit_row;
  u64 hit_cap;
  unsigned long long* cursor;  // counters[0]
  u32* flags;                  // counters[1]
  u32* tile_cnt;
  u64* tile_off;
};

__global__ void __launch_bounds__(PT, 2) k_probe(const ProbeParams P) {
  __shared__ u32 s_bases[PT + HALO_WORDS + 2];
  __shared__ u32 s_bound[BWORDS];
  __shared__ u32 s_short[BWORDS];
  __shared__ u32 s_bpos[MAXB], s_bidx[MAXB];
  __shared__ u32 s_nb;
  __shared__ u32 s_scan[33];
  __shared__ u64 s_base;

  const int t = threadIdx.x;
  const int k = P.k;
  const u64 kmask = (k >= 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  const int topsh = 2 * (k - 1);
  const u64 vmask = (k >= 2) ? ((1ull << (k - 1)) - 1) : 0ull;

  for (u64 tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    const u64 ts = tile * TILE;
    // ---- 1. stage the tile: one 128-bit load per thread, packed to 2 bits per base ----
    s_bases[t] = load_pack16(P.seq, ts + 16ull * t, P.total);
    if (t < HALO_WORDS + 2) s_bases[PT + t] = (t < HALO_WORDS) ? load_pack16(P.seq, ts + TILE + 16ull * t, P.total) : 0u;
    if (t < BWORDS) {
      s_bound[t] = 0;
      s_short[t] = 0;
    }
    if (t == 0) s_nb = 0;
    __syncthreads();
    // ---- 2. read boundaries inside (ts, ts + TILE + k - 2] ----
    const u32 j0 = __ldg(P.tile_first + tile);
    const u64 limit = ts + TILE + (u64)(k > 1 ? k - 1 : 1);
    for (u64 j = (u64)j0 + t; j <= P.n_reads; j += PT) {
      u64 o = __ldg(P.read_off + j);
      if (o >= limit) break;
      u32 rel = (u32)(o - ts);
      atomicOr(&s_bound[rel >> 5], 1u << (rel & 31));
      if (rel < TILE) {
        u32 slot = atomicAdd(&s_nb, 1u);
        if (slot < MAXB) {
          s_bpos[slot] = rel;
          s_bidx[slot] = (u32)j;
        }
        // Q6: a read of length k-1 still yields one window (its bases + the NUL terminator read as A)
        if (k >= 2 && j < P.n_reads && __ldg(P.read_off + j + 1) - o == (u64)(k - 1))
          atomicOr(&s_short[rel >> 5], 1u << (rel & 31));
      }
    }
    if (t == 0 && k >= 2) {  // the read that starts exactly at the tile start
      u64 o = __ldg(P.read_off + j0 - 1);
      if (o == ts && __ldg(P.read_off + j0) - o == (u64)(k - 1)) atomicOr(&s_short[0], 1u);
    }
    __syncthreads();
    // ---- 3. this thread's WPT windows ----
    const u32 w0 = s_bases[t], w1 = s_bases[t + 1], w2 = s_bases[t + 2];
    const u64 lo = (u64)w0 | ((u64)w1 << 32);
    u64 B, S;
    {
      int idx = t >> 1, sh = (t & 1) * 16;
      u64 b01 = (u64)s_bound[idx] | ((u64)s_bound[idx + 1] << 32);
      u64 s01 = (u64)s_short[idx] | ((u64)s_short[idx + 1] << 32);
      B = b01 >> sh;
      S = s01 >> sh;
      if (sh) B |= (u64)s_bound[idx + 2] << 48;
      S &= 0xFFFFull;
    }
    u64 f = 0, r = 0;
#pragma unroll 1
    for (int i = 0; i < k - 1; i++) {
      u64 b = (i < 32) ? ((lo >> (2 * i)) & 3) : 0;
      f = (f << 2) | b;
      r = (r >> 2) | ((3 - b) << topsh);
    }
    u64 canon[WPT];
    u32 vm = 0;
    const u64 p0 = ts + 16ull * t;
#pragma unroll
    for (int i = 0; i < WPT; i++) {
      int bi = k - 1 + i;  // 0..46
      u64 b = (bi < 32) ? ((lo >> (2 * bi)) & 3) : (u64)((w2 >> (2 * (bi - 32))) & 3);
      f = ((f << 2) | b) & kmask;
      r = (r >> 2) | ((3 - b) << topsh);
      bool ok = (p0 + i < P.total) && (((B >> (i + 1)) & vmask) == 0);
      u64 c = f < r ? f : r;
      if ((S >> i) & 1) {  // bogus window of a (k-1)-long read: last base forced to A
        u64 f2 = f & ~3ull, r2 = r | (3ull << topsh);
        c = f2 < r2 ? f2 : r2;
        ok = true;
      }
      canon[i] = c;
      vm |= (ok ? 1u : 0u) << i;
    }
    if (k >= 32) vm = 0;  // Q2: k = 32 yields no hits in the reference
    // ---- 4. filter words: WPT independent loads in flight per thread ----
    u64 fw[WPT];
#pragma unroll
    for (int i = 0; i < WPT; i++) {
      u64 h = gvs_mix(canon[i]);
      fw[i] = ((vm >> i) & 1) ? __ldg(P.tab.filt + gvs_filt_word(h, P.tab.filt_words)) : 0ull;
    }
    u32 hm = 0;
    u32 hrow[WPT];
#pragma unroll
    for (int i = 0; i < WPT; i++) {
      hrow[i] = 0;
      u64 h = gvs_mix(canon[i]);
      u64 bits = gvs_filt_bits(h);
      if ((fw[i] & bits) == bits) {
        u32 row = tab_lookup(P.tab, canon[i], h);
        if (row == GVS_ROW_MISSING) {
          atomicOr(P.flags, FLAG_KEYERROR);
        } else if (row < GVS_NOHIT) {
          hm |= 1u << i;
          hrow[i] = row;
        }
      }
    }
    // ---- 5. ordered append of the tile's hits ----
    u32 cnt = __popc(hm), tot;
    u32 ex = block_excl_scan(cnt, OpSum(), &tot, s_scan);
    if (t == 0) {
      u64 base = 0;
      if (tot) base = atomicAdd(P.cursor, (unsigned long long)tot);
      s_base = base;
      P.tile_cnt[tile] = tot;
      P.tile_off[tile] = base;
      if (base + tot > P.hit_cap) atomicOr(P.flags, FLAG_OVERFLOW);
    }
    __syncthreads();
    if (hm) {
      u64 o = s_base + ex;
      u32 nb = s_nb;
#pragma unroll
      for (int i = 0; i < WPT; i++) {
        if ((hm >> i) & 1) {
          if (o < P.hit_cap) {
            u32 prel = 16u * t + i;
            u64 p = ts + prel;
            u32 rd;
            if (nb <= MAXB) {
              // last boundary at or before p; equal positions (empty reads) -> the largest index
              rd = j0 - 1;
              u32 bestpos = 0;
              bool any = false;
              for (u32 q = 0; q < nb; q++) {
                u32 bp = s_bpos[q], bj = s_bidx[q];
                if (bp <= prel && (!any || bp > bestpos || (bp == bestpos && bj > rd))) {
                  any = true;
                  bestpos = bp;
                  rd = bj;
                }
              }
            } else {
              u64 l = 1, h2 = P.n_reads;  // first j with off[j] > p
              while (l < h2) {
                u64 mid = (l + h2) >> 1;
                if (__ldg(P.read_off + mid) > p) h2 = mid; else l = mid + 1;
              }
              rd = (u32)(l - 1);
            }
            P.hit_read[o] = rd;
            P.hit_w[o] = (u32)(p - __ldg(P.read_off + rd));
            P.hit_row[o] = hrow[i];
          }
          o++;
        }
      }
    }
    __syncthreads();
  }
}

// tile hit runs -> global position order
__global__ void __launch_bounds__(256) k_gather_hits(const u32* __restrict__ tile_cnt, const u64* __restrict__ tile_off,
                                                     const u64* __restrict__ tile_dst, u64 n_tiles,
                                                     const u32* __restrict__ a0, const u32* __restrict__ a1,
                                                     const u32* __restrict__ a2, u32* b0, u32* b1, u32* b2) {
  u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  u64 nw = ((u64)gridDim.x * blockDim.x) >> 5;
  for (u64 tile = warp; tile < n_tiles; tile += nw) {
    u32 c = tile_cnt[tile];
    if (!c) continue;
    u64 s = tile_off[tile], d = tile_dst[tile];
    for (u32 i = lane; i < c; i += 32) {
      b0[d + i] = a0[s + i];
      b1[d + i] = a1[s + i];
      b2[d + i] = a2[s + i];
    }
  }
}

__device__ __forceinline__ u32 chunk_of(const u64* __restrict__ chunk_first, u32 n_chunks, u32 read) {
  u32 lo = 0, hi = n_chunks;  // last c with chunk_first[c] <= read
  while (hi - lo > 1) {
    u32 mid = (lo + hi) >> 1;
    if (chunk_first[mid] <= read) lo = mid; else hi = mid;
  }
  return lo;
}

// flags: bit0 = emitted (curLoc != prevLoc, nim:92), bit1 = first hit of its read
__global__ void __launch_bounds__(256) k_emit_flags(const u32* __restrict__ hread, const u32* __restrict__ hrow,
                                                    const u32* __restrict__ loc_gidx, u64 n,
                                                    const u64* __restrict__ chunk_first, u32 n_chunks, u8* flags) {
  u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  u32 rd = hread[j];
  bool head = true, emit = true;
  if (j > 0) {
    u32 rp = hread[j - 1];
    head = rp != rd;
    bool same_chunk = !head || chunk_of(chunk_first, n_chunks, rp) == chunk_of(chunk_first, n_chunks, rd);
    if (same_chunk) emit = loc_gidx[hrow[j]] != loc_gidx[hrow[j - 1]];
  }
  flags[j] = (emit ? 1 : 0) | (head ? 2 : 0);
}

__global__ void __launch_bounds__(256) k_emit_rows(const u32* __restrict__ hread, const u32* __restrict__ hw,
                                                   const u32* __restrict__ hrow, const u8* __restrict__ flags,
                                                   const u64* __restrict__ packed_excl, const u32* __restrict__ seg_base,
                                                   u64 n, const u32* __restrict__ loc_contig,
                                                   const u32* __restrict__ loc_start, const u32* __restrict__ loc_group,
                                                   const u32* __restrict__ loc_gidx, u32* r_read, u32* r_pos,
                                                   u32* r_contig, u32* r_start, u32* r_group, u32* r_gidx) {
  u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (!(flags[j] & 1)) return;
  u64 pe = packed_excl[j];
  u32 out = (u32)(pe >> 32);        // emitted rows before j
  u32 supp = (u32)pe;               // suppressed hits before j (whole batch)
  u32 base = seg_base[j] - 1;       // suppressed hits before the first hit of j's read
  u32 row = hrow[j];
  r_read[out] = hread[j];
  r_pos[out] = hw[j] - (supp - base);  // `continue` skips `inc i` (nim:92,96; Q3)
  r_contig[out] = loc_contig[row];
  r_start[out] = loc_start[row];
  r_group[out] = loc_group[row];
  r_gidx[out] = loc_gidx[row];
}


static int match_once(gvs_ctx* ctx, u64 n_tiles, bool* overflow, u64* need) {
  u64* counters = ctx->counters.as<u64>();
  CK(cudaMemsetAsync(counters, 0, 4 * sizeof(u64), ctx->stream));
  ProbeParams P;
  P.seq = ctx->seq;
  P.total = ctx->total_bases;
  P.read_off = ctx->read_off;
  P.n_reads = ctx->n_reads;
  P.tile_first = ctx->tile_first.as<u32>();
  P.n_tiles = n_tiles;
  P.tab.filt = ctx->filt.as<u64>();
  P.tab.filt_words = ctx->filt_words;
  P.tab.keys = ctx->tab_keys.as<u64>();
  P.tab.rows = ctx->tab_rows.as<u32>();
  P.tab.slots = ctx->tab_slots;
  P.k = ctx->k;
  P.hit_read = ctx->hit_read.as<u32>();
  P.hit_w = ctx->hit_w.as<u32>();
  P.hit_row = ctx->hit_row.as<u32>();
  P.hit_cap = ctx->hit_cap;
  P.cursor = (unsigned long long*)counters;
  P.flags = (u32*)(counters + 1);
  P.tile_cnt = ctx->tile_cnt.as<u32>();
  P.tile_off = ctx->tile_off.as<u64>();
  {
    StageTimer tm(ctx, GVS_ST_PROBE);
    u64 grid = (u64)ctx->n_sm * 2;
    if (grid > n_tiles) grid = n_tiles;
    LAUNCH(k_probe, (unsigned)grid, PT, 0, P);
  }
  u64 h[2];
  CKR(read_dev(ctx, counters, h, 2));
  *need = h[0];
  u32 fl = (u32)h[1];
  if (fl & FLAG_KEYERROR)
    return gvs_fail(ctx, GVS_E_KEYERROR, "KeyError: a read matched a db k-mer that has no .loc row (kmerpos_annot3.nim:90)");
  *overflow = (fl & FLAG_OVERFLOW) != 0;
  return 0;
}

extern "C" int gvs_match(gvs_ctx* ctx, uint64_t* n_rows_out) {
  if (!ctx) return GVS_E_ARG;
  if (!ctx->db_ready) return gvs_fail(ctx, GVS_E_STATE, "gvs_match: no database");
  if (!ctx->reads_ready) return gvs_fail(ctx, GVS_E_STATE, "gvs_match: no reads");
  CK(cudaSetDevice(ctx->device));
  ctx->match_ready = ctx->diag_ready = ctx->val_ready = false;
  ctx->rows.n = 0;
  ctx->n_hits = 0;
  u64 total = ctx->total_bases;
  if (n_rows_out) *n_rows_out = 0;
  if (total == 0 || ctx->n_reads == 0) {
    ctx->match_ready = true;
    return 0;
  }
  if (((uintptr_t)ctx->seq & 15) != 0) return gvs_fail(ctx, GVS_E_ARG, "read buffer must be 16-byte aligned");
  u64 n_tiles = cdiv(total, TILE);
  CKR(gvs_reserve(ctx, ctx->tile_first, n_tiles * 4));
  CKR(gvs_reserve(ctx, ctx->tile_cnt, n_tiles * 4));
  CKR(gvs_reserve(ctx, ctx->tile_off, n_tiles * 8));
  CKR(gvs_reserve(ctx, ctx->tile_dst, n_tiles * 8));
  LAUNCH(k_tile_index, (unsigned)cdiv(n_tiles, 256), 256, 0, ctx->read_off, ctx->n_reads, n_tiles, ctx->tile_first.as<u32>());
  // hit capacity: start at 1/32 of the windows (typical density is <1 %), retry once with the
  // exact need if a batch is denser
  if (ctx->hit_cap == 0) ctx->hit_cap = total / 32 + 4096;
  for (int attempt = 0;; attempt++) {
    CKR(gvs_reserve(ctx, ctx->hit_read, ctx->hit_cap * 4));
    CKR(gvs_reserve(ctx, ctx->hit_w, ctx->hit_cap * 4));
    CKR(gvs_reserve(ctx, ctx->hit_row, ctx->hit_cap * 4));
    bool overflow = false;
    u64 need = 0;
    CKR(match_once(ctx, n_tiles, &overflow, &need));
    if (!overflow) {
      ctx->n_hits = need;
      break;
    }
    if (attempt >= 1) return gvs_fail(ctx, GVS_E_OVERFLOW, "hit buffer overflow after resize");
    ctx->hit_cap = need + 4096;
  }
  u64 nh = ctx->n_hits;
  if (nh == 0) {
    ctx->match_ready = true;
    return 0;
  }
  if (nh >= 0xFFFFFFF0ull) return gvs_fail(ctx, GVS_E_OVERFLOW, "more than 2^32 hits in one batch");
  StageTimer tm(ctx, GVS_ST_EMIT);
  // ---- order the hits ----
  CKR(gvs_reserve(ctx, ctx->ohit_read, nh * 4));
  CKR(gvs_reserve(ctx, ctx->ohit_w, nh * 4));
  CKR(gvs_reserve(ctx, ctx->ohit_row, nh * 4));
  {
    const u32* tc = ctx->tile_cnt.as<u32>();
    u64* td = ctx->tile_dst.as<u64>();
    auto f = [tc] __device__(u64 i) -> u64 { return (u64)tc[i]; };
    auto g = [td] __device__(u64 i, u64 ex, u64 v) { td[i] = ex; };
    CKR((device_scan<u64>(ctx, n_tiles, f, g, OpSum(), (u64*)nullptr)));
    u64 grid = cdiv(n_tiles * 32, 256);
    if (grid > (u64)ctx->n_sm * 32) grid = (u64)ctx->n_sm * 32;
    LAUNCH(k_gather_hits, (unsigned)grid, 256, 0, tc, ctx->tile_off.as<u64>(), td, n_tiles, ctx->hit_read.as<u32>(),
           ctx->hit_w.as<u32>(), ctx->hit_row.as<u32>(), ctx->ohit_read.as<u32>(), ctx->ohit_w.as<u32>(),
           ctx->ohit_row.as<u32>());
  }
  // ---- suppression + position drift ----
  CKR(gvs_reserve(ctx, ctx->flags_a, nh));          // u8 flags
  CKR(gvs_reserve(ctx, ctx->flags_b, nh * 8));      // packed exclusive counts
  CKR(gvs_reserve(ctx, ctx->flags_c, nh * 4));      // per-hit base of its read
  u8* fl = ctx->flags_a.as<u8>();
  u64* pex = ctx->flags_b.as<u64>();
  u32* sb = ctx->flags_c.as<u32>();
  LAUNCH(k_emit_flags, (unsigned)cdiv(nh, 256), 256, 0, ctx->ohit_read.as<u32>(), ctx->ohit_row.as<u32>(),
         ctx->loc_gidx.as<u32>(), nh, ctx->chunk_first.as<u64>(), ctx->n_chunks, fl);
  u64* tot = ctx->counters.as<u64>() + 2;
  {
    auto f = [fl] __device__(u64 i) -> u64 { return (fl[i] & 1) ? (1ull << 32) : 1ull; };
    auto g = [pex] __device__(u64 i, u64 ex, u64 v) { pex[i] = ex; };
    CKR((device_scan<u64>(ctx, nh, f, g, OpSum(), tot)));
  }
  {
    // base of the read = suppressed count before its first hit (+1 so that 0 is the identity);
    // the counts are non-decreasing, so an inclusive prefix max picks the latest head
    auto f = [fl, pex] __device__(u64 i) -> u32 { return (fl[i] & 2) ? (u32)pex[i] + 1u : 0u; };
    auto g = [sb] __device__(u64 i, u32 ex, u32 v) { sb[i] = ex > v ? ex : v; };
    CKR((device_scan<u32>(ctx, nh, f, g, OpMax(), (u32*)nullptr)));
  }
  u64 packed_total = 0;
  CKR(read_dev(ctx, tot, &packed_total));
  u64 n_rows = packed_total >> 32;
  CKR(gvs_reserve_rows(ctx, ctx->rows, n_rows));
  LAUNCH(k_emit_rows, (unsigned)cdiv(nh, 256), 256, 0, ctx->ohit_read.as<u32>(), ctx->ohit_w.as<u32>(),
         ctx->ohit_row.as<u32>(), fl, pex, sb, nh, ctx->loc_contig.as<u32>(), ctx->loc_start.as<u32>(),
         ctx->loc_group.as<u32>(), ctx->loc_gidx.as<u32>(), ctx->rows.read.as<u32>(), ctx->rows.pos.as<u32>(),
         ctx->rows.contig.as<u32>(), ctx->rows.start.as<u32>(), ctx->rows.group.as<u32>(), ctx->rows.gidx.as<u32>());
  ctx->rows.n = n_rows;
  ctx->match_ready = true;
  if (n_rows_out) *n_rows_out = n_rows;
  return 0;
}

extern "C" int gvs_rows_get(gvs_ctx* ctx, int which, uint32_t* read_idx, uint32_t* pos, uint32_t* contig,
                            uint32_t* start, uint32_t* group) {
  if (!ctx) return GVS_E_ARG;
  CK(cudaSetDevice(ctx->device));
  Rows* r;
  if (which == 0) {
    if (!ctx->match_ready) return gvs_fail(ctx, GVS_E_STATE, "gvs_rows_get(0) before gvs_match");
    r = &ctx->rows;
  } else if (which == 1) {
    if (!ctx->diag_ready) return gvs_fail(ctx, GVS_E_STATE, "gvs_rows_get(1) before gvs_diag_filter");
    r = &ctx->kept;
  } else {
    return gvs_fail(ctx, GVS_E_ARG, "which must be 0 or 1");
  }
  u64 n = r->n;
  if (n == 0) return 0;
  if (read_idx) CK(cudaMemcpyAsync(read_idx, r->read.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (pos) CK(cudaMemcpyAsync(pos, r->pos.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (contig) CK(cudaMemcpyAsync(contig, r->contig.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (start) CK(cudaMemcpyAsync(start, r->start.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (group) CK(cudaMemcpyAsync(group, r->group.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
