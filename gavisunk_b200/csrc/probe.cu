// probe.cu -- ★ the hot kernel: every k-mer window of every read against the SUNK table
// (inner loop of workflow/src/kmerpos_annot3.nim:85-96; nim-kmer 0.2.6 encode/slide semantics,
// SURVEY.md Q1/Q2/Q6).
//
// Warp-synchronous design (no block barriers; a warp that takes the rare exact-probe path never
// stalls its neighbours):
//   * each warp owns a contiguous span of 512-window tiles and walks it with a private cursor into
//     read_off (no per-tile binary search)
//   * per tile: ONE coalesced 128-bit load per lane (+2 lanes of halo) of ASCII bases, packed to
//     2 bits per base twice in shared memory: forward (big-endian) and reverse-complement tile, so
//     that both strands of a window are constant-shift funnel extractions (no rolling state, no
//     warm-up of k-1 bases per strip)
//   * canonical = min(fwd, rc); 32-bit hash -> one 4-byte word of the L2-resident blocked Bloom
//     filter per window, 16 independent loads in flight per lane
//   * filter-positive windows (~1 %) are compacted into a per-warp queue in position order and
//     looked up in the exact table (32-byte bucket in HBM) 32 at a time
//   * hits are appended to the global hit list with one atomicAdd per tile; (count, offset) per
//     tile lets a later pass restore global position order
#include "table.cuh"

#define PW_WARPS 8                       // warps per block
#define PW_TILE 512                      // window starts per warp tile
#define PW_NW 34                         // packed words per tile (512 + 32 bases)
#define PW_MAXB 32

#define FLAG_KEYERROR 1u
#define FLAG_OVERFLOW 2u

// ASCII -> 2-bit codes for 4 bytes at once (mapping pinned by tests/golden/kat_bytes: A/a 0, C/c 1,
// G/g 2, T/t/U/u 3, bytes 0x01..0x03 themselves, everything else 0)
__device__ __forceinline__ u32 p_codes4(u32 x) {
  u32 f = ((x >> 1) ^ (x >> 2)) & 0x03030303u;
  u32 l = x | 0x20202020u;
  u32 letter = __vcmpeq4(l, 0x63636363u) | __vcmpeq4(l, 0x67676767u) | __vcmpeq4(l, 0x74747474u) |
               __vcmpeq4(l, 0x75757575u);
  u32 low = __vcmpeq4(x & 0xFCFCFCFCu, 0u);
  return (f & letter) | (x & low & 0x03030303u);
}
// 4 code bytes (first base in byte 0) -> 8 bits big-endian (first base in bits 7:6)
__device__ __forceinline__ u32 p_be8(u32 c) { return (c * 0x40100401u) >> 24; }
__device__ __forceinline__ u32 p_pack16_be(uint4 v) {
  return (p_be8(p_codes4(v.x)) << 24) | (p_be8(p_codes4(v.y)) << 16) | (p_be8(p_codes4(v.z)) << 8) | p_be8(p_codes4(v.w));
}
// reverse complement of 16 packed bases
__device__ __forceinline__ u32 p_rc16(u32 w) {
  u32 x = __brev(~w);
  return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}
__device__ __forceinline__ u32 p_load16(const u8* __restrict__ seq, u64 g, u64 total) {
  uint4 v = make_uint4(0, 0, 0, 0);
  if (g + 16 <= total) {
    v = __ldg((const uint4*)(seq + g));
  } else if (g < total) {
    u32 w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16 && g + i < total; i++) w[i >> 2] |= (u32)seq[g + i] << (8 * (i & 3));
    v = make_uint4(w[0], w[1], w[2], w[3]);
  }
  return p_pack16_be(v);
}
// 2K bits starting at base offset `off` (0..15) of the big-endian string a:b:c
template <int K>
__device__ __forceinline__ u64 p_extract(u32 a, u32 b, u32 c, int off) {
  int s = 2 * off;
  u32 hi = __funnelshift_l(b, a, s);
  u32 lo = __funnelshift_l(c, b, s);
  u64 v = ((u64)hi << 32) | lo;
  return v >> (64 - 2 * K);
}
// canonical k-mer of window p (0..511) from the shared tiles, optionally with the last base forced
// to A (the bogus window of a (k-1)-long read, Q6)
template <int K>
__device__ __forceinline__ u64 p_canon_at(const u32* fw, const u32* rc, u32 p, bool force_last_a) {
  u32 wi = p >> 4;
  u64 f = p_extract<K>(fw[wi], fw[wi + 1], fw[wi + 2], p & 15);
  u32 q = PW_NW * 16 - p - K;
  u32 qi = q >> 4;
  u64 r = p_extract<K>(rc[qi], rc[qi + 1], rc[qi + 2], q & 15);
  if (force_last_a) {
    f &= ~3ull;
    r |= 3ull << (2 * (K - 1));
  }
  return f < r ? f : r;
}

struct Probe2Params {
  const u8* __restrict__ seq;
  u64 total;
  const u64* __restrict__ read_off;
  u64 n_reads;
  u64 n_tiles;
  u64 tiles_per_warp;
  const u32* __restrict__ filt;
  u32 filt_mask;
  TabView tab;
  u32* hit_read;
  u32* hit_w;
  u32* hit_row;
  u64 hit_cap;
  unsigned long long* cursor;
  u32* flags;
  u32* tile_cnt;
  u64* tile_off;
};

struct WarpSmem {
  u32 fw[PW_NW + 4];
  u32 rc[PW_NW + 4];
  u32 bound[20];
  u32 shrt[20];
  u32 bpos[PW_MAXB];
  u32 bidx[PW_MAXB];
  u32 q_row[PW_TILE];
  u16 q_p[PW_TILE];
};

template <int K>
__global__ void __launch_bounds__(PW_WARPS * 32, 4) k_probe2(const Probe2Params P) {
  __shared__ WarpSmem sm_all[PW_WARPS];
  const int lane = threadIdx.x & 31;
  WarpSmem& sm = sm_all[threadIdx.x >> 5];
  const u64 warp = (u64)blockIdx.x * PW_WARPS + (threadIdx.x >> 5);
  u64 tile = warp * P.tiles_per_warp;
  u64 tile_end = tile + P.tiles_per_warp;
  if (tile_end > P.n_tiles) tile_end = P.n_tiles;
  if (tile >= tile_end) return;
  constexpr int CMAX = (K + 15 + 15) / 16;
  constexpr u32 LT_MASK_ALL = 0xFFFFFFFFu;

  // cursor: first boundary index j >= 1 with read_off[j] > tile start
  u64 cur;
  {
    u64 ts0 = tile * PW_TILE;
    u64 lo = 1, hi = P.n_reads;
    while (lo < hi) {
      u64 mid = (lo + hi) >> 1;
      if (__ldg(P.read_off + mid) > ts0) hi = mid; else lo = mid + 1;
    }
    cur = lo;
  }
  u64 next_off = __ldg(P.read_off + cur);
  if (lane < 4) { sm.rc[PW_NW + lane] = 0; sm.fw[PW_NW + lane] = 0; }

  for (; tile < tile_end; tile++) {
    const u64 ts = tile * PW_TILE;
    // ---- stage: ASCII -> forward and reverse-complement packed tiles ----
    {
      u32 w = p_load16(P.seq, ts + 16ull * lane, P.total);
      sm.fw[lane] = w;
      sm.rc[PW_NW - 1 - lane] = p_rc16(w);
      if (lane < 2) {
        u32 w2 = p_load16(P.seq, ts + 512 + 16ull * lane, P.total);
        sm.fw[32 + lane] = w2;
        sm.rc[PW_NW - 1 - 32 - lane] = p_rc16(w2);
      }
    }
    // ---- read boundaries in (ts, ts + 512 + K - 2] ----
    const u64 limit = ts + PW_TILE + (K > 1 ? K - 1 : 1);
    const u64 cur0 = cur;
    u32 nb = 0;
    bool has_bound = next_off < limit;
    bool start_short = false;
    if (has_bound) {
      if (lane < 20) { sm.bound[lane] = 0; sm.shrt[lane] = 0; }
      __syncwarp();
      u64 j = cur + lane;
      for (;;) {
        u64 o = (j <= P.n_reads) ? __ldg(P.read_off + j) : ~0ull;
        bool inr = o < limit;
        bool intile = inr && (o - ts) < PW_TILE;
        u32 bal_tile = __ballot_sync(LT_MASK_ALL, intile);
        if (inr) {
          u32 rel = (u32)(o - ts);
          atomicOr(&sm.bound[rel >> 5], 1u << (rel & 31));
          if (intile) {
            u32 slot = nb + __popc(bal_tile & ((1u << lane) - 1));
            if (slot < PW_MAXB) { sm.bpos[slot] = rel; sm.bidx[slot] = (u32)j; }
            if (K >= 2 && j < P.n_reads && __ldg(P.read_off + j + 1) - o == (u64)(K - 1))
              atomicOr(&sm.shrt[rel >> 5], 1u << (rel & 31));
          }
        }
        nb += __popc(bal_tile);
        u32 bal_adv = __ballot_sync(LT_MASK_ALL, o <= ts + PW_TILE);
        cur += __popc(bal_adv);
        u32 bal_in = __ballot_sync(LT_MASK_ALL, inr);
        if (bal_in != LT_MASK_ALL) break;
        j += 32;
      }
      next_off = (cur <= P.n_reads) ? __ldg(P.read_off + cur) : ~0ull;
    }
    if (K >= 2 && lane == 0) {  // the read that starts exactly at ts (its boundary belongs to the previous tile)
      u64 o = __ldg(P.read_off + cur0 - 1);
      if (o == ts && __ldg(P.read_off + cur0) - o == (u64)(K - 1)) start_short = true;
    }
    __syncwarp();

    // ---- per lane: 16 windows, both strands by constant-shift extraction ----
    u32 cm = 0;  // candidate (filter-positive) windows of this lane
    {
      const u32 f0 = sm.fw[lane], f1 = sm.fw[lane + 1], f2 = sm.fw[lane + 2];
      const int rbase = PW_NW - lane - CMAX;
      const u32 r0 = sm.rc[rbase], r1 = sm.rc[rbase + 1], r2 = sm.rc[rbase + 2], r3 = sm.rc[rbase + 3];
      // validity: no boundary inside (p, p+K-1], p < total
      u32 inval = 0, S16 = 0;
      if (has_bound) {
        int idx = lane >> 1, sh = (lane & 1) * 16;
        u64 B = ((u64)sm.bound[idx] | ((u64)sm.bound[idx + 1] << 32)) >> sh;
        if (sh) B |= (u64)sm.bound[idx + 2] << 48;
        u64 x = 0;
        if (K >= 2) {
          x = B >> 1;
          int c = 1;
#pragma unroll
          for (int it = 0; it < 6; it++) {
            if (c < K - 1) {
              int step = (K - 1 - c) < c ? (K - 1 - c) : c;
              x |= x >> step;
              c += step;
            }
          }
        }
        inval = (u32)x & 0xFFFFu;
        S16 = (sm.shrt[idx] >> sh) & 0xFFFFu;
      }
      if (start_short && lane == 0) S16 |= 1u;
      {
        u64 p0 = ts + 16ull * lane;
        if (p0 + 16 > P.total) {
          u32 nv = p0 < P.total ? (u32)(P.total - p0) : 0u;
          inval |= 0xFFFFu & ~((1u << nv) - 1);
        }
      }
      // Filter blocks: J consecutive windows share the (K-J+1)-mer that starts at the last of them, and
      // every SUNK was inserted into the block of each of its J sub-mers (table.cu), so ONE 16-byte
      // block load serves J windows: 0.25 scattered sectors per base instead of 1.
      constexpr int J = GVS_FJ(K);
      constexpr int L = K - J + 1;
      constexpr int NG = 16 / J;
      uint4 blk[NG];
#pragma unroll
      for (int g = 0; g < NG; g++) {
        const int o = J * g + (J - 1);  // lane-relative base offset of the shared sub-mer
        u64 sf = p_extract<L>(f0, f1, f2, o);
        const int e = o + L;            // = J*g + K, inside [K, K+15] like the windows' own ends
        const int c = (e + 15) / 16;
        const int j0 = CMAX - c;
        const int off = 16 * c - e;
        u64 sr = (j0 == 0) ? p_extract<L>(r0, r1, r2, off) : p_extract<L>(r1, r2, r3, off);
        u32 hb = gvs_bhash(sf < sr ? sf : sr);
        const u32 gm = (1u << J) - 1;
        bool any_valid = ((inval >> (J * g)) & gm) != gm;
        blk[g] = any_valid ? __ldg((const uint4*)P.filt + (hb & P.filt_mask)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < 16; i++) {
        u64 f = p_extract<K>(f0, f1, f2, i);
        const int e = i + K;
        const int c = (e + 15) / 16;
        const int j0 = CMAX - c;
        const int off = 16 * c - e;
        u64 r = (j0 == 0) ? p_extract<K>(r0, r1, r2, off) : p_extract<K>(r1, r2, r3, off);
        u64 canon = f < r ? f : r;
        u32 h = gvs_fhash(canon);
        const uint4 b4 = blk[i / J];
        u32 t = __funnelshift_r(b4.x, 0u, h) & __funnelshift_r(b4.y, 0u, h >> 5) & __funnelshift_r(b4.z, 0u, h >> 10) &
                __funnelshift_r(b4.w, 0u, h >> 15);
        cm |= (t & 1u) << i;
      }
      cm &= ~inval;
      if (S16) {  // bogus windows of (K-1)-long reads: last base read as A, always "valid"
        for (u32 s = S16; s; s &= s - 1) {
          int i = __ffs(s) - 1;
          u32 p = 16u * lane + i;
          u64 canon = p_canon_at<K>(sm.fw, sm.rc, p, true);
          // its own first sub-mer (real bases only) selects the block
          u64 sub = (p_extract<K>(sm.fw[p >> 4], sm.fw[(p >> 4) + 1], sm.fw[(p >> 4) + 2], p & 15) & ~3ull) >> (2 * (K - L));
          u64 subr = gvs_revcomp(sub, L);
          uint4 b4 = __ldg((const uint4*)P.filt + (gvs_bhash(sub < subr ? sub : subr) & P.filt_mask));
          u32 h = gvs_fhash(canon);
          u32 t = __funnelshift_r(b4.x, 0u, h) & __funnelshift_r(b4.y, 0u, h >> 5) & __funnelshift_r(b4.z, 0u, h >> 10) &
                  __funnelshift_r(b4.w, 0u, h >> 15);
          if (t & 1u) cm |= 1u << i; else cm &= ~(1u << i);
        }
      }
      cm |= (S16 << 16);  // remember which candidates are forced-A windows
    }
    // ---- queue candidates in position order ----
    u32 ncand_lane = __popc(cm & 0xFFFFu);
    u32 incl = ncand_lane;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      u32 o = __shfl_up_sync(LT_MASK_ALL, incl, d);
      if (lane >= d) incl += o;
    }
    const u32 ncand = __shfl_sync(LT_MASK_ALL, incl, 31);
    u32 nh = 0;
    if (ncand) {
      u32 qo = incl - ncand_lane;
      for (u32 s = cm & 0xFFFFu; s; s &= s - 1) {
        int i = __ffs(s) - 1;
        sm.q_p[qo++] = (u16)((16u * lane + i) | (((cm >> (16 + i)) & 1) << 15));
      }
      __syncwarp();
      // ---- exact lookups, 32 candidates per round ----
      for (u32 base = 0; base < ncand; base += 32) {
        u32 row = GVS_NOHIT;
        u32 pp = 0;
        if (base + lane < ncand) {
          u32 e = sm.q_p[base + lane];
          pp = e & 0x7FFFu;
          u64 canon = p_canon_at<K>(sm.fw, sm.rc, pp, (e >> 15) != 0);
          row = tab_lookup(P.tab, canon, gvs_mix(canon));
          if (row == GVS_ROW_MISSING) {
            atomicOr(P.flags, FLAG_KEYERROR);
            row = GVS_NOHIT;
          } else if (row >= GVS_NOHIT) {
            row = GVS_NOHIT;
          }
        }
        u32 bal = __ballot_sync(LT_MASK_ALL, row != GVS_NOHIT);
        __syncwarp();
        if (row != GVS_NOHIT) {
          u32 d = nh + __popc(bal & ((1u << lane) - 1));
          sm.q_p[d] = (u16)pp;
          sm.q_row[d] = row;
        }
        nh += __popc(bal);
        __syncwarp();
      }
    }
    // ---- append the tile's hits ----
    u64 obase = 0;
    if (lane == 0) {
      if (nh) {
        obase = atomicAdd(P.cursor, (unsigned long long)nh);
        if (obase + nh > P.hit_cap) atomicOr(P.flags, FLAG_OVERFLOW);
      }
      P.tile_cnt[tile] = nh;
      P.tile_off[tile] = obase;
    }
    if (nh) {
      obase = __shfl_sync(LT_MASK_ALL, obase, 0);
      for (u32 idx = lane; idx < nh; idx += 32) {
        u64 o = obase + idx;
        if (o >= P.hit_cap) break;
        u32 prel = sm.q_p[idx];
        u64 p = ts + prel;
        u32 rd;
        if (!has_bound || nb <= PW_MAXB) {
          rd = (u32)(cur0 - 1);
          if (has_bound) {
            u32 bestpos = 0;
            bool any = false;
            for (u32 q = 0; q < nb; q++) {
              u32 bp = sm.bpos[q], bj = sm.bidx[q];
              if (bp <= prel && (!any || bp > bestpos || (bp == bestpos && bj > rd))) {
                any = true;
                bestpos = bp;
                rd = bj;
              }
            }
          }
        } else {
          u64 l = 1, h2 = P.n_reads;
          while (l < h2) {
            u64 mid = (l + h2) >> 1;
            if (__ldg(P.read_off + mid) > p) h2 = mid; else l = mid + 1;
          }
          rd = (u32)(l - 1);
        }
        P.hit_read[o] = rd;
        P.hit_w[o] = (u32)(p - __ldg(P.read_off + rd));
        P.hit_row[o] = sm.q_row[idx];
      }
    }
    __syncwarp();
  }
}

typedef void (*probe_fn)(const Probe2Params);
template <int K>
static probe_fn probe_entry() { return k_probe2<K>; }

static probe_fn probe_table(int k) {
  switch (k) {
#define PK(n) case n: return probe_entry<n>();
    PK(1) PK(2) PK(3) PK(4) PK(5) PK(6) PK(7) PK(8) PK(9) PK(10) PK(11) PK(12) PK(13) PK(14) PK(15) PK(16)
    PK(17) PK(18) PK(19) PK(20) PK(21) PK(22) PK(23) PK(24) PK(25) PK(26) PK(27) PK(28) PK(29) PK(30) PK(31)
#undef PK
  }
  return nullptr;
}

// launches the probe over ctx's reads; outputs in ctx->hit_*, tile_cnt, tile_off; counters[0] =
// total hits, counters[1] = flags
int gvs_probe_launch(gvs_ctx* ctx, u64* n_tiles_out) {
  u64 total = ctx->total_bases;
  u64 n_tiles = cdiv(total, PW_TILE);
  *n_tiles_out = n_tiles;
  CKR(gvs_reserve(ctx, ctx->tile_cnt, n_tiles * 4));
  CKR(gvs_reserve(ctx, ctx->tile_off, n_tiles * 8));
  CKR(gvs_reserve(ctx, ctx->tile_dst, n_tiles * 8));
  probe_fn fn = probe_table(ctx->k);
  if (!fn) return gvs_fail(ctx, GVS_E_ARG, "no probe kernel for k=%d", ctx->k);
  u64* counters = ctx->counters.as<u64>();
  Probe2Params P;
  P.seq = ctx->seq;
  P.total = total;
  P.read_off = ctx->read_off;
  P.n_reads = ctx->n_reads;
  P.n_tiles = n_tiles;
  u64 blocks = (u64)ctx->n_sm * 4;
  u64 warps = blocks * PW_WARPS;
  P.tiles_per_warp = cdiv(n_tiles, warps);
  if (P.tiles_per_warp < 4) P.tiles_per_warp = 4;
  blocks = cdiv(cdiv(n_tiles, P.tiles_per_warp), PW_WARPS);
  P.filt = ctx->filt.as<u32>();
  P.filt_mask = (u32)(ctx->filt_words - 1);
  P.tab.keys = ctx->tab_keys.as<u64>();
  P.tab.rows = ctx->tab_rows.as<u32>();
  P.tab.slots = ctx->tab_slots;
  P.hit_read = ctx->hit_read.as<u32>();
  P.hit_w = ctx->hit_w.as<u32>();
  P.hit_row = ctx->hit_row.as<u32>();
  P.hit_cap = ctx->hit_cap;
  P.cursor = (unsigned long long*)counters;
  P.flags = (u32*)(counters + 1);
  P.tile_cnt = ctx->tile_cnt.as<u32>();
  P.tile_off = ctx->tile_off.as<u64>();
  fn<<<(unsigned)blocks, PW_WARPS * 32, 0, ctx->stream>>>(P);
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return gvs_fail(ctx, GVS_E_CUDA, "k_probe2 launch: %s", cudaGetErrorString(e));
  return 0;
}
