// validate.cu -- bad-SUNK histogram (workflow/scripts/badsunks_AR.py:20-103) and the per-read
// inter-SUNK distance validation (workflow/scripts/process-by-contig_lowmem_AR.py:50-207).
// SURVEY.md A.5, A.6 (steps 1-5), quirks Q12-Q14.
//
//   k_hist / k_cnt_hist / k_mode / k_bad_flag
//       rows per (contig, group); per haplotype the smallest mode m of the non-zero counts;
//       bad <=> count > m + 4*sqrt(m) or count < 2 (threshold passed in as an integer floor)
//   k_validate<BIG>
//       one thread block per read.  Rows minus bad groups, stable rank-sort by assembly start,
//       then every pair (i<j) is tested with the integer form of the reference's float64 ratio
//       test (0.9 < dpos/dstart < 1.1  <=>  9*ds < 10*dp < 11*ds, exact for 32-bit inputs),
//       orientation majority, "multipos" clean-up, union-find over group IDs in shared memory,
//       largest component (ties: the component holding the earliest vertex in graph-tool's
//       insertion order), output in vertex order.
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// histogram / bad groups
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_hist(const u32* __restrict__ gidx, u64 n, i32* hist) {
  u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) atomicAdd(&hist[gidx[j]], 1);
}
__global__ void __launch_bounds__(256) k_hist_max(const i32* __restrict__ hist, u64 ng, u32* maxv) {
  u32 m = 0;
  for (u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x; g < ng; g += (u64)gridDim.x * blockDim.x) m = max(m, (u32)hist[g]);
  for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, d));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(maxv, m);
}
__global__ void __launch_bounds__(256) k_cnt_hist(const i32* __restrict__ hist, const u32* __restrict__ grp_contig,
                                                  const u8* __restrict__ contig_hap, u64 ng, u32 M, u32* cnt_hist) {
  u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  u32 c = (u32)hist[g];
  if (!c) return;
  u8 h = contig_hap[grp_contig[g]];
  if (h > 1) return;
  atomicAdd(&cnt_hist[(u64)h * (M + 1) + c], 1u);
}
// pandas .mode()[0]: the smallest of the most frequent values
__global__ void __launch_bounds__(1024) k_mode(const u32* __restrict__ cnt_hist, u32 M, i64* mode) {
  __shared__ unsigned long long best;
  const u32* ch = cnt_hist + (u64)blockIdx.x * (M + 1);
  if (threadIdx.x == 0) best = 0;
  __syncthreads();
  unsigned long long loc = 0;
  for (u32 v = 1 + threadIdx.x; v <= M; v += blockDim.x) {
    u32 f = ch[v];
    if (f) {
      unsigned long long key = ((unsigned long long)f << 32) | (0xFFFFFFFFu - v);  // max freq, then min v
      if (key > loc) loc = key;
    }
  }
  atomicMax(&best, loc);
  __syncthreads();
  if (threadIdx.x == 0) mode[blockIdx.x] = best ? (i64)(0xFFFFFFFFu - (u32)best) : 0;
}
__global__ void __launch_bounds__(256) k_bad_flag(const i32* __restrict__ hist, const u32* __restrict__ grp_contig,
                                                  const u8* __restrict__ contig_hap, u64 ng, i64 lim0, i64 lim1, u8* bad) {
  u64 g = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ng) return;
  i64 c = hist[g];
  u8 h = contig_hap[grp_contig[g]];
  bool b = false;
  if (c > 0 && h <= 1) b = c > (h ? lim1 : lim0) || c < 2;  // badsunks_AR.py:48
  bad[g] = b ? 1 : 0;
}

extern "C" int gvs_contigs_set(gvs_ctx* ctx, const uint8_t* contig_hap, const uint32_t* contig_hash, uint32_t n_contigs) {
  if (!ctx || !contig_hap) return GVS_E_ARG;
  CK(cudaSetDevice(ctx->device));
  if (ctx->n_contigs && n_contigs != ctx->n_contigs) return gvs_fail(ctx, GVS_E_ARG, "n_contigs mismatch (%u vs %u)", n_contigs, ctx->n_contigs);
  ctx->n_contigs = n_contigs;
  CKR(to_dev(ctx, ctx->contig_hap, contig_hap, n_contigs));
  if (contig_hash) CKR(to_dev(ctx, ctx->contig_hash, contig_hash, n_contigs));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int gvs_group_hist(gvs_ctx* ctx, int accumulate, int32_t** hist_dev) {
  if (!ctx) return GVS_E_ARG;
  if (!ctx->diag_ready) return gvs_fail(ctx, GVS_E_STATE, "gvs_group_hist before gvs_diag_filter / gvs_rows_set(1)");
  if (!ctx->groups_ready) return gvs_fail(ctx, GVS_E_STATE, "gvs_group_hist: no group index");
  CK(cudaSetDevice(ctx->device));
  StageTimer tm(ctx, GVS_ST_HIST);
  u64 ng = ctx->n_groups;
  bool fresh = ctx->hist.cap < (ng ? ng : 1) * 4;
  CKR(gvs_reserve(ctx, ctx->hist, (ng ? ng : 1) * 4));
  if (!accumulate || fresh || !ctx->hist_ready) CK(cudaMemsetAsync(ctx->hist.p, 0, (ng ? ng : 1) * 4, ctx->stream));
  u64 n = ctx->kept.n;
  if (n) LAUNCH(k_hist, (unsigned)cdiv(n, 256), 256, 0, ctx->kept.gidx.as<u32>(), n, ctx->hist.as<i32>());
  ctx->hist_ready = true;
  ctx->bad_ready = false;
  if (hist_dev) *hist_dev = ctx->hist.as<i32>();
  return 0;
}

extern "C" int gvs_hist_mode(gvs_ctx* ctx, int64_t mode[2]) {
  if (!ctx || !mode) return GVS_E_ARG;
  if (!ctx->hist_ready) return gvs_fail(ctx, GVS_E_STATE, "gvs_hist_mode before gvs_group_hist");
  if (ctx->contig_hap.cap == 0) return gvs_fail(ctx, GVS_E_STATE, "gvs_hist_mode: contig haplotypes unknown (gvs_contigs_set)");
  CK(cudaSetDevice(ctx->device));
  StageTimer tm(ctx, GVS_ST_HIST);
  u64 ng = ctx->n_groups;
  mode[0] = mode[1] = 0;
  if (ng == 0) return 0;
  u32* mx = (u32*)(ctx->counters.as<u64>() + 16);
  CK(cudaMemsetAsync(mx, 0, 4, ctx->stream));
  u64 grid = cdiv(ng, 256);
  if (grid > (u64)ctx->n_sm * 8) grid = (u64)ctx->n_sm * 8;
  LAUNCH(k_hist_max, (unsigned)grid, 256, 0, ctx->hist.as<i32>(), ng, mx);
  u32 M = 0;
  CKR(read_dev(ctx, mx, &M));
  if (M == 0) return 0;
  if (M > (1u << 28)) return gvs_fail(ctx, GVS_E_OVERFLOW, "group hit count %u too large for the mode table", M);
  CKR(gvs_reserve(ctx, ctx->cnt_hist, 2ull * (M + 1) * 4 + 16));
  CK(cudaMemsetAsync(ctx->cnt_hist.p, 0, 2ull * (M + 1) * 4 + 16, ctx->stream));
  LAUNCH(k_cnt_hist, (unsigned)cdiv(ng, 256), 256, 0, ctx->hist.as<i32>(), ctx->grp_contig.as<u32>(), ctx->contig_hap.as<u8>(), ng,
         M, ctx->cnt_hist.as<u32>());
  i64* dm = (i64*)(ctx->counters.as<u64>() + 18);
  LAUNCH(k_mode, 2, 1024, 0, ctx->cnt_hist.as<u32>(), M, dm);
  i64 hm[2];
  CKR(read_dev(ctx, dm, hm, 2));
  mode[0] = hm[0];
  mode[1] = hm[1];
  return 0;
}

extern "C" int gvs_bad_groups(gvs_ctx* ctx, const int64_t limit_floor[2], uint64_t* n_bad) {
  if (!ctx || !limit_floor) return GVS_E_ARG;
  if (!ctx->hist_ready) return gvs_fail(ctx, GVS_E_STATE, "gvs_bad_groups before gvs_group_hist");
  CK(cudaSetDevice(ctx->device));
  StageTimer tm(ctx, GVS_ST_HIST);
  u64 ng = ctx->n_groups;
  CKR(gvs_reserve(ctx, ctx->bad_flag, ng ? ng : 1));
  CKR(gvs_reserve(ctx, ctx->bad_list, (ng ? ng : 1) * 4));
  ctx->n_bad = 0;
  if (ng) {
    LAUNCH(k_bad_flag, (unsigned)cdiv(ng, 256), 256, 0, ctx->hist.as<i32>(), ctx->grp_contig.as<u32>(), ctx->contig_hap.as<u8>(), ng,
           limit_floor[0], limit_floor[1], ctx->bad_flag.as<u8>());
    const u8* bf = ctx->bad_flag.as<u8>();
    u32* bl = ctx->bad_list.as<u32>();
    u32* tot = (u32*)(ctx->counters.as<u64>() + 20);
    auto f = [bf] __device__(u64 g) -> u32 { return bf[g]; };
    auto g2 = [bl] __device__(u64 g, u32 ex, u32 v) { if (v) bl[ex] = (u32)g; };
    CKR((device_scan<u32>(ctx, ng, f, g2, OpSum(), tot)));
    u32 nb = 0;
    CKR(read_dev(ctx, tot, &nb));
    ctx->n_bad = nb;
  }
  ctx->bad_ready = true;
  if (n_bad) *n_bad = ctx->n_bad;
  return 0;
}

extern "C" int gvs_bad_get(gvs_ctx* ctx, uint32_t* group_index) {
  if (!ctx) return GVS_E_ARG;
  if (!ctx->bad_ready) return gvs_fail(ctx, GVS_E_STATE, "gvs_bad_get before gvs_bad_groups");
  CK(cudaSetDevice(ctx->device));
  if (ctx->n_bad && group_index) {
    CK(cudaMemcpyAsync(group_index, ctx->bad_list.p, ctx->n_bad * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// per-read validation
// ---------------------------------------------------------------------------------------------
#define VCAP 512          // rows per read handled in shared memory
#define VROW_BYTES 80     // bytes of working storage per row
#define NOV 0xFFFFFFFFu
#define NOT64 0xFFFFFFFFFFFFFFFFull

struct VWork {
  u32 *P, *S, *ID, *G;        // rows sorted by assembly start (stable)
  u32 *uP, *uS, *uID, *uG;    // rows in file order (after bad-group removal)
  u32 *deg, *fpart, *rep, *par, *csize;
  u64 *key, *tv, *ct;
  u8 *good, *multi, *pres, *left;
};

__device__ __forceinline__ void vwork_carve(VWork& w, u8* base, u32 cap) {
  u64* q = (u64*)base;
  w.key = q; q += cap;
  w.tv = q; q += cap;
  w.ct = q; q += cap;
  u32* p = (u32*)q;
  w.P = p; p += cap; w.S = p; p += cap; w.ID = p; p += cap; w.G = p; p += cap;
  w.uP = p; p += cap; w.uS = p; p += cap; w.uID = p; p += cap; w.uG = p; p += cap;
  w.deg = p; p += cap; w.fpart = p; p += cap; w.rep = p; p += cap; w.par = p; p += cap; w.csize = p; p += cap;
  u8* b = (u8*)p;
  w.good = b; b += cap; w.multi = b; b += cap; w.pres = b; b += cap; w.left = b; b += cap;
}

// 0.9 < dpos/dstart < 1.1 in float64 (process-by-contig_lowmem_AR.py:145-147) as exact integers (Q12)
__device__ __forceinline__ bool pair_ok(u32 pi, u32 si, u32 pj, u32 sj) {
  u64 ds = si > sj ? si - sj : sj - si;
  u64 dp = pi > pj ? pi - pj : pj - pi;
  return 9 * ds < 10 * dp && 10 * dp < 11 * ds;
}

__device__ __forceinline__ u32 uf_find(const u32* par, u32 x) {
  u32 p = par[x];
  while (p != x) { x = p; p = par[x]; }
  return x;
}
__device__ __forceinline__ void uf_union(u32* par, u32 a, u32 b) {
  for (;;) {
    a = uf_find(par, a);
    b = uf_find(par, b);
    if (a == b) return;
    u32 hi = a > b ? a : b, lo = a > b ? b : a;
    if (atomicCAS(&par[hi], hi, lo) == hi) return;
  }
}

struct ValParams {
  const u32 *read, *pos, *start, *group, *gidx;  // kept rows
  const u32* seg_start;
  u32 n_seg;
  const u8* bad;             // per group index (may be null: nothing is bad)
  const u64* read_off;       // read lengths from offsets ...
  const u32* read_len;       // ... or explicit
  u32 min_len;
  u32* out_id;               // per kept row slot: validated IDs of the read, from its first row on
  u32* out_gidx;
  u32* seg_cnt;              // validated IDs per segment
  u32 big_min;               // segments with more rows than this belong to the BIG launch
  const u32* big_list;       // BIG: segment ids
  u32 n_big;
  u8* gscratch;              // BIG: per block VROW_BYTES * big_cap
  u32 big_cap;
  u32* stats;                // [0] reads whose clean-up removed every edge (reference would raise)
};

template <bool BIG>
__global__ void __launch_bounds__(BIG ? 1024 : 128) k_validate(const ValParams V) {
  extern __shared__ __align__(16) u8 smem[];
  __shared__ u32 s_n0, s_n1, s_m, s_kept;
  __shared__ unsigned long long s_best;
  __shared__ u32 s_bestroot, s_distinct;
  const int NT = BIG ? 1024 : 128;
  const int tid = threadIdx.x;
  VWork w;
  u32 cap;
  if (BIG) {
    cap = V.big_cap;
    vwork_carve(w, V.gscratch + (u64)blockIdx.x * VROW_BYTES * cap, cap);
  } else {
    cap = VCAP;
    vwork_carve(w, smem, cap);
  }
  const u32 n_items = BIG ? V.n_big : V.n_seg;
  for (u32 it = blockIdx.x; it < n_items; it += gridDim.x) {
    const u32 s = BIG ? V.big_list[it] : it;
    const u32 a = V.seg_start[s], b = V.seg_start[s + 1];
    const u32 mrows = b - a;
    if (!BIG && mrows > V.big_min) continue;  // handled by the BIG launch (which also sets seg_cnt)
    __syncthreads();
    if (tid == 0) { s_m = 0; s_n0 = 0; s_n1 = 0; s_kept = 0; s_best = 0; s_bestroot = NOV; s_distinct = 0; }
    __syncthreads();
    // read length filter (hard-coded 10000 in the reference, :106-108; Q13)
    u32 rd = V.read[a];
    u32 rlen = V.read_len ? V.read_len[rd] : (u32)(V.read_off[rd + 1] - V.read_off[rd]);
    bool skip = rlen < V.min_len || mrows < 2;
    // ---- rows minus bad groups, file order (:70-72) ----
    if (!skip) {
      for (u32 base = 0; base < mrows; base += NT) {
        u32 i = base + tid;
        bool keep = i < mrows && !(V.bad && V.bad[V.gidx[a + i]]);
        // ordered compaction: per-warp ballots + running offset through shared memory
        u32 bal = __ballot_sync(0xFFFFFFFFu, keep);
        __shared__ u32 s_wcnt[32];
        int wid = tid >> 5, lane = tid & 31;
        if (lane == 0) s_wcnt[wid] = __popc(bal);
        __syncthreads();
        u32 off = s_m;
        for (int q = 0; q < wid; q++) off += s_wcnt[q];
        if (keep) {
          u32 d = off + __popc(bal & ((1u << lane) - 1));
          w.uP[d] = V.pos[a + i];
          w.uS[d] = V.start[a + i];
          w.uID[d] = V.group[a + i];
          w.uG[d] = V.gidx[a + i];
        }
        __syncthreads();
        if (tid == 0) {
          u32 t = 0;
          for (int q = 0; q < NT / 32; q++) t += s_wcnt[q];
          s_m += t;
        }
        __syncthreads();
      }
    }
    const u32 m = s_m;
    if (skip || m < 2) {
      if (tid == 0) V.seg_cnt[s] = 0;
      continue;
    }
    // ---- stable sort by start (sort_values(['rname','start']), :100) + scratch init ----
    for (u32 i = tid; i < m; i += NT) {
      u32 si = w.uS[i], rank = 0;
      for (u32 j = 0; j < m; j++) {
        u32 sj = w.uS[j];
        rank += (sj < si) || (sj == si && j < i);
      }
      w.P[rank] = w.uP[i];
      w.S[rank] = si;
      w.ID[rank] = w.uID[i];
      w.G[rank] = w.uG[i];
      w.deg[i] = 0; w.fpart[i] = NOV; w.par[i] = i; w.csize[i] = 0;
      w.tv[i] = NOT64; w.ct[i] = NOT64;
      w.good[i] = 0; w.multi[i] = 0; w.pres[i] = 0; w.left[i] = 0;
    }
    __syncthreads();
    // at least two distinct groups (:91-97)
    for (u32 i = tid; i < m; i += NT)
      if (w.ID[i] != w.ID[0]) s_distinct = 1;
    __syncthreads();
    if (!s_distinct) {
      if (tid == 0) V.seg_cnt[s] = 0;
      continue;
    }
    const u64 cells = (u64)m * m;
    // ---- pass A: masked pairs by sign (:140-152) ----
    {
      u32 n0 = 0, n1 = 0;
      u32 i = tid / m, j = tid % m;
      for (u64 c = tid; c < cells; c += NT) {
        if (i < j && pair_ok(w.P[i], w.S[i], w.P[j], w.S[j])) {
          if (w.P[i] > w.P[j]) n1++; else n0++;
        }
        j += NT;
        while (j >= m) { j -= m; i++; }
      }
      for (int d = 16; d; d >>= 1) {
        n0 += __shfl_xor_sync(0xFFFFFFFFu, n0, d);
        n1 += __shfl_xor_sync(0xFFFFFFFFu, n1, d);
      }
      if ((tid & 31) == 0) {
        if (n0) atomicAdd(&s_n0, n0);
        if (n1) atomicAdd(&s_n1, n1);
      }
    }
    __syncthreads();
    if (s_n0 + s_n1 < 1) {  // `if sum(mask) < 1: continue` (:148)
      if (tid == 0) V.seg_cnt[s] = 0;
      continue;
    }
    const bool orient = s_n1 > s_n0;  // np.unique sorted + argmax: a tie keeps 0 (:151-152)
    // ---- pass B: incidence of every row in the oriented edge list (the (ID,pos) multiset M) ----
    {
      u32 i = tid / m, j = tid % m;
      for (u64 c = tid; c < cells; c += NT) {
        if (i < j && pair_ok(w.P[i], w.S[i], w.P[j], w.S[j]) && ((w.P[i] > w.P[j]) == orient)) {
          atomicAdd(&w.deg[i], 1u);
          atomicAdd(&w.deg[j], 1u);
          w.left[i] = 1;
          atomicMin(&w.fpart[j], i);
        }
        j += NT;
        while (j >= m) { j -= m; i++; }
      }
    }
    __syncthreads();
    // order of first appearance in M = [all left ends in edge order] + [all right ends in edge order]
    for (u32 r = tid; r < m; r += NT)
      w.key[r] = w.left[r] ? (u64)r : ((1ull << 63) | ((u64)w.fpart[r] * m + r));
    __syncthreads();
    // ---- multipos (:161-181): IDs seen at more than one read position keep their most frequent one ----
    for (u32 r = tid; r < m; r += NT) {
      u32 id = w.ID[r], dr = w.deg[r];
      u32 first = r;
      bool multi = false, good = dr > 0;
      for (u32 q = 0; q < m; q++) {
        if (w.ID[q] != id) continue;
        if (q < first) first = q;
        if (q == r) continue;
        u32 dq = w.deg[q];
        if (dq == 0) continue;
        multi = true;
        if (dq > dr || (dq == dr && w.key[q] < w.key[r])) good = false;
      }
      w.rep[r] = first;
      w.multi[r] = (multi && dr > 0) ? 1 : 0;
      w.good[r] = good ? 1 : 0;
    }
    __syncthreads();
    // ---- pass C: surviving edges -> graph on IDs (:189-192) ----
    {
      u32 nk = 0;
      u32 i = tid / m, j = tid % m;
      for (u64 c = tid; c < cells; c += NT) {
        if (i < j && pair_ok(w.P[i], w.S[i], w.P[j], w.S[j]) && ((w.P[i] > w.P[j]) == orient)) {
          // dropped iff the left ID is multi-positioned, this is not its good row and the right row is
          // not that ID's good row either (left end only: Q14)
          bool drop = w.multi[i] && !w.good[i] && !(w.good[j] && w.ID[j] == w.ID[i]);
          if (!drop) {
            nk++;
            u32 ri = w.rep[i], rj = w.rep[j];
            u64 e = (u64)i * m + j;
            w.pres[ri] = 1;
            w.pres[rj] = 1;
            atomicMin((unsigned long long*)&w.tv[ri], (unsigned long long)(2 * e));
            atomicMin((unsigned long long*)&w.tv[rj], (unsigned long long)(2 * e + 1));
            if (ri != rj) uf_union(w.par, ri, rj);
          }
        }
        j += NT;
        while (j >= m) { j -= m; i++; }
      }
      if (nk) atomicAdd(&s_kept, nk);
    }
    __syncthreads();
    if (s_kept == 0) {  // graph-tool would raise on the empty graph; counted, read skipped (A.6 step 5)
      if (tid == 0) {
        V.seg_cnt[s] = 0;
        atomicAdd(&V.stats[0], 1u);
      }
      continue;
    }
    // ---- components: size and earliest vertex ----
    for (u32 r = tid; r < m; r += NT) {
      if (w.rep[r] == r && w.pres[r]) {
        u32 root = uf_find(w.par, r);
        atomicAdd(&w.csize[root], 1u);
        atomicMin((unsigned long long*)&w.ct[root], (unsigned long long)w.tv[r]);
      }
    }
    __syncthreads();
    // largest component; ties -> lowest label = the one holding the earliest-inserted vertex
    for (u32 r = tid; r < m; r += NT) {
      if (w.rep[r] == r && w.pres[r] && w.par[r] == r) {
        unsigned long long key = ((unsigned long long)w.csize[r] << 42) | ((1ull << 42) - 1 - w.ct[r]);
        atomicMax(&s_best, key);
      }
    }
    __syncthreads();
    for (u32 r = tid; r < m; r += NT) {
      if (w.rep[r] == r && w.pres[r] && w.par[r] == r) {
        unsigned long long key = ((unsigned long long)w.csize[r] << 42) | ((1ull << 42) - 1 - w.ct[r]);
        if (key == s_best) s_bestroot = r;
      }
    }
    __syncthreads();
    const u32 broot = s_bestroot;
    // ---- output in vertex order (first appearance in the edge list, source before target) ----
    for (u32 r = tid; r < m; r += NT) {
      if (w.rep[r] == r && w.pres[r] && uf_find(w.par, r) == broot) {
        u64 t = w.tv[r];
        u32 rank = 0;
        for (u32 q = 0; q < m; q++)
          if (w.rep[q] == q && w.pres[q] && w.tv[q] < t && uf_find(w.par, q) == broot) rank++;
        V.out_id[a + rank] = w.ID[r];
        V.out_gidx[a + rank] = w.G[r];
      }
    }
    if (tid == 0) V.seg_cnt[s] = w.csize[broot];
  }
}

__global__ void __launch_bounds__(256) k_pairs_compact(const u32* __restrict__ seg_start, const u32* __restrict__ seg_cnt,
                                                       const u32* __restrict__ seg_off, u32 n_seg, const u32* __restrict__ out_id,
                                                       const u32* __restrict__ out_gidx, const u32* __restrict__ read,
                                                       const u32* __restrict__ contig, u32* p_read, u32* p_contig, u32* p_group,
                                                       u32* p_gidx) {
  u32 wv = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (wv >= n_seg) return;
  u32 c = seg_cnt[wv];
  if (!c) return;
  u32 a = seg_start[wv], o = seg_off[wv];
  u32 rd = read[a], ct = contig[a];
  for (u32 i = lane; i < c; i += 32) {
    p_read[o + i] = rd;
    p_contig[o + i] = ct;
    p_group[o + i] = out_id[a + i];
    p_gidx[o + i] = out_gidx[a + i];
  }
}

extern "C" int gvs_validate(gvs_ctx* ctx, uint32_t min_read_len, uint64_t* n_pairs_out) {
  if (!ctx) return GVS_E_ARG;
  if (!ctx->diag_ready) return gvs_fail(ctx, GVS_E_STATE, "gvs_validate before gvs_diag_filter / gvs_rows_set(1)");
  if (!ctx->read_off && !ctx->have_read_len) return gvs_fail(ctx, GVS_E_STATE, "gvs_validate: read lengths unknown");
  CK(cudaSetDevice(ctx->device));
  StageTimer tm(ctx, GVS_ST_VALIDATE);
  ctx->val_ready = false;
  ctx->n_pairs = 0;
  if (n_pairs_out) *n_pairs_out = 0;
  Rows& R = ctx->kept;
  u64 n = R.n;
  if (n == 0) {
    ctx->val_ready = true;
    return 0;
  }
  u64 n_seg = 0;
  CKR(gvs_build_segments(ctx, R.read.as<u32>(), n, ctx->kseg_start, ctx->flags_c, &n_seg));
  ctx->n_kseg = n_seg;
  const u32* seg_start = ctx->kseg_start.as<u32>();
  // scratch: out_id[n], out_gidx[n], seg_cnt[n_seg], seg_off[n_seg], big_list[n_seg]
  CKR(gvs_reserve(ctx, ctx->val_scratch, (2 * n + 3 * n_seg + 16) * 4));
  u32* out_id = ctx->val_scratch.as<u32>();
  u32* out_gidx = out_id + n;
  u32* seg_cnt = out_gidx + n;
  u32* seg_off = seg_cnt + n_seg;
  u32* big_list = seg_off + n_seg;
  // big segments (more rows than fit shared memory)
  u32* nb_dev = (u32*)(ctx->counters.as<u64>() + 21);
  u32* mx_dev = (u32*)(ctx->counters.as<u64>() + 22);
  {
    auto f = [seg_start] __device__(u64 s) -> u32 { return (seg_start[s + 1] - seg_start[s]) > VCAP ? 1u : 0u; };
    auto g = [big_list] __device__(u64 s, u32 ex, u32 v) { if (v) big_list[ex] = (u32)s; };
    CKR((device_scan<u32>(ctx, n_seg, f, g, OpSum(), nb_dev)));
    auto f2 = [seg_start] __device__(u64 s) -> u32 { return seg_start[s + 1] - seg_start[s]; };
    auto g2 = [] __device__(u64 s, u32 ex, u32 v) {};
    CKR((device_scan<u32>(ctx, n_seg, f2, g2, OpMax(), mx_dev)));
  }
  u32 n_big = 0, max_m = 0;
  CKR(read_dev(ctx, nb_dev, &n_big));
  CKR(read_dev(ctx, mx_dev, &max_m));
  u32* stats = (u32*)(ctx->counters.as<u64>() + 23);
  CK(cudaMemsetAsync(stats, 0, 8, ctx->stream));
  ValParams V;
  V.read = R.read.as<u32>(); V.pos = R.pos.as<u32>(); V.start = R.start.as<u32>(); V.group = R.group.as<u32>();
  V.gidx = R.gidx.as<u32>();
  V.seg_start = seg_start; V.n_seg = (u32)n_seg;
  V.bad = ctx->bad_ready ? ctx->bad_flag.as<u8>() : nullptr;
  V.read_off = ctx->read_off; V.read_len = ctx->have_read_len ? ctx->read_len.as<u32>() : nullptr;
  V.min_len = min_read_len;
  V.out_id = out_id; V.out_gidx = out_gidx; V.seg_cnt = seg_cnt;
  V.big_min = VCAP; V.big_list = big_list; V.n_big = n_big; V.gscratch = nullptr; V.big_cap = 0; V.stats = stats;
  {
    static bool attr_set = false;
    size_t sm = (size_t)VROW_BYTES * VCAP;
    if (!attr_set) {
      CK(cudaFuncSetAttribute(k_validate<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      attr_set = true;
    }
    u64 grid = n_seg;
    u64 cap = (u64)ctx->n_sm * 5;
    if (grid > cap) grid = cap;
    LAUNCH(k_validate<false>, (unsigned)grid, 128, sm, V);
  }
  if (n_big) {
    u32 blocks = n_big < (u32)ctx->n_sm ? n_big : (u32)ctx->n_sm;
    u32 bcap = (max_m + 15) & ~15u;
    CKR(gvs_reserve(ctx, ctx->scan_tmp2, (u64)blocks * VROW_BYTES * bcap));
    V.gscratch = ctx->scan_tmp2.as<u8>();
    V.big_cap = bcap;
    LAUNCH(k_validate<true>, blocks, 1024, 0, V);
  }
  // compaction of the validated (ID, read) pairs
  u32* tot = (u32*)(ctx->counters.as<u64>() + 25);
  {
    auto f = [seg_cnt] __device__(u64 s) -> u32 { return seg_cnt[s]; };
    auto g = [seg_off] __device__(u64 s, u32 ex, u32 v) { seg_off[s] = ex; };
    CKR((device_scan<u32>(ctx, n_seg, f, g, OpSum(), tot)));
  }
  u32 np = 0;
  CKR(read_dev(ctx, tot, &np));
  CKR(gvs_reserve(ctx, ctx->pair_read, (u64)np * 4));
  CKR(gvs_reserve(ctx, ctx->pair_contig, (u64)np * 4));
  CKR(gvs_reserve(ctx, ctx->pair_group, (u64)np * 4));
  CKR(gvs_reserve(ctx, ctx->pair_gidx, (u64)np * 4));
  if (np)
    LAUNCH(k_pairs_compact, (unsigned)cdiv(n_seg * 32, 256), 256, 0, seg_start, seg_cnt, seg_off, (u32)n_seg, out_id, out_gidx,
           R.read.as<u32>(), R.contig.as<u32>(), ctx->pair_read.as<u32>(), ctx->pair_contig.as<u32>(),
           ctx->pair_group.as<u32>(), ctx->pair_gidx.as<u32>());
  ctx->n_pairs = np;
  ctx->val_ready = true;
  if (n_pairs_out) *n_pairs_out = np;
  return 0;
}

extern "C" int gvs_pairs_get(gvs_ctx* ctx, uint32_t* read_idx, uint32_t* contig, uint32_t* group, uint32_t* group_index) {
  if (!ctx) return GVS_E_ARG;
  if (!ctx->val_ready) return gvs_fail(ctx, GVS_E_STATE, "gvs_pairs_get before gvs_validate");
  CK(cudaSetDevice(ctx->device));
  u64 n = ctx->n_pairs;
  if (n == 0) return 0;
  if (read_idx) CK(cudaMemcpyAsync(read_idx, ctx->pair_read.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (contig) CK(cudaMemcpyAsync(contig, ctx->pair_contig.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (group) CK(cudaMemcpyAsync(group, ctx->pair_group.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (group_index) CK(cudaMemcpyAsync(group_index, ctx->pair_gidx.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
