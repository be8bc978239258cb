// table.cuh -- device-side view of the SUNK probe structures (shared by build and probe kernels)
//
// Layout in HBM (DESIGN.md "Data layout"):
//   filt      u32[filt_words]   word-blocked Bloom filter, 2 bits per key inside ONE 32-bit word;
//                               sized to stay L2-resident (<= 64 MiB), probed once per read base
//   tab_keys  u64[tab_slots]    open-addressed canonical k-mers, buckets of 4 slots = one 32-byte
//                               sector, load factor <= 0.5, EMPTY = all ones (k <= 31 => < 2^62)
//   tab_gidx  u64[tab_slots]    low word: .loc row of the key (or GVS_ROW_MISSING / GVS_ROW_NOTINDB), high
//                               word: the row's dense group index; one 8-byte load per key match
//                               (tab_rows u32[tab_slots] only exists while the table is built);
//                               touched only on a key match
#pragma once
#include "common.cuh"

struct TabView {
  const u64* __restrict__ keys;
  const u64* __restrict__ val;  // low word: .loc row (or GVS_ROW_*), high word: dense group index of the row
  u64 slots;
};

#define GVS_NOHIT 0xFFFFFFFDu

// 16 bytes of a table bucket / one value word; .L2::64B: a miss fills half a line, not the whole 128 bytes (the buckets
// are 32-byte sectors scattered over gigabytes: profiles/microbench/fetch_granularity.cu)
__device__ __forceinline__ ulonglong2 tab_ld16(const ulonglong2* p) {
  ulonglong2 v;
  asm volatile("ld.global.nc.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ u64 tab_ld8(const u64* p) {
  u64 v;
  asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
// exact lookup; returns the slot of the key or ~0
__device__ __forceinline__ u64 tab_find(const TabView& t, u64 key, u64 h) {
  u64 nb = t.slots >> 2;
  u64 b = gvs_tab_bucket(h, t.slots);
  for (u64 it = 0; it < nb; it++) {
    const ulonglong2* p = (const ulonglong2*)(t.keys + (b << 2));
    ulonglong2 a = tab_ld16(p), c = tab_ld16(p + 1);
    if (a.x == key) return (b << 2) + 0;
    if (a.y == key) return (b << 2) + 1;
    if (c.x == key) return (b << 2) + 2;
    if (c.y == key) return (b << 2) + 3;
    if (a.x == GVS_EMPTY_KEY || a.y == GVS_EMPTY_KEY || c.x == GVS_EMPTY_KEY || c.y == GVS_EMPTY_KEY) return ~0ull;
    b = (b + 1) & (nb - 1);
  }
  return ~0ull;
}
// probe-side lookup: the bucket's four value words are requested together with its keys, so that a hit costs ONE
// round trip to HBM instead of two dependent ones (the probe's warps wait on exactly this chain); a miss -- the
// common case, a false positive of the filters -- reads one sector more than it needs
__device__ __forceinline__ u32 tab_lookup_spec(const TabView& t, u64 key, u64 h, u32* gidx) {
  const u64 nb = t.slots >> 2;
  u64 b = gvs_tab_bucket(h, t.slots);
  for (u64 it = 0; it < nb; it++) {
    const ulonglong2* p = (const ulonglong2*)(t.keys + (b << 2));
    const ulonglong2* q = (const ulonglong2*)(t.val + (b << 2));
    const ulonglong2 a = tab_ld16(p), c = tab_ld16(p + 1);
    const ulonglong2 va = tab_ld16(q), vc = tab_ld16(q + 1);
    u64 v;
    bool found = true;
    if (a.x == key) v = va.x;
    else if (a.y == key) v = va.y;
    else if (c.x == key) v = vc.x;
    else if (c.y == key) v = vc.y;
    else found = false;
    if (found) {
      *gidx = (u32)(v >> 32);
      return (u32)v;
    }
    if (a.x == GVS_EMPTY_KEY || a.y == GVS_EMPTY_KEY || c.x == GVS_EMPTY_KEY || c.y == GVS_EMPTY_KEY) return GVS_NOHIT;
    b = (b + 1) & (nb - 1);
  }
  return GVS_NOHIT;
}
// returns row, GVS_ROW_MISSING, or GVS_NOHIT; *gidx = group index of the row (valid for a real row)
__device__ __forceinline__ u32 tab_lookup(const TabView& t, u64 key, u64 h, u32* gidx) {
  u64 s = tab_find(t, key, h);
  if (s == ~0ull) return GVS_NOHIT;
  const u64 v = tab_ld8(t.val + s);
  *gidx = (u32)(v >> 32);
  return (u32)v;
}

// find-or-insert; returns slot index (never fails while load < 1)
__device__ __forceinline__ u64 tab_insert(u64* keys, u64 slots, u64 key, u64 h) {
  u64 nb = slots >> 2;
  u64 b = gvs_tab_bucket(h, slots);
  for (;;) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      u64 s = (b << 2) + j;
      u64 cur = keys[s];
      if (cur == key) return s;
      if (cur == GVS_EMPTY_KEY) {
        u64 old = atomicCAS((unsigned long long*)&keys[s], (unsigned long long)GVS_EMPTY_KEY,
                            (unsigned long long)key);
        if (old == GVS_EMPTY_KEY || old == key) return s;
      }
    }
    b = (b + 1) & (nb - 1);
  }
}
