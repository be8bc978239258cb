// ctx.cu -- context life cycle, stream / profiling control, read-batch registration.
#include "common.cuh"

extern "C" const char* gvs_version(void) { return "gavisunk_b200 0.1 (sm_100a)"; }

extern "C" gvs_ctx* gvs_create(int device, int k) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return nullptr;  // no CPU fallback
  if (k < 1 || k > 32) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  gvs_ctx* c = new gvs_ctx();
  c->device = device;
  c->k = k;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->n_sm = prop.multiProcessorCount;
  for (int i = 0; i < GVS_ST_COUNT; i++) {
    cudaEventCreate(&c->ev0[i]);
    cudaEventCreate(&c->ev1[i]);
    c->ev_valid[i] = false;
  }
  if (gvs_reserve(c, c->counters, 64 * sizeof(u64)) != 0) {
    delete c;
    return nullptr;
  }
  return c;
}

static void release_rows(Rows& r) {
  gvs_release(r.read); gvs_release(r.pos); gvs_release(r.contig);
  gvs_release(r.start); gvs_release(r.group); gvs_release(r.gidx);
}

extern "C" void gvs_destroy(gvs_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  DevBuf* all[] = {&c->loc_kmer, &c->loc_contig, &c->loc_start, &c->loc_group, &c->loc_gidx, &c->grp_contig,
                   &c->grp_start, &c->tab_keys, &c->tab_rows, &c->tab_gidx, &c->hit_gidx, &c->hit_nf, &c->ohit_gidx, &c->ohit_nf, &c->filt, &c->filt1, &c->contig_hap, &c->contig_hash,
                   &c->contig_len, &c->own_seq, &c->own_off, &c->chunk_first, &c->chunk_hap, &c->tile_first,
                   &c->tile_cnt, &c->tile_off, &c->tile_dst, &c->hit_read, &c->hit_w, &c->hit_row, &c->ohit_read,
                   &c->ohit_w, &c->ohit_row, &c->counters, &c->scan_tmp, &c->scan_tmp2, &c->flags_a, &c->flags_b,
                   &c->flags_c, &c->seg_start, &c->seg_ndist, &c->seg_cap, &c->seg_best, &c->seg_good, &c->seg_dir,
                   &c->diag_scratch, &c->best_read, &c->best_contig, &c->best_good, &c->best_dir, &c->hist,
                   &c->cnt_hist, &c->bad_flag, &c->bad_list, &c->kseg_start, &c->val_scratch, &c->val_cnt,
                   &c->val_off, &c->pair_read, &c->pair_contig, &c->pair_group, &c->pair_gidx, &c->parent,
                   &c->present, &c->comp_min, &c->comp_max, &c->comp_cnt, &c->iv_contig, &c->iv_start, &c->iv_end,
                   &c->gap_contig, &c->gap_start, &c->gap_end, &c->nodata_contig, &c->read_len, &c->gt_keys, &c->gt_val};
  for (DevBuf* b : all) gvs_release(*b);
  release_rows(c->rows);
  release_rows(c->kept);
  for (int i = 0; i < GVS_ST_COUNT; i++) {
    cudaEventDestroy(c->ev0[i]);
    cudaEventDestroy(c->ev1[i]);
  }
  for (int i = 0; i < 2; i++) {
    if (c->aux[i]) {
      cudaStreamSynchronize(c->aux[i]);
      cudaStreamDestroy(c->aux[i]);
    }
    if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  for (cudaEvent_t e : c->seg_ev) cudaEventDestroy(e);
  if (c->ev_reads_free) cudaEventDestroy(c->ev_reads_free);
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
  }
  delete c;
}

extern "C" const char* gvs_last_error(gvs_ctx* c) { return c ? c->err.c_str() : "null context (no CUDA device?)"; }

extern "C" int gvs_set_stream(gvs_ctx* ctx, void* s) {
  if (!ctx) return GVS_E_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->stream = (cudaStream_t)s;
  return 0;
}

extern "C" int gvs_sync(gvs_ctx* ctx) {
  if (!ctx) return GVS_E_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int gvs_set_profiling(gvs_ctx* ctx, int on) {
  if (!ctx) return GVS_E_ARG;
  ctx->profiling = on != 0;
  for (int i = 0; i < GVS_ST_COUNT; i++) ctx->ev_valid[i] = false;
  return 0;
}

extern "C" int gvs_stage_ms(gvs_ctx* ctx, int stage, float* ms) {
  if (!ctx || stage < 0 || stage >= GVS_ST_COUNT || !ms) return GVS_E_ARG;
  if (!ctx->ev_valid[stage]) return gvs_fail(ctx, GVS_E_STATE, "stage %d was not timed", stage);
  CK(cudaSetDevice(ctx->device));
  CK(cudaEventSynchronize(ctx->ev1[stage]));
  CK(cudaEventElapsedTime(ms, ctx->ev0[stage], ctx->ev1[stage]));
  return 0;
}

extern "C" int gvs_set_copy_pipeline(gvs_ctx* ctx, uint64_t min_bytes, uint32_t segments) {
  if (!ctx) return GVS_E_ARG;
  if (segments > 4096) return gvs_fail(ctx, GVS_E_ARG, "at most 4096 copy segments");
  ctx->seg_min_bytes = min_bytes;
  ctx->seg_count = segments;
  return 0;
}

extern "C" int gvs_host_alloc(uint64_t bytes, void** p) {
  if (!p) return GVS_E_ARG;
  *p = nullptr;
  if (cudaHostAlloc(p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    *p = nullptr;
    return GVS_E_NOMEM;
  }
  return 0;
}
extern "C" void gvs_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

extern "C" uint64_t gvs_launch_count(gvs_ctx* ctx) { return ctx ? ctx->launches : 0; }

// packed: seq holds 2-bit words (16 bases per big-endian u32, zero-padded to a whole word) instead of ASCII
static int reads_set_impl(gvs_ctx* ctx, const uint8_t* seq, bool packed, const uint64_t* read_off, uint64_t n_reads,
                          const uint64_t* chunk_first, const uint8_t* chunk_hap, uint32_t n_chunks, int on_device) {
  if (!ctx) return GVS_E_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!read_off || !chunk_first || !chunk_hap || n_chunks == 0) return gvs_fail(ctx, GVS_E_ARG, "null read batch arrays");
  if (n_reads >= 0xFFFFFFF0ull) return gvs_fail(ctx, GVS_E_OVERFLOW, "more than 2^32 reads in one batch");
  ctx->reads_ready = false;
  ctx->match_ready = ctx->diag_ready = ctx->val_ready = false;
  if (chunk_first[0] != 0 || chunk_first[n_chunks] != n_reads) return gvs_fail(ctx, GVS_E_ARG, "chunk_first must span [0, n_reads]");
  for (u32 c = 0; c < n_chunks; c++)
    if (chunk_first[c] > chunk_first[c + 1]) return gvs_fail(ctx, GVS_E_ARG, "chunk_first not monotone");
  ctx->h_chunk_first.assign(chunk_first, chunk_first + n_chunks + 1);
  ctx->h_chunk_hap.assign(chunk_hap, chunk_hap + n_chunks);
  ctx->n_chunks = n_chunks;
  CKR(to_dev(ctx, ctx->chunk_first, chunk_first, (size_t)n_chunks + 1));
  CKR(to_dev(ctx, ctx->chunk_hap, chunk_hap, (size_t)n_chunks));
  u64 total = 0;
  if (on_device) {
    if (ctx->copy_stream) CK(cudaStreamSynchronize(ctx->copy_stream));
    ctx->seg_tile_end.clear();
    CKR(read_dev(ctx, read_off + n_reads, &total));
    ctx->seq = seq;
    ctx->read_off = read_off;
  } else {
    total = read_off[n_reads];
    const u64 nbytes = packed ? 4 * cdiv(total, 16) : total;                  // bytes of the sequence buffer
    const u64 tile_bytes = packed ? GVS_TILE_BASES / 4 : GVS_TILE_BASES;      // ... per probe tile
    // 64 bytes of slack: the probe kernel stages 16-byte vectors and may touch the tail
    CKR(gvs_reserve(ctx, ctx->own_seq, nbytes + 64));
    CKR(gvs_reserve(ctx, ctx->own_off, (n_reads + 1) * sizeof(u64)));
    CK(cudaMemcpyAsync(ctx->own_off.p, read_off, (n_reads + 1) * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    ctx->seg_tile_end.clear();
    const u64 n_tiles = cdiv(total, GVS_TILE_BASES);
    u64 n_seg = (nbytes >= ctx->seg_min_bytes && ctx->seg_count > 1) ? ctx->seg_count : 1;
    if (n_seg > n_tiles) n_seg = n_tiles ? n_tiles : 1;
    if (n_seg == 1) {
      if (nbytes) CK(cudaMemcpyAsync(ctx->own_seq.p, seq, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    } else {
      if (!ctx->copy_stream) CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
      if (!ctx->ev_reads_free) CK(cudaEventCreateWithFlags(&ctx->ev_reads_free, cudaEventDisableTiming));
      while (ctx->seg_ev.size() < n_seg) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->seg_ev.push_back(e);
      }
      // the previous batch may still be read by work queued on the compute stream
      CK(cudaEventRecord(ctx->ev_reads_free, ctx->stream));
      CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_reads_free, 0));
      u64 c0 = 0;
      for (u64 s = 0; s < n_seg; s++) {
        u64 t1 = (s + 1 == n_seg) ? n_tiles : (s + 1) * n_tiles / n_seg;
        // the probe prefetches two tiles past the end of its span (the halo of the last windows)
        u64 c1 = (s + 1 == n_seg) ? nbytes : (t1 + 2) * tile_bytes;
        if (c1 > nbytes) c1 = nbytes;
        if (c1 > c0) CK(cudaMemcpyAsync((u8*)ctx->own_seq.p + c0, seq + c0, c1 - c0, cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->seg_ev[s], ctx->copy_stream));
        ctx->seg_tile_end.push_back(t1);
        c0 = c1;
      }
    }
    ctx->seq = ctx->own_seq.as<u8>();
    ctx->read_off = ctx->own_off.as<u64>();
  }
  ctx->seq_packed = packed;
  ctx->n_reads = n_reads;
  ctx->total_bases = total;
  ctx->reads_ready = true;
  return 0;
}

extern "C" int gvs_reads_set(gvs_ctx* ctx, const uint8_t* seq, const uint64_t* read_off, uint64_t n_reads,
                             const uint64_t* chunk_first, const uint8_t* chunk_hap, uint32_t n_chunks, int on_device) {
  return reads_set_impl(ctx, seq, false, read_off, n_reads, chunk_first, chunk_hap, n_chunks, on_device);
}
extern "C" int gvs_reads_set_packed(gvs_ctx* ctx, const uint32_t* words, const uint64_t* read_off, uint64_t n_reads,
                                    const uint64_t* chunk_first, const uint8_t* chunk_hap, uint32_t n_chunks, int on_device) {
  return reads_set_impl(ctx, (const uint8_t*)words, true, read_off, n_reads, chunk_first, chunk_hap, n_chunks, on_device);
}
