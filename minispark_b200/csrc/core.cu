// core.cu -- context lifetime, stream-ordered device memory, error word, relation handles.
#include "common.cuh"

static void release_cached_blocks(msc_ctx* ctx, size_t keep_bytes) {
  while (ctx->big_free_bytes > keep_bytes && !ctx->big_free.empty()) {
    auto it = std::prev(ctx->big_free.end());  // largest first
    cudaFreeAsync(it->second, ctx->stream);
    ctx->big_free_bytes -= it->first;
    ctx->big_free.erase(it);
  }
}

int msc_alloc(msc_ctx* ctx, size_t nbytes, void** out) {
  if (nbytes == 0) nbytes = 16;
  if (nbytes >= msc_ctx::kBigBlock) {
    auto it = ctx->big_free.lower_bound(nbytes);
    if (it != ctx->big_free.end() && it->first <= nbytes + nbytes / 8) {
      *out = it->second;
      ctx->big_live[it->second] = it->first;
      ctx->big_free_bytes -= it->first;
      ctx->big_free.erase(it);
      ctx->stats.device_bytes += nbytes;
      return MSC_OK;
    }
  }
  cudaError_t e = cudaMallocAsync(out, nbytes, ctx->stream);
  if (e == cudaErrorMemoryAllocation && !ctx->big_free.empty()) {  // give the cache back and try once more
    cudaGetLastError();
    release_cached_blocks(ctx, 0);
    cudaStreamSynchronize(ctx->stream);
    e = cudaMallocAsync(out, nbytes, ctx->stream);
  }
  MSC_CUDA(ctx, e);
  if (nbytes >= msc_ctx::kBigBlock) ctx->big_live[*out] = nbytes;
  ctx->stats.device_bytes += nbytes;
  return MSC_OK;
}

int msc_free(msc_ctx* ctx, void* p, size_t nbytes) {
  if (!p) return MSC_OK;
  if (nbytes == 0) nbytes = 16;
  if (!ctx->run_index.empty()) {  // what was derived from this memory goes with it
    auto it = ctx->run_index.lower_bound(p);
    while (it != ctx->run_index.end() && static_cast<const char*>(it->first) < static_cast<const char*>(p) + nbytes) {
      void* offsets = it->second.offsets;
      const size_t bytes = it->second.offsets_bytes;
      it = ctx->run_index.erase(it);
      msc_free(ctx, offsets, bytes);  // (a small block: it cannot itself be the key of an entry)
      it = ctx->run_index.lower_bound(p);
    }
  }
  ctx->stats.device_bytes -= nbytes;
  auto live = ctx->big_live.find(p);
  if (live != ctx->big_live.end()) {
    const size_t real = live->second;
    ctx->big_live.erase(live);
    ctx->big_free.emplace(real, p);
    ctx->big_free_bytes += real;
    if (ctx->big_free_bytes > ctx->big_free_cap) release_cached_blocks(ctx, ctx->big_free_cap / 2);
    return MSC_OK;
  }
  MSC_CUDA(ctx, cudaFreeAsync(p, ctx->stream));
  return MSC_OK;
}

int msc_alloc_rows(msc_ctx* ctx, uint64_t nrows, size_t width, void** out, size_t* bytes_out) {
  const size_t bytes = static_cast<size_t>(msc_round_up(nrows ? nrows : 1, MSC_ROW_PAD)) * width + 256;
  MSC_TRY(msc_alloc(ctx, bytes, out));
  // the padding is bulk-copied by the scan kernel; keep it defined (rows past nrows are masked)
  const size_t used = static_cast<size_t>(nrows) * width;
  MSC_CUDA(ctx, cudaMemsetAsync(static_cast<char*>(*out) + used, 0, bytes - used, ctx->stream));
  if (bytes_out) *bytes_out = bytes;
  return MSC_OK;
}

int msc_device_error_rc(msc_ctx* ctx, int e) {
  if (e == 0) return MSC_OK;
  if (e & MSC_DEVERR_DIV_ZERO) return ctx->fail(MSC_ERR_DIV_ZERO, "division by zero");
  if (e & MSC_DEVERR_OVERFLOW) return ctx->fail(MSC_ERR_OVERFLOW, "int too big to convert");
  if (e & MSC_DEVERR_COLLISION) return ctx->fail(MSC_ERR_COLLISION, "string hash collision in dictionary");
  if (e & MSC_DEVERR_STRLEN) return ctx->fail(MSC_ERR_STRLEN, "string longer than 255 bytes");
  if (e & MSC_DEVERR_PEER_TIMEOUT) return ctx->fail(MSC_ERR_PEER, "a peer GPU did not deliver its part of an exchange in time");
  if (e & MSC_DEVERR_IO) return ctx->fail(MSC_ERR_IO, "malformed BlockFile: a string block's lengths exceed its bytes");
  return ctx->fail(MSC_ERR_ARG, "hash table full");
}

int msc_check_device_error(msc_ctx* ctx) {
  MSC_CUDA(ctx, cudaMemcpyAsync(ctx->h_err, ctx->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int e = *ctx->h_err;
  if (e == 0) return MSC_OK;
  MSC_CUDA(ctx, cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
  return msc_device_error_rc(ctx, e);
}

extern "C" int msc_abi_version(void) { return MSC_ABI_VERSION; }

extern "C" int msc_create(int device, msc_ctx** out) {
  if (!out) return MSC_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return MSC_ERR_CUDA;
  msc_ctx* ctx = new msc_ctx();
  ctx->device = device;
  auto bail = [&](const char* what) {
    fprintf(stderr, "msc_create: %s failed: %s\n", what, cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return MSC_ERR_CUDA;
  };
  if (cudaSetDevice(device) != cudaSuccess) return bail("cudaSetDevice");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail("cudaGetDeviceProperties");
  if (prop.major != 10) {
    fprintf(stderr, "msc_create: device %d is sm_%d%d; this library is built for sm_100a only\n", device, prop.major, prop.minor);
    delete ctx;
    return MSC_ERR_CUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->big_free_cap = prop.totalGlobalMem / 2;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("cudaStreamCreate");
  for (auto& s : ctx->copy)
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return bail("cudaStreamCreate");
  if (cudaEventCreate(&ctx->ev_a) != cudaSuccess || cudaEventCreate(&ctx->ev_b) != cudaSuccess) return bail("cudaEventCreate");
  if (cudaEventCreate(&ctx->ev_s0) != cudaSuccess || cudaEventCreate(&ctx->ev_s1) != cudaSuccess) return bail("cudaEventCreate");
  if (cudaEventCreate(&ctx->ev_t0) != cudaSuccess || cudaEventCreate(&ctx->ev_t1) != cudaSuccess) return bail("cudaEventCreate");
  for (auto& e : ctx->ring_ev)
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return bail("cudaEventCreate");
  // keep freed blocks in the pool: relations are created and dropped on every query
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t threshold = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
  if (cudaMalloc(&ctx->d_err, sizeof(int)) != cudaSuccess) return bail("cudaMalloc");
  if (cudaMemset(ctx->d_err, 0, sizeof(int)) != cudaSuccess) return bail("cudaMemset");
  if (cudaHostAlloc(&ctx->h_err, sizeof(int), cudaHostAllocDefault) != cudaSuccess) return bail("cudaHostAlloc");
  if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_scratch), 16 * sizeof(unsigned long long), cudaHostAllocDefault) != cudaSuccess)
    return bail("cudaHostAlloc");
  *out = ctx;
  return MSC_OK;
}

extern "C" void msc_destroy(msc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (auto& e : ctx->run_index)
    if (e.second.offsets) cudaFreeAsync(e.second.offsets, ctx->stream);
  ctx->run_index.clear();
  release_cached_blocks(ctx, 0);
  cudaStreamSynchronize(ctx->stream);
  for (auto& r : ctx->ring)
    if (r) cudaFreeHost(r);
  for (auto& e : ctx->ring_ev)
    if (e) cudaEventDestroy(e);
  if (ctx->d_ticket) cudaFreeAsync(ctx->d_ticket, ctx->stream);
  if (ctx->d_err) cudaFree(ctx->d_err);
  if (ctx->h_err) cudaFreeHost(ctx->h_err);
  if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  if (ctx->ev_s0) cudaEventDestroy(ctx->ev_s0);
  if (ctx->ev_s1) cudaEventDestroy(ctx->ev_s1);
  if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
  if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
  for (auto& s : ctx->copy)
    if (s) cudaStreamDestroy(s);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* msc_last_error(msc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int msc_sync(msc_ctx* ctx) {
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MSC_OK;
}

extern "C" int msc_stream_handle(msc_ctx* ctx, void** stream) {
  if (!ctx || !stream) return MSC_ERR_ARG;
  *stream = ctx->stream;
  return MSC_OK;
}

extern "C" int msc_timer_start(msc_ctx* ctx) {
  if (!ctx) return MSC_ERR_ARG;
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
  return MSC_OK;
}

extern "C" int msc_timer_stop(msc_ctx* ctx, double* ms) {
  if (!ctx || !ms) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_t1, ctx->stream));
  MSC_CUDA(ctx, cudaEventSynchronize(ctx->ev_t1));
  float f = 0;
  MSC_CUDA(ctx, cudaEventElapsedTime(&f, ctx->ev_t0, ctx->ev_t1));
  *ms = f;
  return MSC_OK;
}

extern "C" int msc_get_stats(msc_ctx* ctx, msc_stats* out) {
  if (!ctx || !out) return MSC_ERR_ARG;
  *out = ctx->stats;
  return MSC_OK;
}

extern "C" int msc_host_alloc(msc_ctx* ctx, size_t nbytes, void** out) {
  MSC_CUDA(ctx, cudaHostAlloc(out, nbytes ? nbytes : 16, cudaHostAllocDefault));
  return MSC_OK;
}

extern "C" int msc_host_free(msc_ctx* ctx, void* p) {
  MSC_CUDA(ctx, cudaFreeHost(p));
  return MSC_OK;
}

extern "C" int msc_dev_alloc(msc_ctx* ctx, size_t nbytes, void** out) {
  MSC_TRY(msc_alloc(ctx, nbytes ? nbytes : 16, out));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MSC_OK;
}

extern "C" int msc_dev_free(msc_ctx* ctx, void* p) {
  if (!p) return MSC_OK;
  return msc_free(ctx, p, 0);
}

extern "C" int msc_memcpy_d2h(msc_ctx* ctx, void* host_dst, const void* dev_src, size_t nbytes) {
  if (nbytes == 0) return MSC_OK;
  MSC_CUDA(ctx, cudaMemcpyAsync(host_dst, dev_src, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MSC_OK;
}

extern "C" int msc_memcpy_h2d(msc_ctx* ctx, void* dev_dst, const void* host_src, size_t nbytes) {
  if (nbytes == 0) return MSC_OK;
  MSC_CUDA(ctx, cudaMemcpyAsync(dev_dst, host_src, nbytes, cudaMemcpyHostToDevice, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MSC_OK;
}

// ---- relations -----------------------------------------------------------------------------------
extern "C" int msc_rel_info(msc_rel* r, uint64_t* nrows, int32_t* ncols) {
  if (!r) return MSC_ERR_ARG;
  if (nrows) *nrows = r->nrows;
  if (ncols) *ncols = static_cast<int32_t>(r->cols.size());
  return MSC_OK;
}

extern "C" int msc_rel_col(msc_rel* r, int32_t col, void** dev_ptr, int32_t* phys) {
  if (!r || col < 0 || col >= static_cast<int32_t>(r->cols.size())) return MSC_ERR_ARG;
  if (dev_ptr) *dev_ptr = r->cols[col].data;
  if (phys) *phys = r->cols[col].phys;
  return MSC_OK;
}

extern "C" int msc_rel_cols(msc_rel* r, msc_colbind* cols, int32_t ncols) {
  if (!r || !cols || ncols != static_cast<int32_t>(r->cols.size())) return MSC_ERR_ARG;
  for (int32_t i = 0; i < ncols; ++i) {
    cols[i].data = r->cols[i].data;
    cols[i].phys = r->cols[i].phys;
  }
  return MSC_OK;
}

extern "C" void msc_rel_free(msc_rel* r) {
  if (!r) return;
  for (auto& c : r->cols)
    if (c.owned && c.data) msc_free(r->ctx, c.data, c.bytes);
  if (r->d_meta) msc_free(r->ctx, r->d_meta, 3 * sizeof(unsigned long long));
  delete r;
}

// ---- peer mailboxes: device memory other ranks' GPUs write into over NVLink (CUDA IPC, one process per GPU) ----------
extern "C" int msc_peer_alloc(msc_ctx* ctx, size_t nbytes, void** dev_ptr, void* handle64) {
  if (!ctx || !dev_ptr || !handle64 || nbytes == 0) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  MSC_CUDA(ctx, cudaSetDevice(ctx->device));
  MSC_CUDA(ctx, cudaMalloc(dev_ptr, nbytes));  // IPC needs a cudaMalloc allocation, not one from the stream-ordered pool
  MSC_CUDA(ctx, cudaMemset(*dev_ptr, 0, nbytes));
  cudaIpcMemHandle_t h;
  MSC_CUDA(ctx, cudaIpcGetMemHandle(&h, *dev_ptr));
  memcpy(handle64, &h, sizeof(h));
  return MSC_OK;
}

extern "C" int msc_peer_open(msc_ctx* ctx, const void* handle64, void** peer_ptr) {
  if (!ctx || !handle64 || !peer_ptr) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  MSC_CUDA(ctx, cudaSetDevice(ctx->device));
  MSC_CUDA(ctx, cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MSC_OK;
}

extern "C" int msc_peer_close(msc_ctx* ctx, void* peer_ptr) {
  if (!ctx) return MSC_ERR_ARG;
  if (peer_ptr) MSC_CUDA(ctx, cudaIpcCloseMemHandle(peer_ptr));
  return MSC_OK;
}

extern "C" int msc_peer_free(msc_ctx* ctx, void* dev_ptr) {
  if (!ctx) return MSC_ERR_ARG;
  if (dev_ptr) MSC_CUDA(ctx, cudaFree(dev_ptr));
  return MSC_OK;
}

extern "C" int msc_rel_nrows_dev(msc_rel* r, const uint64_t** nrows_dev) {
  if (!r || !nrows_dev || !r->d_meta) return MSC_ERR_ARG;
  *nrows_dev = reinterpret_cast<const uint64_t*>(r->d_meta);
  return MSC_OK;
}

extern "C" int msc_rel_settle(msc_ctx* ctx, msc_rel* const* rels, int32_t nrels, int32_t* nonfinite) {
  if (!ctx || (nrels && !rels) || nrels < 0 || nrels > 5) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  unsigned long long* h = ctx->h_scratch;  // 16 pinned words: 3 per relation
  for (int i = 0; i < nrels; ++i)
    if (rels[i] && rels[i]->pending)
      MSC_CUDA(ctx, cudaMemcpyAsync(h + 3 * i, rels[i]->d_meta, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int err = 0, nf = 0;
  for (int i = 0; i < nrels; ++i) {
    if (!rels[i] || !rels[i]->pending) continue;
    rels[i]->nrows = h[3 * i];
    nf |= h[3 * i + 1] != 0;
    err |= static_cast<int>(h[3 * i + 2]);
    rels[i]->pending = false;
  }
  if (nonfinite) *nonfinite = nf;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b) == cudaSuccess) ctx->stats.last_kernel_ms = ms;
  if (cudaEventElapsedTime(&ms, ctx->ev_s0, ctx->ev_s1) == cudaSuccess) ctx->stats.last_scan_ms = ms;
  return msc_device_error_rc(ctx, err);
}

extern "C" int msc_rel_alloc(msc_ctx* ctx, uint64_t nrows, const int32_t* phys, int32_t ncols, msc_rel** out) {
  if (!ctx || !out || ncols < 0 || (ncols && !phys)) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  msc_rel* r = new msc_rel();
  r->ctx = ctx;
  r->nrows = nrows;
  for (int i = 0; i < ncols; ++i) {
    msc_col c;
    c.phys = phys[i];
    const size_t w = msc_phys_width(phys[i]);
    int rc = w ? msc_alloc_rows(ctx, nrows, w, &c.data, &c.bytes) : ctx->fail(MSC_ERR_ARG, "bad physical type");
    if (rc == MSC_OK && cudaMemsetAsync(c.data, 0, c.bytes, ctx->stream) != cudaSuccess) rc = ctx->fail(MSC_ERR_CUDA, "memset failed");
    if (rc != MSC_OK) {
      msc_rel_free(r);
      return rc;
    }
    r->cols.push_back(c);
  }
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out = r;
  return MSC_OK;
}

extern "C" int msc_rel_wrap(msc_ctx* ctx, uint64_t nrows, const msc_colbind* cols, int32_t ncols, msc_rel** out) {
  if (!ctx || !out || ncols < 0) return MSC_ERR_ARG;
  msc_rel* r = new msc_rel();
  r->ctx = ctx;
  r->nrows = nrows;
  for (int i = 0; i < ncols; ++i) {
    msc_col c;
    c.data = const_cast<void*>(cols[i].data);
    c.phys = cols[i].phys;
    c.owned = false;
    r->cols.push_back(c);
  }
  *out = r;
  return MSC_OK;
}

extern "C" int msc_rel_copy_column(msc_ctx* ctx, msc_rel* r, int32_t col, void* host_dst, size_t cap_bytes) {
  if (!r || col < 0 || col >= static_cast<int32_t>(r->cols.size())) return ctx->fail(MSC_ERR_ARG, "bad column");
  const size_t need = static_cast<size_t>(r->nrows) * msc_phys_width(r->cols[col].phys);
  if (need > cap_bytes) return ctx->fail(MSC_ERR_ARG, "host buffer too small");
  return msc_memcpy_d2h(ctx, host_dst, r->cols[col].data, need);
}
