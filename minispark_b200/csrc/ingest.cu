// ingest.cu -- BlockFile -> device-resident columns.
//
// Replaces the reference's read path: BlockFile._deserialize_block / _deserialize_block_column
// (src/mini_spark/io.py:112-163, one f.read per value and one pass per column),
// LoadTableBlockTask.generate_chunks (tasks.py:117-121) and the Zig twin ColumnData.readColumn /
// Block.readBlock / LoadTableBlockProducer.next (zig-src/src/block_file.zig:225-268,297-306;
// tasks.zig:212-222).  Differences by design:
//   * column pruning: only the requested columns' byte ranges are touched (the reference reads all);
//   * INTEGER / FLOAT / TIMESTAMP payloads are already in device-native layout, so in the native
//     layout they are DMA'd straight into their final column buffers (no decode pass at all);
//     the wide layout (i64 / f64) runs a widen kernel per block;
//   * STRING payloads (u8 lengths + bytes) are turned into offsets by a device prefix sum and
//     dictionary-encoded on the device; the codes are narrowed to u8/u16 once the load is complete.
// Host->device traffic goes through cudaMemcpyAsync on two copy streams, from the caller's (pinned)
// BlockFile image or, for files, through a pinned staging ring filled by pread.
#include <errno.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>

#include "strings.cuh"

struct BlockInfo {
  uint64_t start = 0;
  uint32_t rows = 0;
  std::vector<uint64_t> col_off;    // file offset of each column payload
  std::vector<uint64_t> col_bytes;  // payload bytes
};

struct msc_table {
  msc_ctx* ctx = nullptr;
  int fd = -1;
  const uint8_t* image = nullptr;
  uint64_t size = 0;
  std::vector<std::string> names;
  std::vector<int> types;
  std::vector<BlockInfo> blocks;
  uint64_t nrows = 0;
};

namespace {

int read_at(msc_table* t, uint64_t off, void* dst, size_t n) {
  if (n > t->size || off > t->size - n) return t->ctx->fail(MSC_ERR_IO, "BlockFile truncated");  // (overflow-safe: both come from the file)
  if (t->image) {
    memcpy(dst, t->image + off, n);
    return MSC_OK;
  }
  size_t done = 0;
  while (done < n) {
    const ssize_t r = pread(t->fd, static_cast<char*>(dst) + done, n - done, static_cast<off_t>(off + done));
    if (r <= 0) return t->ctx->fail(MSC_ERR_IO, std::string("pread failed: ") + strerror(errno));
    done += static_cast<size_t>(r);
  }
  return MSC_OK;
}

int parse_table(msc_table* t) {
  msc_ctx* ctx = t->ctx;
  if (t->size < 5) return ctx->fail(MSC_ERR_IO, "BlockFile too small");
  uint8_t ncols = 0;
  MSC_TRY(read_at(t, 0, &ncols, 1));
  uint64_t off = 1;
  for (int c = 0; c < ncols; ++c) {
    uint8_t hdr[2];
    MSC_TRY(read_at(t, off, hdr, 2));
    off += 2;
    if (hdr[0] > 3) return ctx->fail(MSC_ERR_IO, "unknown column type ordinal");
    std::string name(hdr[1], '\0');
    if (hdr[1]) MSC_TRY(read_at(t, off, &name[0], hdr[1]));
    off += hdr[1];
    t->types.push_back(hdr[0]);
    t->names.push_back(name);
  }
  uint32_t nblocks = 0;
  MSC_TRY(read_at(t, t->size - 4, &nblocks, 4));
  if (static_cast<uint64_t>(nblocks) * 8 + 4 > t->size) return ctx->fail(MSC_ERR_IO, "bad BlockFile footer");
  std::vector<uint64_t> starts(nblocks);
  if (nblocks) MSC_TRY(read_at(t, t->size - 4 - 8ULL * nblocks, starts.data(), 8ULL * nblocks));
  for (uint32_t b = 0; b < nblocks; ++b) {
    BlockInfo bi;
    bi.start = starts[b];
    if (bi.start >= t->size) return ctx->fail(MSC_ERR_IO, "block start beyond end of file");
    MSC_TRY(read_at(t, bi.start, &bi.rows, 4));
    uint64_t p = bi.start + 4;
    for (int c = 0; c < ncols; ++c) {
      uint64_t nbytes = 0;
      MSC_TRY(read_at(t, p, &nbytes, 8));
      p += 8;
      if (nbytes > t->size || p > t->size - nbytes) return ctx->fail(MSC_ERR_IO, "column payload runs past end of file");
      const int ty = t->types[c];
      if (ty != MSC_T_STRING) {
        const uint64_t w = (ty == MSC_T_TIMESTAMP) ? 8 : 4;
        if (nbytes != w * bi.rows) return ctx->fail(MSC_ERR_IO, "column payload size does not match row count");
      } else if (nbytes < bi.rows) {
        return ctx->fail(MSC_ERR_IO, "string column shorter than its length prefix");
      }
      bi.col_off.push_back(p);
      bi.col_bytes.push_back(nbytes);
      p += nbytes;
    }
    t->nrows += bi.rows;
    t->blocks.push_back(std::move(bi));
  }
  return MSC_OK;
}

// a STRING block's u8 lengths must add up to no more than the bytes that follow them (else the encoder would read past
// the staged block): offsets[rows] is their sum
__global__ void check_string_total_kernel(const uint64_t* total, uint64_t limit, int* err) {
  if (*total > limit) atomicOr(err, MSC_DEVERR_IO);
}

__global__ void widen_i32_kernel(const int* in, long long* out, uint64_t n) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    out[i] = in[i];
}
__global__ void widen_f32_kernel(const float* in, double* out, uint64_t n) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    out[i] = static_cast<double>(in[i]);
}
template <class TOut>
__global__ void narrow_codes_kernel(const uint32_t* in, TOut* out, uint64_t n) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    out[i] = static_cast<TOut>(in[i]);
}

struct Loader {
  msc_ctx* ctx;
  msc_table* t;
  int next_stream = 0;
  int next_ring = 0;
  uint64_t bytes = 0;

  // host (file or image) -> device, asynchronously on a copy stream
  int copy_in(uint64_t file_off, uint64_t nbytes, void* dev_dst, cudaStream_t* used) {
    cudaStream_t cs = ctx->copy[next_stream];
    next_stream ^= 1;
    *used = cs;
    bytes += nbytes;
    if (nbytes == 0) return MSC_OK;
    if (t->image) {
      MSC_CUDA(ctx, cudaMemcpyAsync(dev_dst, t->image + file_off, nbytes, cudaMemcpyHostToDevice, cs));
      return MSC_OK;
    }
    if (ctx->ring_bytes == 0) {
      ctx->ring_bytes = 32ULL << 20;
      for (auto& r : ctx->ring) MSC_CUDA(ctx, cudaHostAlloc(&r, ctx->ring_bytes, cudaHostAllocDefault));
    }
    uint64_t done = 0;
    while (done < nbytes) {
      const uint64_t chunk = std::min<uint64_t>(ctx->ring_bytes, nbytes - done);
      const int slot = next_ring;
      next_ring = (next_ring + 1) % msc_ctx::kRing;
      MSC_CUDA(ctx, cudaEventSynchronize(ctx->ring_ev[slot]));  // previous DMA out of this slot is done
      MSC_TRY(read_at(t, file_off + done, ctx->ring[slot], chunk));
      MSC_CUDA(ctx, cudaMemcpyAsync(static_cast<char*>(dev_dst) + done, ctx->ring[slot], chunk, cudaMemcpyHostToDevice, cs));
      MSC_CUDA(ctx, cudaEventRecord(ctx->ring_ev[slot], cs));
      done += chunk;
    }
    return MSC_OK;
  }
};

}  // namespace

extern "C" int msc_table_open(msc_ctx* ctx, const char* path, msc_table** out) {
  if (!ctx || !path || !out) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  msc_table* t = new msc_table();
  t->ctx = ctx;
  t->fd = open(path, O_RDONLY);
  if (t->fd < 0) {
    const int rc = ctx->fail(MSC_ERR_IO, std::string("cannot open ") + path + ": " + strerror(errno));
    delete t;
    return rc;
  }
  struct stat st;
  if (fstat(t->fd, &st) != 0) {
    close(t->fd);
    delete t;
    return ctx->fail(MSC_ERR_IO, "fstat failed");
  }
  t->size = static_cast<uint64_t>(st.st_size);
  const int rc = parse_table(t);
  if (rc != MSC_OK) {
    msc_table_close(t);
    return rc;
  }
  *out = t;
  return MSC_OK;
}

extern "C" int msc_table_open_mem(msc_ctx* ctx, const void* image, size_t nbytes, msc_table** out) {
  if (!ctx || !image || !out) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  msc_table* t = new msc_table();
  t->ctx = ctx;
  t->image = static_cast<const uint8_t*>(image);
  t->size = nbytes;
  const int rc = parse_table(t);
  if (rc != MSC_OK) {
    delete t;
    return rc;
  }
  *out = t;
  return MSC_OK;
}

extern "C" void msc_table_close(msc_table* t) {
  if (!t) return;
  if (t->fd >= 0) close(t->fd);
  delete t;
}

extern "C" int msc_table_info(msc_table* t, int32_t* ncols, int32_t* nblocks, uint64_t* nrows) {
  if (!t) return MSC_ERR_ARG;
  if (ncols) *ncols = static_cast<int32_t>(t->types.size());
  if (nblocks) *nblocks = static_cast<int32_t>(t->blocks.size());
  if (nrows) *nrows = t->nrows;
  return MSC_OK;
}

extern "C" int msc_table_col_info(msc_table* t, int32_t col, int32_t* type, char* name, int32_t name_cap) {
  if (!t || col < 0 || col >= static_cast<int32_t>(t->types.size())) return MSC_ERR_ARG;
  if (type) *type = t->types[col];
  if (name && name_cap > 0) {
    strncpy(name, t->names[col].c_str(), name_cap - 1);
    name[name_cap - 1] = '\0';
  }
  return MSC_OK;
}

extern "C" int msc_table_block_rows(msc_table* t, int32_t block, uint32_t* rows) {
  if (!t || !rows || block < 0 || block >= static_cast<int32_t>(t->blocks.size())) return MSC_ERR_ARG;
  *rows = t->blocks[block].rows;
  return MSC_OK;
}

extern "C" int msc_table_load(msc_ctx* ctx, msc_table* t, const int32_t* cols, int32_t ncols, const int32_t* blocks,
                              int32_t nblocks, int32_t layout, msc_dict** dicts, msc_rel** out) {
  if (!ctx || !t || !out || ncols < 0 || nblocks < 0 || (ncols && !cols) || (nblocks && !blocks))
    return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  const auto t0 = std::chrono::steady_clock::now();
  static const bool trace = getenv("MSC_TRACE") != nullptr;  // phase timings of the load on stderr
  auto mark = [&](const char* what) {
    if (trace) fprintf(stderr, "[msc_table_load] %-28s +%.2f ms\n", what,
                       std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  };
  uint64_t total = 0, max_rows = 0, max_str_bytes = 0;
  for (int b = 0; b < nblocks; ++b) {
    if (blocks[b] < 0 || blocks[b] >= static_cast<int32_t>(t->blocks.size())) return ctx->fail(MSC_ERR_ARG, "bad block id");
    const BlockInfo& bi = t->blocks[blocks[b]];
    total += bi.rows;
    max_rows = std::max<uint64_t>(max_rows, bi.rows);
    for (int c = 0; c < ncols; ++c) {
      if (cols[c] < 0 || cols[c] >= static_cast<int32_t>(t->types.size())) return ctx->fail(MSC_ERR_ARG, "bad column id");
      if (t->types[cols[c]] == MSC_T_STRING) max_str_bytes = std::max(max_str_bytes, bi.col_bytes[cols[c]]);
    }
  }
  if (total >= 0xFFFFFFFFULL) return ctx->fail(MSC_ERR_ARG, "more than 2^32-1 rows per GPU is not supported");
  const bool wide = layout == MSC_LAYOUT_WIDE;

  msc_rel* rel = new msc_rel();
  rel->ctx = ctx;
  rel->nrows = total;
  auto fail = [&](int rc) {
    cudaStreamSynchronize(ctx->copy[0]);
    cudaStreamSynchronize(ctx->copy[1]);
    cudaStreamSynchronize(ctx->stream);
    msc_rel_free(rel);
    return rc;
  };
  bool any_string = false, any_widen = false;
  for (int c = 0; c < ncols; ++c) {
    const int ty = t->types[cols[c]];
    msc_col col;
    switch (ty) {
      case MSC_T_INTEGER: col.phys = wide ? MSC_P_I64 : MSC_P_I32; any_widen |= wide; break;
      case MSC_T_FLOAT: col.phys = wide ? MSC_P_F64 : MSC_P_F32; any_widen |= wide; break;
      case MSC_T_TIMESTAMP: col.phys = MSC_P_I64; break;
      default: col.phys = MSC_P_U32; any_string = true; break;
    }
    const int rc = msc_alloc_rows(ctx, total, msc_phys_width(col.phys), &col.data, &col.bytes);
    if (rc != MSC_OK) return fail(rc);
    rel->cols.push_back(col);
    if (ty == MSC_T_STRING && dicts && dicts[c] == nullptr) {
      const int drc = msc_dict_create(ctx, &dicts[c]);
      if (drc != MSC_OK) return fail(drc);
    }
    if (ty == MSC_T_STRING && !dicts) return fail(ctx->fail(MSC_ERR_ARG, "string column needs a dictionary slot"));
  }
  // device staging (two sets, alternated per block so DMA of block b+1 overlaps decode of block b)
  DevTmp stage[2] = {DevTmp(ctx), DevTmp(ctx)};
  DevTmp offs[2] = {DevTmp(ctx), DevTmp(ctx)};
  cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  const uint64_t stage_bytes = std::max<uint64_t>(any_string ? max_str_bytes : 0, any_widen ? max_rows * 4 : 0);
  if (stage_bytes) {
    for (int i = 0; i < 2; ++i) {
      int rc = stage[i].alloc(stage_bytes + 256);
      if (rc == MSC_OK && any_string) rc = offs[i].alloc((max_rows + 1) * sizeof(uint64_t));
      if (rc != MSC_OK) return fail(rc);
      if (cudaEventCreateWithFlags(&ev_copy[i], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming) != cudaSuccess)
        return fail(ctx->fail(MSC_ERR_CUDA, "cudaEventCreate failed"));
    }
  }
  // all allocations above are stream-ordered on ctx->stream: make them visible to the copy streams
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(ctx->fail(MSC_ERR_CUDA, "sync failed"));

  mark("allocated + synced");
  Loader ld{ctx, t};
  int flip = 0;
  uint64_t row_off = 0;
  int rc = MSC_OK;
  for (int b = 0; b < nblocks && rc == MSC_OK; ++b) {
    const BlockInfo& bi = t->blocks[blocks[b]];
    for (int c = 0; c < ncols && rc == MSC_OK; ++c) {
      const int fc = cols[c];
      const int ty = t->types[fc];
      msc_col& col = rel->cols[c];
      cudaStream_t cs;
      const bool direct = (ty == MSC_T_TIMESTAMP) || (!wide && ty != MSC_T_STRING);
      if (direct) {
        rc = ld.copy_in(bi.col_off[fc], bi.col_bytes[fc], static_cast<char*>(col.data) + row_off * msc_phys_width(col.phys), &cs);
        continue;
      }
      // staged path: wait until the previous user of this staging set has finished
      const int sset = flip;
      flip ^= 1;
      cs = ctx->copy[ld.next_stream];
      if (cudaStreamWaitEvent(cs, ev_done[sset], 0) != cudaSuccess) { rc = ctx->fail(MSC_ERR_CUDA, "cudaStreamWaitEvent failed"); break; }
      rc = ld.copy_in(bi.col_off[fc], bi.col_bytes[fc], stage[sset].p, &cs);
      if (rc != MSC_OK) break;
      cudaEventRecord(ev_copy[sset], cs);
      cudaStreamWaitEvent(ctx->stream, ev_copy[sset], 0);
      const unsigned grid = static_cast<unsigned>(ctx->sm_count * 8);
      if (ty == MSC_T_INTEGER) {
        widen_i32_kernel<<<grid, 256, 0, ctx->stream>>>(stage[sset].as<int>(), static_cast<long long*>(col.data) + row_off, bi.rows);
        ctx->stats.launches += 1;
      } else if (ty == MSC_T_FLOAT) {
        widen_f32_kernel<<<grid, 256, 0, ctx->stream>>>(stage[sset].as<float>(), static_cast<double*>(col.data) + row_off, bi.rows);
        ctx->stats.launches += 1;
      } else {
        const uint8_t* lens = stage[sset].as<uint8_t>();
        const uint8_t* body = lens + bi.rows;
        rc = msc_exclusive_scan_u8_u64(ctx, lens, offs[sset].as<uint64_t>(), bi.rows);
        if (rc == MSC_OK) {
          check_string_total_kernel<<<1, 1, 0, ctx->stream>>>(offs[sset].as<uint64_t>() + bi.rows, bi.col_bytes[fc] - bi.rows, ctx->d_err);
          ctx->stats.launches += 1;
        }
        if (rc == MSC_OK)
          rc = msc_dict_encode_u8_async(ctx, dicts[c], offs[sset].as<uint64_t>(), lens, body, bi.rows, bi.col_bytes[fc] - bi.rows,
                                        static_cast<uint32_t*>(col.data) + row_off);
      }
      cudaEventRecord(ev_done[sset], ctx->stream);
    }
    row_off += bi.rows;
  }
  mark("block loop issued");
  // join the copy streams into the compute stream
  for (int i = 0; i < 2; ++i) {
    if (cudaStreamSynchronize(ctx->copy[i]) != cudaSuccess && rc == MSC_OK) rc = ctx->fail(MSC_ERR_CUDA, "copy stream failed");
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == MSC_OK) rc = ctx->fail(MSC_ERR_CUDA, "compute stream failed");
  for (int i = 0; i < 2; ++i) {
    if (ev_copy[i]) cudaEventDestroy(ev_copy[i]);
    if (ev_done[i]) cudaEventDestroy(ev_done[i]);
  }
  for (int c = 0; c < ncols && rc == MSC_OK; ++c)
    if (t->types[cols[c]] == MSC_T_STRING) rc = msc_dict_settle(ctx, dicts[c]);  // host mirrors + the device error word
  if (rc != MSC_OK) return fail(rc);
  mark("copies + decode complete");

  // narrow dictionary codes (native layout): u8 for <=256 entries, u16 for <=65536
  if (!wide) {
    for (int c = 0; c < ncols; ++c) {
      if (t->types[cols[c]] != MSC_T_STRING) continue;
      const uint32_t n = dicts[c]->n;
      const int phys = n <= 256 ? MSC_P_U8 : (n <= 65536 ? MSC_P_U16 : MSC_P_U32);
      if (phys == MSC_P_U32) continue;
      msc_col narrow;
      narrow.phys = phys;
      rc = msc_alloc_rows(ctx, total, msc_phys_width(phys), &narrow.data, &narrow.bytes);
      if (rc != MSC_OK) return fail(rc);
      const unsigned grid = static_cast<unsigned>(ctx->sm_count * 8);
      if (phys == MSC_P_U8)
        narrow_codes_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(static_cast<uint32_t*>(rel->cols[c].data), static_cast<uint8_t*>(narrow.data), total);
      else
        narrow_codes_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(static_cast<uint32_t*>(rel->cols[c].data), static_cast<uint16_t*>(narrow.data), total);
      ctx->stats.launches += 1;
      msc_free(ctx, rel->cols[c].data, rel->cols[c].bytes);
      rel->cols[c] = narrow;
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(ctx->fail(MSC_ERR_CUDA, "narrow failed"));
  }
  mark("codes narrowed");
  const auto t1 = std::chrono::steady_clock::now();
  ctx->stats.last_ingest_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  ctx->stats.last_ingest_bytes = ld.bytes;
  *out = rel;
  return MSC_OK;
}
