// result.cu -- device relation -> BlockFile on disk.
//
// Replaces WriteToLocalFileTask.write (src/mini_spark/tasks.py:399-410) -> BlockFile.append_data
// (io.py:231-252, one struct.pack per value) and zig BlockFile.appendData
// (zig-src/src/block_file.zig:413-456).  Values are narrowed to the on-disk types on the device
// (INTEGER -> i32 with an overflow check, the reference raises OverflowError at io.py:90;
// FLOAT -> f32, io.py:91-94), strings are gathered from their dictionary into the
// u8-length-prefixed layout (io.py:100-104), then each column payload is copied to the host once
// and written with its u64 size prefix.  The reader is the reference-format
// `BlockFile.read_data_rows` used by `ExecutionEngine.collect_results` (execution.py:47-55).
#include <errno.h>

#include "strings.cuh"

namespace {

__device__ __forceinline__ long long read_int(const void* col, int phys, uint64_t i) {
  switch (phys) {
    case MSC_P_U8: return static_cast<const uint8_t*>(col)[i];
    case MSC_P_U16: return static_cast<const uint16_t*>(col)[i];
    case MSC_P_U32: return static_cast<const uint32_t*>(col)[i];
    case MSC_P_I32: return static_cast<const int*>(col)[i];
    default: return static_cast<const long long*>(col)[i];
  }
}

__global__ void to_i32_kernel(const void* col, int phys, uint64_t lo, uint64_t n, int* out, int* err) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const long long v = read_int(col, phys, lo + i);
    if (v > 2147483647LL || v < -2147483648LL) atomicOr(err, MSC_DEVERR_OVERFLOW);
    out[i] = static_cast<int>(v);
  }
}
__global__ void to_i64_kernel(const void* col, int phys, uint64_t lo, uint64_t n, long long* out) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    out[i] = read_int(col, phys, lo + i);
}
__global__ void to_f32_kernel(const void* col, int phys, uint64_t lo, uint64_t n, float* out) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    out[i] = (phys == MSC_P_F32) ? static_cast<const float*>(col)[lo + i]
                                 : static_cast<float>(static_cast<const double*>(col)[lo + i]);
}
__global__ void str_len_kernel(const void* codes, int phys, uint64_t lo, uint64_t n, const uint32_t* ent_len, uint8_t* lens) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    lens[i] = static_cast<uint8_t>(ent_len[read_int(codes, phys, lo + i)]);
}
__global__ void str_copy_kernel(const void* codes, int phys, uint64_t lo, uint64_t n, const uint64_t* ent_start,
                                const uint32_t* ent_len, const uint8_t* heap, const uint64_t* starts, uint8_t* out) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const long long c = read_int(codes, phys, lo + i);
    const uint8_t* src = heap + ent_start[c];
    uint8_t* dst = out + starts[i];
    for (uint32_t b = 0; b < ent_len[c]; ++b) dst[b] = src[b];
  }
}

struct FileCloser {
  FILE* f;
  ~FileCloser() {
    if (f) fclose(f);
  }
};

}  // namespace

extern "C" int msc_write_blockfile(msc_ctx* ctx, msc_rel* r, const msc_out_col* cols, int32_t ncols, const char* path,
                                   uint32_t rows_per_block) {
  if (!ctx || !r || !cols || ncols < 1 || ncols > 254 || !path || rows_per_block == 0)
    return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  for (int c = 0; c < ncols; ++c) {
    if (cols[c].rel_col < 0 || cols[c].rel_col >= static_cast<int32_t>(r->cols.size())) return ctx->fail(MSC_ERR_ARG, "bad relation column");
    if (cols[c].type == MSC_T_STRING && !cols[c].dict) return ctx->fail(MSC_ERR_ARG, "string column without dictionary");
    if (!cols[c].name || strlen(cols[c].name) >= 255) return ctx->fail(MSC_ERR_ARG, "bad column name");
    const int phys = r->cols[cols[c].rel_col].phys;
    const bool is_f = phys == MSC_P_F32 || phys == MSC_P_F64;
    if ((cols[c].type == MSC_T_FLOAT) != is_f) return ctx->fail(MSC_ERR_ARG, "column type does not match its physical type");
  }
  FileCloser fc{fopen(path, "wb")};
  if (!fc.f) return ctx->fail(MSC_ERR_IO, std::string("cannot create ") + path + ": " + strerror(errno));
  // schema header
  std::vector<uint8_t> hdr;
  hdr.push_back(static_cast<uint8_t>(ncols));
  for (int c = 0; c < ncols; ++c) {
    hdr.push_back(static_cast<uint8_t>(cols[c].type));
    const size_t nl = strlen(cols[c].name);
    hdr.push_back(static_cast<uint8_t>(nl));
    hdr.insert(hdr.end(), cols[c].name, cols[c].name + nl);
  }
  if (fwrite(hdr.data(), 1, hdr.size(), fc.f) != hdr.size()) return ctx->fail(MSC_ERR_IO, "write failed");
  uint64_t file_pos = hdr.size();
  std::vector<uint64_t> block_starts;
  std::vector<uint8_t> host;
  const unsigned grid = static_cast<unsigned>(ctx->sm_count * 8);
  for (uint64_t lo = 0; lo < r->nrows; lo += rows_per_block) {
    const uint64_t n = std::min<uint64_t>(rows_per_block, r->nrows - lo);
    block_starts.push_back(file_pos);
    const uint32_t n32 = static_cast<uint32_t>(n);
    if (fwrite(&n32, 4, 1, fc.f) != 1) return ctx->fail(MSC_ERR_IO, "write failed");
    file_pos += 4;
    for (int c = 0; c < ncols; ++c) {
      const msc_col& col = r->cols[cols[c].rel_col];
      uint64_t payload = 0;
      DevTmp dev(ctx), starts(ctx), body(ctx);
      switch (cols[c].type) {
        case MSC_T_INTEGER:
          payload = n * 4;
          MSC_TRY(dev.alloc(payload));
          to_i32_kernel<<<grid, 256, 0, ctx->stream>>>(col.data, col.phys, lo, n, dev.as<int>(), ctx->d_err);
          break;
        case MSC_T_TIMESTAMP:
          payload = n * 8;
          MSC_TRY(dev.alloc(payload));
          to_i64_kernel<<<grid, 256, 0, ctx->stream>>>(col.data, col.phys, lo, n, dev.as<long long>());
          break;
        case MSC_T_FLOAT:
          payload = n * 4;
          MSC_TRY(dev.alloc(payload));
          to_f32_kernel<<<grid, 256, 0, ctx->stream>>>(col.data, col.phys, lo, n, dev.as<float>());
          break;
        default: {
          msc_dict* d = cols[c].dict;
          MSC_TRY(dev.alloc(n));
          MSC_TRY(starts.alloc((n + 1) * 8));
          str_len_kernel<<<grid, 256, 0, ctx->stream>>>(col.data, col.phys, lo, n, d->ent_len, dev.as<uint8_t>());
          MSC_TRY(msc_exclusive_scan_u8_u64(ctx, dev.as<uint8_t>(), starts.as<uint64_t>(), n));
          uint64_t nbytes = 0;
          MSC_TRY(msc_memcpy_d2h(ctx, &nbytes, starts.as<uint64_t>() + n, 8));
          MSC_TRY(body.alloc(nbytes + 16));
          str_copy_kernel<<<grid, 256, 0, ctx->stream>>>(col.data, col.phys, lo, n, d->ent_start, d->ent_len, d->heap,
                                                        starts.as<uint64_t>(), body.as<uint8_t>());
          ctx->stats.launches += 1;
          payload = n + nbytes;
          host.resize(payload);
          MSC_TRY(msc_memcpy_d2h(ctx, host.data(), dev.p, n));
          MSC_TRY(msc_memcpy_d2h(ctx, host.data() + n, body.p, nbytes));
        } break;
      }
      ctx->stats.launches += 1;
      MSC_CUDA(ctx, cudaGetLastError());
      if (cols[c].type != MSC_T_STRING) {
        host.resize(payload);
        MSC_TRY(msc_memcpy_d2h(ctx, host.data(), dev.p, payload));
      }
      if (fwrite(&payload, 8, 1, fc.f) != 1 || (payload && fwrite(host.data(), 1, payload, fc.f) != payload))
        return ctx->fail(MSC_ERR_IO, "write failed");
      file_pos += 8 + payload;
    }
  }
  MSC_TRY(msc_check_device_error(ctx));
  const uint32_t nb = static_cast<uint32_t>(block_starts.size());
  if ((nb && fwrite(block_starts.data(), 8, nb, fc.f) != nb) || fwrite(&nb, 4, 1, fc.f) != 1) return ctx->fail(MSC_ERR_IO, "write failed");
  if (fflush(fc.f) != 0) return ctx->fail(MSC_ERR_IO, "flush failed");
  return MSC_OK;
}


// ---- single rows of a relation, as raw 64-bit values: what the boundary merge of sorted partial aggregates needs (the one
// group that straddles two ranks) -- reading a row costs one tiny kernel and one host wait, folding one a tiny kernel.
namespace {

struct RowCols {
  const void* data[MSC_VM_MAX_OUT + 1];
  int phys[MSC_VM_MAX_OUT + 1];
  int kind[MSC_VM_MAX_OUT + 1];
  int ncols;
};

__device__ __forceinline__ long long row_value(const void* col, int phys, uint64_t i) {
  switch (phys) {
    case MSC_P_U8: return static_cast<const uint8_t*>(col)[i];
    case MSC_P_U16: return static_cast<const uint16_t*>(col)[i];
    case MSC_P_U32: return static_cast<const uint32_t*>(col)[i];
    case MSC_P_I32: return static_cast<const int*>(col)[i];
    case MSC_P_F32: return __double_as_longlong(static_cast<double>(static_cast<const float*>(col)[i]));
    default: return static_cast<const long long*>(col)[i];
  }
}

__global__ void read_rows_kernel(RowCols c, const uint64_t* rows, int nreq, long long* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nreq * c.ncols) return;
  out[i] = row_value(c.data[i % c.ncols], c.phys[i % c.ncols], rows[i / c.ncols]);
}

// accumulator columns are 64-bit (I64 / F64); kind < 0: leave the column alone (the key)
__global__ void fold_row_kernel(RowCols c, uint64_t row, const long long* values) {
  const int i = threadIdx.x;
  if (i >= c.ncols || c.kind[i] < 0) return;
  long long* cell = static_cast<long long*>(const_cast<void*>(c.data[i])) + row;
  const long long cur = *cell, v = values[i];
  long long r = cur;
  switch (c.kind[i]) {
    case MSC_AGG_SUM_F: r = __double_as_longlong(__longlong_as_double(cur) + __longlong_as_double(v)); break;
    case MSC_AGG_SUM_I: r = cur + v; break;
    case MSC_AGG_MIN_F: r = __longlong_as_double(v) < __longlong_as_double(cur) ? v : cur; break;
    case MSC_AGG_MAX_F: r = __longlong_as_double(v) > __longlong_as_double(cur) ? v : cur; break;
    case MSC_AGG_MIN_I: r = v < cur ? v : cur; break;
    default: r = v > cur ? v : cur; break;
  }
  *cell = r;
}

int row_cols(msc_ctx* ctx, msc_rel* r, RowCols* c) {
  if (r->cols.size() > MSC_VM_MAX_OUT + 1) return ctx->fail(MSC_ERR_ARG, "too many columns");
  c->ncols = static_cast<int>(r->cols.size());
  for (int i = 0; i < c->ncols; ++i) {
    c->data[i] = r->cols[i].data;
    c->phys[i] = r->cols[i].phys;
    c->kind[i] = -1;
  }
  return MSC_OK;
}

}  // namespace

extern "C" int msc_rel_read_rows(msc_ctx* ctx, msc_rel* r, const uint64_t* rows, int32_t nreq, int64_t* out) {
  if (!ctx || !r || !rows || !out || nreq < 1 || nreq > 4) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  for (int i = 0; i < nreq; ++i)
    if (rows[i] >= r->nrows) return ctx->fail(MSC_ERR_ARG, "row out of range");
  RowCols c;
  MSC_TRY(row_cols(ctx, r, &c));
  const int n = nreq * c.ncols;
  DevTmp d_rows(ctx), d_out(ctx);
  MSC_TRY(d_rows.alloc(sizeof(uint64_t) * 4));
  MSC_TRY(d_out.alloc(sizeof(long long) * n));
  unsigned long long* h = ctx->h_scratch;  // pinned, 16 words
  for (int i = 0; i < nreq; ++i) h[i] = rows[i];
  MSC_CUDA(ctx, cudaMemcpyAsync(d_rows.p, h, sizeof(uint64_t) * nreq, cudaMemcpyHostToDevice, ctx->stream));
  read_rows_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(c, d_rows.as<uint64_t>(), nreq, d_out.as<long long>());
  ctx->stats.launches += 1;
  MSC_CUDA(ctx, cudaMemcpyAsync(out, d_out.p, sizeof(long long) * n, cudaMemcpyDeviceToHost, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  MSC_CUDA(ctx, cudaGetLastError());
  return MSC_OK;
}

extern "C" int msc_rel_fold_row(msc_ctx* ctx, msc_rel* r, uint64_t row, const int64_t* values, const int32_t* kinds) {
  if (!ctx || !r || !values || !kinds || row >= r->nrows) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  RowCols c;
  MSC_TRY(row_cols(ctx, r, &c));
  for (int i = 0; i < c.ncols; ++i) {
    c.kind[i] = kinds[i];
    if (kinds[i] >= 0 && r->cols[i].phys != MSC_P_I64 && r->cols[i].phys != MSC_P_F64) return ctx->fail(MSC_ERR_ARG, "accumulator columns are 64-bit");
    if (kinds[i] > MSC_AGG_MAX_I) return ctx->fail(MSC_ERR_ARG, "bad aggregate kind");
  }
  DevTmp d_vals(ctx);
  MSC_TRY(d_vals.alloc(sizeof(long long) * c.ncols));
  MSC_CUDA(ctx, cudaMemcpyAsync(d_vals.p, values, sizeof(long long) * c.ncols, cudaMemcpyHostToDevice, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // (`values` is the caller's pageable memory)
  fold_row_kernel<<<1, 32, 0, ctx->stream>>>(c, row, d_vals.as<long long>());
  ctx->stats.launches += 1;
  MSC_CUDA(ctx, cudaGetLastError());
  return MSC_OK;
}
