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
