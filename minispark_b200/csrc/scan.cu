// scan.cu -- host side of the fused scan: program validation, launch planning, msc_scan_aggregate /
// msc_scan_project, and the small helper kernels (prefix sums, table initialisation / compaction).
// The scan kernel itself lives in scan_kernel.cuh and is instantiated per (rows per lane, mode) in
// scan_inst_*.cu so the specialisations compile in parallel.
#define MSCAN_DECL_ONLY  // the kernel is instantiated in scan_inst_*.cu
#include <algorithm>
#include <chrono>
#include <memory>

#include "jit.h"
#include "scan_kernel.cuh"
#include "scan_regvm.h"

using namespace mscan;

namespace {

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
// the accumulator columns of a streaming aggregate, each filled with its identity, in ONE launch (16-byte stores)
struct FillCols {
  unsigned long long* col[MSC_VM_MAX_AGGS];
  unsigned long long init[MSC_VM_MAX_AGGS];
  int n;
};
__global__ void fill_cols_kernel(const __grid_constant__ FillCols f, uint64_t rows) {
  const uint64_t pairs = rows / 2, stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (int a = 0; a < f.n; ++a) {
    const ulonglong2 v = make_ulonglong2(f.init[a], f.init[a]);
    ulonglong2* p = reinterpret_cast<ulonglong2*>(f.col[a]);  // (column allocations are 256-byte aligned)
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < pairs; i += stride) p[i] = v;
    if ((rows & 1) && blockIdx.x == 0 && threadIdx.x == 0) f.col[a][rows - 1] = f.init[a];
  }
}

// after the specialised streaming aggregate: carry[a][t] = what tile t added to the run that was open when it began (run number
// tile_offsets[t] - 1); several tiles may carry into one long run, so the additions are atomic
struct CarryFold {
  unsigned long long* col[MSC_VM_MAX_AGGS];
  unsigned long long init[MSC_VM_MAX_AGGS];
  int kind[MSC_VM_MAX_AGGS];
  int n;
};
__device__ __forceinline__ void carry_atomic_fold(int kind, unsigned long long* addr, unsigned long long v) {
  switch (kind) {
    case MSC_AGG_SUM_F: atomicAdd(reinterpret_cast<double*>(addr), __longlong_as_double(static_cast<long long>(v))); break;
    case MSC_AGG_SUM_I: atomicAdd(addr, v); break;
    case MSC_AGG_MIN_I: atomicMin(reinterpret_cast<long long*>(addr), static_cast<long long>(v)); break;
    case MSC_AGG_MAX_I: atomicMax(reinterpret_cast<long long*>(addr), static_cast<long long>(v)); break;
    default: {  // f64 min / max: CAS loop
      unsigned long long old = *addr;
      while (true) {
        const double a = __longlong_as_double(static_cast<long long>(old)), b = __longlong_as_double(static_cast<long long>(v));
        const double m = kind == MSC_AGG_MIN_F ? (b < a ? b : a) : (b > a ? b : a);
        const unsigned long long merged = static_cast<unsigned long long>(__double_as_longlong(m));
        if (merged == old) break;
        const unsigned long long prev = atomicCAS(addr, old, merged);
        if (prev == old) break;
        old = prev;
      }
    }
  }
}
__global__ void carry_fold_kernel(const unsigned long long* carry, uint64_t cstride, const uint64_t* tile_offsets, uint64_t ntiles,
                                  const __grid_constant__ CarryFold f) {
  for (uint64_t t = 1 + static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < ntiles; t += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t before = tile_offsets[t];
    if (before == 0) continue;
    for (int a = 0; a < f.n; ++a) {
      const unsigned long long c = carry[static_cast<uint64_t>(a) * cstride + t];
      if (c != f.init[a]) carry_atomic_fold(f.kind[a], f.col[a] + (before - 1), c);
    }
  }
}

// identities, kinds and output columns of a dense aggregate travel as kernel parameters (no staging copies)
struct DenseMeta {
  long long init[MSC_VM_MAX_AGGS + 1];
  int kinds[MSC_VM_MAX_AGGS + 1];
  unsigned long long* out_acc[MSC_VM_MAX_AGGS];
};

__global__ void dense_init_kernel(unsigned long long* out, int ngroups, int naggs, const __grid_constant__ DenseMeta m) {
  const int n = ngroups * naggs;
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = static_cast<unsigned long long>(m.init[i % naggs]);
}

// compact the dense table: one output row per group whose row counter (slot `count_slot`) is non-zero, carrying
// the first `nexport` accumulators; out_n[0] = groups, out_n[1] = 1 when a SUM_F came out non-finite, out_n[2] =
// the device error word (fetched and cleared)
__global__ void dense_finalize_kernel(const unsigned long long* table, int ngroups, int stride, int nexport, int count_slot,
                                      const __grid_constant__ DenseMeta m, uint32_t* out_key, unsigned long long* out_n, int* err) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long n = 0, nonfinite = 0;
  for (int g = 0; g < ngroups; ++g) {
    for (int a = 0; a < nexport; ++a)
      if (m.kinds[a] == MSC_AGG_SUM_F && !isfinite(__longlong_as_double(static_cast<long long>(table[g * stride + a])))) nonfinite = 1;
    if (table[g * stride + count_slot] == 0) continue;
    out_key[n] = g;
    for (int a = 0; a < nexport; ++a) m.out_acc[a][n] = table[g * stride + a];
    ++n;
  }
  out_n[0] = n;
  out_n[1] = nonfinite;  // a masked regvm variant must not be trusted then (v * 0.0 = NaN leaks across groups)
  out_n[2] = static_cast<unsigned long long>(*err);  // the device error word rides along with the group count
  *err = 0;
}

// {row count, 0, device error word (fetched and cleared)} of a pending relation
__global__ void pending_meta_kernel(const unsigned long long* count, unsigned long long* meta, int* err) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  meta[0] = *count;
  meta[1] = 0;
  meta[2] = static_cast<unsigned long long>(*err);
  *err = 0;
}

// the part of dense_finalize_kernel a caller needs when the table goes on to a cross-rank merge instead
__global__ void dense_check_kernel(const unsigned long long* table, int ngroups, int stride, const __grid_constant__ DenseMeta m,
                                   unsigned long long* out_n, int* err) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long nonfinite = 0;
  for (int g = 0; g < ngroups; ++g)
    for (int a = 0; a < stride; ++a)
      if (m.kinds[a] == MSC_AGG_SUM_F && !isfinite(__longlong_as_double(static_cast<long long>(table[g * stride + a])))) nonfinite = 1;
  out_n[0] = 0;
  out_n[1] = nonfinite;
  out_n[2] = static_cast<unsigned long long>(*err);
  *err = 0;
}

constexpr int HTILE = 1024;  // slots per block in the hash-table compaction
// slots of the hash aggregate are [key, accumulators...] padded to 1 << shift words (scan_kernel.cuh)
__global__ void hash_init_kernel(unsigned long long* tbl, uint64_t cap, uint32_t shift, int naggs, const __grid_constant__ DenseMeta m) {
  const uint64_t words = cap << shift, mask = (1ull << shift) - 1;
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < words; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t w = i & mask;
    tbl[i] = w == 0 ? HASH_EMPTY : (w <= static_cast<uint64_t>(naggs) ? static_cast<unsigned long long>(m.init[w - 1]) : 0ull);
  }
}

__global__ void hash_count_kernel(const unsigned long long* tbl, uint64_t cap, uint32_t shift, uint32_t* tile_counts) {
  __shared__ uint32_t wsum[8];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * HTILE;
  uint32_t c = 0;
  for (int i = threadIdx.x; i < HTILE; i += blockDim.x) c += (base + i < cap && tbl[(base + i) << shift] != HASH_EMPTY);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
    for (int w = 0; w < blockDim.x / 32; ++w) tot += wsum[w];
    tile_counts[blockIdx.x] = tot;
  }
}

__global__ void hash_emit_kernel(const unsigned long long* tbl, uint64_t cap, uint32_t shift, int naggs, const uint64_t* tile_offsets,
                                 long long* out_key, unsigned long long* const* out_acc) {
  // 256 threads, HTILE slots: each thread owns 4 consecutive slots so output order is slot order
  __shared__ uint32_t scratch[8];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * HTILE + threadIdx.x * 4;
  uint32_t c = 0;
  unsigned long long k[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    k[i] = (base + i < cap) ? tbl[(base + i) << shift] : HASH_EMPTY;
    c += (k[i] != HASH_EMPTY);
  }
  // block exclusive scan (256 threads)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = c;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) scratch[warp] = inc;
  __syncthreads();
  uint32_t wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += scratch[w];
  uint64_t pos = tile_offsets[blockIdx.x] + wbase + inc - c;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (k[i] == HASH_EMPTY) continue;
    out_key[pos] = static_cast<long long>(k[i]);
    const unsigned long long* slot = tbl + ((base + i) << shift) + 1;
    for (int a = 0; a < naggs; ++a) out_acc[a][pos] = slot[a];
    ++pos;
  }
}

// Runs of equal adjacent values of a column, counted per 256-row tile (one block per tile), plus the number of
// descents.  The total bounds the column's distinct values (every value starts at least one run; exact for clustered
// keys such as lineitem's l_orderkey) and sizes the hash table of a GROUP BY on that column; with no descent the column
// is sorted, every run IS a group, and the aggregate needs no table at all (MODE_RUNS).
constexpr int RUN_TILE = 256;
// one warp per tile, 8 consecutive rows per lane (two 128-bit loads for 4-byte keys), the row before a lane's first through
// a shuffle: one element is loaded once (the first version loaded every element twice, from one thread per row: 178 us for
// 240 MB)
template <class T>
__global__ void run_heads_kernel(const T* col, uint64_t n, uint64_t ntiles, uint32_t* tile_counts, unsigned long long* descents) {
  const int lane = threadIdx.x & 31;
  const uint64_t warps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
  uint32_t my_desc = 0;
  for (uint64_t tile = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; tile < ntiles; tile += warps) {
    const uint64_t row0 = tile * RUN_TILE + static_cast<uint64_t>(lane) * 8;
    T v[8];
    if (row0 + 8 <= n) {
      if constexpr (sizeof(T) == 4) {
        const uint4 a = reinterpret_cast<const uint4*>(col + row0)[0], b = reinterpret_cast<const uint4*>(col + row0)[1];
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = static_cast<T>(w[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = col[row0 + i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = row0 + i < n ? col[row0 + i] : T(0);
    }
    T prev = __shfl_up_sync(0xffffffffu, v[7], 1);
    if (lane == 0 && row0 > 0) prev = col[row0 - 1];
    uint32_t heads = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool valid = row0 + i < n;
      const T before = i ? v[i ? i - 1 : 0] : prev;
      heads += valid && (row0 + i == 0 || v[i] != before);
      my_desc += valid && row0 + i > 0 && v[i] < before;
    }
    for (int o = 16; o > 0; o >>= 1) heads += __shfl_xor_sync(0xffffffffu, heads, o);
    if (lane == 0) tile_counts[tile] = heads;
  }
  for (int o = 16; o > 0; o >>= 1) my_desc += __shfl_xor_sync(0xffffffffu, my_desc, o);
  if (lane == 0 && my_desc) atomicAdd(descents, static_cast<unsigned long long>(my_desc));
}

// ---- generic exclusive scan: 3 kernels, CHUNK elements per block ---------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

template <class TIn>
__global__ void scan_block_sums_kernel(const TIn* in, uint64_t n, uint64_t* bsum) {
  __shared__ uint64_t wsum[SCAN_THREADS / 32];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SCAN_CHUNK;
  uint64_t c = 0;
  for (int i = threadIdx.x; i < SCAN_CHUNK; i += SCAN_THREADS)
    if (base + i < n) c += in[base + i];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t tot = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) tot += wsum[w];
    bsum[blockIdx.x] = tot;
  }
}

// single block: in-place exclusive scan of bsum[nb]; total -> *total_out
__global__ void scan_spine_kernel(uint64_t* bsum, uint64_t nb, uint64_t* total_out) {
  __shared__ uint64_t wsum[32];
  __shared__ uint64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint64_t base = 0; base < nb; base += blockDim.x) {
    const uint64_t i = base + threadIdx.x;
    const uint64_t v = (i < nb) ? bsum[i] : 0;
    uint64_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t nn = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += nn;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint64_t wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += wsum[w];
    const uint64_t carry = carry_s;
    if (i < nb) bsum[i] = carry + wbase + inc - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = carry + wbase + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry_s;
}

template <class TIn>
__global__ void scan_apply_kernel(const TIn* in, uint64_t n, const uint64_t* bsum, uint64_t* out) {
  __shared__ uint64_t wsum[SCAN_THREADS / 32];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * SCAN_CHUNK + static_cast<uint64_t>(threadIdx.x) * SCAN_ITEMS;
  uint64_t v[SCAN_ITEMS];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < n) ? static_cast<uint64_t>(in[base + i]) : 0;
    c += v[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t inc = c;
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t nn = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += nn;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint64_t wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += wsum[w];
  uint64_t pos = bsum[blockIdx.x] + wbase + inc - c;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = pos;
    pos += v[i];
  }
}

template <class TIn>
int exclusive_scan_impl(msc_ctx* ctx, const TIn* in, uint64_t* out, uint64_t n) {
  if (n == 0) {
    MSC_CUDA(ctx, cudaMemsetAsync(out, 0, sizeof(uint64_t), ctx->stream));
    return MSC_OK;
  }
  const uint64_t nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  DevTmp bsum(ctx);
  MSC_TRY(bsum.alloc(nb * sizeof(uint64_t)));
  scan_block_sums_kernel<TIn><<<static_cast<unsigned>(nb), SCAN_THREADS, 0, ctx->stream>>>(in, n, bsum.as<uint64_t>());
  scan_spine_kernel<<<1, 1024, 0, ctx->stream>>>(bsum.as<uint64_t>(), nb, out + n);
  scan_apply_kernel<TIn><<<static_cast<unsigned>(nb), SCAN_THREADS, 0, ctx->stream>>>(in, n, bsum.as<uint64_t>(), out);
  ctx->stats.launches += 3;
  MSC_CUDA(ctx, cudaGetLastError());
  return MSC_OK;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int validate_operand(msc_ctx* ctx, const msc_scan_desc* sd, uint32_t operand, bool allow_lut) {
  const int kind = (operand >> 12) & 7, idx = operand & 0xfff;
  switch (kind) {
    case MSC_SRC_NONE: return MSC_OK;
    case MSC_SRC_TEMP: return idx < sd->ntemps ? MSC_OK : ctx->fail(MSC_ERR_ARG, "operand: bad temporary");
    case MSC_SRC_STAGED: return idx < sd->nstaged ? MSC_OK : ctx->fail(MSC_ERR_ARG, "operand: bad staged column");
    case MSC_SRC_CONST: return idx < sd->nconsts ? MSC_OK : ctx->fail(MSC_ERR_ARG, "operand: bad constant");
    case MSC_SRC_GATHER:
      if ((idx & 63) >= sd->ngather || (idx >> 6) >= sd->nstaged) return ctx->fail(MSC_ERR_ARG, "operand: bad gather column");
      if (sd->staged[idx >> 6].phys != MSC_P_U32) return ctx->fail(MSC_ERR_ARG, "operand: index vector must be U32");
      return MSC_OK;
    case MSC_SRC_GATHER_T:
      if ((idx & 63) >= sd->ngather || (idx >> 6) >= sd->ntemps) return ctx->fail(MSC_ERR_ARG, "operand: bad gather column / index temporary");
      return MSC_OK;
    case MSC_SRC_LUT:
      if (!allow_lut) return ctx->fail(MSC_ERR_ARG, "operand: LUT reference outside a LUT instruction");
      return (idx & ~MSC_PROBE_COMPACT) < sd->nluts ? MSC_OK : ctx->fail(MSC_ERR_ARG, "operand: bad LUT");
    default: return ctx->fail(MSC_ERR_ARG, "operand: unknown kind");
  }
}

// physical type a fast source kind expects of a STAGED operand (-1: not a staged kind)
int fk_phys(int fk) {
  switch (fk) {
    case MSC_FK_F32: return MSC_P_F32;
    case MSC_FK_F64: return MSC_P_F64;
    case MSC_FK_I32F: case MSC_FK_I32: return MSC_P_I32;
    case MSC_FK_I64: return MSC_P_I64;
    case MSC_FK_U8: return MSC_P_U8;
    case MSC_FK_U16: return MSC_P_U16;
    case MSC_FK_U32: return MSC_P_U32;
    default: return -1;
  }
}

// A fast id promises operand kinds / physical types; check the promise against the bound columns
// (a wrong id would make the kernel reinterpret bytes).  fk < 0: operand unused by the shape.
int check_fast_operand(msc_ctx* ctx, const msc_scan_desc* sd, uint32_t operand, int fk) {
  const int kind = (operand >> 12) & 15, idx = operand & 0xfff;
  if (fk == MSC_FK_TEMP) return kind == MSC_SRC_TEMP ? MSC_OK : ctx->fail(MSC_ERR_ARG, "fast shape: operand is not a temporary");
  if (fk == MSC_FK_CONST) return kind == MSC_SRC_CONST ? MSC_OK : ctx->fail(MSC_ERR_ARG, "fast shape: operand is not a constant");
  const int want_kind = (fk == MSC_FK_I32F) ? (MSC_SRC_STAGED | MSC_SRC_I2F) : MSC_SRC_STAGED;
  if (kind != want_kind || idx >= sd->nstaged) return ctx->fail(MSC_ERR_ARG, "fast shape: operand is not a staged column");
  const int phys = sd->staged[idx].phys;
  const bool ok = phys == fk_phys(fk) || (fk == MSC_FK_I64 && phys == MSC_P_I64) || (fk == MSC_FK_F64 && phys == MSC_P_F64);
  return ok ? MSC_OK : ctx->fail(MSC_ERR_ARG, "fast shape: staged column has a different physical type");
}

int validate_fast(msc_ctx* ctx, const msc_scan_desc* sd, int fast, uint32_t w0, uint32_t w1, const int32_t* agg_kinds,
                  const int32_t* out_phys) {
  const int op = w0 & 0x3f, dkind = (w0 >> 6) & 7, tee = (w0 >> 9) & 0xf, dst = (w0 >> 13) & 0x7f;
  const uint32_t a = w1 & 0xffffu, b = w1 >> 16;
  if (fast >= MSC_FAST_ARITH && fast < MSC_FAST_AGGMOV) {
    const int id = fast - MSC_FAST_ARITH, dk = id % 3, bk = (id / 3) % 5, ak = (id / 15) % 5, opi = id / 75;
    if (op != MSC_OP_ADD_F + opi) return ctx->fail(MSC_ERR_ARG, "fast shape: opcode mismatch");
    MSC_TRY(check_fast_operand(ctx, sd, a, ak));
    MSC_TRY(check_fast_operand(ctx, sd, b, bk));
    if (dk == 0) return (dkind == MSC_DST_TEMP && tee == 0) ? MSC_OK : ctx->fail(MSC_ERR_ARG, "fast shape: destination mismatch");
    if (dkind != MSC_DST_AGG || !agg_kinds || agg_kinds[dst] != MSC_AGG_SUM_F || (dk == 2) != (tee != 0))
      return ctx->fail(MSC_ERR_ARG, "fast shape: destination mismatch");
    return MSC_OK;
  }
  if (fast >= MSC_FAST_AGGMOV && fast < MSC_FAST_CMP) {
    const int id = fast - MSC_FAST_AGGMOV, fk = id % 10, kind = id / 10;
    if (op != MSC_OP_MOV || dkind != MSC_DST_AGG || tee != 0 || !agg_kinds || agg_kinds[dst] != kind)
      return ctx->fail(MSC_ERR_ARG, "fast shape: AGGMOV mismatch");
    return check_fast_operand(ctx, sd, a, fk);
  }
  if (fast >= MSC_FAST_CMP && fast < MSC_FAST_GROUP) {
    const int id = fast - MSC_FAST_CMP, fk = id % 10, cmpi = id / 10;
    const bool is_f = fk == MSC_FK_F32 || fk == MSC_FK_F64 || fk == MSC_FK_I32F;
    if (op != (is_f ? MSC_OP_LT_F : MSC_OP_LT_I) + cmpi || dkind != MSC_DST_FILTER || tee != 0)
      return ctx->fail(MSC_ERR_ARG, "fast shape: CMP mismatch");
    MSC_TRY(check_fast_operand(ctx, sd, a, fk));
    return check_fast_operand(ctx, sd, b, MSC_FK_CONST);
  }
  if (fast >= MSC_FAST_GROUP && fast < MSC_FAST_OUT) {
    if (op != MSC_OP_MOV || dkind != MSC_DST_GROUP || tee != 0) return ctx->fail(MSC_ERR_ARG, "fast shape: GROUP mismatch");
    return check_fast_operand(ctx, sd, a, fast - MSC_FAST_GROUP);
  }
  if (fast >= MSC_FAST_OUT && fast < MSC_FAST_OUT + 20) {
    const int id = fast - MSC_FAST_OUT, u32 = id % 2, fk = id / 2;
    if (op != MSC_OP_MOV || dkind != MSC_DST_OUT || tee != 0 || !out_phys || (out_phys[dst] == MSC_P_U32) != (u32 != 0))
      return ctx->fail(MSC_ERR_ARG, "fast shape: OUT mismatch");
    return check_fast_operand(ctx, sd, a, fk);
  }
  return ctx->fail(MSC_ERR_ARG, "unknown fast shape id");
}

int validate_program(msc_ctx* ctx, const msc_scan_desc* sd, int mode, const int32_t* agg_kinds, int naggs, const int32_t* out_phys,
                     int nout) {
  if (sd->ncode <= 0 || sd->ncode > MSC_VM_MAX_CODE) return ctx->fail(MSC_ERR_ARG, "program length out of range");
  if (sd->nstaged < 0 || sd->nstaged > MSC_VM_MAX_STAGED) return ctx->fail(MSC_ERR_ARG, "too many staged columns");
  if (sd->ngather < 0 || sd->ngather > MSC_VM_MAX_GATHER) return ctx->fail(MSC_ERR_ARG, "too many gather columns");
  if (sd->nconsts < 0 || sd->nconsts > MSC_VM_MAX_CONSTS) return ctx->fail(MSC_ERR_ARG, "too many constants");
  if (sd->nluts < 0 || sd->nluts > MSC_VM_MAX_LUTS) return ctx->fail(MSC_ERR_ARG, "too many LUTs");
  if (sd->ntemps < 0 || sd->ntemps > MSC_VM_MAX_TEMPS) return ctx->fail(MSC_ERR_ARG, "too many temporaries");
  bool ended = false;
  for (int pc = 0; pc < sd->ncode; pc += 2) {
    const uint32_t w0 = sd->code[pc];
    const int op = w0 & 0x3f;
    if (op == MSC_OP_END) {
      ended = true;
      break;
    }
    if (pc + 1 >= sd->ncode) return ctx->fail(MSC_ERR_ARG, "truncated instruction");
    const uint32_t w1 = sd->code[pc + 1];
    if (op >= MSC_OP__COUNT) return ctx->fail(MSC_ERR_ARG, "unknown opcode");
    if (op == MSC_OP_RANK) continue;
    const bool lut = op == MSC_OP_LUT8 || op == MSC_OP_LUT32 || op == MSC_OP_PROBE;
    MSC_TRY(validate_operand(ctx, sd, w1 & 0xffffu, false));
    MSC_TRY(validate_operand(ctx, sd, w1 >> 16, lut));
    if (lut && ((w1 >> 28) & 7) != MSC_SRC_LUT) return ctx->fail(MSC_ERR_ARG, "LUT / PROBE instruction needs a LUT operand");
    const int tee = (w0 >> 9) & 0xf, dkind = (w0 >> 6) & 7, dst = (w0 >> 13) & 0x7f;
    if (tee > sd->ntemps) return ctx->fail(MSC_ERR_ARG, "tee: bad temporary");
    switch (dkind) {
      case MSC_DST_TEMP:
        if (dst >= sd->ntemps) return ctx->fail(MSC_ERR_ARG, "destination: bad temporary");
        break;
      case MSC_DST_FILTER: case MSC_DST_NONE: break;
      case MSC_DST_GROUP:
        if (mode != MODE_DENSE && mode != MODE_HASH && mode != MODE_RUNS && mode != MODE_BUILD) return ctx->fail(MSC_ERR_ARG, "GROUP outside an aggregate scan");
        break;
      case MSC_DST_AGG:
        if (mode != MODE_DENSE && mode != MODE_HASH && mode != MODE_RUNS) return ctx->fail(MSC_ERR_ARG, "AGG outside an aggregate scan");
        if (dst >= naggs) return ctx->fail(MSC_ERR_ARG, "AGG: bad accumulator");
        break;
      case MSC_DST_OUT:
        if (dst >= nout) return ctx->fail(MSC_ERR_ARG, "OUT: bad column");
        break;
      default: return ctx->fail(MSC_ERR_ARG, "unknown destination kind");
    }
    const int fast = w0 >> 20;
    if (fast) MSC_TRY(validate_fast(ctx, sd, fast, w0, w1, agg_kinds, out_phys));
  }
  if (!ended) return ctx->fail(MSC_ERR_ARG, "program has no END");
  return MSC_OK;
}

// Is this fast id one of the handlers scan_kernel.cuh compiles (run_fast: aggmov_valid / cmp_valid / group_valid /
// out_valid)?  An id outside that set falls back to the generic path IN the kernel, which reads operand indices as slot
// numbers -- so such an instruction must keep its slot numbers (no patching below) and lose its fast id.
bool fast_shape_compiled(int fast) {
  auto is_float = [](int fk) { return fk == MSC_FK_F32 || fk == MSC_FK_F64 || fk == MSC_FK_I32F; };
  auto is_int_col = [](int fk) { return fk == MSC_FK_I32 || fk == MSC_FK_I64; };
  auto is_code = [](int fk) { return fk == MSC_FK_U8 || fk == MSC_FK_U16 || fk == MSC_FK_U32; };
  if (fast >= MSC_FAST_ARITH && fast < MSC_FAST_AGGMOV) return true;
  if (fast >= MSC_FAST_AGGMOV && fast < MSC_FAST_CMP) {
    const int id = fast - MSC_FAST_AGGMOV, fk = id % 10, kind = id / 10;
    const bool fkind = kind == MSC_AGG_SUM_F || kind == MSC_AGG_MIN_F || kind == MSC_AGG_MAX_F;
    if (fk == MSC_FK_TEMP) return true;
    if (fkind) return is_float(fk);
    return is_int_col(fk) || (fk == MSC_FK_CONST && kind == MSC_AGG_SUM_I);
  }
  if (fast >= MSC_FAST_CMP && fast < MSC_FAST_GROUP) {
    const int fk = (fast - MSC_FAST_CMP) % 10;
    return fk != MSC_FK_TEMP && fk != MSC_FK_CONST;
  }
  if (fast >= MSC_FAST_GROUP && fast < MSC_FAST_OUT) {
    const int fk = fast - MSC_FAST_GROUP;
    return fk == MSC_FK_TEMP || is_int_col(fk) || is_code(fk);
  }
  if (fast >= MSC_FAST_OUT && fast < MSC_FAST_OUT + 20) {
    const int id = fast - MSC_FAST_OUT, u32 = id % 2, fk = id / 2;
    if (fk == MSC_FK_CONST) return false;
    return u32 ? (is_code(fk) || fk == MSC_FK_TEMP) : !is_code(fk);
  }
  return false;
}

// Build kernel params + launch geometry.  extra_smem = CTA-wide bytes after the warp regions.
int plan_launch(msc_ctx* ctx, const msc_scan_desc* sd, int R, size_t extra_smem, LaunchPlan* lp, int ntemps_override = -1,
                int max_ctas = 4) {
  ScanParams& p = lp->p;
  memset(&p, 0, sizeof(p));
  const uint32_t wt = 32 * R;  // rows per warp tile
  p.nrows = sd->nrows;
  p.ntiles = static_cast<uint32_t>((sd->nrows + wt - 1) / wt);
  p.nstaged = sd->nstaged;
  const int ntemps = ntemps_override >= 0 ? ntemps_override : sd->ntemps;
  p.ntemps = ntemps;
  uint32_t off = 0;
  for (int c = 0; c < sd->nstaged; ++c) {
    const size_t w = msc_phys_width(sd->staged[c].phys);
    if (w == 0 || sd->staged[c].data == nullptr) return ctx->fail(MSC_ERR_ARG, "bad staged column");
    if ((reinterpret_cast<uintptr_t>(sd->staged[c].data) & 15) != 0) return ctx->fail(MSC_ERR_ARG, "staged column not 16B aligned");
    p.staged[c].base = static_cast<const unsigned char*>(sd->staged[c].data);
    p.staged[c].width = static_cast<uint32_t>(w);
    p.staged[c].smem_off = off;
    p.staged[c].phys = sd->staged[c].phys;
    p.staged[c].tile_bytes = static_cast<uint32_t>(w * wt);
    p.tile_tx_bytes += static_cast<uint32_t>(w * wt);
    off += static_cast<uint32_t>(msc_round_up(w * wt, 128));
  }
  p.stage_bytes = off ? off : 128;
  if (p.stage_bytes > 0xFFF0) return ctx->fail(MSC_ERR_ARG, "scan stages too many bytes per warp tile");
  for (int c = 0; c < sd->ngather; ++c) {
    if (msc_phys_width(sd->gather[c].phys) == 0 || sd->gather[c].data == nullptr) return ctx->fail(MSC_ERR_ARG, "bad gather column");
    p.gather[c] = sd->gather[c].data;
    p.gather_phys[c] = sd->gather[c].phys;
  }
  for (int c = 0; c < sd->nluts; ++c) p.luts[c] = sd->luts[c];
  memcpy(p.code, sd->code, sizeof(uint32_t) * sd->ncode);
  if (sd->ncode < MSC_VM_MAX_CODE) p.code[sd->ncode] = MSC_OP_END;
  // fast instructions address staged columns by (offset in the warp stage) / 16 instead of by slot
  for (int pc = 0; pc + 1 < sd->ncode; pc += 2) {
    const uint32_t w0 = p.code[pc];
    if ((w0 & 0x3f) == MSC_OP_END) break;
    if ((w0 >> 20) == 0) continue;
    if (!fast_shape_compiled(static_cast<int>(w0 >> 20))) {  // (e.g. a dictionary code widened into an I64 output column)
      p.code[pc] = w0 & 0xFFFFFu;
      continue;
    }
    uint32_t w1 = p.code[pc + 1];
    for (int side = 0; side < 2; ++side) {
      const uint32_t operand = (w1 >> (16 * side)) & 0xffffu;
      if (((operand >> 12) & 7) != MSC_SRC_STAGED) continue;
      const uint32_t patched = (operand & 0xf000u) | (p.staged[operand & 0xfff].smem_off >> 4);
      w1 = (w1 & ~(0xffffu << (16 * side))) | (patched << (16 * side));
    }
    p.code[pc + 1] = w1;
  }
  memcpy(p.consts, sd->consts, sizeof(int64_t) * sd->nconsts);
  p.err = ctx->d_err;
  const size_t temps_bytes = static_cast<size_t>(ntemps) * wt * sizeof(long long);
  // ring depth: aim at ~3 CTAs per SM (about 72 KB each); MSC_SCAN_STAGES overrides
  static const int forced = getenv("MSC_SCAN_STAGES") ? atoi(getenv("MSC_SCAN_STAGES")) : 0;
  const size_t budget = 72 * 1024;
  const size_t fixed = SMEM_HEADER + extra_smem + NW * temps_bytes;
  uint32_t ns = budget > fixed ? static_cast<uint32_t>((budget - fixed) / (static_cast<size_t>(NW) * p.stage_bytes)) : 0;
  if (ns > 4) ns = 4;
  if (ns < 2) ns = 2;
  if (ntemps_override == 0) {
    // regvm: registers allow `max_ctas` CTAs per SM (128 registers -> 4, the 168 of the 3- and 4-group masked
    // variants -> 3).  A second ring slot hides the HBM latency of a warp's next tile (prof_r1h: 18 % of the stall
    // samples were the wait for a single-slot ring), but resident warps matter more: take it only when it is free.
    auto ctas = [&](uint32_t n) {
      const size_t per_cta = fixed + static_cast<size_t>(NW) * n * p.stage_bytes + 1024;
      return std::min<size_t>(static_cast<size_t>(max_ctas), (227 * 1024) / per_cta);
    };
    ns = ctas(2) >= ctas(1) ? 2 : 1;
  }
  if (forced >= 1 && forced <= MAX_STAGES) ns = static_cast<uint32_t>(forced);
  p.nstages = ns;
  p.warp_bytes = static_cast<uint32_t>(ns * p.stage_bytes + temps_bytes);
  lp->R = R;
  lp->smem = SMEM_HEADER + static_cast<size_t>(NW) * p.warp_bytes + extra_smem;
  if (lp->smem > 227 * 1024) return ctx->fail(MSC_ERR_ARG, "scan needs more shared memory than an SM has");
  lp->grid = 0;
  return MSC_OK;
}

// Validate the regvm encoding against the bound columns / accumulators and resolve its column
// operands to shared-memory offsets.  Returns MSC_OK and fills `out`, or an error.
int build_regvm(msc_ctx* ctx, const msc_scan_desc* sd, const ScanParams& p, const int32_t* agg_kinds, int naggs, RegvmProgram* out,
                bool* generic_only) {
  *generic_only = false;
  if (sd->ncode2 <= 0 || sd->ncode2 > MSC_RV_MAX_CODE) return ctx->fail(MSC_ERR_ARG, "regvm program length out of range");
  memset(out, 0, sizeof(*out));
  int depth = 0;
  bool grouped = false, ended = false;
  for (int pc = 0; pc < sd->ncode2; ++pc) {
    const uint32_t w = sd->code2[pc];
    const uint32_t id = w & 0xff;
    uint32_t a[2] = {(w >> 8) & 0xff, (w >> 16) & 0xff};
    if (id >= MSC_RV__COUNT || (w >> 24) != 0) return ctx->fail(MSC_ERR_ARG, "regvm: unknown handler");
    const msc_rv_info& info = MSC_RV_INFO[id];
    if (id == MSC_RV_END) {
      if (depth != 0) return ctx->fail(MSC_ERR_ARG, "regvm: unbalanced stack at END");
      out->code[pc] = w;
      out->n = pc + 1;
      ended = true;
      break;
    }
    if (info.depth >= 0 && depth != info.depth) return ctx->fail(MSC_ERR_ARG, "regvm: stack depth mismatch");
    depth += info.delta;
    if (depth < 0 || depth > MSC_RV_MAX_DEPTH) return ctx->fail(MSC_ERR_ARG, "regvm: stack depth out of range");
    // GROUP freezes the rows' groups (masked-out rows go to the trash group), so filters must come first
    if ((info.flags & MSC_RV_F_FILTER) && grouped) return ctx->fail(MSC_ERR_ARG, "regvm: filter after GROUP");
    if ((info.flags & MSC_RV_F_GROUP) && grouped) return ctx->fail(MSC_ERR_ARG, "regvm: more than one GROUP");
    if (info.flags & MSC_RV_F_GENERIC_ONLY) *generic_only = true;
    const signed char kinds[2] = {info.a1, info.a2};
    for (int k = 0; k < 2; ++k) {
      switch (kinds[k]) {
        case MSC_RV_ARG_COL:
          // a spanning handler reads `span` columns at fixed strides: they must be staged back to back
          for (int j = 0; j < info.span; ++j) {
            const int c = static_cast<int>(a[k]) + j;
            if (c >= sd->nstaged || sd->staged[c].phys != info.phys)
              return ctx->fail(MSC_ERR_ARG, "regvm: column operand has the wrong physical type");
            if (p.staged[c].smem_off != p.staged[a[k]].smem_off + j * p.staged[a[k]].tile_bytes)
              return ctx->fail(MSC_ERR_ARG, "regvm: spanned columns are not adjacent in the stage");
          }
          a[k] = p.staged[a[k]].smem_off / MSC_RV_COL_UNIT;
          if (a[k] > 0xff) return ctx->fail(MSC_ERR_ARG, "regvm: stage too large");
          break;
        case MSC_RV_ARG_CONST:
          if (static_cast<int>(a[k]) >= sd->nconsts) return ctx->fail(MSC_ERR_ARG, "regvm: bad constant");
          break;
        case MSC_RV_ARG_SLOT:
          for (int j = 0; j < info.span; ++j)
            if (static_cast<int>(a[k]) + j >= naggs || agg_kinds[a[k] + j] != info.agg)
              return ctx->fail(MSC_ERR_ARG, "regvm: accumulator kind mismatch");
          if (!grouped) return ctx->fail(MSC_ERR_ARG, "regvm: aggregate before GROUP");
          break;
        default: a[k] = 0; break;
      }
    }
    if (info.flags & MSC_RV_F_GROUP) grouped = true;
    out->code[pc] = id | (a[0] << 8) | (a[1] << 16);
  }
  if (!ended) return ctx->fail(MSC_ERR_ARG, "regvm: program has no END");
  if (!grouped) return ctx->fail(MSC_ERR_ARG, "regvm: program has no GROUP");
  return MSC_OK;
}

bool regvm_enabled() {
  static const bool on = !(getenv("MSC_SCAN_REGVM") && atoi(getenv("MSC_SCAN_REGVM")) == 0);
  return on;
}

template <int MODE>
int launch_scan_r(msc_ctx* ctx, LaunchPlan* lp) {
  if constexpr (MODE == MODE_RUNS) return launch_scan<8, MODE>(ctx, lp);  // its per-tile run counts are for 256-row tiles
  if constexpr (MODE == MODE_DENSE || MODE == MODE_HASH) {
    if (lp->R == 8) return launch_scan<8, MODE>(ctx, lp);
  }
  return launch_scan<4, MODE>(ctx, lp);
}

// rows per lane: 8 amortises the per-instruction dispatch over twice the rows (aggregate scans of
// large relations); small relations and project scans use 4.  MSC_SCAN_R=4|8 overrides.
int pick_rows_per_thread(uint64_t nrows) {
  static const int forced = getenv("MSC_SCAN_R") ? atoi(getenv("MSC_SCAN_R")) : 0;
  if (forced == 4 || forced == 8) return forced;
  return nrows >= (1u << 20) ? 8 : 4;
}

msc_rel* new_rel(msc_ctx* ctx, uint64_t nrows) {
  msc_rel* r = new msc_rel();
  r->ctx = ctx;
  r->nrows = nrows;
  return r;
}

int add_col(msc_ctx* ctx, msc_rel* rel, int phys, uint64_t nrows) {
  msc_col c;
  c.phys = phys;
  MSC_TRY(msc_alloc_rows(ctx, nrows, msc_phys_width(phys), &c.data, &c.bytes));
  rel->cols.push_back(c);
  return MSC_OK;
}

// All columns of a result relation.  Small results (a GROUP BY's handful of rows) share one allocation and one
// memset: per-column calls cost a few microseconds each on the stream, which adds up to more than the scan of a
// small table.  The first column owns the block.
int add_cols(msc_ctx* ctx, msc_rel* rel, const int* phys, int ncols, uint64_t nrows) {
  const uint64_t padded = msc_round_up(nrows ? nrows : 1, MSC_ROW_PAD);
  if (padded > MSC_ROW_PAD || ncols == 0) {
    for (int i = 0; i < ncols; ++i) MSC_TRY(add_col(ctx, rel, phys[i], nrows));
    return MSC_OK;
  }
  size_t total = 0;
  std::vector<size_t> off(ncols);
  for (int i = 0; i < ncols; ++i) {
    off[i] = total;
    total += msc_round_up(padded * msc_phys_width(phys[i]) + 256, 256);
  }
  void* base = nullptr;
  MSC_TRY(msc_alloc(ctx, total, &base));
  MSC_CUDA(ctx, cudaMemsetAsync(base, 0, total, ctx->stream));
  for (int i = 0; i < ncols; ++i) {
    msc_col c;
    c.phys = phys[i];
    c.data = static_cast<char*>(base) + off[i];
    c.bytes = i == 0 ? total : 0;
    c.owned = i == 0;
    rel->cols.push_back(c);
  }
  return MSC_OK;
}

}  // namespace

int msc_exclusive_scan_u8_u64(msc_ctx* ctx, const uint8_t* in, uint64_t* out, uint64_t n) {
  return exclusive_scan_impl<uint8_t>(ctx, in, out, n);
}
int msc_exclusive_scan_u32_u64(msc_ctx* ctx, const uint32_t* in, uint64_t* out, uint64_t n) {
  return exclusive_scan_impl<uint32_t>(ctx, in, out, n);
}

// =================================================================================================
// dense aggregation, in pieces: identities / layout, init + scan into a [ngroups][stride] table, merge of several
// ranks' tables, compaction into a relation.  msc_scan_aggregate chains scan + compaction; the multi-GPU path
// (execution.py) puts an all-gather and a merge between them.
namespace {

struct DensePlan {
  long long init[MSC_VM_MAX_AGGS + 1];
  int kinds[MSC_VM_MAX_AGGS + 1];
  bool use_regvm;
  int stride;      // accumulators per group in the table: naggs, or naggs + 1 with the hidden row counter
  int count_slot;  // slot whose value tells whether the group received rows
};

int agg_identity(msc_ctx* ctx, int kind, long long* out) {
  switch (kind) {
    case MSC_AGG_SUM_F: *out = __builtin_bit_cast(long long, 0.0); return MSC_OK;
    case MSC_AGG_SUM_I: *out = 0; return MSC_OK;
    // MIN/MAX are seeded with the reference's MAX_INT / MIN_INT sentinels (tasks.py:303-310, constants.py:14-15)
    case MSC_AGG_MIN_F: *out = __builtin_bit_cast(long long, 2147483647.0); return MSC_OK;
    case MSC_AGG_MAX_F: *out = __builtin_bit_cast(long long, -2147483648.0); return MSC_OK;
    case MSC_AGG_MIN_I: *out = 2147483647LL; return MSC_OK;
    case MSC_AGG_MAX_I: *out = -2147483648LL; return MSC_OK;
    default: return ctx->fail(MSC_ERR_ARG, "bad aggregate kind");
  }
}

// All dense kernels but the masked regvm variants fold filtered rows into a trash group.  The C++ kernel keeps a
// hidden per-group row counter to tell which groups received rows; the regvm kernels use the query's own COUNT
// accumulator (count_slot2) for that.
int dense_plan(msc_ctx* ctx, const msc_scan_desc* sd, const int32_t* agg_kinds, int naggs, DensePlan* dp) {
  if (naggs < 0 || naggs > MSC_VM_MAX_AGGS) return ctx->fail(MSC_ERR_ARG, "bad arguments");
  for (int a = 0; a < naggs; ++a) {
    dp->kinds[a] = agg_kinds[a];
    MSC_TRY(agg_identity(ctx, agg_kinds[a], &dp->init[a]));
  }
  dp->kinds[naggs] = MSC_AGG_SUM_I;
  dp->init[naggs] = 0;
  dp->use_regvm = sd->ncode2 > 0 && regvm_enabled() && sd->count_slot2 >= 0 && sd->count_slot2 < naggs &&
                  agg_kinds[sd->count_slot2] == MSC_AGG_SUM_I;
  dp->stride = dp->use_regvm ? naggs : naggs + 1;
  dp->count_slot = dp->use_regvm ? sd->count_slot2 : naggs;
  return MSC_OK;
}

DenseMeta dense_meta(const DensePlan& dp) {
  DenseMeta meta;
  memset(&meta, 0, sizeof(meta));
  memcpy(meta.init, dp.init, sizeof(long long) * dp.stride);
  memcpy(meta.kinds, dp.kinds, sizeof(int) * dp.stride);
  return meta;
}

// init + scan into `table` ([ngroups][stride], device).  allow_masked: a masked regvm variant may run (the caller
// must then look at the sums: non-finite ones mean "run again with allow_masked = false", gen_regvm.py).
int dense_scan_into(msc_ctx* ctx, const msc_scan_desc* sd, int ngroups, const int32_t* agg_kinds, int naggs, const DensePlan& dp,
                    bool allow_masked, unsigned long long* table, bool* was_masked, bool want_jit = false) {
  static const bool masked_enabled = !(getenv("MSC_SCAN_MASKED") && atoi(getenv("MSC_SCAN_MASKED")) == 0);
  // MSC_SCAN_JIT: 0 = never specialise, 1 (default) = when asked (MSC_DENSE_JIT) or already compiled, 2 = always
  static const int jit_mode = getenv("MSC_SCAN_JIT") ? atoi(getenv("MSC_SCAN_JIT")) : 1;
  const int ntot = dp.stride;
  if (jit_mode > 0 && jit_dense_supported(sd, ngroups, ntot) &&
      (want_jit || sd->want_jit != 0 || jit_mode > 1 || jit_dense_cached(ctx, sd, ngroups, naggs, ntot, dp.kinds, dp.init, allow_masked && masked_enabled))) {
    *was_masked = allow_masked && masked_enabled;
    dense_init_kernel<<<1, 256, 0, ctx->stream>>>(table, ngroups, ntot, dense_meta(dp));
    ctx->stats.launches += 1;
    const int jrc = jit_dense_launch(ctx, sd, ngroups, naggs, ntot, dp.kinds, dp.init, table, true, was_masked);
    // No compiler on this machine, or a program the generator declines: the interpreters below run the scan instead
    // (still on the GPU -- there is no CPU path); anything else is a real error.
    if (!(jrc == MSC_ERR_ARG && ctx->err.rfind("jit:", 0) == 0) && !(jrc == MSC_ERR_CUDA && ctx->err.rfind("jit: lib", 0) == 0)) return jrc;
    static bool warned = false;
    if (!warned) {
      fprintf(stderr, "minispark_cuda: scan not specialised (%s); using the interpreter kernels\n", ctx->err.c_str());
      warned = true;
    }
    *was_masked = false;
  }
  LaunchPlan lp;
  RegvmProgram rv;
  int variant = 0;
  if (dp.use_regvm) {
    // validate against a provisional plan first: whether the masked variants apply depends on the program
    bool generic_only = false;
    MSC_TRY(plan_launch(ctx, sd, MSC_RV_ROWS, 0, &lp, 0));
    MSC_TRY(build_regvm(ctx, sd, lp.p, agg_kinds, naggs, &rv, &generic_only));
    if (allow_masked && masked_enabled && !generic_only && ngroups <= MSC_RV_MAX_NG) variant = ngroups;
  }
  *was_masked = variant > 0;
  const int smem_groups = variant > 0 ? ngroups : ngroups + 1;  // + the trash group
  const size_t acc_bytes = static_cast<size_t>(smem_groups) * ntot * NT * sizeof(long long);
  if (acc_bytes > 96 * 1024) return ctx->fail(MSC_ERR_ARG, "dense aggregate: groups x aggregates too large; use hash mode");
  const size_t regvm_bytes = (MSC_RV_MAX_CODE + 2) * sizeof(uint32_t) + MSC_VM_MAX_CONSTS * sizeof(long long);
  if (dp.use_regvm) MSC_TRY(plan_launch(ctx, sd, MSC_RV_ROWS, acc_bytes + regvm_bytes, &lp, 0, variant >= 3 ? 3 : 4));
  else MSC_TRY(plan_launch(ctx, sd, pick_rows_per_thread(sd->nrows), acc_bytes, &lp));
  lp.p.ngroups = ngroups;
  lp.p.naggs = ntot;
  memcpy(lp.p.agg_init, dp.init, sizeof(long long) * ntot);
  memcpy(lp.p.agg_kind, dp.kinds, sizeof(int) * ntot);
  dense_init_kernel<<<1, 256, 0, ctx->stream>>>(table, ngroups, ntot, dense_meta(dp));
  ctx->stats.launches += 1;
  lp.p.dense_out = table;
  if (sd->nrows == 0) return MSC_OK;
  if (!dp.use_regvm) return launch_scan_r<MODE_DENSE>(ctx, &lp);
  switch (variant) {
    case 1: return launch_regvm_dense_ng1(ctx, &lp, &rv);
    case 2: return launch_regvm_dense_ng2(ctx, &lp, &rv);
    case 3: return launch_regvm_dense_ng3(ctx, &lp, &rv);
    case 4: return launch_regvm_dense_ng4(ctx, &lp, &rv);
    default: return launch_regvm_dense_ng0(ctx, &lp, &rv);
  }
}

// compact `table` into a relation: group id (U32) + the first naggs accumulators of every group that received rows.
// Synchronous form: one host wait; reports a non-finite SUM_F and the device error word along with the group count.
// async: nothing is waited for -- the relation comes back pending (its counts stay in rel->d_meta, msc_rel_settle).
int dense_compact(msc_ctx* ctx, const unsigned long long* table, int ngroups, int naggs, const DensePlan& dp, msc_rel** out,
                  bool* nonfinite, bool async = false) {
  msc_rel* rel = new_rel(ctx, async ? ngroups : 0);
  int physes[MSC_VM_MAX_AGGS + 1];
  physes[0] = MSC_P_U32;
  for (int a = 0; a < naggs; ++a)
    physes[1 + a] = (dp.kinds[a] == MSC_AGG_SUM_F || dp.kinds[a] == MSC_AGG_MIN_F || dp.kinds[a] == MSC_AGG_MAX_F) ? MSC_P_F64 : MSC_P_I64;
  int rc = add_cols(ctx, rel, physes, 1 + naggs, ngroups);
  if (rc == MSC_OK) rc = msc_alloc(ctx, 3 * sizeof(unsigned long long), reinterpret_cast<void**>(&rel->d_meta));
  if (rc != MSC_OK) {
    msc_rel_free(rel);
    return rc;
  }
  DenseMeta meta = dense_meta(dp);
  for (int a = 0; a < naggs; ++a) meta.out_acc[a] = static_cast<unsigned long long*>(rel->cols[1 + a].data);
  dense_finalize_kernel<<<1, 32, 0, ctx->stream>>>(table, ngroups, dp.stride, naggs, dp.count_slot, meta,
                                                  static_cast<uint32_t*>(rel->cols[0].data), rel->d_meta, ctx->d_err);
  ctx->stats.launches += 1;
  if (async) {
    rel->pending = true;
    *out = rel;
    return MSC_OK;
  }
  cudaEventRecord(ctx->ev_b, ctx->stream);
  unsigned long long* n = ctx->h_scratch;  // pinned: a plain stack buffer would make the copy synchronous twice over
  if (cudaMemcpyAsync(n, rel->d_meta, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    msc_rel_free(rel);
    return ctx->fail(MSC_ERR_CUDA, "dense aggregate failed");
  }
  const int drc = msc_device_error_rc(ctx, static_cast<int>(n[2]));
  if (drc != MSC_OK) {
    msc_rel_free(rel);
    return drc;
  }
  rel->nrows = n[0];
  *nonfinite = n[1] != 0;
  *out = rel;
  return MSC_OK;
}

void note_times(msc_ctx* ctx, bool scanned) {
  float ms = 0;
  if (cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b) == cudaSuccess) ctx->stats.last_kernel_ms = ms;
  if (scanned && cudaEventElapsedTime(&ms, ctx->ev_s0, ctx->ev_s1) == cudaSuccess) ctx->stats.last_scan_ms = ms;
}

__device__ __forceinline__ long long agg_combine_dev(int kind, long long cur, long long v) {  // = mscan::agg_combine
  switch (kind) {
    case MSC_AGG_SUM_F: return __double_as_longlong(__longlong_as_double(cur) + __longlong_as_double(v));
    case MSC_AGG_SUM_I: return cur + v;
    case MSC_AGG_MIN_F: return (__longlong_as_double(v) < __longlong_as_double(cur)) ? v : cur;
    case MSC_AGG_MAX_F: return (__longlong_as_double(v) > __longlong_as_double(cur)) ? v : cur;
    case MSC_AGG_MIN_I: return (v < cur) ? v : cur;
    default: return (v > cur) ? v : cur;  // MSC_AGG_MAX_I
  }
}

// fold `world` tables ([gmax][stride] each, rank-major) into out[ngroups_out][stride]; perm[r * gmax + g] is the
// output group of rank r's group g (< 0: rank r has no such group).  Rank order, so every rank computes the same bits.
__global__ void dense_merge_kernel(const unsigned long long* tables, int world, int gmax, int stride, const int* perm, int ngroups_out,
                                   const __grid_constant__ DenseMeta m, unsigned long long* out) {
  const int cells = ngroups_out * stride;
  for (int i = threadIdx.x; i < cells; i += blockDim.x) out[i] = static_cast<unsigned long long>(m.init[i % stride]);
  __syncthreads();
  for (int a = threadIdx.x; a < stride; a += blockDim.x) {  // one thread per accumulator column: no races between ranks
    for (int r = 0; r < world; ++r)
      for (int g = 0; g < gmax; ++g) {
        const int dst = perm[r * gmax + g];
        if (dst < 0 || dst >= ngroups_out) continue;
        unsigned long long* cell = out + dst * stride + a;
        *cell = static_cast<unsigned long long>(
            agg_combine_dev(m.kinds[a], static_cast<long long>(*cell), static_cast<long long>(tables[(static_cast<size_t>(r) * gmax + g) * stride + a])));
      }
  }
}

}  // namespace

extern "C" int msc_dense_layout(msc_ctx* ctx, const msc_scan_desc* sd, const int32_t* agg_kinds, int32_t naggs, int32_t* stride,
                                int32_t* count_slot) {
  if (!ctx || !sd || !stride || !count_slot) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  DensePlan dp;
  MSC_TRY(dense_plan(ctx, sd, agg_kinds, naggs, &dp));
  *stride = dp.stride;
  *count_slot = dp.count_slot;
  return MSC_OK;
}

extern "C" int msc_scan_dense_table(msc_ctx* ctx, const msc_scan_desc* sd, int32_t ngroups, const int32_t* agg_kinds, int32_t naggs,
                                    void* table, int32_t flags) {
  if (!ctx || !sd || !table || ngroups <= 0) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  MSC_TRY(validate_program(ctx, sd, MODE_DENSE, agg_kinds, naggs, nullptr, 0));
  DensePlan dp;
  MSC_TRY(dense_plan(ctx, sd, agg_kinds, naggs, &dp));
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  const bool exact_only = (flags & MSC_DENSE_EXACT) != 0, want_jit = (flags & MSC_DENSE_JIT) != 0;
  if (flags & MSC_DENSE_ASYNC) {  // enqueue only: the caller checks the sums after its merge (msc_dense_merge_compact)
    bool masked = false;
    return dense_scan_into(ctx, sd, ngroups, agg_kinds, naggs, dp, !exact_only, static_cast<unsigned long long*>(table), &masked, want_jit);
  }
  DevTmp d_n(ctx);
  MSC_TRY(d_n.alloc(3 * sizeof(unsigned long long)));
  for (int attempt = exact_only ? 1 : 0; attempt < 2; ++attempt) {
    bool masked = false;
    MSC_TRY(dense_scan_into(ctx, sd, ngroups, agg_kinds, naggs, dp, attempt == 0, static_cast<unsigned long long*>(table), &masked, want_jit));
    dense_check_kernel<<<1, 32, 0, ctx->stream>>>(static_cast<const unsigned long long*>(table), ngroups, dp.stride, dense_meta(dp),
                                                 d_n.as<unsigned long long>(), ctx->d_err);
    ctx->stats.launches += 1;
    MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
    unsigned long long* n = ctx->h_scratch;
    MSC_CUDA(ctx, cudaMemcpyAsync(n, d_n.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MSC_TRY(msc_device_error_rc(ctx, static_cast<int>(n[2])));
    if (masked && n[1] != 0) continue;  // non-finite SUM out of a masked variant: redo it the exact way
    note_times(ctx, sd->nrows > 0);
    return MSC_OK;
  }
  return ctx->fail(MSC_ERR_ARG, "dense aggregate: unreachable");
}

namespace {
int plan_for_table(msc_ctx* ctx, const int32_t* agg_kinds, int naggs, int stride, int count_slot, DensePlan* dp) {
  if (naggs < 0 || naggs > MSC_VM_MAX_AGGS || (stride != naggs && stride != naggs + 1) || count_slot < 0 || count_slot >= stride)
    return ctx->fail(MSC_ERR_ARG, "bad dense table layout");
  for (int a = 0; a < naggs; ++a) {
    dp->kinds[a] = agg_kinds[a];
    MSC_TRY(agg_identity(ctx, agg_kinds[a], &dp->init[a]));
  }
  dp->kinds[naggs] = MSC_AGG_SUM_I;
  dp->init[naggs] = 0;
  dp->stride = stride;
  dp->count_slot = count_slot;
  dp->use_regvm = false;
  return MSC_OK;
}
}  // namespace

extern "C" int msc_dense_merge(msc_ctx* ctx, const void* tables, int32_t world, int32_t gmax, int32_t stride, const int32_t* agg_kinds,
                               int32_t naggs, const int32_t* perm, int32_t ngroups_out, void* out_table) {
  if (!ctx || !tables || !perm || !out_table || world <= 0 || gmax <= 0 || ngroups_out <= 0)
    return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  DensePlan dp;
  MSC_TRY(plan_for_table(ctx, agg_kinds, naggs, stride, 0, &dp));
  DevTmp d_perm(ctx);
  MSC_TRY(d_perm.alloc(sizeof(int32_t) * world * gmax));
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  MSC_CUDA(ctx, cudaMemcpyAsync(d_perm.p, perm, sizeof(int32_t) * world * gmax, cudaMemcpyHostToDevice, ctx->stream));
  dense_merge_kernel<<<1, 64, 0, ctx->stream>>>(static_cast<const unsigned long long*>(tables), world, gmax, stride, d_perm.as<int>(),
                                               ngroups_out, dense_meta(dp), static_cast<unsigned long long*>(out_table));
  ctx->stats.launches += 1;
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  MSC_CUDA(ctx, cudaGetLastError());
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `perm` is the caller's
  note_times(ctx, false);
  return MSC_OK;
}

extern "C" int msc_dense_merge_compact(msc_ctx* ctx, const void* tables, int32_t world, int32_t gmax, int32_t stride,
                                       const int32_t* agg_kinds, int32_t naggs, const int32_t* perm_dev, int32_t ngroups_out,
                                       int32_t count_slot, void* scratch_table, msc_rel** out, int32_t* nonfinite) {
  if (!ctx || !tables || !perm_dev || !scratch_table || !out || !nonfinite || world <= 0 || gmax <= 0 || ngroups_out <= 0)
    return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  DensePlan dp;
  MSC_TRY(plan_for_table(ctx, agg_kinds, naggs, stride, count_slot, &dp));
  // no event here: ev_a still marks the start of the msc_scan_dense_table(ASYNC) this call completes
  dense_merge_kernel<<<1, 64, 0, ctx->stream>>>(static_cast<const unsigned long long*>(tables), world, gmax, stride, perm_dev, ngroups_out,
                                               dense_meta(dp), static_cast<unsigned long long*>(scratch_table));
  ctx->stats.launches += 1;
  MSC_CUDA(ctx, cudaGetLastError());
  bool nf = false;
  MSC_TRY(dense_compact(ctx, static_cast<const unsigned long long*>(scratch_table), ngroups_out, naggs, dp, out, &nf));
  *nonfinite = nf ? 1 : 0;
  note_times(ctx, true);
  return MSC_OK;
}

extern "C" int msc_dense_merge_compact_async(msc_ctx* ctx, const void* tables, int32_t world, int32_t gmax, int32_t stride,
                                             const int32_t* agg_kinds, int32_t naggs, const int32_t* perm_dev, int32_t ngroups_out,
                                             int32_t count_slot, void* scratch_table, msc_rel** out) {
  if (!ctx || !tables || !perm_dev || !scratch_table || !out || world <= 0 || gmax <= 0 || ngroups_out <= 0)
    return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  DensePlan dp;
  MSC_TRY(plan_for_table(ctx, agg_kinds, naggs, stride, count_slot, &dp));
  dense_merge_kernel<<<1, 64, 0, ctx->stream>>>(static_cast<const unsigned long long*>(tables), world, gmax, stride, perm_dev, ngroups_out,
                                               dense_meta(dp), static_cast<unsigned long long*>(scratch_table));
  ctx->stats.launches += 1;
  MSC_CUDA(ctx, cudaGetLastError());
  bool nf = false;
  return dense_compact(ctx, static_cast<const unsigned long long*>(scratch_table), ngroups_out, naggs, dp, out, &nf, true);
}

extern "C" int msc_dense_compact_async(msc_ctx* ctx, const void* table, int32_t ngroups, int32_t stride, const int32_t* agg_kinds,
                                       int32_t naggs, int32_t count_slot, msc_rel** out) {
  if (!ctx || !table || !out || ngroups <= 0) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  DensePlan dp;
  MSC_TRY(plan_for_table(ctx, agg_kinds, naggs, stride, count_slot, &dp));
  bool nf = false;
  return dense_compact(ctx, static_cast<const unsigned long long*>(table), ngroups, naggs, dp, out, &nf, true);
}

// One pass of a prepared dense aggregate in ONE call: scan into `table` -> compaction -> final projection bound to the
// compacted columns -> a single host wait.  What execution.py's PreparedAggregate.run() otherwise drives through six
// entry points (each costs a few microseconds of ctypes marshalling on a 0.4 ms step).
extern "C" int msc_dense_chain(msc_ctx* ctx, const msc_scan_desc* scan, int32_t ngroups, const int32_t* agg_kinds, int32_t naggs, void* table,
                               int32_t stride, int32_t count_slot, int32_t flags, msc_scan_desc* final_scan, const int32_t* final_cols,
                               const int32_t* out_phys, int32_t nout, msc_rel** raw_out, msc_rel** final_out, int32_t* nonfinite) {
  if (!ctx || !scan || !table || !final_scan || !final_cols || !raw_out || !final_out || !nonfinite)
    return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  *raw_out = nullptr;
  *final_out = nullptr;
  MSC_TRY(msc_scan_dense_table(ctx, scan, ngroups, agg_kinds, naggs, table, flags | MSC_DENSE_ASYNC));
  msc_rel* raw = nullptr;
  MSC_TRY(msc_dense_compact_async(ctx, table, ngroups, stride, agg_kinds, naggs, count_slot, &raw));
  for (int sl = 0; sl < final_scan->nstaged; ++sl) {
    const int c = final_cols[sl];
    if (c < 0 || c >= static_cast<int>(raw->cols.size())) {
      msc_rel_free(raw);
      return ctx->fail(MSC_ERR_ARG, "final projection refers to a column the aggregate does not have");
    }
    final_scan->staged[sl].data = raw->cols[c].data;
  }
  final_scan->nrows = static_cast<uint64_t>(ngroups);
  final_scan->nrows_dev = reinterpret_cast<const uint64_t*>(raw->d_meta);
  msc_rel* fin = nullptr;
  int rc = msc_scan_project(ctx, final_scan, out_phys, nout, &fin);
  if (rc != MSC_OK) {
    msc_rel_free(raw);
    return rc;
  }
  msc_rel* rels[2] = {raw, fin};
  rc = msc_rel_settle(ctx, rels, 2, nonfinite);
  if (rc != MSC_OK) {
    msc_rel_free(raw);
    msc_rel_free(fin);
    return rc;
  }
  *raw_out = raw;
  *final_out = fin;
  return MSC_OK;
}

namespace {
int dense_fused_impl(msc_ctx* ctx, const msc_scan_desc* scan, int32_t ngroups, const int32_t* agg_kinds, int32_t naggs, void* table, int32_t flags,
                     const msc_scan_desc* final_scan, const int32_t* final_cols, const int32_t* out_phys, int32_t nout,
                     const msc_peer_spec* peer, msc_rel** final_out, int32_t* nonfinite) {
  if (!ctx || !scan || !table || !final_scan || !final_cols || !final_out || !nonfinite || ngroups <= 0 || nout <= 0 || nout > MSC_VM_MAX_OUT)
    return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  *final_out = nullptr;
  *nonfinite = 0;
  static const int jit_mode = getenv("MSC_SCAN_JIT") ? atoi(getenv("MSC_SCAN_JIT")) : 1;
  static const bool masked_enabled = !(getenv("MSC_SCAN_MASKED") && atoi(getenv("MSC_SCAN_MASKED")) == 0);
  const int out_groups = peer ? peer->nglobal : ngroups;
  if (jit_mode == 0 || !(flags & MSC_DENSE_JIT) || ngroups > 32 || out_groups > 32 || out_groups <= 0) return MSC_OK;
  if (!peer && scan->nrows == 0) return MSC_OK;
  if (peer && (peer->world < 1 || peer->world > MSC_PEER_MAX_WORLD || peer->rank < 0 || peer->rank >= peer->world || peer->nlocal != ngroups ||
               peer->gmax < ngroups || !peer->inv))
    return ctx->fail(MSC_ERR_ARG, "bad peer specification");
  MSC_TRY(validate_program(ctx, scan, MODE_DENSE, agg_kinds, naggs, nullptr, 0));
  MSC_TRY(validate_program(ctx, final_scan, MODE_PROJECT, nullptr, 0, out_phys, nout));
  DensePlan dp;
  MSC_TRY(dense_plan(ctx, scan, agg_kinds, naggs, &dp));
  if (!jit_dense_supported(scan, ngroups, dp.stride)) return MSC_OK;
  if (!ctx->d_ticket) {
    MSC_TRY(msc_alloc(ctx, sizeof(uint32_t), reinterpret_cast<void**>(&ctx->d_ticket)));
    MSC_CUDA(ctx, cudaMemsetAsync(ctx->d_ticket, 0, sizeof(uint32_t), ctx->stream));
  }
  const bool compile_only = peer && peer->compile_only;
  msc_rel* rel = new_rel(ctx, static_cast<uint64_t>(out_groups));
  int rc = add_cols(ctx, rel, out_phys, nout, static_cast<uint64_t>(out_groups));
  if (rc == MSC_OK) rc = msc_alloc(ctx, 3 * sizeof(unsigned long long), reinterpret_cast<void**>(&rel->d_meta));
  if (rc != MSC_OK) {
    msc_rel_free(rel);
    return rc;
  }
  void* outs[MSC_VM_MAX_OUT];
  for (int i = 0; i < nout; ++i) outs[i] = rel->cols[i].data;
  JitFinish fin{final_scan, final_cols, out_phys, nout, dp.count_slot, outs, rel->d_meta, ctx->d_ticket, peer};
  if (!compile_only) {
    MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
    dense_init_kernel<<<1, 256, 0, ctx->stream>>>(static_cast<unsigned long long*>(table), ngroups, dp.stride, dense_meta(dp));
    ctx->stats.launches += 1;
  }
  bool masked = !(flags & MSC_DENSE_EXACT) && masked_enabled;
  rc = jit_dense_launch(ctx, scan, ngroups, naggs, dp.stride, dp.kinds, dp.init, static_cast<unsigned long long*>(table), true, &masked, &fin);
  if (rc != MSC_OK) {
    msc_rel_free(rel);
    // the generator declined, or there is no compiler on this machine: not an error, just not fused
    return (rc == MSC_ERR_ARG && ctx->err.rfind("jit:", 0) == 0) || (rc == MSC_ERR_CUDA && ctx->err.rfind("jit: lib", 0) == 0) ? MSC_OK : rc;
  }
  if (compile_only) {  // the kernel exists now; hand back an empty relation as the "yes"
    rel->nrows = 0;
    *final_out = rel;
    return MSC_OK;
  }
  rel->pending = true;
  msc_rel* rels[1] = {rel};
  rc = msc_rel_settle(ctx, rels, 1, nonfinite);
  if (rc != MSC_OK) {
    msc_rel_free(rel);
    return rc;
  }
  *final_out = rel;
  return MSC_OK;
}
}  // namespace

// The same pass with NO launch after the scan: the specialised kernel's last CTA compacts the groups and runs the final
// projection (jit.cu emit_finish).  *final_out stays null when this query cannot be fused (no specialised kernel for it,
// lookup tables in the final projection, more than 32 groups, empty input): the caller then uses msc_dense_chain.
extern "C" int msc_dense_fused(msc_ctx* ctx, const msc_scan_desc* scan, int32_t ngroups, const int32_t* agg_kinds, int32_t naggs, void* table,
                               int32_t flags, const msc_scan_desc* final_scan, const int32_t* final_cols, const int32_t* out_phys, int32_t nout,
                               msc_rel** final_out, int32_t* nonfinite) {
  return dense_fused_impl(ctx, scan, ngroups, agg_kinds, naggs, table, flags, final_scan, final_cols, out_phys, nout, nullptr, final_out, nonfinite);
}

// ... and across ranks: the last CTA also exchanges the partial tables over NVLink peer memory and merges them
// (include/minispark_cuda.h, "fused scan + cross-GPU merge")
extern "C" int msc_dense_fused_peer(msc_ctx* ctx, const msc_scan_desc* scan, const int32_t* agg_kinds, int32_t naggs, void* table, int32_t flags,
                                    const msc_scan_desc* final_scan, const int32_t* final_cols, const int32_t* out_phys, int32_t nout,
                                    const msc_peer_spec* peer, msc_rel** final_out, int32_t* nonfinite) {
  if (!peer) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  return dense_fused_impl(ctx, scan, peer->nlocal, agg_kinds, naggs, table, flags, final_scan, final_cols, out_phys, nout, peer, final_out, nonfinite);
}

// ---- prepared dense aggregate: everything a pass needs, decided and allocated once --------------------------------
struct msc_prepared {
  msc_ctx* ctx = nullptr;
  msc_scan_desc scan, fin;
  int ngroups = 0, naggs = 0, nout = 0;
  int32_t kinds[MSC_VM_MAX_AGGS + 1], out_phys[MSC_VM_MAX_OUT], fin_cols[MSC_VM_MAX_STAGED];
  DensePlan dp;
  unsigned long long* table = nullptr;  // [ngroups][stride]; holds the identities between passes (the finish resets it)
  size_t table_bytes = 0;
  // A ring of result relations: passes may be enqueued ahead of the host (msc_prepared_enqueue) and collected in order
  // (msc_prepared_wait); pass k writes ring[k % kRing], so a result lives until kRing later passes have been enqueued.
  static constexpr int kRing = 4;
  msc_rel* ring[kRing] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t done[kRing] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t k0[kRing] = {nullptr, nullptr, nullptr, nullptr}, k1[kRing] = {nullptr, nullptr, nullptr, nullptr};  // around each pass's kernel
  unsigned long long* h_meta = nullptr;  // pinned, 3 words per ring entry: {rows, non-finite SUM, device error word}
  uint64_t issued = 0, collected = 0;
  bool table_clean = false;
  bool has_peer = false;
  msc_peer_spec peer;
};

extern "C" int msc_prepared_create(msc_ctx* ctx, const msc_scan_desc* scan, int32_t ngroups, const int32_t* agg_kinds, int32_t naggs,
                                   const msc_scan_desc* final_scan, const int32_t* final_cols, const int32_t* out_phys, int32_t nout,
                                   const msc_peer_spec* peer, msc_prepared** out) {
  if (!ctx || !scan || !agg_kinds || !final_scan || !final_cols || !out_phys || !out || ngroups <= 0 || naggs < 0 || naggs > MSC_VM_MAX_AGGS ||
      nout <= 0 || nout > MSC_VM_MAX_OUT)
    return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  *out = nullptr;
  static const int jit_mode = getenv("MSC_SCAN_JIT") ? atoi(getenv("MSC_SCAN_JIT")) : 1;
  const int out_groups = peer ? peer->nglobal : ngroups;
  if (jit_mode == 0 || ngroups > 32 || out_groups > 32 || out_groups <= 0 || (!peer && scan->nrows == 0)) return MSC_OK;
  if (peer && (peer->world < 1 || peer->world > MSC_PEER_MAX_WORLD || peer->rank < 0 || peer->rank >= peer->world || peer->nlocal != ngroups ||
               peer->gmax < ngroups || !peer->inv))
    return ctx->fail(MSC_ERR_ARG, "bad peer specification");
  MSC_TRY(validate_program(ctx, scan, MODE_DENSE, agg_kinds, naggs, nullptr, 0));
  MSC_TRY(validate_program(ctx, final_scan, MODE_PROJECT, nullptr, 0, out_phys, nout));
  auto p = std::make_unique<msc_prepared>();
  p->ctx = ctx;
  p->scan = *scan;
  p->fin = *final_scan;
  p->ngroups = ngroups;
  p->naggs = naggs;
  p->nout = nout;
  memcpy(p->kinds, agg_kinds, sizeof(int32_t) * naggs);
  memcpy(p->out_phys, out_phys, sizeof(int32_t) * nout);
  memcpy(p->fin_cols, final_cols, sizeof(int32_t) * final_scan->nstaged);
  MSC_TRY(dense_plan(ctx, scan, agg_kinds, naggs, &p->dp));
  if (!jit_dense_supported(scan, ngroups, p->dp.stride)) return MSC_OK;
  if (peer) {
    p->has_peer = true;
    p->peer = *peer;
  }
  if (!ctx->d_ticket) {
    MSC_TRY(msc_alloc(ctx, sizeof(uint32_t), reinterpret_cast<void**>(&ctx->d_ticket)));
    MSC_CUDA(ctx, cudaMemsetAsync(ctx->d_ticket, 0, sizeof(uint32_t), ctx->stream));
  }
  // compile both variants' worth lazily: the masked one now (the common case), the exact one if a pass ever needs it
  int rc = MSC_OK;
  auto drop = [&]() {
    for (int i = 0; i < msc_prepared::kRing; ++i) {
      if (p->ring[i]) msc_rel_free(p->ring[i]);
      if (p->done[i]) cudaEventDestroy(p->done[i]);
      if (p->k0[i]) cudaEventDestroy(p->k0[i]);
      if (p->k1[i]) cudaEventDestroy(p->k1[i]);
      p->ring[i] = nullptr;
      p->done[i] = p->k0[i] = p->k1[i] = nullptr;
    }
    if (p->h_meta) cudaFreeHost(p->h_meta);
    p->h_meta = nullptr;
  };
  for (int i = 0; rc == MSC_OK && i < msc_prepared::kRing; ++i) {
    p->ring[i] = new_rel(ctx, static_cast<uint64_t>(out_groups));
    rc = add_cols(ctx, p->ring[i], out_phys, nout, static_cast<uint64_t>(out_groups));
    if (rc == MSC_OK) rc = msc_alloc(ctx, 3 * sizeof(unsigned long long), reinterpret_cast<void**>(&p->ring[i]->d_meta));
    if (rc == MSC_OK && (cudaEventCreateWithFlags(&p->done[i], cudaEventDisableTiming) != cudaSuccess || cudaEventCreate(&p->k0[i]) != cudaSuccess ||
                         cudaEventCreate(&p->k1[i]) != cudaSuccess))
      rc = ctx->fail(MSC_ERR_CUDA, "cudaEventCreate failed");
  }
  if (rc == MSC_OK && cudaHostAlloc(reinterpret_cast<void**>(&p->h_meta), sizeof(unsigned long long) * 3 * msc_prepared::kRing, cudaHostAllocDefault) != cudaSuccess)
    rc = ctx->fail(MSC_ERR_CUDA, "cudaHostAlloc failed");
  p->table_bytes = sizeof(unsigned long long) * ngroups * p->dp.stride;
  if (rc == MSC_OK) rc = msc_alloc(ctx, p->table_bytes, reinterpret_cast<void**>(&p->table));
  if (rc != MSC_OK) {
    drop();
    return rc;
  }
  msc_rel* rel = p->ring[0];
  void* outs[MSC_VM_MAX_OUT];
  for (int i = 0; i < nout; ++i) outs[i] = rel->cols[i].data;
  msc_peer_spec probe;
  memset(&probe, 0, sizeof(probe));
  if (peer) probe = *peer;
  probe.compile_only = 1;
  JitFinish fin{&p->fin, p->fin_cols, p->out_phys, nout, p->dp.count_slot, outs, rel->d_meta, ctx->d_ticket, &probe};
  static const bool masked_enabled = !(getenv("MSC_SCAN_MASKED") && atoi(getenv("MSC_SCAN_MASKED")) == 0);
  bool masked = masked_enabled;
  if (!peer) fin.peer = nullptr;
  fin.compile_only = true;  // find out NOW whether the generator accepts the pair of programs and a compiler exists
  rc = jit_dense_launch(ctx, scan, ngroups, naggs, p->dp.stride, p->dp.kinds, p->dp.init, p->table, false, &masked, &fin);
  if (rc != MSC_OK) {
    const bool declined = (rc == MSC_ERR_ARG && ctx->err.rfind("jit:", 0) == 0) || (rc == MSC_ERR_CUDA && ctx->err.rfind("jit: lib", 0) == 0);
    drop();
    msc_free(ctx, p->table, p->table_bytes);
    return declined ? MSC_OK : rc;
  }
  *out = p.release();
  return MSC_OK;
}

extern "C" int msc_prepared_enqueue(msc_prepared* p, int32_t flags, uint64_t epoch) {
  if (!p) return MSC_ERR_ARG;
  msc_ctx* ctx = p->ctx;
  if (p->issued - p->collected >= static_cast<uint64_t>(msc_prepared::kRing)) return ctx->fail(MSC_ERR_ARG, "too many passes in flight: msc_prepared_wait first");
  static const bool masked_enabled = !(getenv("MSC_SCAN_MASKED") && atoi(getenv("MSC_SCAN_MASKED")) == 0);
  const int slot = static_cast<int>(p->issued % msc_prepared::kRing);
  msc_rel* rel = p->ring[slot];
  void* outs[MSC_VM_MAX_OUT];
  for (int i = 0; i < p->nout; ++i) outs[i] = rel->cols[i].data;
  p->peer.epoch = epoch;
  p->peer.compile_only = 0;
  // the finish writes {rows, non-finite, error word} straight into pinned host memory (unified addressing): no copy
  // between two passes' kernels, the host only waits for the pass's event
  JitFinish fin{&p->fin, p->fin_cols, p->out_phys, p->nout, p->dp.count_slot, outs, p->h_meta + 3 * slot, ctx->d_ticket, p->has_peer ? &p->peer : nullptr};
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  if (!p->table_clean) {
    dense_init_kernel<<<1, 256, 0, ctx->stream>>>(p->table, p->ngroups, p->dp.stride, dense_meta(p->dp));
    ctx->stats.launches += 1;
  }
  bool masked = !(flags & MSC_DENSE_EXACT) && masked_enabled;
  p->table_clean = false;
  MSC_CUDA(ctx, cudaEventRecord(p->k0[slot], ctx->stream));
  MSC_TRY(jit_dense_launch(ctx, &p->scan, p->ngroups, p->naggs, p->dp.stride, p->dp.kinds, p->dp.init, p->table, true, &masked, &fin));
  MSC_CUDA(ctx, cudaEventRecord(p->k1[slot], ctx->stream));
  p->table_clean = true;  // (stream order: the finish has left the identities behind by the time the next pass starts)
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  MSC_CUDA(ctx, cudaEventRecord(p->done[slot], ctx->stream));
  p->issued += 1;
  return MSC_OK;
}

extern "C" int msc_prepared_wait(msc_prepared* p, msc_rel** result, uint64_t* nrows, int32_t* nonfinite) {
  if (!p || !result || !nrows || !nonfinite) return MSC_ERR_ARG;
  msc_ctx* ctx = p->ctx;
  if (p->collected == p->issued) return ctx->fail(MSC_ERR_ARG, "no pass in flight");
  const int slot = static_cast<int>(p->collected % msc_prepared::kRing);
  MSC_CUDA(ctx, cudaEventSynchronize(p->done[slot]));
  p->collected += 1;
  const unsigned long long* h = p->h_meta + 3 * slot;
  msc_rel* rel = p->ring[slot];
  rel->nrows = h[0];
  rel->pending = false;
  *nonfinite = h[1] != 0;
  *result = rel;
  *nrows = rel->nrows;
  {
    float ms = 0;  // this pass's kernel alone (its own event pair: later passes may be in flight already)
    if (cudaEventElapsedTime(&ms, p->k0[slot], p->k1[slot]) == cudaSuccess) {
      ctx->stats.last_scan_ms = ms;
      ctx->stats.last_kernel_ms = ms;
    }
  }
  if (h[2]) {
    MSC_CUDA(ctx, cudaMemsetAsync(ctx->d_err, 0, sizeof(int), ctx->stream));
    return msc_device_error_rc(ctx, static_cast<int>(h[2]));
  }
  return MSC_OK;
}

extern "C" int msc_prepared_run(msc_prepared* p, int32_t flags, uint64_t epoch, msc_rel** result, uint64_t* nrows, int32_t* nonfinite) {
  if (!p || !result || !nrows || !nonfinite) return MSC_ERR_ARG;
  if (p->issued != p->collected) return p->ctx->fail(MSC_ERR_ARG, "passes in flight: collect them with msc_prepared_wait first");
  MSC_TRY(msc_prepared_enqueue(p, flags, epoch));
  return msc_prepared_wait(p, result, nrows, nonfinite);
}

extern "C" void msc_prepared_free(msc_prepared* p) {
  if (!p) return;
  cudaStreamSynchronize(p->ctx->stream);
  for (int i = 0; i < msc_prepared::kRing; ++i) {
    if (p->ring[i]) msc_rel_free(p->ring[i]);
    if (p->done[i]) cudaEventDestroy(p->done[i]);
    if (p->k0[i]) cudaEventDestroy(p->k0[i]);
    if (p->k1[i]) cudaEventDestroy(p->k1[i]);
  }
  if (p->h_meta) cudaFreeHost(p->h_meta);
  if (p->table) msc_free(p->ctx, p->table, p->table_bytes);
  delete p;
}

extern "C" int msc_jit_dense_source(const msc_scan_desc* sd, int32_t ngroups, const int32_t* agg_kinds, int32_t naggs, int32_t masked, char* buf,
                                    size_t cap, size_t* len) {
  if (!sd || !agg_kinds || !len || ngroups <= 0 || naggs < 0 || naggs > MSC_VM_MAX_AGGS) return MSC_ERR_ARG;
  msc_ctx scratch;  // only carries the error text of dense_plan
  DensePlan dp;
  if (dense_plan(&scratch, sd, agg_kinds, naggs, &dp) != MSC_OK) return MSC_ERR_ARG;
  std::string source, err;
  if (!jit_dense_supported(sd, ngroups, dp.stride)) err = "groups x accumulators exceed the register budget of a specialised kernel";
  else jit_dense_source(sd, ngroups, naggs, dp.stride, dp.kinds, dp.init, masked != 0, &source, &err);
  const std::string& text = source.empty() ? err : source;
  *len = text.size();
  if (buf && cap) {
    const size_t n = std::min(cap - 1, text.size());
    memcpy(buf, text.data(), n);
    buf[n] = 0;
  }
  return source.empty() ? MSC_ERR_ARG : MSC_OK;
}

extern "C" int msc_jit_dense_fused_source(const msc_scan_desc* sd, int32_t ngroups, const int32_t* agg_kinds, int32_t naggs, int32_t masked,
                                          const msc_scan_desc* final_scan, const int32_t* final_cols, const int32_t* out_phys, int32_t nout,
                                          int32_t peer, char* buf, size_t cap, size_t* len) {
  if (!sd || !agg_kinds || !len || !final_scan || !final_cols || !out_phys || ngroups <= 0 || naggs < 0 || naggs > MSC_VM_MAX_AGGS) return MSC_ERR_ARG;
  msc_ctx scratch;
  DensePlan dp;
  if (dense_plan(&scratch, sd, agg_kinds, naggs, &dp) != MSC_OK) return MSC_ERR_ARG;
  std::string source, err;
  msc_peer_spec dummy;
  memset(&dummy, 0, sizeof(dummy));
  JitFinish fin{final_scan, final_cols, out_phys, nout, dp.count_slot, nullptr, nullptr, nullptr, peer ? &dummy : nullptr};
  if (!jit_dense_supported(sd, ngroups, dp.stride)) err = "groups x accumulators exceed the register budget of a specialised kernel";
  else jit_dense_source(sd, ngroups, naggs, dp.stride, dp.kinds, dp.init, masked != 0, &source, &err, &fin);
  const std::string& text = source.empty() ? err : source;
  *len = text.size();
  if (buf && cap) {
    const size_t n = std::min(cap - 1, text.size());
    memcpy(buf, text.data(), n);
    buf[n] = 0;
  }
  return source.empty() ? MSC_ERR_ARG : MSC_OK;
}

extern "C" int msc_jit_runs_source(const msc_scan_desc* sd, const int32_t* agg_kinds, int32_t naggs, int32_t key_col, char* buf, size_t cap, size_t* len) {
  if (!sd || !agg_kinds || !len || naggs < 0 || naggs > MSC_VM_MAX_AGGS || key_col < 0 || key_col >= sd->nstaged) return MSC_ERR_ARG;
  std::string source, err;
  jit_runs_source(sd, naggs, agg_kinds, key_col, &source, &err);
  const std::string& text = source.empty() ? err : source;
  *len = text.size();
  if (buf && cap) {
    const size_t n = std::min(cap - 1, text.size());
    memcpy(buf, text.data(), n);
    buf[n] = 0;
  }
  return source.empty() ? MSC_ERR_ARG : MSC_OK;
}

extern "C" int msc_jit_project_source(const msc_scan_desc* sd, int32_t count_only, const int32_t* out_phys, int32_t nout, char* buf, size_t cap,
                                      size_t* len) {
  if (!sd || !len || nout < 0 || nout > MSC_VM_MAX_OUT || (nout && !out_phys)) return MSC_ERR_ARG;
  std::string source, err;
  jit_project_source(sd, count_only != 0, out_phys, nout, &source, &err);
  const std::string& text = source.empty() ? err : source;
  *len = text.size();
  if (buf && cap) {
    const size_t n = std::min(cap - 1, text.size());
    memcpy(buf, text.data(), n);
    buf[n] = 0;
  }
  return source.empty() ? MSC_ERR_ARG : MSC_OK;
}

extern "C" int msc_jit_compile(const char* source, void* cubin, size_t cap, size_t* len, char* log, size_t log_cap) {
  if (!source || !len) return MSC_ERR_ARG;
  std::vector<char> bin;
  std::string err;
  const int rc = jit_compile_source(source, &bin, &err);
  if (log && log_cap) snprintf(log, log_cap, "%s", err.c_str());
  *len = bin.size();
  if (rc == MSC_OK && cubin && cap >= bin.size()) memcpy(cubin, bin.data(), bin.size());
  return rc;
}

extern "C" int msc_dense_compact(msc_ctx* ctx, const void* table, int32_t ngroups, int32_t stride, const int32_t* agg_kinds,
                                 int32_t naggs, int32_t count_slot, msc_rel** out) {
  if (!ctx || !table || !out || ngroups <= 0) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  DensePlan dp;
  MSC_TRY(plan_for_table(ctx, agg_kinds, naggs, stride, count_slot, &dp));
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  bool nonfinite = false;
  MSC_TRY(dense_compact(ctx, static_cast<const unsigned long long*>(table), ngroups, naggs, dp, out, &nonfinite));
  note_times(ctx, false);
  return MSC_OK;
}

extern "C" int msc_scan_aggregate(msc_ctx* ctx, const msc_scan_desc* sd, int32_t ngroups, const int32_t* agg_kinds,
                                  int32_t naggs, uint64_t hash_capacity_hint, msc_rel** out) {
  if (!ctx || !sd || !out || naggs < 0 || naggs > MSC_VM_MAX_AGGS) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  const bool dense = ngroups > 0;
  MSC_TRY(validate_program(ctx, sd, dense ? MODE_DENSE : MODE_HASH, agg_kinds, naggs, nullptr, 0));
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));

  if (dense) {
    DensePlan dp;
    MSC_TRY(dense_plan(ctx, sd, agg_kinds, naggs, &dp));
    DevTmp table(ctx);
    MSC_TRY(table.alloc(sizeof(unsigned long long) * ngroups * dp.stride));
    // attempt 0 may use a masked regvm variant (exactly `ngroups` groups, per-group reduction in registers); if a SUM
    // comes back non-finite that variant cannot be trusted (gen_regvm.py) and attempt 1 reruns on the generic kernel
    for (int attempt = 0; attempt < 2; ++attempt) {
      bool masked = false, nonfinite = false;
      MSC_TRY(dense_scan_into(ctx, sd, ngroups, agg_kinds, naggs, dp, attempt == 0, table.as<unsigned long long>(), &masked));
      msc_rel* rel = nullptr;
      MSC_TRY(dense_compact(ctx, table.as<unsigned long long>(), ngroups, naggs, dp, &rel, &nonfinite));
      if (masked && nonfinite) {
        msc_rel_free(rel);
        continue;
      }
      note_times(ctx, sd->nrows > 0);
      *out = rel;
      return MSC_OK;
    }
    return ctx->fail(MSC_ERR_ARG, "dense aggregate: unreachable");
  }

  long long init[MSC_VM_MAX_AGGS + 1];
  int kinds[MSC_VM_MAX_AGGS + 1];
  for (int a = 0; a < naggs; ++a) {
    kinds[a] = agg_kinds[a];
    MSC_TRY(agg_identity(ctx, agg_kinds[a], &init[a]));
  }
  const int R = pick_rows_per_thread(sd->nrows);
  ctx->stats.last_agg_runs = 0;

  // ---- hash mode ----
  // a hint with MSC_HASH_HINT_SOFT set is what an earlier run of the same scan produced: the first table is sized for it, but
  // nothing is taken on trust -- more groups than that only cost a second scan
  const uint64_t soft_groups = (hash_capacity_hint & MSC_HASH_HINT_SOFT) ? (hash_capacity_hint & ~MSC_HASH_HINT_SOFT) : 0;
  if (hash_capacity_hint & MSC_HASH_HINT_SOFT) hash_capacity_hint = 0;
  uint64_t want = hash_capacity_hint ? hash_capacity_hint : sd->nrows;
  static const uint64_t runs_min_rows = getenv("MSC_SCAN_RUNS_MIN_ROWS") ? strtoull(getenv("MSC_SCAN_RUNS_MIN_ROWS"), nullptr, 10) : (1u << 16);
  if (!hash_capacity_hint && sd->nrows >= runs_min_rows && !sd->nrows_dev) {
    // No hint.  When the key is a plain integer column, one pass over it counts its runs of equal adjacent values:
    //  * their number bounds the number of groups (sizing for nrows made the sf10 l_orderkey table 4.3 GB for 15 M groups);
    //  * a column without descents is sorted, so every run is exactly one group: a streaming aggregate (MODE_RUNS) writes
    //    group r's key and accumulators straight to row r of the result -- no table, no probing, no compaction.
    static const bool runs_enabled = !(getenv("MSC_SCAN_RUNS") && atoi(getenv("MSC_SCAN_RUNS")) == 0);
    int key_col = -1;
    bool filtered = false;
    for (int pc = 0; pc + 1 < sd->ncode; pc += 2) {
      const uint32_t w0 = sd->code[pc], a = sd->code[pc + 1] & 0xffffu;
      if ((w0 & 0x3f) == MSC_OP_END) break;
      if (((w0 >> 6) & 7) == MSC_DST_FILTER) filtered = true;
      if (((w0 >> 6) & 7) == MSC_DST_GROUP && (w0 & 0x3f) == MSC_OP_MOV && ((a >> 12) & 15) == MSC_SRC_STAGED) key_col = a & 0xfff;
    }
    const int kphys = key_col >= 0 ? sd->staged[key_col].phys : -1;
    if (kphys == MSC_P_U8 || kphys == MSC_P_U16 || kphys == MSC_P_U32 || kphys == MSC_P_I32 || kphys == MSC_P_I64) {
      const void* kdata = sd->staged[key_col].data;
      const uint64_t ntiles = (sd->nrows + RUN_TILE - 1) / RUN_TILE;
      static const bool index_enabled = !(getenv("MSC_RUN_INDEX") && atoi(getenv("MSC_RUN_INDEX")) == 0);
      auto cached = (sd->table_columns && index_enabled) ? ctx->run_index.find(kdata) : ctx->run_index.end();
      if (cached != ctx->run_index.end() && (cached->second.nrows != sd->nrows || cached->second.ntiles != ntiles)) cached = ctx->run_index.end();
      DevTmp d_desc(ctx), rcounts(ctx), roffsets(ctx);
      uint64_t runs = 0;
      bool sorted = false;
      uint64_t* run_offsets = nullptr;
      if (cached != ctx->run_index.end()) {  // this table column has been looked at before (msc_ctx::RunIndex)
        runs = cached->second.runs;
        sorted = cached->second.sorted;
        run_offsets = static_cast<uint64_t*>(cached->second.offsets);
        ctx->stats.last_run_index_hit = 1;
      } else {
      ctx->stats.last_run_index_hit = 0;
      MSC_TRY(d_desc.alloc(sizeof(unsigned long long)));
      MSC_TRY(rcounts.alloc(sizeof(uint32_t) * ntiles));
      MSC_TRY(roffsets.alloc(sizeof(uint64_t) * (ntiles + 1)));
      MSC_CUDA(ctx, cudaMemsetAsync(d_desc.p, 0, sizeof(unsigned long long), ctx->stream));
      const unsigned grid = static_cast<unsigned>(std::min<uint64_t>((ntiles + 7) / 8, static_cast<uint64_t>(ctx->sm_count) * 16));
      unsigned long long* dd = d_desc.as<unsigned long long>();
      uint32_t* rc_ = rcounts.as<uint32_t>();
      switch (kphys) {
        case MSC_P_U8: run_heads_kernel<<<grid, RUN_TILE, 0, ctx->stream>>>(static_cast<const uint8_t*>(kdata), sd->nrows, ntiles, rc_, dd); break;
        case MSC_P_U16: run_heads_kernel<<<grid, RUN_TILE, 0, ctx->stream>>>(static_cast<const uint16_t*>(kdata), sd->nrows, ntiles, rc_, dd); break;
        case MSC_P_U32: run_heads_kernel<<<grid, RUN_TILE, 0, ctx->stream>>>(static_cast<const uint32_t*>(kdata), sd->nrows, ntiles, rc_, dd); break;
        case MSC_P_I32: run_heads_kernel<<<grid, RUN_TILE, 0, ctx->stream>>>(static_cast<const int32_t*>(kdata), sd->nrows, ntiles, rc_, dd); break;
        default: run_heads_kernel<<<grid, RUN_TILE, 0, ctx->stream>>>(static_cast<const long long*>(kdata), sd->nrows, ntiles, rc_, dd); break;
      }
      ctx->stats.launches += 1;
      MSC_TRY(msc_exclusive_scan_u32_u64(ctx, rc_, roffsets.as<uint64_t>(), ntiles));
      unsigned long long* h = ctx->h_scratch;
      MSC_CUDA(ctx, cudaMemcpyAsync(h, roffsets.as<uint64_t>() + ntiles, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
      MSC_CUDA(ctx, cudaMemcpyAsync(h + 1, d_desc.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
      MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      runs = h[0];
      sorted = h[1] == 0;
      run_offsets = roffsets.as<uint64_t>();
      if (sd->table_columns && index_enabled) {  // keep it with the table column (a handful of columns: the oldest entry makes room)
        if (ctx->run_index.size() >= 8) {
          auto victim = ctx->run_index.begin();
          void* vo = victim->second.offsets;
          const size_t vb = victim->second.offsets_bytes;
          ctx->run_index.erase(victim);
          msc_free(ctx, vo, vb);
        }
        msc_ctx::RunIndex ri;
        ri.nrows = sd->nrows;
        ri.runs = runs;
        ri.ntiles = ntiles;
        ri.sorted = sorted;
        ri.offsets = roffsets.p;
        ri.offsets_bytes = roffsets.n;
        roffsets.p = nullptr;  // (ownership moves to the index)
        roffsets.n = 0;
        ctx->run_index[kdata] = ri;
      }
      }
      if (runs > 0 && runs < want) want = runs;
      if (sorted && !filtered && runs_enabled && runs > 0 && runs < (1ull << 31)) {
        // ---- streaming aggregate over the runs of a sorted key ----
        LaunchPlan lp;
        MSC_TRY(plan_launch(ctx, sd, 8, 0, &lp));
        lp.p.naggs = naggs;
        memcpy(lp.p.agg_kind, kinds, sizeof(int) * naggs);
        msc_rel* rel = new_rel(ctx, runs);
        int rc = add_col(ctx, rel, MSC_P_I64, runs);
        for (int a = 0; rc == MSC_OK && a < naggs; ++a)
          rc = add_col(ctx, rel, (kinds[a] == MSC_AGG_SUM_F || kinds[a] == MSC_AGG_MIN_F || kinds[a] == MSC_AGG_MAX_F) ? MSC_P_F64 : MSC_P_I64, runs);
        if (rc != MSC_OK) {
          msc_rel_free(rel);
          return rc;
        }
        lp.p.out[0] = rel->cols[0].data;
        for (int a = 0; a < naggs; ++a) lp.p.out[1 + a] = rel->cols[1 + a].data;
        lp.p.tile_offsets = run_offsets;
        lp.p.run_key_col = key_col;
        static const int jit_mode = getenv("MSC_SCAN_JIT") ? atoi(getenv("MSC_SCAN_JIT")) : 1;
        bool jitted = false;
        if (jit_mode > 0 && naggs + 1 <= MSC_VM_MAX_OUT && (sd->want_jit != 0 || jit_mode > 1 || jit_runs_cached(ctx, sd, naggs, kinds, key_col))) {
          // The specialised kernel STORES every run's cell before anything is added to it (jit_prelude.inc fold_store_heads): the
          // result columns need no identity fill.  What a tile adds to the run that was open when it began goes to a carry
          // cell per (accumulator, tile), folded into the runs by carry_fold_kernel afterwards.
          DevTmp carry(ctx);
          const uint64_t cstride = (ntiles + 1) & ~1ull;  // (a row per accumulator, of even length: 16-byte stores fill it)
          MSC_TRY(carry.alloc(sizeof(unsigned long long) * std::max<uint64_t>(2, static_cast<uint64_t>(naggs) * cstride)));
          FillCols fc;
          CarryFold cf;
          fc.n = cf.n = naggs;
          for (int a = 0; a < naggs; ++a) {
            fc.col[a] = carry.as<unsigned long long>() + static_cast<uint64_t>(a) * cstride;
            fc.init[a] = cf.init[a] = static_cast<unsigned long long>(init[a]);
            cf.col[a] = static_cast<unsigned long long*>(rel->cols[1 + a].data);
            cf.kind[a] = kinds[a];
          }
          if (naggs > 0) {
            fill_cols_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(fc, ntiles);
            ctx->stats.launches += 1;
          }
          rc = jit_runs_launch(ctx, sd, naggs, kinds, key_col, run_offsets, lp.p.out, carry.as<unsigned long long>(), true);
          jitted = rc == MSC_OK;
          const bool declined = (rc == MSC_ERR_ARG && ctx->err.rfind("jit:", 0) == 0) || (rc == MSC_ERR_CUDA && ctx->err.rfind("jit: lib", 0) == 0);
          if (rc != MSC_OK && !declined) {
            msc_rel_free(rel);
            return rc;
          }
          if (jitted && naggs > 0 && ntiles > 1) {
            carry_fold_kernel<<<static_cast<unsigned>(std::min<uint64_t>((ntiles + 255) / 256, static_cast<uint64_t>(ctx->sm_count) * 8)), 256, 0, ctx->stream>>>(
                carry.as<unsigned long long>(), cstride, run_offsets, ntiles, cf);
            ctx->stats.launches += 1;
          }
        }
        if (!jitted) {  // the interpreter accumulates every run with atomics: all cells start from the identity
          FillCols fc;
          fc.n = naggs;
          for (int a = 0; a < naggs; ++a) {
            fc.col[a] = static_cast<unsigned long long*>(rel->cols[1 + a].data);
            fc.init[a] = static_cast<unsigned long long>(init[a]);
          }
          if (naggs > 0) {
            fill_cols_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(fc, runs);
            ctx->stats.launches += 1;
          }
          rc = launch_scan_r<MODE_RUNS>(ctx, &lp);
          if (rc != MSC_OK) {
            msc_rel_free(rel);
            return rc;
          }
          ctx->stats.last_scan_kind = MSC_SCAN_KIND_RUNS;
        }
        ctx->stats.last_agg_runs = 1;
        MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
        const int drc = msc_check_device_error(ctx);  // synchronises the stream
        if (drc != MSC_OK) {
          msc_rel_free(rel);
          return drc;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
        ctx->stats.last_kernel_ms = ms;
        if (cudaEventElapsedTime(&ms, ctx->ev_s0, ctx->ev_s1) == cudaSuccess) ctx->stats.last_scan_ms = ms;
        *out = rel;
        return MSC_OK;
      }
    }
  }
  if (want < 16) want = 16;
  uint32_t shift = 0;
  while ((1u << shift) < static_cast<uint32_t>(1 + naggs)) ++shift;
  // Nothing above said how many groups there are (an expression or floating-point key, a filtered scan): sizing the table for
  // "every row its own group" cost 16 GB and 3.8 ms of initialisation and compaction for the 1000 groups of sf10
  // `GROUP BY l_orderkey % 1000`.  Start small instead; the scan counts the keys it inserts and gives up as soon as the table
  // is half full, and the next attempt is 64 times larger (2^15 groups, then 2^21, then the bound itself).
  static const bool optimistic = !(getenv("MSC_HASH_OPTIMISTIC") && atoi(getenv("MSC_HASH_OPTIMISTIC")) == 0);
  // CTA-local pre-aggregation (scan_kernel.cuh, ScanParams::lcap): as many slots as fit ~40 KB of shared memory
  static const int local_slots_env = getenv("MSC_HASH_LOCAL_SLOTS") ? atoi(getenv("MSC_HASH_LOCAL_SLOTS")) : -1;
  DevTmp tbl(ctx), counts(ctx), offsets(ctx), d_ptrs(ctx), hstate(ctx);
  uint64_t cap = 64;
  uint64_t ladder[4];
  int nladder = 0;
  if (optimistic && !hash_capacity_hint) {
    if (soft_groups > 0 && soft_groups + soft_groups / 4 + 16 < want) ladder[nladder++] = soft_groups + soft_groups / 4 + 16;
    for (const uint64_t g : {1ull << 15, 1ull << 21})
      if (g < want && (nladder == 0 || g > ladder[nladder - 1])) ladder[nladder++] = g;
  }
  ladder[nladder++] = want;
  uint64_t tried_cap = 0;
  int scans = 0;
  for (int attempt = 0;; ++attempt) {
    const uint64_t groups = ladder[attempt];
    const bool last = attempt == nladder - 1;
    {
      uint64_t c = sd->nrows > (1u << 16) ? (1ull << 16) : 64;
      while (c < groups * 2) c <<= 1;
      if (!last && c == tried_cap) continue;  // (a table of this size has just overflowed)
      tried_cap = c;
    }
    ++scans;
    const bool known = hash_capacity_hint != 0 || (soft_groups > 0 && attempt == 0);  // `groups` is (about) the number of groups, not a guess
    // (never fewer than 2^16 slots: the 1000 groups of sf10 `GROUP BY l_orderkey % 1000` take 2.6 ms in a 4096-slot table and
    // 1.5 ms spread over 65536 slots -- the same number of hot lines, but more L2 slices share their atomics)
    // (a scan of a few thousand rows -- the final aggregate over merged partial results -- keeps a table of its own size:
    // there is no contention to spread, and initialising / compacting 65536 slots would be most of its time)
    const uint64_t min_cap = sd->nrows > (1u << 16) ? (1ull << 16) : 64;
    cap = min_cap;
    while (cap < groups * 2) cap <<= 1;
    if (cap > (1ULL << 31)) return ctx->fail(MSC_ERR_ARG, "hash aggregate: more than 2^30 groups per GPU is not supported");
    uint32_t lcap = 0;
    if (naggs >= 1 && groups <= (1ull << 21)) {  // (beyond that almost no row would find its key in a few hundred local slots)
      lcap = 4096;
      while (lcap > 64 && static_cast<size_t>(lcap) * (1 + naggs) * 8 > 40 * 1024) lcap >>= 1;
      // more groups than local slots: the local tables would overflow and only cost occupancy (sf10, 1000 groups: 1.94 ms with
      // 512 slots, 1.64 ms without); with an unknown count they stay -- a handful of hot groups is the case that must not
      // happen (50 groups: 18.4 ms without, 2.4 ms with)
      if (known && groups > lcap - lcap / 4) lcap = 0;
      if (local_slots_env >= 0) {
        lcap = 0;
        if (local_slots_env > 0) {
          lcap = 64;
          while (lcap < static_cast<uint32_t>(local_slots_env) && lcap < 8192) lcap <<= 1;
        }
      }
    }
    const size_t local_bytes = lcap ? static_cast<size_t>(lcap) * (1 + naggs) * 8 + 16 : 0;
    LaunchPlan lp;
    MSC_TRY(plan_launch(ctx, sd, R, local_bytes, &lp));
    lp.p.naggs = naggs;
    memcpy(lp.p.agg_kind, kinds, sizeof(int) * naggs);
    memcpy(lp.p.agg_init, init, sizeof(long long) * naggs);
    tbl.release();
    MSC_TRY(tbl.alloc((cap << shift) * sizeof(unsigned long long)));
    {
      DenseMeta meta;
      memset(&meta, 0, sizeof(meta));
      memcpy(meta.init, init, sizeof(long long) * naggs);
      hash_init_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(tbl.as<unsigned long long>(), cap, shift, naggs, meta);
    }
    ctx->stats.launches += 1;
    lp.p.htbl = tbl.as<unsigned long long>();
    lp.p.hshift = shift;
    lp.p.hcap = cap;
    lp.p.lcap = lcap;
    if (!last) {
      if (!hstate.p) MSC_TRY(hstate.alloc(2 * sizeof(unsigned long long)));
      MSC_CUDA(ctx, cudaMemsetAsync(hstate.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
      lp.p.hstate = hstate.as<unsigned long long>();
      lp.p.hlimit = cap / 2;
    }
    if (sd->nrows > 0) MSC_TRY(launch_scan_r<MODE_HASH>(ctx, &lp));
    ctx->stats.last_hash_local_slots = static_cast<int32_t>(lcap);
    ctx->stats.last_hash_attempts = scans;
    if (last) break;
    unsigned long long* h = ctx->h_scratch;
    MSC_CUDA(ctx, cudaMemcpyAsync(h, hstate.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h[1] == 0) break;  // every group found room
  }
  // compaction of occupied slots
  const uint64_t nht = (cap + HTILE - 1) / HTILE;
  MSC_TRY(counts.alloc(nht * sizeof(uint32_t)));
  MSC_TRY(offsets.alloc((nht + 1) * sizeof(uint64_t)));
  hash_count_kernel<<<static_cast<unsigned>(nht), 256, 0, ctx->stream>>>(tbl.as<unsigned long long>(), cap, shift, counts.as<uint32_t>());
  ctx->stats.launches += 1;
  MSC_TRY(msc_exclusive_scan_u32_u64(ctx, counts.as<uint32_t>(), offsets.as<uint64_t>(), nht));
  uint64_t ngrp = 0;
  MSC_CUDA(ctx, cudaMemcpyAsync(&ngrp, offsets.as<uint64_t>() + nht, sizeof(ngrp), cudaMemcpyDeviceToHost, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  int drc = msc_check_device_error(ctx);
  if (drc != MSC_OK) return drc;
  msc_rel* rel = new_rel(ctx, ngrp);
  int rc = add_col(ctx, rel, MSC_P_I64, ngrp);
  for (int a = 0; rc == MSC_OK && a < naggs; ++a)
    rc = add_col(ctx, rel, (kinds[a] == MSC_AGG_SUM_F || kinds[a] == MSC_AGG_MIN_F || kinds[a] == MSC_AGG_MAX_F) ? MSC_P_F64 : MSC_P_I64, ngrp);
  if (rc != MSC_OK) {
    msc_rel_free(rel);
    return rc;
  }
  std::vector<unsigned long long*> ptrs;
  for (int a = 0; a < naggs; ++a) ptrs.push_back(static_cast<unsigned long long*>(rel->cols[1 + a].data));
  MSC_TRY(d_ptrs.alloc(sizeof(void*) * (naggs + 1)));
  if (naggs) MSC_CUDA(ctx, cudaMemcpyAsync(d_ptrs.p, ptrs.data(), sizeof(void*) * naggs, cudaMemcpyHostToDevice, ctx->stream));
  hash_emit_kernel<<<static_cast<unsigned>(nht), 256, 0, ctx->stream>>>(tbl.as<unsigned long long>(), cap, shift, naggs, offsets.as<uint64_t>(),
                                                                      static_cast<long long*>(rel->cols[0].data), d_ptrs.as<unsigned long long*>());
  ctx->stats.launches += 1;
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
  ctx->stats.last_kernel_ms = ms;
  if (sd->nrows > 0 && cudaEventElapsedTime(&ms, ctx->ev_s0, ctx->ev_s1) == cudaSuccess) ctx->stats.last_scan_ms = ms;
  *out = rel;
  return MSC_OK;
}

// The build half of a hash join as ONE scan over the build side's base rows: filters, then GROUP <- key; rows that pass
// insert (key, their row number) into a compact table.  A count pass sizes the table when the scan filters.
extern "C" int msc_scan_join_build(msc_ctx* ctx, const msc_scan_desc* sd, msc_rel** out_table, int32_t* usable, uint64_t* nkeys_out) {
  if (!ctx || !sd || !out_table || !usable) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  MSC_TRY(validate_program(ctx, sd, MODE_BUILD, nullptr, 0, nullptr, 0));
  if (sd->nrows >= 0xFFFFFFFFull) return ctx->fail(MSC_ERR_ARG, "join side exceeds 2^32-1 rows");
  *usable = 0;
  *out_table = nullptr;
  int rank_pc = -1;
  bool grouped = false;
  for (int pc = 0; pc + 1 < sd->ncode; pc += 2) {
    const int op = sd->code[pc] & 0x3f, dk = (sd->code[pc] >> 6) & 7;
    if (op == MSC_OP_END) break;
    if (op == MSC_OP_RANK) rank_pc = pc;
    if (dk == MSC_DST_GROUP) grouped = true;
    if (dk == MSC_DST_AGG || dk == MSC_DST_OUT) return ctx->fail(MSC_ERR_ARG, "a join build scan has filters and a GROUP (the key) only");
  }
  if (!grouped) return ctx->fail(MSC_ERR_ARG, "join build scan without a key");
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  LaunchPlan lp;
  MSC_TRY(plan_launch(ctx, sd, 4, 0, &lp));
  uint64_t nkeys = sd->nrows;
  if (rank_pc >= 0 && sd->nrows > 0) {  // filtered: count the survivors first (the table must be sized for them, not for the base rows)
    LaunchPlan cp = lp;
    cp.p.code[rank_pc + 2] = MSC_OP_END;
    DevTmp counts(ctx), offsets(ctx);
    MSC_TRY(counts.alloc(sizeof(uint32_t) * cp.p.ntiles));
    MSC_TRY(offsets.alloc(sizeof(uint64_t) * (cp.p.ntiles + 1)));
    cp.p.tile_counts = counts.as<uint32_t>();
    cp.timed = false;
    MSC_TRY(launch_scan_r<MODE_COUNT>(ctx, &cp));
    MSC_TRY(msc_exclusive_scan_u32_u64(ctx, counts.as<uint32_t>(), offsets.as<uint64_t>(), cp.p.ntiles));
    MSC_CUDA(ctx, cudaMemcpyAsync(&nkeys, offsets.as<uint64_t>() + cp.p.ntiles, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    const int drc = msc_check_device_error(ctx);  // synchronises the stream
    if (drc != MSC_OK) return drc;
  }
  if (nkeys_out) *nkeys_out = nkeys;
  uint64_t cap = 64, bitmap_bits = 1024;
  while (cap < nkeys + nkeys / 2) cap <<= 1;
  while (bitmap_bits < nkeys * 16 && bitmap_bits < (1ull << 32)) bitmap_bits <<= 1;
  msc_rel* rel = new_rel(ctx, cap);
  msc_col c;
  c.phys = MSC_P_U8;
  c.bytes = sizeof(MscJoinTableHeader) + bitmap_bits / 8 + cap * 8;
  int rc = msc_alloc(ctx, c.bytes, &c.data);
  if (rc != MSC_OK) {
    msc_rel_free(rel);
    return rc;
  }
  rel->cols.push_back(c);
  MscJoinTableHeader* header = static_cast<MscJoinTableHeader*>(c.data);
  MscJoinTableHeader* h = reinterpret_cast<MscJoinTableHeader*>(ctx->h_scratch);  // pinned, 16 words
  *h = MscJoinTableHeader{cap, 0, 8, 0, bitmap_bits, {0, 0, 0}};
  MSC_CUDA(ctx, cudaMemcpyAsync(header, h, sizeof(*h), cudaMemcpyHostToDevice, ctx->stream));
  char* body = static_cast<char*>(c.data) + sizeof(MscJoinTableHeader);
  MSC_CUDA(ctx, cudaMemsetAsync(body, 0, bitmap_bits / 8, ctx->stream));
  MSC_CUDA(ctx, cudaMemsetAsync(body + bitmap_bits / 8, 0xFF, cap * 8, ctx->stream));  // MSC_J_EMPTY8 = all ones
  lp.p.jheader = header;
  lp.p.jbitmap = reinterpret_cast<uint32_t*>(body);
  lp.p.jslots = reinterpret_cast<unsigned long long*>(body + bitmap_bits / 8);
  lp.p.hcap = cap;
  if (sd->nrows > 0) {
    rc = launch_scan_r<MODE_BUILD>(ctx, &lp);
    if (rc != MSC_OK) {
      msc_rel_free(rel);
      return rc;
    }
  }
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  MSC_CUDA(ctx, cudaMemcpyAsync(h, header, sizeof(*h), cudaMemcpyDeviceToHost, ctx->stream));
  const int drc = msc_check_device_error(ctx);  // synchronises the stream
  if (drc != MSC_OK) {
    msc_rel_free(rel);
    return drc;
  }
  note_times(ctx, sd->nrows > 0);
  if (h->duplicates || h->wide_keys) {  // not a lookup table after all: the caller builds the general way
    msc_rel_free(rel);
    return MSC_OK;
  }
  *usable = 1;
  *out_table = rel;
  return MSC_OK;
}

extern "C" int msc_scan_project(msc_ctx* ctx, const msc_scan_desc* sd, const int32_t* out_phys, int32_t nout, msc_rel** out) {
  if (!ctx || !sd || !out || nout < 0 || nout > MSC_VM_MAX_OUT) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  MSC_TRY(validate_program(ctx, sd, MODE_PROJECT, nullptr, 0, out_phys, nout));
  for (int i = 0; i < nout; ++i)
    if (out_phys[i] != MSC_P_I64 && out_phys[i] != MSC_P_F64 && out_phys[i] != MSC_P_U32)
      return ctx->fail(MSC_ERR_ARG, "project output must be I64, F64 or U32");
  const int R = 4;
  bool has_filter = false;
  int rank_pc = -1;
  for (int pc = 0; pc + 1 < sd->ncode; pc += 2) {
    const int op = sd->code[pc] & 0x3f;
    if (op == MSC_OP_END) break;
    if (((sd->code[pc] >> 6) & 7) == MSC_DST_FILTER && op != MSC_OP_RANK) has_filter = true;
    if (op == MSC_OP_RANK) rank_pc = pc;
  }
  if (has_filter && rank_pc < 0) return ctx->fail(MSC_ERR_ARG, "filtered projection needs a RANK instruction");
  // pending mode (sd->nrows_dev): the input's row count is still on the device; sd->nrows bounds it.  Nothing is waited
  // for: outputs are sized for the bound, the kernels mask by the device's count, and the result comes back pending.
  const bool pending = sd->nrows_dev != nullptr;
  static const bool trace = getenv("MSC_TRACE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (trace) fprintf(stderr, "[msc_scan_project pending=%d] %-20s +%.1f us\n", pending ? 1 : 0, what,
                       std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
  };
  if (!pending) MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));  // a pending chain keeps the first call's start mark
  // specialised kernels (jit.cu) for both passes, or the interpreter for both: the two paths tile the rows differently
  static const int jit_mode = getenv("MSC_SCAN_JIT") ? atoi(getenv("MSC_SCAN_JIT")) : 1;
  bool use_jit = jit_mode > 0 && sd->nstaged >= 1 && sd->nrows > 0 && nout > 0 &&
                 (sd->want_jit != 0 || jit_mode > 1 || jit_project_cached(ctx, sd, out_phys, nout));
  if (use_jit) {  // both passes or none: no compiler here, or a program the generator declines -> the interpreter runs the scan
    const int jrc = jit_project_compile(ctx, sd, has_filter, out_phys, nout);
    if (jrc != MSC_OK) {
      if (!(jrc == MSC_ERR_ARG && ctx->err.rfind("jit:", 0) == 0) && !(jrc == MSC_ERR_CUDA && ctx->err.rfind("jit: lib", 0) == 0)) return jrc;
      use_jit = false;
    }
  }
  LaunchPlan lp;
  MSC_TRY(plan_launch(ctx, sd, R, 0, &lp));
  if (use_jit) lp.p.ntiles = static_cast<uint32_t>((sd->nrows + 255) / 256);
  lp.p.nrows_dev = reinterpret_cast<const unsigned long long*>(sd->nrows_dev);
  lp.timed = !pending;  // a pending chain reports the scan it started with (the aggregate), not this follow-up
  uint64_t nout_rows = sd->nrows;
  DevTmp counts(ctx), offsets(ctx);
  if (has_filter && sd->nrows > 0) {
    // pass 1: rows surviving per tile (program truncated after RANK)
    LaunchPlan cp = lp;
    cp.p.code[rank_pc + 2] = MSC_OP_END;
    MSC_TRY(counts.alloc(sizeof(uint32_t) * cp.p.ntiles));
    MSC_TRY(offsets.alloc(sizeof(uint64_t) * (cp.p.ntiles + 1)));
    cp.p.tile_counts = counts.as<uint32_t>();
    if (use_jit) MSC_TRY(jit_project_launch(ctx, sd, true, out_phys, nout, counts.as<uint32_t>(), nullptr, nullptr, cp.timed));
    else MSC_TRY(launch_scan_r<MODE_COUNT>(ctx, &cp));
    MSC_TRY(msc_exclusive_scan_u32_u64(ctx, counts.as<uint32_t>(), offsets.as<uint64_t>(), cp.p.ntiles));
    if (!pending) {
      MSC_CUDA(ctx, cudaMemcpyAsync(&nout_rows, offsets.as<uint64_t>() + cp.p.ntiles, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
      int drc = msc_check_device_error(ctx);  // synchronises the stream
      if (drc != MSC_OK) return drc;
    }
    lp.p.tile_offsets = offsets.as<uint64_t>();
  }
  mark("planned");
  msc_rel* rel = new_rel(ctx, nout_rows);
  {
    int rc = add_cols(ctx, rel, out_phys, nout, nout_rows);
    if (rc == MSC_OK && pending) rc = msc_alloc(ctx, 3 * sizeof(unsigned long long), reinterpret_cast<void**>(&rel->d_meta));
    if (rc != MSC_OK) {
      msc_rel_free(rel);
      return rc;
    }
  }
  mark("allocated");
  for (int i = 0; i < nout; ++i) {
    lp.p.out[i] = rel->cols[i].data;
    lp.p.out_phys[i] = out_phys[i];
  }
  if (sd->nrows > 0 && nout_rows > 0 && nout > 0) {
    int rc = use_jit ? jit_project_launch(ctx, sd, false, out_phys, nout, nullptr, lp.p.tile_offsets, lp.p.out, lp.timed)
                     : launch_scan_r<MODE_PROJECT>(ctx, &lp);
    if (rc != MSC_OK) {
      msc_rel_free(rel);
      return rc;
    }
  }
  mark("launched");
  if (pending) {
    const unsigned long long* count = (has_filter && sd->nrows > 0) ? offsets.as<unsigned long long>() + lp.p.ntiles
                                                                   : reinterpret_cast<const unsigned long long*>(sd->nrows_dev);
    pending_meta_kernel<<<1, 32, 0, ctx->stream>>>(count, rel->d_meta, ctx->d_err);
    ctx->stats.launches += 1;
    MSC_CUDA(ctx, cudaGetLastError());
    rel->pending = true;
    *out = rel;
    return MSC_OK;
  }
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  const int drc = msc_check_device_error(ctx);  // synchronises the stream
  if (drc != MSC_OK) {
    msc_rel_free(rel);
    return drc;
  }
  note_times(ctx, sd->nrows > 0);
  *out = rel;
  return MSC_OK;
}
