// join.cu -- device hash join (build + probe) and shuffle partitioning.
//
// msc_hash_join replaces BroadcastHashJoinTask.generate_chunks (src/mini_spark/tasks.py:201-240:
// dict key -> [left row idx], then per right row emit left-cols ++ right-cols) and the Zig
// JoinProducer (zig-src/src/tasks.zig:21-196).  It is late-materialising: the result is a pair of
// row-index vectors, and the columns of both sides are gathered by the scan kernel (LOADG_*) only
// where a later operator needs them.  Unlike the reference (whose row indices restart per chunk,
// tasks.py:216-218) the build side may be of any size.
//
// msc_partition replaces WriteToShufflePartitions.write (tasks.py:347-375) / zig fill_buckets
// (task_utils.zig:53-98): rows are routed by hash(key) % nparts into partition-contiguous order (stable: input order
// inside every partition), ready for the exchange between ranks (shuffle.cu).
#include <stdlib.h>

#include "common.cuh"
#include "join_table.cuh"

namespace {

constexpr unsigned long long J_EMPTY = MSC_J_EMPTY;
constexpr uint32_t NIL = MSC_J_NIL;
using JoinSlot = MscJoinSlot;

__device__ __forceinline__ unsigned long long norm_key(long long k) { return msc_join_norm_key(k); }

__global__ void join_init_kernel(JoinSlot* slots, uint64_t cap) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < cap;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    JoinSlot s;
    s.key = J_EMPTY;
    s.head = NIL;
    s.len = 0;
    slots[i] = s;
  }
}

// build: key -> chain of left rows (slot.head -> next[row] -> ...)
__global__ void join_build_kernel(const long long* keys, uint32_t n, JoinSlot* slots, uint32_t* next, uint64_t cap,
                                  unsigned long long* duplicates = nullptr, uint32_t* bitmap = nullptr, uint64_t bitmap_bits = 0) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = norm_key(keys[i]);
  const uint64_t mask = cap - 1;
  const uint64_t hash = msc_mix64(k);
  if (bitmap) {
    const uint64_t bit = (hash >> 32) & (bitmap_bits - 1);
    atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));
  }
  uint64_t pos = hash & mask;
  while (true) {  // find-or-insert in one atomic per step: the old value says "inserted", "found" or "someone else's"
    const unsigned long long prev = atomicCAS(&slots[pos].key, J_EMPTY, k);
    if (prev == J_EMPTY || prev == k) break;
    pos = (pos + 1) & mask;
  }
  const uint32_t before = atomicExch(&slots[pos].head, i);
  if (next) next[i] = before;
  if (atomicAdd(&slots[pos].len, 1u) != 0 && duplicates) *duplicates = 1;  // (any writer stores the same value)
}

// compact table (8-byte slots, 32-bit keys, no chains: a second row with the same key only raises the flag)
__global__ void join_init8_kernel(unsigned long long* slots, uint64_t cap) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < cap; i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    slots[i] = MSC_J_EMPTY8;
}

__global__ void join_build8_kernel(const long long* keys, uint32_t n, unsigned long long* slots, uint64_t cap, MscJoinTableHeader* header,
                                   uint32_t* bitmap, uint64_t bitmap_bits) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long key = keys[i];
  if (!msc_join_key_is_32bit(key)) {
    header->wide_keys = 1;
    return;
  }
  const uint32_t k32 = static_cast<uint32_t>(key);
  const unsigned long long mine = (static_cast<unsigned long long>(i) << 32) | k32;
  const uint64_t mask = cap - 1;
  const uint64_t bit = msc_fmix32(k32 ^ 0x9e3779b9u) & (bitmap_bits - 1);
  atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));
  uint64_t pos = msc_fmix32(k32) & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(&slots[pos], MSC_J_EMPTY8, mine);
    if (prev == MSC_J_EMPTY8) break;
    if (static_cast<uint32_t>(prev) == k32) {
      header->duplicates = 1;
      break;
    }
    pos = (pos + 1) & mask;
  }
}

// chain head of key k (NIL: no match) and the chain's length
__device__ __forceinline__ uint32_t join_find(const JoinSlot* slots, uint64_t cap, unsigned long long k, uint32_t* len) {
  const uint64_t mask = cap - 1;
  uint64_t pos = msc_mix64(k) & mask;
  while (true) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(slots + pos));
    const unsigned long long cur = (static_cast<unsigned long long>(raw.y) << 32) | raw.x;
    if (cur == k) {
      *len = raw.w;
      return raw.z;
    }
    if (cur == J_EMPTY) {
      *len = 0;
      return NIL;
    }
    pos = (pos + 1) & mask;
  }
}

// Probe ONCE: the chain head a right row finds is remembered (first[j]), its matches are counted, and the block's total
// goes to block_counts -- the emit pass neither hashes nor probes again, and the prefix sum runs over blocks, not rows
// (probing in both passes and a 64-bit offset per probe row cost 0.33 ms of the sf10 join's 0.62 ms).
constexpr int JBLOCK = 256;

__global__ void join_count_kernel(const long long* rkeys, uint32_t nr, const JoinSlot* slots, const uint32_t* next, uint64_t cap,
                                  uint32_t* first, uint32_t* counts, uint32_t* block_counts) {
  __shared__ uint32_t wsum[JBLOCK / 32];
  const uint32_t j = blockIdx.x * JBLOCK + threadIdx.x;
  uint32_t c = 0;
  if (j < nr) {
    const uint32_t f = join_find(slots, cap, norm_key(rkeys[j]), &c);
    first[j] = f;
    counts[j] = c;
  }
  uint32_t s = c;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
    for (int w = 0; w < JBLOCK / 32; ++w) tot += wsum[w];
    block_counts[blockIdx.x] = tot;
  }
}

// pairs in right-row order (the reference emits right-row major, tasks.py:229-240): block offset + rank inside the block
__global__ void join_emit_kernel(uint32_t nr, const uint32_t* first, const uint32_t* counts, const uint32_t* next,
                                 const uint64_t* block_offsets, uint32_t* out_l, uint32_t* out_r) {
  __shared__ uint32_t wsum[JBLOCK / 32];
  const uint32_t j = blockIdx.x * JBLOCK + threadIdx.x;
  const uint32_t c = j < nr ? counts[j] : 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = c;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint32_t wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += wsum[w];
  if (c == 0) return;
  uint64_t o = block_offsets[blockIdx.x] + wbase + inc - c;
  uint32_t l = first[j];
  for (uint32_t m = 0; m < c; ++m) {  // (a unique build key -- the usual case -- never reads next[])
    out_l[o] = l;
    out_r[o] = j;
    ++o;
    if (m + 1 < c) l = next[l];
  }
}

// ---- partitioning ---------------------------------------------------------------------------------
constexpr int PBLOCK = 4096;  // rows per histogram block
constexpr int PTHREADS = 256;
constexpr int MAX_PARTS = 64;

__device__ __forceinline__ long long read_key(const void* col, int phys, uint64_t i) {
  switch (phys) {
    case MSC_P_U8: return static_cast<const uint8_t*>(col)[i];
    case MSC_P_U16: return static_cast<const uint16_t*>(col)[i];
    case MSC_P_U32: return static_cast<const uint32_t*>(col)[i];
    case MSC_P_I32: return static_cast<const int*>(col)[i];
    case MSC_P_F32: return __double_as_longlong(static_cast<double>(static_cast<const float*>(col)[i]));
    default: return static_cast<const long long*>(col)[i];
  }
}

// Range partitioning: partition p holds the keys in [lo[p], lo[p + 1]) (lo[0] = -inf); use == 0: hash partitioning.
struct PartBounds {
  long long lo[MAX_PARTS];
  int use;
};

// partition id from the HIGH hash bits (the aggregation / join tables index with the low bits), or by key range
__device__ __forceinline__ int part_of(long long key, int nparts, const PartBounds& b) {
  if (b.use) {
    int p = 0;
    for (int i = 1; i < nparts; ++i) p += key >= b.lo[i];
    return p;
  }
  return static_cast<int>((msc_mix64(norm_key(key)) >> 32) % static_cast<unsigned>(nparts));
}

// per-block histogram; *disorder is raised when some row's partition is lower than its predecessor's (rows that are not
// yet partition-contiguous)
__global__ void part_hist_kernel(const void* key, int phys, uint64_t n, int nparts, uint32_t nblocks, uint32_t* hist,
                                 const __grid_constant__ PartBounds bounds, int* disorder) {
  __shared__ uint32_t h[MAX_PARTS];
  if (threadIdx.x < MAX_PARTS) h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * PBLOCK;
  bool down = false;
  for (int i = threadIdx.x; i < PBLOCK; i += PTHREADS) {
    const uint64_t r = base + i;
    if (r < n) {
      const int p = part_of(read_key(key, phys, r), nparts, bounds);
      atomicAdd(&h[p], 1u);
      if (r > 0 && part_of(read_key(key, phys, r - 1), nparts, bounds) > p) down = true;
    }
  }
  if (down) *disorder = 1;
  __syncthreads();
  if (threadIdx.x < nparts) hist[static_cast<uint64_t>(threadIdx.x) * nblocks + blockIdx.x] = h[threadIdx.x];
}

// Output position of every row, STABLE: rows of one partition keep their input order (the reference appends the rows of
// a chunk to each bucket in row order, tasks.py:357-368 / task_utils.zig:60-98).  A block walks its rows 256 at a time;
// inside a warp equal partitions find each other with match.any, the lane's rank among them is a popcount, the warps'
// counts are prefix-summed in shared memory, and the per-partition cursor moves on once per step.
__global__ void part_pos_kernel(const void* key, int phys, uint64_t n, int nparts, uint32_t nblocks,
                                const uint64_t* offsets, uint32_t* pos, uint8_t* part, const __grid_constant__ PartBounds bounds) {
  constexpr int NW = PTHREADS / 32;
  __shared__ unsigned long long cursor[MAX_PARTS];
  __shared__ uint32_t wcount[NW][MAX_PARTS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < nparts) cursor[tid] = offsets[static_cast<uint64_t>(tid) * nblocks + blockIdx.x];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * PBLOCK;
  for (int it = 0; it < PBLOCK / PTHREADS; ++it) {
    const uint64_t r = base + static_cast<uint64_t>(it) * PTHREADS + tid;
    if (base + static_cast<uint64_t>(it) * PTHREADS >= n) break;  // (uniform per block)
    for (int i = tid; i < NW * MAX_PARTS; i += PTHREADS) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const bool valid = r < n;
    const int p = valid ? part_of(read_key(key, phys, r), nparts, bounds) : MAX_PARTS;  // rows past the end group among themselves
    const unsigned same = __match_any_sync(0xffffffffu, p);
    const uint32_t rank = __popc(same & ((1u << lane) - 1u));
    if (valid && rank == 0) wcount[warp][p] = __popc(same);
    __syncthreads();
    uint32_t total = 0;
    if (tid < nparts) {
      for (int w = 0; w < NW; ++w) {
        const uint32_t c = wcount[w][tid];
        wcount[w][tid] = total;
        total += c;
      }
    }
    __syncthreads();
    if (valid) {
      pos[r] = static_cast<uint32_t>(cursor[p] + wcount[warp][p] + rank);
      if (part) part[r] = static_cast<uint8_t>(p);
    }
    __syncthreads();
    if (tid < nparts) cursor[tid] += total;
  }
}

template <class T>
__global__ void scatter_kernel(const T* in, T* out, const uint32_t* pos, uint64_t n) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    out[pos[i]] = in[i];
}

inline unsigned grid_for(uint64_t n, int block) { return static_cast<unsigned>((n + block - 1) / block); }

}  // namespace

extern "C" int msc_hash_join(msc_ctx* ctx, const int64_t* left_keys, uint64_t nleft, const int64_t* right_keys,
                             uint64_t nright, msc_rel** out_pairs) {
  if (!ctx || !out_pairs) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  if (nleft >= NIL || nright >= NIL) return ctx->fail(MSC_ERR_ARG, "join side exceeds 2^32-1 rows");
  msc_rel* rel = new msc_rel();
  rel->ctx = ctx;
  auto finish_empty = [&]() {
    for (int i = 0; i < 2; ++i) {
      msc_col c;
      c.phys = MSC_P_U32;
      msc_alloc_rows(ctx, 0, 4, &c.data, &c.bytes);
      rel->cols.push_back(c);
    }
    *out_pairs = rel;
    return MSC_OK;
  };
  if (nleft == 0 || nright == 0) return finish_empty();
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  uint64_t cap = 64;
  while (cap < nleft * 2) cap <<= 1;
  DevTmp slots(ctx), next(ctx), counts(ctx), offsets(ctx), first(ctx), bcounts(ctx);
  auto fail = [&](int rc) {
    msc_rel_free(rel);
    return rc;
  };
  int rc;
  if ((rc = slots.alloc(cap * sizeof(JoinSlot))) != MSC_OK || (rc = next.alloc(nleft * 4)) != MSC_OK ||
      (rc = counts.alloc(nright * 4)) != MSC_OK || (rc = first.alloc(nright * 4)) != MSC_OK)
    return fail(rc);
  const uint64_t nblocks = (nright + JBLOCK - 1) / JBLOCK;
  if ((rc = bcounts.alloc(nblocks * 4)) != MSC_OK || (rc = offsets.alloc((nblocks + 1) * 8)) != MSC_OK) return fail(rc);
  const uint32_t nl = static_cast<uint32_t>(nleft), nr = static_cast<uint32_t>(nright);
  join_init_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(slots.as<JoinSlot>(), cap);
  join_build_kernel<<<grid_for(nl, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const long long*>(left_keys), nl, slots.as<JoinSlot>(),
                                                               next.as<uint32_t>(), cap);
  join_count_kernel<<<static_cast<unsigned>(nblocks), JBLOCK, 0, ctx->stream>>>(reinterpret_cast<const long long*>(right_keys), nr,
                                                                                slots.as<JoinSlot>(), next.as<uint32_t>(), cap, first.as<uint32_t>(),
                                                                                counts.as<uint32_t>(), bcounts.as<uint32_t>());
  ctx->stats.launches += 3;
  if ((rc = msc_exclusive_scan_u32_u64(ctx, bcounts.as<uint32_t>(), offsets.as<uint64_t>(), nblocks)) != MSC_OK) return fail(rc);
  uint64_t npairs = 0;
  if ((rc = msc_memcpy_d2h(ctx, &npairs, offsets.as<uint64_t>() + nblocks, 8)) != MSC_OK) return fail(rc);
  if (npairs >= NIL) return fail(ctx->fail(MSC_ERR_ARG, "join result exceeds 2^32-1 rows"));
  rel->nrows = npairs;
  for (int i = 0; i < 2; ++i) {
    msc_col c;
    c.phys = MSC_P_U32;
    if ((rc = msc_alloc_rows(ctx, npairs, 4, &c.data, &c.bytes)) != MSC_OK) return fail(rc);
    rel->cols.push_back(c);
  }
  if (npairs) {
    join_emit_kernel<<<static_cast<unsigned>(nblocks), JBLOCK, 0, ctx->stream>>>(nr, first.as<uint32_t>(), counts.as<uint32_t>(), next.as<uint32_t>(),
                                                                                offsets.as<uint64_t>(), static_cast<uint32_t*>(rel->cols[0].data),
                                                                                static_cast<uint32_t*>(rel->cols[1].data));
    ctx->stats.launches += 1;
  }
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  MSC_CUDA(ctx, cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
  ctx->stats.last_kernel_ms = ms;
  *out_pairs = rel;
  return MSC_OK;
}

// The build half alone, for scans that probe per row (MSC_OP_PROBE): a 1-column relation that owns [header][slots].
// Compact 8-byte slots first (32-bit keys: INTEGER columns, dictionary codes); a key outside that range rebuilds wide.
extern "C" int msc_join_build(msc_ctx* ctx, const int64_t* keys, uint64_t nkeys, msc_rel** out_table, int32_t* unique, int32_t* slot_bytes) {
  if (!ctx || !out_table || !unique || !slot_bytes || (nkeys && !keys)) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  if (nkeys >= NIL) return ctx->fail(MSC_ERR_ARG, "join side exceeds 2^32-1 rows");
  uint64_t cap = 64;
  while (cap < nkeys + nkeys / 2) cap <<= 1;  // load <= 2/3: the smaller the table, the more of it stays in L2
  static const bool compact_enabled = !(getenv("MSC_JOIN_COMPACT") && atoi(getenv("MSC_JOIN_COMPACT")) == 0);
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  for (int attempt = compact_enabled ? 0 : 1; attempt < 2; ++attempt) {
    const bool compact = attempt == 0;
    msc_rel* rel = new msc_rel();
    rel->ctx = ctx;
    rel->nrows = cap;
    msc_col c;
    c.phys = MSC_P_U8;
    uint64_t bitmap_bits = 1024;
    while (bitmap_bits < nkeys * 16 && bitmap_bits < (1ull << 32)) bitmap_bits <<= 1;
    c.bytes = sizeof(MscJoinTableHeader) + bitmap_bits / 8 + cap * (compact ? 8 : sizeof(JoinSlot));
    int rc = msc_alloc(ctx, c.bytes, &c.data);
    if (rc != MSC_OK) {
      delete rel;
      return rc;
    }
    rel->cols.push_back(c);
    MscJoinTableHeader* header = static_cast<MscJoinTableHeader*>(c.data);
    MscJoinTableHeader* h = reinterpret_cast<MscJoinTableHeader*>(ctx->h_scratch);  // pinned, 16 words
    *h = MscJoinTableHeader{cap, 0, compact ? 8ull : 16ull, 0, bitmap_bits, {0, 0, 0}};
    MSC_CUDA(ctx, cudaMemcpyAsync(header, h, sizeof(*h), cudaMemcpyHostToDevice, ctx->stream));
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(static_cast<char*>(c.data) + sizeof(MscJoinTableHeader));
    MSC_CUDA(ctx, cudaMemsetAsync(bitmap, 0, bitmap_bits / 8, ctx->stream));
    void* slots = static_cast<char*>(c.data) + sizeof(MscJoinTableHeader) + bitmap_bits / 8;
    const uint32_t n = static_cast<uint32_t>(nkeys);
    if (compact) {
      join_init8_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(static_cast<unsigned long long*>(slots), cap);
      if (n) join_build8_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const long long*>(keys), n, static_cast<unsigned long long*>(slots), cap, header, bitmap, bitmap_bits);
    } else {
      join_init_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(static_cast<JoinSlot*>(slots), cap);
      if (n) join_build_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const long long*>(keys), n, static_cast<JoinSlot*>(slots), nullptr, cap, &header->duplicates, bitmap, bitmap_bits);
    }
    ctx->stats.launches += n ? 2 : 1;
    MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
    MSC_CUDA(ctx, cudaMemcpyAsync(h, header, sizeof(*h), cudaMemcpyDeviceToHost, ctx->stream));
    MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MSC_CUDA(ctx, cudaGetLastError());
    if (compact && h->wide_keys) {  // (INTEGER arithmetic, TIMESTAMP or FLOAT keys)
      msc_rel_free(rel);
      continue;
    }
    *unique = h->duplicates == 0;
    *slot_bytes = compact ? 8 : 16;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b) == cudaSuccess) ctx->stats.last_kernel_ms = ms;
    *out_table = rel;
    return MSC_OK;
  }
  return ctx->fail(MSC_ERR_ARG, "join build: unreachable");
}

static int partition_impl(msc_ctx* ctx, msc_rel* in, int32_t key_col, int32_t nparts, const int64_t* lower_bounds, uint64_t* counts_host,
                          msc_rel** out);

extern "C" int msc_partition(msc_ctx* ctx, msc_rel* in, int32_t key_col, int32_t nparts, uint64_t* counts_host, msc_rel** out) {
  return partition_impl(ctx, in, key_col, nparts, nullptr, counts_host, out);
}

extern "C" int msc_partition_range(msc_ctx* ctx, msc_rel* in, int32_t key_col, int32_t nparts, const int64_t* lower_bounds, uint64_t* counts_host,
                                   msc_rel** out) {
  if (!lower_bounds) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  return partition_impl(ctx, in, key_col, nparts, lower_bounds, counts_host, out);
}

static int partition_impl(msc_ctx* ctx, msc_rel* in, int32_t key_col, int32_t nparts, const int64_t* lower_bounds, uint64_t* counts_host,
                          msc_rel** out) {
  if (!ctx || !in || !out || !counts_host || nparts < 1 || nparts > MAX_PARTS || key_col < 0 ||
      key_col >= static_cast<int32_t>(in->cols.size()))
    return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  PartBounds pb;
  memset(&pb, 0, sizeof(pb));
  if (lower_bounds) {
    pb.use = 1;
    for (int p = 1; p < nparts; ++p) {
      if (p > 1 && lower_bounds[p] < lower_bounds[p - 1]) return ctx->fail(MSC_ERR_ARG, "partition bounds must not decrease");
      pb.lo[p] = lower_bounds[p];
    }
  }
  const uint64_t n = in->nrows;
  msc_rel* rel = new msc_rel();
  rel->ctx = ctx;
  rel->nrows = n;
  auto fail = [&](int rc) {
    msc_rel_free(rel);
    return rc;
  };
  int rc;
  for (auto& src : in->cols) {
    msc_col c;
    c.phys = src.phys;
    if ((rc = msc_alloc_rows(ctx, n, msc_phys_width(c.phys), &c.data, &c.bytes)) != MSC_OK) return fail(rc);
    rel->cols.push_back(c);
  }
  for (int p = 0; p < nparts; ++p) counts_host[p] = 0;
  if (n == 0) {
    *out = rel;
    return MSC_OK;
  }
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  const uint32_t nblocks = static_cast<uint32_t>((n + PBLOCK - 1) / PBLOCK);
  const uint64_t cells = static_cast<uint64_t>(nblocks) * nparts;
  DevTmp hist(ctx), offsets(ctx), pos(ctx), disorder(ctx);
  if ((rc = hist.alloc(cells * 4)) != MSC_OK || (rc = offsets.alloc((cells + 1) * 8)) != MSC_OK || (rc = disorder.alloc(sizeof(int))) != MSC_OK)
    return fail(rc);
  const msc_col& key = in->cols[key_col];
  MSC_CUDA(ctx, cudaMemsetAsync(disorder.p, 0, sizeof(int), ctx->stream));
  part_hist_kernel<<<nblocks, PTHREADS, 0, ctx->stream>>>(key.data, key.phys, n, nparts, nblocks, hist.as<uint32_t>(), pb, disorder.as<int>());
  ctx->stats.launches += 1;
  if ((rc = msc_exclusive_scan_u32_u64(ctx, hist.as<uint32_t>(), offsets.as<uint64_t>(), cells)) != MSC_OK) return fail(rc);
  // rows per partition = difference of the partition-major offsets
  std::vector<uint64_t> bounds(nparts + 1);
  for (int p = 0; p <= nparts; ++p) {
    const uint64_t idx = static_cast<uint64_t>(p) * nblocks;
    MSC_CUDA(ctx, cudaMemcpyAsync(&bounds[p], offsets.as<uint64_t>() + idx, 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  int h_disorder = 0;
  MSC_CUDA(ctx, cudaMemcpyAsync(&h_disorder, disorder.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int p = 0; p < nparts; ++p) counts_host[p] = bounds[p + 1] - bounds[p];
  if (!h_disorder) {
    // The rows are partition-contiguous as they are (one partition, or a sorted key routed by range): the result SHARES the
    // input's columns instead of copying them -- the input relation must stay alive as long as the result is used.
    for (size_t c = 0; c < in->cols.size(); ++c) {
      if (rel->cols[c].owned && rel->cols[c].data) msc_free(ctx, rel->cols[c].data, rel->cols[c].bytes);
      rel->cols[c].data = in->cols[c].data;
      rel->cols[c].bytes = in->cols[c].bytes;
      rel->cols[c].owned = false;
    }
    MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
    MSC_CUDA(ctx, cudaGetLastError());
    ctx->stats.last_kernel_ms = 0;
    *out = rel;
    return MSC_OK;
  }
  if ((rc = pos.alloc(n * 4)) != MSC_OK) return fail(rc);
  part_pos_kernel<<<nblocks, PTHREADS, 0, ctx->stream>>>(key.data, key.phys, n, nparts, nblocks, offsets.as<uint64_t>(), pos.as<uint32_t>(), nullptr, pb);
  ctx->stats.launches += 1;
  const unsigned grid = static_cast<unsigned>(ctx->sm_count * 8);
  for (size_t c = 0; c < in->cols.size(); ++c) {
    const void* src = in->cols[c].data;
    void* dst = rel->cols[c].data;
    switch (msc_phys_width(in->cols[c].phys)) {
      case 1: scatter_kernel<uint8_t><<<grid, 256, 0, ctx->stream>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), pos.as<uint32_t>(), n); break;
      case 2: scatter_kernel<uint16_t><<<grid, 256, 0, ctx->stream>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), pos.as<uint32_t>(), n); break;
      case 4: scatter_kernel<uint32_t><<<grid, 256, 0, ctx->stream>>>(static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst), pos.as<uint32_t>(), n); break;
      default: scatter_kernel<uint64_t><<<grid, 256, 0, ctx->stream>>>(static_cast<const uint64_t*>(src), static_cast<uint64_t*>(dst), pos.as<uint32_t>(), n); break;
    }
    ctx->stats.launches += 1;
  }
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  MSC_CUDA(ctx, cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
  ctx->stats.last_kernel_ms = ms;
  *out = rel;
  return MSC_OK;
}
