// strings.cu -- device string dictionaries: encode, literal lookup, LIKE, translation, concat.
//
// STRING columns arrive as u8-length-prefixed byte runs (reference src/mini_spark/io.py:100-104)
// and are dictionary-encoded on the device at ingest, so the scan kernel only ever sees integer
// codes.  String predicates are evaluated once per *dictionary entry* and applied per row through
// a lookup table:
//   '=' / '!=' against a literal  -> msc_dict_lookup + integer compare
//   LIKE (sql.py:178-179,192-194; zig-regex at templates/plan.zig:66-68) -> msc_dict_like -> LUT8
//   column = column / join keys   -> msc_dict_translate -> LUT32
//   '+' concat (sql.py:331-333; zig concatStrings utils.zig:118-131) -> msc_str_concat
#include "strings.cuh"

namespace {

constexpr uint32_t NO_CODE = 0xFFFFFFFFu;

__device__ __forceinline__ uint64_t hash_bytes(const uint8_t* p, uint32_t len, uint64_t seed) {
  uint64_t h = 0xcbf29ce484222325ULL ^ seed;  // FNV-1a, then a murmur finaliser
  for (uint32_t i = 0; i < len; ++i) {
    h ^= p[i];
    h *= 0x100000001b3ULL;
  }
  h = msc_mix64(h ^ len);
  return h ? h : 1;  // 0 marks an empty slot
}

// pass A: find (or claim) the hash slot of every batch string
template <class TLen>
__global__ void dict_probe_kernel(const uint64_t* starts, const TLen* lens, const uint8_t* bytes, uint64_t n,
                                  unsigned long long* hkeys, const int32_t* hcode, uint32_t* hrep, uint64_t hcap,
                                  uint64_t seed, int insert, uint32_t* slot_out, int* err) {
  const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint64_t h = hash_bytes(bytes + starts[i], lens[i], seed);
  const uint64_t mask = hcap - 1;
  uint64_t pos = h & mask;
  uint32_t slot = NO_CODE;
  for (uint64_t probe = 0; probe < hcap; ++probe) {
    const unsigned long long cur = hkeys[pos];
    if (cur == h) {
      slot = static_cast<uint32_t>(pos);
      break;
    }
    if (cur == 0) {
      if (!insert) break;
      const unsigned long long prev = atomicCAS(hkeys + pos, 0ULL, static_cast<unsigned long long>(h));
      if (prev == 0 || prev == h) {
        slot = static_cast<uint32_t>(pos);
        break;
      }
    }
    pos = (pos + 1) & mask;
  }
  if (insert && slot == NO_CODE) atomicOr(err, MSC_DEVERR_TABLE_FULL);
  if (insert && slot != NO_CODE && hcode[slot] < 0) {
    // one atomic per distinct slot and warp (the lowest lane holds the smallest row): a low-cardinality column
    // would otherwise serialise millions of atomics on a handful of addresses in its first batch
    const unsigned peers = __match_any_sync(__activemask(), slot);
    if ((__ffs(peers) - 1) == static_cast<int>(threadIdx.x & 31)) atomicMin(hrep + slot, static_cast<uint32_t>(i));
  }
  slot_out[i] = slot;
}

// pass B: the representative row of every new slot appends the string to the dictionary
template <class TLen>
__global__ void dict_assign_kernel(const uint64_t* starts, const TLen* lens, const uint8_t* bytes, uint64_t n,
                                   const uint32_t* slot_of, int32_t* hcode, uint32_t* hrep, uint64_t* ent_start,
                                   uint32_t* ent_len, uint8_t* heap, unsigned long long* counters) {
  const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint32_t slot = slot_of[i];
  if (slot == NO_CODE || hrep[slot] != static_cast<uint32_t>(i)) return;
  const uint32_t len = lens[i];
  const unsigned long long code = atomicAdd(counters + 0, 1ULL);
  const unsigned long long off = atomicAdd(counters + 1, static_cast<unsigned long long>(len));
  ent_start[code] = off;
  ent_len[code] = len;
  const uint8_t* src = bytes + starts[i];
  for (uint32_t b = 0; b < len; ++b) heap[off + b] = src[b];
  hcode[slot] = static_cast<int32_t>(code);
  hrep[slot] = NO_CODE;
}

// pass C: read the codes back and verify the bytes (a 64-bit hash collision is reported, never ignored)
template <class TLen>
__global__ void dict_resolve_kernel(const uint64_t* starts, const TLen* lens, const uint8_t* bytes, uint64_t n,
                                    const uint32_t* slot_of, const int32_t* hcode, const uint64_t* ent_start,
                                    const uint32_t* ent_len, const uint8_t* heap, uint32_t* codes_out, int* err) {
  const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint32_t slot = slot_of[i];
  uint32_t code = NO_CODE;
  if (slot != NO_CODE && hcode[slot] >= 0) {
    code = static_cast<uint32_t>(hcode[slot]);
    const uint32_t len = lens[i];
    bool same = ent_len[code] == len;
    const uint8_t* a = bytes + starts[i];
    const uint8_t* b = heap + ent_start[code];
    for (uint32_t k = 0; same && k < len; ++k) same = a[k] == b[k];
    if (!same) {
      atomicOr(err, MSC_DEVERR_COLLISION);
      code = NO_CODE;
    }
  }
  codes_out[i] = code;
}

__global__ void dict_rehash_kernel(const uint64_t* ent_start, const uint32_t* ent_len, const uint8_t* heap, uint32_t n,
                                   unsigned long long* hkeys, int32_t* hcode, uint64_t hcap, uint64_t seed) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint64_t h = hash_bytes(heap + ent_start[e], ent_len[e], seed);
  const uint64_t mask = hcap - 1;
  uint64_t pos = h & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(hkeys + pos, 0ULL, static_cast<unsigned long long>(h));
    if (prev == 0) {
      hcode[pos] = static_cast<int32_t>(e);
      return;
    }
    pos = (pos + 1) & mask;
  }
}

__global__ void fill_i32_kernel(int32_t* p, int32_t v, uint64_t n) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}

// SQL LIKE with '%' (any run) and '_' (any one byte), anchored at both ends
__device__ bool like_match(const uint8_t* s, int n, const uint8_t* p, int m) {
  int i = 0, j = 0, star = -1, mark = 0;
  while (i < n) {
    if (j < m && p[j] != '%' && (p[j] == '_' || p[j] == s[i])) {
      ++i;
      ++j;
    } else if (j < m && p[j] == '%') {
      star = j++;
      mark = i;
    } else if (star >= 0) {
      j = star + 1;
      i = ++mark;
    } else {
      return false;
    }
  }
  while (j < m && p[j] == '%') ++j;
  return j == m;
}

__global__ void dict_like_kernel(const uint64_t* ent_start, const uint32_t* ent_len, const uint8_t* heap, uint32_t n,
                                 const uint8_t* pattern, int plen, uint8_t* lut) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  lut[e] = like_match(heap + ent_start[e], static_cast<int>(ent_len[e]), pattern, plen) ? 1 : 0;
}

struct ConcatPart {
  const void* codes;  // nullptr -> literal
  int phys;
  const uint64_t* ent_start;
  const uint32_t* ent_len;
  const uint8_t* heap;
  uint32_t lit_off;
  uint32_t lit_len;
};
struct ConcatArgs {
  ConcatPart parts[8];
  int nparts;
  const uint8_t* lits;
};

__device__ __forceinline__ uint32_t read_code(const void* codes, int phys, uint64_t i) {
  if (phys == MSC_P_U8) return static_cast<const uint8_t*>(codes)[i];
  if (phys == MSC_P_U16) return static_cast<const uint16_t*>(codes)[i];
  return static_cast<const uint32_t*>(codes)[i];
}

__global__ void concat_len_kernel(ConcatArgs a, uint64_t n, uint32_t* lens, int* err) {
  const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t total = 0;
  for (int k = 0; k < a.nparts; ++k) {
    const ConcatPart& p = a.parts[k];
    total += p.codes ? p.ent_len[read_code(p.codes, p.phys, i)] : p.lit_len;
  }
  if (total > 255) atomicOr(err, MSC_DEVERR_STRLEN);
  lens[i] = total;
}

__global__ void concat_copy_kernel(ConcatArgs a, uint64_t n, const uint64_t* starts, uint8_t* out) {
  const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint8_t* dst = out + starts[i];
  for (int k = 0; k < a.nparts; ++k) {
    const ConcatPart& p = a.parts[k];
    const uint8_t* src;
    uint32_t len;
    if (p.codes) {
      const uint32_t c = read_code(p.codes, p.phys, i);
      src = p.heap + p.ent_start[c];
      len = p.ent_len[c];
    } else {
      src = a.lits + p.lit_off;
      len = p.lit_len;
    }
    for (uint32_t b = 0; b < len; ++b) dst[b] = src[b];
    dst += len;
  }
}

__global__ void pack_entries_kernel(const uint64_t* ent_start, const uint32_t* ent_len, const uint8_t* heap, uint32_t n,
                                    const uint64_t* out_starts, uint8_t* out) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint8_t* src = heap + ent_start[e];
  uint8_t* dst = out + out_starts[e];
  for (uint32_t b = 0; b < ent_len[e]; ++b) dst[b] = src[b];
}

inline unsigned grid_for(uint64_t n, int block) { return static_cast<unsigned>((n + block - 1) / block); }

// Re-home the dictionary into arrays of the given capacities (entries / heap bytes / hash slots, the last a power
// of two); unchanged capacities keep their arrays.  Needs exact host mirrors (no batches in flight).
int dict_resize(msc_ctx* ctx, msc_dict* d, uint64_t ent_cap, uint64_t heap_cap, uint64_t hcap) {
  if (ent_cap != d->ent_cap) {
    uint64_t* ns = nullptr;
    uint32_t* nl = nullptr;
    MSC_TRY(msc_alloc(ctx, ent_cap * sizeof(uint64_t), reinterpret_cast<void**>(&ns)));
    MSC_TRY(msc_alloc(ctx, ent_cap * sizeof(uint32_t), reinterpret_cast<void**>(&nl)));
    if (d->n) {
      MSC_CUDA(ctx, cudaMemcpyAsync(ns, d->ent_start, d->n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
      MSC_CUDA(ctx, cudaMemcpyAsync(nl, d->ent_len, d->n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    msc_free(ctx, d->ent_start, d->ent_cap * sizeof(uint64_t));
    msc_free(ctx, d->ent_len, d->ent_cap * sizeof(uint32_t));
    d->ent_start = ns;
    d->ent_len = nl;
    d->ent_cap = ent_cap;
  }
  if (heap_cap != d->heap_cap) {
    uint8_t* nh = nullptr;
    MSC_TRY(msc_alloc(ctx, heap_cap, reinterpret_cast<void**>(&nh)));
    if (d->nbytes) MSC_CUDA(ctx, cudaMemcpyAsync(nh, d->heap, d->nbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    msc_free(ctx, d->heap, d->heap_cap);
    d->heap = nh;
    d->heap_cap = heap_cap;
  }
  if (hcap != d->hcap) {
    if (hcap > (1ULL << 31)) return ctx->fail(MSC_ERR_ARG, "dictionary too large");
    msc_free(ctx, d->hkeys, d->hcap * sizeof(uint64_t));
    msc_free(ctx, d->hcode, d->hcap * sizeof(int32_t));
    msc_free(ctx, d->hrep, d->hcap * sizeof(uint32_t));
    MSC_TRY(msc_alloc(ctx, hcap * sizeof(uint64_t), reinterpret_cast<void**>(&d->hkeys)));
    MSC_TRY(msc_alloc(ctx, hcap * sizeof(int32_t), reinterpret_cast<void**>(&d->hcode)));
    MSC_TRY(msc_alloc(ctx, hcap * sizeof(uint32_t), reinterpret_cast<void**>(&d->hrep)));
    d->hcap = hcap;
    MSC_CUDA(ctx, cudaMemsetAsync(d->hkeys, 0, hcap * sizeof(uint64_t), ctx->stream));
    MSC_CUDA(ctx, cudaMemsetAsync(d->hrep, 0xFF, hcap * sizeof(uint32_t), ctx->stream));
    fill_i32_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(d->hcode, -1, hcap);
    ctx->stats.launches += 1;
    if (d->n) {
      dict_rehash_kernel<<<grid_for(d->n, 256), 256, 0, ctx->stream>>>(
          d->ent_start, d->ent_len, d->heap, d->n, reinterpret_cast<unsigned long long*>(d->hkeys), d->hcode, hcap, d->seed);
      ctx->stats.launches += 1;
    }
    MSC_CUDA(ctx, cudaGetLastError());
  }
  return MSC_OK;
}

uint64_t pow2_at_least(uint64_t v, uint64_t floor_) {
  uint64_t c = floor_;
  while (c < v) c *= 2;
  return c;
}

// grow entry arrays / heap / hash table so that `add_entries` new strings totalling `add_bytes` fit
int dict_reserve(msc_ctx* ctx, msc_dict* d, uint64_t add_entries, uint64_t add_bytes) {
  const uint64_t need_ent = d->n + add_entries, need_bytes = d->nbytes + add_bytes;
  const uint64_t ent_cap = need_ent > d->ent_cap ? pow2_at_least(need_ent, d->ent_cap ? d->ent_cap : 1024) : d->ent_cap;
  const uint64_t heap_cap = need_bytes > d->heap_cap ? pow2_at_least(need_bytes, d->heap_cap ? d->heap_cap : 4096) : d->heap_cap;
  const uint64_t hcap = need_ent * 2 > d->hcap ? pow2_at_least(need_ent * 2, d->hcap ? d->hcap : 1024) : d->hcap;
  return dict_resize(ctx, d, ent_cap, heap_cap, hcap);
}

// fold completed counter snapshots into the host mirrors; returns the rows / bytes still unaccounted for
void dict_poll(msc_dict* d, bool wait, uint64_t* pend_rows, uint64_t* pend_bytes) {
  while (d->ring_count > 0) {
    msc_dict::Pending& e = d->ring[d->ring_head];
    if (wait) cudaEventSynchronize(e.ev);
    else if (cudaEventQuery(e.ev) != cudaSuccess) break;
    d->n = static_cast<uint32_t>(e.host[0]);
    d->nbytes = e.host[1];
    d->ring_head = (d->ring_head + 1) % msc_dict::kPending;
    --d->ring_count;
  }
  uint64_t r = 0, b = 0;
  for (int i = 0; i < d->ring_count; ++i) {
    const msc_dict::Pending& e = d->ring[(d->ring_head + i) % msc_dict::kPending];
    r += e.rows;
    b += e.bytes;
  }
  *pend_rows = r;
  *pend_bytes = b;
}

template <class TLen>
int dict_encode_impl(msc_ctx* ctx, msc_dict* d, const uint64_t* starts, const TLen* lens, const uint8_t* bytes, uint64_t n,
                     uint64_t batch_bytes, int insert, uint32_t* codes_out, bool async) {
  if (n == 0) return MSC_OK;
  if (n > 0xFFFFFFF0ULL) return ctx->fail(MSC_ERR_ARG, "dictionary batch too large");
  uint64_t pend_rows = 0, pend_bytes = 0;
  dict_poll(d, !async, &pend_rows, &pend_bytes);
  if (!insert && d->n == 0 && pend_rows == 0) {  // nothing can match an empty dictionary
    MSC_CUDA(ctx, cudaMemsetAsync(codes_out, 0xFF, n * sizeof(uint32_t), ctx->stream));
    return MSC_OK;
  }
  if (insert) {
    // capacity must cover what the in-flight batches may still add; growing needs exact mirrors, so settle first
    const bool fits = d->n + pend_rows + n <= d->ent_cap && d->nbytes + pend_bytes + batch_bytes <= d->heap_cap &&
                      (d->n + pend_rows + n) * 2 <= d->hcap;
    if (!fits) {
      dict_poll(d, true, &pend_rows, &pend_bytes);
      // pipelined loads reserve for two further batches so that steady state never has to settle
      MSC_TRY(dict_reserve(ctx, d, async ? 3 * n : n, async ? 3 * batch_bytes : batch_bytes));
    }
  }
  DevTmp slots(ctx);
  MSC_TRY(slots.alloc(n * sizeof(uint32_t)));
  const unsigned grid = grid_for(n, 256);
  auto* hk = reinterpret_cast<unsigned long long*>(d->hkeys);
  dict_probe_kernel<TLen><<<grid, 256, 0, ctx->stream>>>(starts, lens, bytes, n, hk, d->hcode, d->hrep, d->hcap, d->seed,
                                                         insert, slots.as<uint32_t>(), ctx->d_err);
  if (insert)
    dict_assign_kernel<TLen><<<grid, 256, 0, ctx->stream>>>(starts, lens, bytes, n, slots.as<uint32_t>(), d->hcode, d->hrep,
                                                            d->ent_start, d->ent_len, d->heap, d->d_counters);
  dict_resolve_kernel<TLen><<<grid, 256, 0, ctx->stream>>>(starts, lens, bytes, n, slots.as<uint32_t>(), d->hcode,
                                                           d->ent_start, d->ent_len, d->heap, codes_out, ctx->d_err);
  ctx->stats.launches += insert ? 3 : 2;
  MSC_CUDA(ctx, cudaGetLastError());
  if (insert) {
    if (d->ring_count == msc_dict::kPending) dict_poll(d, true, &pend_rows, &pend_bytes);
    msc_dict::Pending& e = d->ring[(d->ring_head + d->ring_count) % msc_dict::kPending];
    if (!e.ev) {
      MSC_CUDA(ctx, cudaEventCreateWithFlags(&e.ev, cudaEventDisableTiming));
      MSC_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&e.host), 2 * sizeof(unsigned long long), cudaHostAllocDefault));
    }
    e.rows = n;
    e.bytes = batch_bytes;
    MSC_CUDA(ctx, cudaMemcpyAsync(e.host, d->d_counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    MSC_CUDA(ctx, cudaEventRecord(e.ev, ctx->stream));
    ++d->ring_count;
  }
  if (async) return MSC_OK;
  dict_poll(d, true, &pend_rows, &pend_bytes);
  return msc_check_device_error(ctx);
}

}  // namespace

int msc_dict_encode_u8(msc_ctx* ctx, msc_dict* d, const uint64_t* starts, const uint8_t* lens, const uint8_t* bytes, uint64_t n,
                       uint64_t batch_bytes, int insert, uint32_t* codes_out) {
  return dict_encode_impl<uint8_t>(ctx, d, starts, lens, bytes, n, batch_bytes, insert, codes_out, false);
}
int msc_dict_encode_u8_async(msc_ctx* ctx, msc_dict* d, const uint64_t* starts, const uint8_t* lens, const uint8_t* bytes, uint64_t n,
                             uint64_t batch_bytes, uint32_t* codes_out) {
  return dict_encode_impl<uint8_t>(ctx, d, starts, lens, bytes, n, batch_bytes, 1, codes_out, true);
}
int msc_dict_settle(msc_ctx* ctx, msc_dict* d) {
  uint64_t r = 0, b = 0;
  dict_poll(d, true, &r, &b);
  MSC_TRY(msc_check_device_error(ctx));
  // pipelined loads reserve for whole batches of distinct strings; give back what a low-cardinality column never used
  const uint64_t ent_cap = pow2_at_least(d->n, 1024), heap_cap = pow2_at_least(d->nbytes, 4096);
  const uint64_t hcap = pow2_at_least(static_cast<uint64_t>(d->n) * 2, 1024);
  if (d->ent_cap > 4 * ent_cap || d->hcap > 4 * hcap || d->heap_cap > 4 * heap_cap) MSC_TRY(dict_resize(ctx, d, ent_cap, heap_cap, hcap));
  return MSC_OK;
}
int msc_dict_encode_u32(msc_ctx* ctx, msc_dict* d, const uint64_t* starts, const uint32_t* lens, const uint8_t* bytes, uint64_t n,
                        uint64_t batch_bytes, int insert, uint32_t* codes_out) {
  return dict_encode_impl<uint32_t>(ctx, d, starts, lens, bytes, n, batch_bytes, insert, codes_out, false);
}

extern "C" int msc_dict_create(msc_ctx* ctx, msc_dict** out) {
  if (!ctx || !out) return MSC_ERR_ARG;
  msc_dict* d = new msc_dict();
  d->ctx = ctx;
  int rc = msc_alloc(ctx, 2 * sizeof(unsigned long long), reinterpret_cast<void**>(&d->d_counters));
  if (rc != MSC_OK) {
    delete d;
    return rc;
  }
  MSC_CUDA(ctx, cudaMemsetAsync(d->d_counters, 0, 2 * sizeof(unsigned long long), ctx->stream));
  *out = d;
  return MSC_OK;
}

extern "C" void msc_dict_free(msc_dict* d) {
  if (!d) return;
  msc_ctx* ctx = d->ctx;
  msc_free(ctx, d->ent_start, d->ent_cap * sizeof(uint64_t));
  msc_free(ctx, d->ent_len, d->ent_cap * sizeof(uint32_t));
  msc_free(ctx, d->heap, d->heap_cap);
  msc_free(ctx, d->hkeys, d->hcap * sizeof(uint64_t));
  msc_free(ctx, d->hcode, d->hcap * sizeof(int32_t));
  msc_free(ctx, d->hrep, d->hcap * sizeof(uint32_t));
  msc_free(ctx, d->d_counters, 2 * sizeof(unsigned long long));
  for (auto& e : d->ring) {
    if (e.ev) {
      cudaEventSynchronize(e.ev);
      cudaEventDestroy(e.ev);
    }
    if (e.host) cudaFreeHost(e.host);
  }
  delete d;
}

extern "C" int msc_dict_size(msc_dict* d, uint32_t* nentries, uint64_t* nbytes) {
  if (!d) return MSC_ERR_ARG;
  if (nentries) *nentries = d->n;
  if (nbytes) *nbytes = d->nbytes;
  return MSC_OK;
}

extern "C" int msc_dict_lookup(msc_ctx* ctx, msc_dict* d, const char* s, size_t len, int32_t insert, int64_t* code) {
  if (!ctx || !d || !code || (!s && len)) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  if (len > 255) {
    if (insert) return ctx->fail(MSC_ERR_STRLEN, "string longer than 255 bytes");
    *code = -1;
    return MSC_OK;
  }
  if (!insert && d->n == 0) {
    *code = -1;
    return MSC_OK;
  }
  DevTmp buf(ctx);
  MSC_TRY(buf.alloc(512));
  // layout: [0,8) start=16 | [8,12) len | [12,16) code out | [16,..) bytes
  unsigned char host[512];
  memset(host, 0, sizeof(host));
  const uint64_t start = 16;
  const uint32_t l32 = static_cast<uint32_t>(len);
  memcpy(host, &start, 8);
  memcpy(host + 8, &l32, 4);
  if (len) memcpy(host + 16, s, len);
  MSC_CUDA(ctx, cudaMemcpyAsync(buf.p, host, 16 + len, cudaMemcpyHostToDevice, ctx->stream));
  auto* base = buf.as<uint8_t>();
  MSC_TRY(msc_dict_encode_u32(ctx, d, reinterpret_cast<const uint64_t*>(base), reinterpret_cast<const uint32_t*>(base + 8), base, 1,
                              len, insert, reinterpret_cast<uint32_t*>(base + 12)));
  uint32_t c = 0;
  MSC_TRY(msc_memcpy_d2h(ctx, &c, base + 12, 4));
  *code = (c == NO_CODE) ? -1 : static_cast<int64_t>(c);
  return MSC_OK;
}

extern "C" int msc_dict_like(msc_ctx* ctx, msc_dict* d, const char* pattern, size_t len, void** lut_dev) {
  if (!ctx || !d || !lut_dev || len > 4096) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  void* lut = nullptr;
  MSC_TRY(msc_alloc(ctx, static_cast<size_t>(d->n) + 16, &lut));
  DevTmp pat(ctx);
  MSC_TRY(pat.alloc(len + 16));
  if (len) MSC_CUDA(ctx, cudaMemcpyAsync(pat.p, pattern, len, cudaMemcpyHostToDevice, ctx->stream));
  if (d->n) {
    dict_like_kernel<<<grid_for(d->n, 128), 128, 0, ctx->stream>>>(d->ent_start, d->ent_len, d->heap, d->n, pat.as<uint8_t>(),
                                                                  static_cast<int>(len), static_cast<uint8_t*>(lut));
    ctx->stats.launches += 1;
    MSC_CUDA(ctx, cudaGetLastError());
  }
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `pattern` and pat are released on return
  *lut_dev = lut;
  return MSC_OK;
}

extern "C" int msc_dict_translate(msc_ctx* ctx, msc_dict* from, msc_dict* to, int32_t insert, void** lut_dev) {
  if (!ctx || !from || !to || !lut_dev) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  void* lut = nullptr;
  MSC_TRY(msc_alloc(ctx, (static_cast<size_t>(from->n) + 4) * sizeof(uint32_t), &lut));
  if (from->n) {
    if (from == to) return ctx->fail(MSC_ERR_ARG, "translate: dictionaries must differ");
    if (!insert && to->n == 0) {
      MSC_CUDA(ctx, cudaMemsetAsync(lut, 0xFF, static_cast<size_t>(from->n) * sizeof(uint32_t), ctx->stream));
    } else {
      int rc = msc_dict_encode_u32(ctx, to, from->ent_start, from->ent_len, from->heap, from->n, from->nbytes, insert,
                                   static_cast<uint32_t*>(lut));
      if (rc != MSC_OK) {
        msc_free(ctx, lut, 0);
        return rc;
      }
    }
  }
  *lut_dev = lut;
  return MSC_OK;
}

extern "C" int msc_dict_load(msc_ctx* ctx, msc_dict* d, const uint32_t* lens, const uint8_t* bytes, uint32_t nentries) {
  if (!ctx || !d || (nentries && (!lens || !bytes))) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  MSC_TRY(msc_dict_settle(ctx, d));
  if (d->n != 0) return ctx->fail(MSC_ERR_ARG, "msc_dict_load needs an empty dictionary");
  if (nentries == 0) return MSC_OK;
  std::vector<uint64_t> starts(nentries);
  uint64_t total = 0;
  for (uint32_t i = 0; i < nentries; ++i) {
    if (lens[i] > 255) return ctx->fail(MSC_ERR_STRLEN, "string longer than 255 bytes");
    starts[i] = total;
    total += lens[i];
  }
  // arrays first (n == 0: nothing to carry over or rehash), then the entries, then one rehash that assigns code i to entry i
  MSC_TRY(dict_resize(ctx, d, pow2_at_least(nentries, 1024), pow2_at_least(total, 4096), d->hcap));
  MSC_CUDA(ctx, cudaMemcpyAsync(d->ent_start, starts.data(), nentries * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  MSC_CUDA(ctx, cudaMemcpyAsync(d->ent_len, lens, nentries * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (total) MSC_CUDA(ctx, cudaMemcpyAsync(d->heap, bytes, total, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned long long counters[2] = {nentries, total};
  MSC_CUDA(ctx, cudaMemcpyAsync(d->d_counters, counters, sizeof(counters), cudaMemcpyHostToDevice, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host buffers above may go away after return
  d->n = nentries;
  d->nbytes = total;
  const uint64_t hcap = pow2_at_least(static_cast<uint64_t>(nentries) * 2, 1024);
  if (hcap == d->hcap) {  // force the rebuild: dict_resize only rehashes when the slot count changes
    msc_free(ctx, d->hkeys, d->hcap * sizeof(uint64_t));
    msc_free(ctx, d->hcode, d->hcap * sizeof(int32_t));
    msc_free(ctx, d->hrep, d->hcap * sizeof(uint32_t));
    d->hkeys = nullptr;
    d->hcode = nullptr;
    d->hrep = nullptr;
    d->hcap = 0;
  }
  MSC_TRY(dict_resize(ctx, d, d->ent_cap, d->heap_cap, hcap));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MSC_OK;
}

extern "C" int msc_dict_export(msc_ctx* ctx, msc_dict* d, uint32_t* lens, uint8_t* bytes) {
  if (!ctx || !d) return MSC_ERR_ARG;
  if (d->n == 0) return MSC_OK;
  DevTmp starts(ctx), packed(ctx);
  MSC_TRY(starts.alloc((static_cast<size_t>(d->n) + 1) * sizeof(uint64_t)));
  MSC_TRY(packed.alloc(d->nbytes + 16));
  MSC_TRY(msc_exclusive_scan_u32_u64(ctx, d->ent_len, starts.as<uint64_t>(), d->n));
  pack_entries_kernel<<<grid_for(d->n, 128), 128, 0, ctx->stream>>>(d->ent_start, d->ent_len, d->heap, d->n,
                                                                   starts.as<uint64_t>(), packed.as<uint8_t>());
  ctx->stats.launches += 1;
  MSC_CUDA(ctx, cudaGetLastError());
  if (lens) MSC_TRY(msc_memcpy_d2h(ctx, lens, d->ent_len, static_cast<size_t>(d->n) * sizeof(uint32_t)));
  if (bytes) MSC_TRY(msc_memcpy_d2h(ctx, bytes, packed.p, d->nbytes));
  return MSC_OK;
}

extern "C" int msc_str_concat(msc_ctx* ctx, const msc_concat_part* parts, int32_t nparts, uint64_t nrows, msc_dict* out_dict,
                              msc_rel** out) {
  if (!ctx || !parts || nparts < 1 || nparts > 8 || !out_dict || !out) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  ConcatArgs a;
  memset(&a, 0, sizeof(a));
  a.nparts = nparts;
  std::vector<uint8_t> lits;
  for (int k = 0; k < nparts; ++k) {
    ConcatPart& p = a.parts[k];
    if (parts[k].codes.data) {
      if (!parts[k].dict) return ctx->fail(MSC_ERR_ARG, "concat: code column without dictionary");
      p.codes = parts[k].codes.data;
      p.phys = parts[k].codes.phys;
      p.ent_start = parts[k].dict->ent_start;
      p.ent_len = parts[k].dict->ent_len;
      p.heap = parts[k].dict->heap;
    } else {
      p.lit_off = static_cast<uint32_t>(lits.size());
      p.lit_len = static_cast<uint32_t>(parts[k].literal_len);
      lits.insert(lits.end(), parts[k].literal, parts[k].literal + parts[k].literal_len);
    }
  }
  msc_rel* rel = new msc_rel();
  rel->ctx = ctx;
  rel->nrows = nrows;
  msc_col col;
  col.phys = MSC_P_U32;
  int rc = msc_alloc_rows(ctx, nrows, sizeof(uint32_t), &col.data, &col.bytes);
  if (rc != MSC_OK) {
    delete rel;
    return rc;
  }
  rel->cols.push_back(col);
  if (nrows == 0) {
    *out = rel;
    return MSC_OK;
  }
  DevTmp d_lits(ctx), lens(ctx), starts(ctx), bytes(ctx);
  auto fail = [&](int code) {
    msc_rel_free(rel);
    return code;
  };
  if ((rc = d_lits.alloc(lits.size() + 16)) != MSC_OK) return fail(rc);
  if (!lits.empty()) {
    cudaError_t e = cudaMemcpyAsync(d_lits.p, lits.data(), lits.size(), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) return fail(ctx->fail(MSC_ERR_CUDA, cudaGetErrorString(e)));
  }
  a.lits = d_lits.as<uint8_t>();
  if ((rc = lens.alloc(nrows * sizeof(uint32_t))) != MSC_OK) return fail(rc);
  if ((rc = starts.alloc((nrows + 1) * sizeof(uint64_t))) != MSC_OK) return fail(rc);
  const unsigned grid = grid_for(nrows, 256);
  concat_len_kernel<<<grid, 256, 0, ctx->stream>>>(a, nrows, lens.as<uint32_t>(), ctx->d_err);
  ctx->stats.launches += 1;
  if ((rc = msc_exclusive_scan_u32_u64(ctx, lens.as<uint32_t>(), starts.as<uint64_t>(), nrows)) != MSC_OK) return fail(rc);
  uint64_t total = 0;
  if ((rc = msc_memcpy_d2h(ctx, &total, starts.as<uint64_t>() + nrows, sizeof(total))) != MSC_OK) return fail(rc);
  if ((rc = msc_check_device_error(ctx)) != MSC_OK) return fail(rc);
  if ((rc = bytes.alloc(total + 16)) != MSC_OK) return fail(rc);
  concat_copy_kernel<<<grid, 256, 0, ctx->stream>>>(a, nrows, starts.as<uint64_t>(), bytes.as<uint8_t>());
  ctx->stats.launches += 1;
  // the literals must outlive the kernels: synchronise before `lits` goes out of scope
  if ((rc = msc_dict_encode_u32(ctx, out_dict, starts.as<uint64_t>(), lens.as<uint32_t>(), bytes.as<uint8_t>(), nrows, total, 1,
                                static_cast<uint32_t*>(col.data))) != MSC_OK)
    return fail(rc);
  *out = rel;
  return MSC_OK;
}
