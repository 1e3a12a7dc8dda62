// common.cuh -- context, device buffers and small device helpers shared by all translation units
// of libminispark_cuda.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/minispark_cuda.h"

// Rows of padding every column allocation is rounded up to: the scan kernel bulk-copies whole
// tiles (<= 4096 rows), so reading past nrows inside the padding is always in-bounds.
static constexpr uint64_t MSC_ROW_PAD = 8192;

struct msc_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;      // compute
  cudaStream_t copy[2] = {nullptr, nullptr};
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  cudaEvent_t ev_s0 = nullptr, ev_s1 = nullptr;  // bracket the fused scan kernel alone
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;  // msc_timer_start / msc_timer_stop
  int* d_err = nullptr;               // device error word (bit flags written by kernels)
  int* h_err = nullptr;               // pinned mirror
  unsigned long long* h_scratch = nullptr;  // pinned, 16 words: small result read-backs
  uint32_t* d_ticket = nullptr;       // "CTAs done" counter of fused scan + finish kernels (zero between launches)
  std::string err;
  msc_stats stats{};
  // pinned staging ring for file ingest
  static constexpr int kRing = 4;
  void* ring[kRing] = {nullptr, nullptr, nullptr, nullptr};
  size_t ring_bytes = 0;
  cudaEvent_t ring_ev[kRing] = {nullptr, nullptr, nullptr, nullptr};

  // Large-block cache in front of the driver's pool: a query (and every re-ingest of a table) asks for
  // the same multi-hundred-MB column sizes again and again, and growing the driver pool by gigabytes
  // costs 50 ms to > 1 s.  Freed blocks >= kBigBlock wait here, keyed by size; everything is ordered
  // on `stream`, so a cached block may be handed out again immediately.
  static constexpr size_t kBigBlock = 1 << 20;
  // Run index of a table's key column (scan.cu, hash aggregate without a hint): how many runs of equal adjacent values it has,
  // whether it ascends, and the number of runs before every warp tile -- what a GROUP BY on that column needs before it can
  // stream over the runs.  Derived from immutable table data only (msc_scan_desc.table_columns) and dropped when the memory it
  // describes is freed (msc_free), like the dictionaries of string columns it is computed once per table, not once per query.
  struct RunIndex {
    uint64_t nrows = 0, runs = 0, ntiles = 0;
    bool sorted = false;
    void* offsets = nullptr;  // u64[ntiles + 1]
    size_t offsets_bytes = 0;
  };
  std::map<const void*, RunIndex> run_index;
  std::multimap<size_t, void*> big_free;
  std::unordered_map<void*, size_t> big_live;
  size_t big_free_bytes = 0;
  size_t big_free_cap = 0;  // set from the device's memory size in msc_create

  int fail(int code, const std::string& msg) {
    err = msg;
    return code;
  }
};

#define MSC_CUDA(ctx, expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      char _buf[512];                                                                        \
      snprintf(_buf, sizeof(_buf), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
               cudaGetErrorString(_e));                                                      \
      return (ctx)->fail(MSC_ERR_CUDA, _buf);                                                \
    }                                                                                        \
  } while (0)

#define MSC_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != MSC_OK) return _rc; \
  } while (0)

static inline size_t msc_phys_width(int phys) {
  switch (phys) {
    case MSC_P_U8: return 1;
    case MSC_P_U16: return 2;
    case MSC_P_U32: case MSC_P_I32: case MSC_P_F32: return 4;
    case MSC_P_I64: case MSC_P_F64: return 8;
    default: return 0;
  }
}

static inline uint64_t msc_round_up(uint64_t v, uint64_t m) { return (v + m - 1) / m * m; }

// Stream-ordered device allocation on the ctx compute stream (pool keeps memory across calls).
int msc_alloc(msc_ctx* ctx, size_t nbytes, void** out);
int msc_free(msc_ctx* ctx, void* p, size_t nbytes);
// Allocation sized for `nrows` elements of `width` bytes plus tile padding.
int msc_alloc_rows(msc_ctx* ctx, uint64_t nrows, size_t width, void** out, size_t* bytes_out);
// Read + clear the device error word; returns an MSC_ERR_* or MSC_OK.
int msc_check_device_error(msc_ctx* ctx);
// MSC_ERR_* (with the message set) for a device error word a kernel already fetched and cleared.
int msc_device_error_rc(msc_ctx* ctx, int e);

struct msc_col {
  void* data = nullptr;
  int phys = 0;
  size_t bytes = 0;
  bool owned = true;
};

struct msc_rel {
  msc_ctx* ctx = nullptr;
  uint64_t nrows = 0;  // while `pending`: the upper bound the columns were allocated for
  std::vector<msc_col> cols;
  // pending relation: d_meta (device, owned) = {row count, 1 if a SUM_F came out non-finite, device error word}
  unsigned long long* d_meta = nullptr;
  bool pending = false;
};

// RAII temporary device buffer (freed stream-ordered on scope exit).
struct DevTmp {
  msc_ctx* ctx;
  void* p = nullptr;
  size_t n = 0;
  explicit DevTmp(msc_ctx* c) : ctx(c) {}
  int alloc(size_t bytes) {
    n = bytes ? bytes : 16;
    return msc_alloc(ctx, n, &p);
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
  void release() {
    if (p) msc_free(ctx, p, n);
    p = nullptr;
    n = 0;
  }
  ~DevTmp() {
    if (p) msc_free(ctx, p, n);
  }
  DevTmp(const DevTmp&) = delete;
  DevTmp& operator=(const DevTmp&) = delete;
};

// device error flag bits
#define MSC_DEVERR_DIV_ZERO 1
#define MSC_DEVERR_OVERFLOW 2
#define MSC_DEVERR_COLLISION 4
#define MSC_DEVERR_STRLEN 8
#define MSC_DEVERR_TABLE_FULL 16
#define MSC_DEVERR_PEER_TIMEOUT 32
#define MSC_DEVERR_IO 64

// ---- prefix sums (scan.cu exports these for the other translation units) -------------------
// out[i] = sum_{j<i} in[j] for i in [0, n]; out has n+1 elements (out[n] = total).
int msc_exclusive_scan_u8_u64(msc_ctx* ctx, const uint8_t* in, uint64_t* out, uint64_t n);
int msc_exclusive_scan_u32_u64(msc_ctx* ctx, const uint32_t* in, uint64_t* out, uint64_t n);

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t msc_mix64(uint64_t k) {  // murmur3 fmix64
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}
#endif
