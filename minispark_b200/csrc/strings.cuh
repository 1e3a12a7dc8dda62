// strings.cuh -- device string dictionary shared by ingest.cu, strings.cu and result.cu.
#pragma once
#include "common.cuh"

// One dictionary per STRING column.  Entries live on the device as (start, len) into a byte
// heap; a device open-addressing table maps 64-bit string hashes to entry codes.  Codes are dense
// 0..n-1 in insertion order, so a low-cardinality column's codes double as dense group ids.
struct msc_dict {
  msc_ctx* ctx = nullptr;
  uint32_t n = 0;         // entries (host mirror)
  uint64_t nbytes = 0;    // heap bytes used (host mirror)
  uint64_t* ent_start = nullptr;
  uint32_t* ent_len = nullptr;
  uint64_t ent_cap = 0;
  uint8_t* heap = nullptr;
  uint64_t heap_cap = 0;
  uint64_t* hkeys = nullptr;  // 0 = empty
  int32_t* hcode = nullptr;   // -1 = not assigned yet
  uint32_t* hrep = nullptr;   // smallest batch row that hit an unassigned slot
  uint64_t hcap = 0;          // power of two
  unsigned long long* d_counters = nullptr;  // [0] = n, [1] = nbytes
  uint64_t seed = 0x9E3779B97F4A7C15ULL;
  // Pipelined inserts (msc_dict_encode_u8_async): the host mirrors n / nbytes lag behind the device counters.
  // Every batch snapshots the counters into a pinned slot behind an event; a later batch folds the completed
  // snapshots in and bounds the rest by the rows / bytes still in flight, so the host only blocks when that
  // upper bound no longer fits the reserved capacity.
  struct Pending {
    cudaEvent_t ev = nullptr;
    unsigned long long* host = nullptr;  // pinned [2]
    uint64_t rows = 0, bytes = 0;
  };
  static constexpr int kPending = 8;
  Pending ring[kPending];
  int ring_head = 0, ring_count = 0;  // oldest entry, entries in flight
};

// Encode a batch of n strings given as starts[i] / lens[i] into `bytes`; codes_out[i] receives the
// dictionary code (0xFFFFFFFF when insert == 0 and the string is absent).
int msc_dict_encode_u8(msc_ctx* ctx, msc_dict* d, const uint64_t* starts, const uint8_t* lens, const uint8_t* bytes,
                       uint64_t n, uint64_t batch_bytes, int insert, uint32_t* codes_out);
// Pipelined insert-encode: no host synchronisation unless capacity has to grow; call msc_dict_settle before using
// d->n / d->nbytes or reading the device error word.
int msc_dict_encode_u8_async(msc_ctx* ctx, msc_dict* d, const uint64_t* starts, const uint8_t* lens, const uint8_t* bytes,
                             uint64_t n, uint64_t batch_bytes, uint32_t* codes_out);
int msc_dict_settle(msc_ctx* ctx, msc_dict* d);
int msc_dict_encode_u32(msc_ctx* ctx, msc_dict* d, const uint64_t* starts, const uint32_t* lens, const uint8_t* bytes,
                        uint64_t n, uint64_t batch_bytes, int insert, uint32_t* codes_out);
