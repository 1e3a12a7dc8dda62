// join_table.cuh -- layout of a device join hash table, shared by the build / probe kernels (join.cu) and by the fused
// scan kernels that probe it per row (scan_kernel.cuh, MSC_OP_PROBE).
#pragma once
#include <stdint.h>

constexpr unsigned long long MSC_J_EMPTY = 0x8000000000000000ULL;
constexpr uint32_t MSC_J_NIL = 0xFFFFFFFFu;

// One 16-byte slot per key: the key and the head of its chain of build rows share a sector, so an insert (CAS on the key,
// exchange on the head) and a probe (one 128-bit load) touch one random sector each instead of two.
struct __align__(16) MscJoinSlot {
  unsigned long long key;
  uint32_t head;
  uint32_t len;  // build rows with this key: a probe knows its match count without walking the chain
};

// A table handed out by msc_join_build: this header, then `cap` slots (cap a power of two).
struct __align__(16) MscJoinTableHeader {
  unsigned long long cap;
  unsigned long long duplicates;  // != 0: some key has more than one build row
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long msc_join_norm_key(long long k) {
  return (static_cast<unsigned long long>(k) == MSC_J_EMPTY) ? 0ULL : static_cast<unsigned long long>(k);
}
#endif
