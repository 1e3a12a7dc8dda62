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
//   slot_bytes == 16: MscJoinSlot, any 64-bit key
//   slot_bytes == 8 : u64 = build row << 32 | (u32)key for keys that are sign-extended 32-bit values (INTEGER columns,
//                     dictionary codes); 0xFFFFFFFFFFFFFFFF = empty.  Half the bytes per slot: the table of a few million
//                     build rows stays in L2, and a probe is one 8-byte read.
//   bitmap_bits != 0: between the header and the slots lies a bitmap of that many bits (a power of two <= 2^32, ~16 per build
//                     key) with one bit set per build key: a probe whose bit is clear has no partner and need not walk the
//                     slots (most probes of a filtered build side end there).
// Hashes: 16-byte slots: h = mix64(key), slot = h & (cap - 1), bit = (h >> 32) & (bitmap_bits - 1);
//         8-byte slots:  slot = fmix32(k32) & (cap - 1), bit = fmix32(k32 ^ 0x9e3779b9) & (bitmap_bits - 1)  (murmur3 finalisers)
struct __align__(16) MscJoinTableHeader {
  unsigned long long cap;
  unsigned long long duplicates;  // != 0: some key has more than one build row
  unsigned long long slot_bytes;
  unsigned long long wide_keys;   // != 0 (compact build only): a key does not fit 32 bits -- rebuild with 16-byte slots
  unsigned long long bitmap_bits;
  unsigned long long _pad[3];
};
static_assert(sizeof(MscJoinTableHeader) == 64, "join table header is 64 bytes (jit_prelude.inc reads it by offset)");
constexpr unsigned long long MSC_J_EMPTY8 = 0xFFFFFFFFFFFFFFFFULL;

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long msc_join_norm_key(long long k) {
  return (static_cast<unsigned long long>(k) == MSC_J_EMPTY) ? 0ULL : static_cast<unsigned long long>(k);
}
__device__ __forceinline__ bool msc_join_key_is_32bit(long long k) { return k == static_cast<long long>(static_cast<int>(k)); }
__device__ __forceinline__ uint32_t msc_fmix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
// the hashed 64-bit pattern of a key: the compact format hashes the sign-extended value, the wide one the normalised key
__device__ __forceinline__ const void* msc_join_slots(const MscJoinTableHeader* h) {
  return reinterpret_cast<const unsigned char*>(h + 1) + (h->bitmap_bits >> 3);
}
#endif
