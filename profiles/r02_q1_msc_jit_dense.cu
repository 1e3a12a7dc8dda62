#define MINCTAS 4

// jit_prelude.inc -- fixed part of the source handed to NVRTC by jit.cu (a C++ raw string literal body).
//
// The specialised dense aggregate scan ("JIT scan"): the same persistent, warp-private TMA pipeline as
// scan_regvm_impl.cuh, but the per-row program is no interpreter -- jit.cu translates the query's three-address
// program (include/minispark_cuda.h) into straight-line CUDA C++ appended to this text, the way the reference's
// ThreadEngine renders templates/plan.zig per query (src/mini_spark/codegen.py).  Self-contained: no #include, so
// a compile takes ~0.1-0.2 s.  Replaces tasks.py:167-177 (filter), :270-310 (aggregate), sql.py:262-266 (expressions).
typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;

struct JitParams {  // must match struct JitParams in jit.cu
  u64 nrows;
  const u64* nrows_dev;  // when set: the row count is the device's (nrows is an upper bound)
  u32 ntiles;
  u32 _pad;
  const unsigned char* col[24];
  const void* gather[16];
  const void* luts[8];
  i64 consts[32];
  u64* dense_out;  // [ngroups][stride]
  int* err;
  u32* tile_counts;         // count scans: surviving rows of every warp tile
  const u64* tile_offsets;  // project scans with a filter: output position of every warp tile's first surviving row
  void* out[24];            // project scans: output columns (u32, or raw 64-bit); fused finish: columns of the final relation
  i64 fconsts[32];          // fused finish: constants of the final projection
  u64* fmeta;               // fused finish: {result rows, 1 if a SUM came out non-finite, device error word}
  u32* ticket;              // fused finish: CTAs done (the last one finishes; it resets the counter)
  // fused finish across ranks (NVLink peer memory): every rank's mailbox = u64 flag[2][world], u64 slot[2][world][cells]
  u64* mailbox[8];
  const int* inv;           // [world][32]: rank r's local group of merged group G, or -1
  u64 epoch;
  int rank, world, nlocal, slot_cells, nglobal, _pad3;
};

__device__ __forceinline__ void st_release_sys(u64* addr, u64 v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire_sys(const u64* addr) {
  u64 v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ u64 ld_volatile(const u64* addr) {
  u64 v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
  return v;
}

constexpr int NT = 128, NW = 4, R = 8, WT = 256;
constexpr int SMEM_HEADER = NW * 8 * 8;  // mbarriers: full[warp][stage], up to 8 stages

__device__ __forceinline__ u32 smem_u32(const void* p) { return static_cast<u32>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, u32 bytes, u64* bar) {
#ifdef MSC_STREAM_EVICT_FIRST
  // the scanned columns are read once: they must not push the join table this scan probes out of L2
  u64 policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
#else
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
#endif
}
__device__ __forceinline__ bool mbar_try_wait(u64* bar, u32 parity) {
  u32 ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ double l2d(i64 v) { return __longlong_as_double(v); }
__device__ __forceinline__ i64 d2l(double v) { return __double_as_longlong(v); }

// ---- scalar semantics that follow Python (the oracle is PythonExecutionEngine, sql.py:262-266) ----
static __device__ __noinline__ i64 py_floordiv_i(i64 a, i64 b) {
  if (b == 0) return 0;
  i64 q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}
static __device__ __noinline__ i64 py_mod_i(i64 a, i64 b) {
  if (b == 0) return 0;
  i64 m = a % b;
  if (m != 0 && ((m < 0) != (b < 0))) m += b;
  return m;
}
// CPython float_divmod (Objects/floatobject.c)
static __device__ __noinline__ void py_divmod_f(double vx, double wx, double* fd, double* md) {
  if (wx == 0.0) {
    *fd = 0.0;
    *md = 0.0;
    return;
  }
  double mod = fmod(vx, wx);
  double div = (vx - mod) / wx;
  if (mod != 0.0) {
    if ((wx < 0) != (mod < 0)) {
      mod += wx;
      div -= 1.0;
    }
  } else {
    mod = copysign(0.0, wx);
  }
  double floordiv;
  if (div != 0.0) {
    floordiv = floor(div);
    if (div - floordiv > 0.5) floordiv += 1.0;
  } else {
    floordiv = copysign(0.0, vx / wx);
  }
  *fd = floordiv;
  *md = mod;
}
__device__ __forceinline__ i64 py_floordiv_f(i64 a, i64 b) {
  double fd, md;
  py_divmod_f(l2d(a), l2d(b), &fd, &md);
  return d2l(fd);
}
__device__ __forceinline__ i64 py_mod_f(i64 a, i64 b) {
  double fd, md;
  py_divmod_f(l2d(a), l2d(b), &fd, &md);
  return d2l(md);
}

// ---- a lane's 8 rows of a staged column: rows 4*lane..4*lane+3 of each 128-row half of the warp tile, as raw 64-bit
// values (integers sign-/zero-extended, f32 widened to f64: lossless) ----
__device__ __forceinline__ void ld_f32(const unsigned char* col, int lane, i64 (&v)[R]) {
  const float4 a = *reinterpret_cast<const float4*>(col + lane * 16), b = *reinterpret_cast<const float4*>(col + 512 + lane * 16);
  v[0] = d2l((double)a.x); v[1] = d2l((double)a.y); v[2] = d2l((double)a.z); v[3] = d2l((double)a.w);
  v[4] = d2l((double)b.x); v[5] = d2l((double)b.y); v[6] = d2l((double)b.z); v[7] = d2l((double)b.w);
}
__device__ __forceinline__ void ld_i32(const unsigned char* col, int lane, i64 (&v)[R]) {
  const int4 a = *reinterpret_cast<const int4*>(col + lane * 16), b = *reinterpret_cast<const int4*>(col + 512 + lane * 16);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld_u32(const unsigned char* col, int lane, i64 (&v)[R]) {
  const uint4 a = *reinterpret_cast<const uint4*>(col + lane * 16), b = *reinterpret_cast<const uint4*>(col + 512 + lane * 16);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld_64(const unsigned char* col, int lane, i64 (&v)[R]) {
  const longlong2* lo = reinterpret_cast<const longlong2*>(col + lane * 32);
  const longlong2* hi = reinterpret_cast<const longlong2*>(col + 1024 + lane * 32);
  const longlong2 a = lo[0], b = lo[1], c = hi[0], d = hi[1];
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void ld_u16(const unsigned char* col, int lane, i64 (&v)[R]) {
  const uint2 a = *reinterpret_cast<const uint2*>(col + lane * 8), b = *reinterpret_cast<const uint2*>(col + 256 + lane * 8);
  v[0] = a.x & 0xffffu; v[1] = a.x >> 16; v[2] = a.y & 0xffffu; v[3] = a.y >> 16;
  v[4] = b.x & 0xffffu; v[5] = b.x >> 16; v[6] = b.y & 0xffffu; v[7] = b.y >> 16;
}
__device__ __forceinline__ void ld_u8(const unsigned char* col, int lane, i64 (&v)[R]) {
  const u32 a = *reinterpret_cast<const u32*>(col + lane * 4), b = *reinterpret_cast<const u32*>(col + 128 + lane * 4);
  v[0] = a & 0xffu; v[1] = (a >> 8) & 0xffu; v[2] = (a >> 16) & 0xffu; v[3] = a >> 24;
  v[4] = b & 0xffu; v[5] = (b >> 8) & 0xffu; v[6] = (b >> 16) & 0xffu; v[7] = b >> 24;
}

// ---- one element of a staged column by its row index inside the warp tile (the compacted tail of a probing scan) ----
template <int PHYS>
__device__ __forceinline__ i64 ldrow(const unsigned char* col, u32 row) {
  if (PHYS == 0) return col[row];
  if (PHYS == 1) return reinterpret_cast<const unsigned short*>(col)[row];
  if (PHYS == 2) return reinterpret_cast<const u32*>(col)[row];
  if (PHYS == 3) return reinterpret_cast<const int*>(col)[row];
  if (PHYS == 5) return d2l((double)reinterpret_cast<const float*>(col)[row]);
  return reinterpret_cast<const i64*>(col)[row];
}
// exclusive prefix sum of one u32 per lane across the warp; the warp's total in *total
__device__ __forceinline__ u32 warp_exclusive_scan(u32 v, int lane, u32* total) {
  u32 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const u32 n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  *total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}

// ---- gathered column element (join outputs read through an index vector) ----
template <int PHYS>
__device__ __forceinline__ i64 gather_at(const void* col, i64 index, bool valid) {
  if (!valid) return 0;
  const u32 i = static_cast<u32>(index);
  if (PHYS == 0) return __ldg(reinterpret_cast<const unsigned char*>(col) + i);
  if (PHYS == 1) return __ldg(reinterpret_cast<const unsigned short*>(col) + i);
  if (PHYS == 2) return __ldg(reinterpret_cast<const u32*>(col) + i);
  if (PHYS == 3) return __ldg(reinterpret_cast<const int*>(col) + i);
  if (PHYS == 5) return d2l((double)__ldg(reinterpret_cast<const float*>(col) + i));
  return __ldg(reinterpret_cast<const i64*>(col) + i);
}

// ---- MSC_OP_PROBE: build-side row of `key` in a join table (csrc/join_table.cuh), or -1.  Header (64 bytes): {cap,
// duplicates, slot_bytes, wide_keys, bitmap_bits}; then a bitmap with one bit set per build key; then 16-byte slots {key,
// head, len}, or -- COMPACT: slot_bytes == 8, keys that are sign-extended 32-bit values -- u64 slots = row << 32 | (u32)key
// with all ones for "empty".  The program names the format (MSC_PROBE_COMPACT in the instruction), so only one is compiled.
// The probe comes in two halves: `issue` computes the position and reads the key's bitmap word and its first slot, `resolve`
// looks at them and walks on if it must; the dense aggregate scan issues the reads of all of a lane's rows before it
// resolves any.  A clear bit ends a probe without a walk.
template <bool COMPACT>
struct JoinProbe;
template <>
struct JoinProbe<true> {
  u32 key, pos, lo, hi, present;
};
template <>
struct JoinProbe<false> {
  u64 key, pos;
  uint4 raw;
  u32 present;
};
__device__ __forceinline__ u64 mix64(u64 h) {
  h ^= h >> 33; h *= 0xff51afd7ed558ccdull; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull; h ^= h >> 33;
  return h;
}
__device__ __forceinline__ u32 fmix32(u32 h) {
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
// table reads ask L2 to keep their lines (the streamed columns go through with evict-first, see bulk_g2s)
__device__ __forceinline__ u64 l2_keep_policy() {
  u64 policy;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}
__device__ __forceinline__ u32 ldg_keep_u32(const void* ptr, u64 policy) {
  u32 v;
  asm("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(policy));
  return v;
}
__device__ __forceinline__ uint2 ldg_keep_v2(const void* ptr, u64 policy) {
  uint2 v;
  asm("ld.global.nc.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(ptr), "l"(policy));
  return v;
}
__device__ __forceinline__ uint4 ldg_keep_v4(const void* ptr, u64 policy) {
  uint4 v;
  asm("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr), "l"(policy));
  return v;
}
__device__ __forceinline__ void join_probe_issue(const void* table, i64 key, bool& valid, JoinProbe<true>& q, u64 policy) {
  const u64* h = reinterpret_cast<const u64*>(table);
  const u32 mask = static_cast<u32>(h[0]) - 1u, bmask = static_cast<u32>(h[4]) - 1u;  // (cap, bitmap_bits <= 2^32)
  const unsigned char* bitmap = reinterpret_cast<const unsigned char*>(table) + 64;
  const unsigned char* slots = bitmap + (h[4] >> 3);
  q.lo = q.hi = 0xffffffffu;
  q.present = 0;
  valid = valid && key == static_cast<i64>(static_cast<int>(key));  // outside the compact table's key range: no match
  q.key = static_cast<u32>(key);
  q.pos = fmix32(q.key) & mask;
  if (valid) {
    const u32 bit = fmix32(q.key ^ 0x9e3779b9u) & bmask;
    q.present = (ldg_keep_u32(bitmap + 4 * (bit >> 5), policy) >> (bit & 31)) & 1u;
    const uint2 v = ldg_keep_v2(slots + 8ull * q.pos, policy);
    q.lo = v.x;
    q.hi = v.y;
  }
}
__device__ __forceinline__ i64 join_probe_resolve(const void* table, JoinProbe<true> q, bool valid, u64 policy) {
  if (!(valid && q.present)) return -1;
  const u64* h = reinterpret_cast<const u64*>(table);
  const u32 mask = static_cast<u32>(h[0]) - 1u;
  const unsigned char* slots = reinterpret_cast<const unsigned char*>(table) + 64 + (h[4] >> 3);
  while (true) {
    if ((q.lo & q.hi) == 0xffffffffu) return -1;
    if (q.lo == q.key) return static_cast<i64>(q.hi);
    q.pos = (q.pos + 1) & mask;
    const uint2 v = ldg_keep_v2(slots + 8ull * q.pos, policy);
    q.lo = v.x;
    q.hi = v.y;
  }
}
__device__ __forceinline__ void join_probe_issue(const void* table, i64 key, bool& valid, JoinProbe<false>& q, u64 policy) {
  const u64* h = reinterpret_cast<const u64*>(table);
  const u64 mask = h[0] - 1, bits = h[4];
  const unsigned char* bitmap = reinterpret_cast<const unsigned char*>(table) + 64;
  const unsigned char* slots = bitmap + (bits >> 3);
  q.raw = make_uint4(0u, 0x80000000u, 0u, 0u);
  q.present = 0;
  q.key = (static_cast<u64>(key) == 0x8000000000000000ull) ? 0ull : static_cast<u64>(key);
  const u64 hash = mix64(q.key);
  q.pos = hash & mask;
  if (valid) {
    const u64 bit = (hash >> 32) & (bits - 1);
    q.present = (ldg_keep_u32(bitmap + 4 * (bit >> 5), policy) >> (bit & 31)) & 1u;
    q.raw = ldg_keep_v4(slots + 16ull * q.pos, policy);
  }
}
__device__ __forceinline__ i64 join_probe_resolve(const void* table, JoinProbe<false> q, bool valid, u64 policy) {
  if (!(valid && q.present)) return -1;
  const u64* h = reinterpret_cast<const u64*>(table);
  const u64 mask = h[0] - 1;
  const unsigned char* slots = reinterpret_cast<const unsigned char*>(table) + 64 + (h[4] >> 3);
  while (true) {
    const u64 cur = (static_cast<u64>(q.raw.y) << 32) | q.raw.x;
    if (cur == q.key) return static_cast<i64>(q.raw.z);
    if (cur == 0x8000000000000000ull) return -1;
    q.pos = (q.pos + 1) & mask;
    q.raw = ldg_keep_v4(slots + 16ull * q.pos, policy);
  }
}
template <bool COMPACT>
__device__ __forceinline__ void join_probe_issue(const void* table, i64 key, bool& valid, JoinProbe<COMPACT>& q, u64 policy);
template <bool COMPACT>
__device__ __forceinline__ i64 join_probe(const void* table, i64 key, bool valid, u64 policy) {
  JoinProbe<COMPACT> q;
  join_probe_issue(table, key, valid, q, policy);
  return join_probe_resolve(table, q, valid, policy);
}

// ---- accumulator kinds (MSC_AGG_*): 0 SUM_F 1 SUM_I 2 MIN_F 3 MAX_F 4 MIN_I 5 MAX_I ----
template <int KIND>
__device__ __forceinline__ i64 agg_combine(i64 cur, i64 v) {
  if (KIND == 0) return d2l(l2d(cur) + l2d(v));
  if (KIND == 1) return cur + v;
  if (KIND == 2) return (l2d(v) < l2d(cur)) ? v : cur;
  if (KIND == 3) return (l2d(v) > l2d(cur)) ? v : cur;
  if (KIND == 4) return (v < cur) ? v : cur;
  return (v > cur) ? v : cur;
}
__device__ __forceinline__ i64 agg_combine_k(int kind, i64 cur, i64 v) {
  switch (kind) {
    case 0: return agg_combine<0>(cur, v);
    case 1: return agg_combine<1>(cur, v);
    case 2: return agg_combine<2>(cur, v);
    case 3: return agg_combine<3>(cur, v);
    case 4: return agg_combine<4>(cur, v);
    default: return agg_combine<5>(cur, v);
  }
}
// ---- predicated accumulator updates: "if (gsel == G) acc += v" as ONE predicated instruction.  Written in PTX because
// nvcc turns the C++ form into an unconditional add plus two selects (3 instructions, and it parks the 24 row x group
// predicates of a tile in a bit mask); ptxas shares the setp of one row and group between its accumulators. ----
template <int G>
__device__ __forceinline__ void addf_if(double& acc, double v, int gsel) {
  asm("{\n\t.reg .pred q;\n\tsetp.eq.s32 q, %2, %3;\n\t@q add.rn.f64 %0, %0, %1;\n\t}" : "+d"(acc) : "d"(v), "r"(gsel), "n"(G));
}
template <int G>
__device__ __forceinline__ void addi_if(i64& acc, i64 v, int gsel) {
  asm("{\n\t.reg .pred q;\n\tsetp.eq.s32 q, %2, %3;\n\t@q add.s64 %0, %0, %1;\n\t}" : "+l"(acc) : "l"(v), "r"(gsel), "n"(G));
}
template <int G>
__device__ __forceinline__ void inc_if(u32& n, int gsel) {
  asm("{\n\t.reg .pred q;\n\tsetp.eq.s32 q, %1, %2;\n\t@q add.u32 %0, %0, 1;\n\t}" : "+r"(n) : "r"(gsel), "n"(G));
}

template <int KIND>
__device__ __forceinline__ i64 warp_fold(i64 v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = agg_combine<KIND>(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void atomic_fold(int kind, u64* addr, i64 v) {
  switch (kind) {
    case 0: atomicAdd(reinterpret_cast<double*>(addr), l2d(v)); break;
    case 1: atomicAdd(addr, static_cast<u64>(v)); break;
    case 4: atomicMin(reinterpret_cast<i64*>(addr), v); break;
    case 5: atomicMax(reinterpret_cast<i64*>(addr), v); break;
    default: {  // f64 min / max: CAS loop
      u64 old = *addr;
      while (true) {
        const i64 merged = agg_combine_k(kind, static_cast<i64>(old), v);
        if (static_cast<u64>(merged) == old) break;
        const u64 prev = atomicCAS(addr, old, static_cast<u64>(merged));
        if (prev == old) break;
        old = prev;
      }
    }
  }
}

// ---- streaming aggregate over the runs of a sorted key (msc_jit_runs): fold the rows of one 4-row segment that belong
// to the same run in registers, then one atomic per (run, accumulator) -- a run may continue in the next lane or tile ----
template <int KIND>
__device__ __forceinline__ void fold_segment(u64* col, const int* idx, const i64* v) {
  i64 run = v[0];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const bool same_next = r < 3 && idx[r] >= 0 && idx[r + 1] == idx[r];
    if (same_next) {
      run = agg_combine<KIND>(run, v[r + 1]);
    } else {
      if (idx[r] >= 0) atomic_fold(KIND, col + idx[r], run);
      run = v[r < 3 ? r + 1 : r];
    }
  }
}
constexpr int NG = 3, NGP = 4, STRIDE = 6, NSTAGES = 2;
constexpr u32 STAGE_BYTES = 6400, TX_BYTES = 6400, NSTAGED = 6;
__device__ const u32 COL_OFF[6] = {0, 2048, 2304, 3328, 4352, 5376};
__device__ const u32 COL_BYTES[6] = {2048, 256, 1024, 1024, 1024, 1024};
__device__ const int KIND[STRIDE] = {0, 0, 0, 0, 0, 1};
__device__ const i64 INIT[STRIDE] = {0ll, 0ll, 0ll, 0ll, 0ll, 0ll};

__device__ __forceinline__ void issue_tile(const JitParams& p, unsigned char* stages, u64* full, u32 stage, u64 tile, int lane) {
  if (lane == 0) mbar_expect_tx(&full[stage], TX_BYTES);
  __syncwarp();
  if (lane < (int)NSTAGED) bulk_g2s(stages + stage * STAGE_BYTES + COL_OFF[lane], p.col[lane] + tile * COL_BYTES[lane], COL_BYTES[lane], &full[stage]);
}

extern "C" __global__ void __launch_bounds__(NT, MINCTAS) msc_jit_dense(const __grid_constant__ JitParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u64* full = reinterpret_cast<u64*>(smem) + warp * 8;
  unsigned char* stages = smem + SMEM_HEADER + warp * (NSTAGES * STAGE_BYTES);
  i64* red = reinterpret_cast<i64*>(smem + SMEM_HEADER + NW * (NSTAGES * STAGE_BYTES));  // [NW][NG * STRIDE]
  // one-hot f64 masks: row e = 0 is "no group" (filtered out / past the end / code out of range), row g + 1 selects group g
  double* mlut = reinterpret_cast<double*>(red + NW * NG * STRIDE);  // [NG + 1][NGP]
  u32* queue = reinterpret_cast<u32*>(&mlut[(NG + 1) * NGP]) + warp * (2 * WT);  // (probing scans only: survivors of a tile)
  if (lane == 0) {
    for (u32 st = 0; st < NSTAGES; ++st) mbar_init(&full[st], 1);
    mbar_fence_init();
  }
  for (int i = tid; i < (NG + 1) * NGP; i += NT) mlut[i] = (i / NGP >= 1 && i % NGP == i / NGP - 1) ? 1.0 : 0.0;
  __syncthreads();
  const u32 gw = blockIdx.x * NW + warp, nw = gridDim.x * NW;
  const u32 ntiles_w = (p.ntiles > gw) ? (p.ntiles - gw + nw - 1) / nw : 0;
  {
    const u32 pre = ntiles_w < NSTAGES ? ntiles_w : NSTAGES;
    for (u32 k = 0; k < pre; ++k) issue_tile(p, stages, full, k, gw + (u64)k * nw, lane);
  }
  const u64 nrows = p.nrows_dev ? *p.nrows_dev : p.nrows;
  const u64 keep_policy = l2_keep_policy();  // (join-table reads ask L2 to keep their lines)
  bool bad = false;
  double a0_0 = l2d(INIT[0]);
  double a0_1 = l2d(INIT[1]);
  double a0_2 = l2d(INIT[2]);
  double a0_3 = l2d(INIT[3]);
  double a0_4 = l2d(INIT[4]);
  i64 a0_5 = INIT[5]; double n0_5 = 0;
  double a1_0 = l2d(INIT[0]);
  double a1_1 = l2d(INIT[1]);
  double a1_2 = l2d(INIT[2]);
  double a1_3 = l2d(INIT[3]);
  double a1_4 = l2d(INIT[4]);
  i64 a1_5 = INIT[5]; double n1_5 = 0;
  double a2_0 = l2d(INIT[0]);
  double a2_1 = l2d(INIT[1]);
  double a2_2 = l2d(INIT[2]);
  double a2_3 = l2d(INIT[3]);
  double a2_4 = l2d(INIT[4]);
  i64 a2_5 = INIT[5]; double n2_5 = 0;
  u32 stage = 0, parity = 0;
  for (u32 k = 0; k < ntiles_w; ++k) {
    const u64 tile = gw + (u64)k * nw;
    const unsigned char* sb = stages + stage * STAGE_BYTES;
    while (!mbar_try_wait(&full[stage], parity)) {
    }
    u32 vmask = 0xffu;
    const u64 tile_row0 = tile * WT;
    if (tile_row0 + WT > nrows) {
      vmask = 0;
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (tile_row0 + (r / 4) * 128 + 4 * lane + (r % 4) < nrows) vmask |= 1u << r;
    }
    i64 c0[R]; ld_64(sb + 0, lane, c0);
    i64 c1[R]; ld_u8(sb + 2048, lane, c1);
    i64 c2[R]; ld_f32(sb + 2304, lane, c2);
    i64 c3[R]; ld_f32(sb + 3328, lane, c3);
    i64 c4[R]; ld_f32(sb + 4352, lane, c4);
    i64 c5[R]; ld_f32(sb + 5376, lane, c5);
    if (vmask != 0xffu) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (!((vmask >> r) & 1u)) c1[r] = -1;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      bool valid = true;
      int grp = -1;
      double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
      i64 t0 = 0;
      i64 t1 = 0;
      {  // instruction 0
        const i64 x = (((c0[r]) <= (p.consts[0])) ? 1ll : 0ll);
        valid = valid && (x != 0);
      }
      {  // instruction 1
        const i64 x = c1[r];
        grp = (x >= 0 && x < NG) ? (int)x : -1;
        { const double2* mrow = reinterpret_cast<const double2*>(mlut + (valid ? grp + 1 : 0) * NGP);
          { const double2 mm = mrow[0]; m0 = mm.x; m1 = mm.y; }
          { const double2 mm = mrow[1]; m2 = mm.x; m3 = mm.y; }
        }
      }
      {  // instruction 2
        const i64 x = c2[r];
        { const int gsel = valid ? grp : -1;
          a0_0 = fma(l2d(x), m0, a0_0);
          a1_0 = fma(l2d(x), m1, a1_0);
          a2_0 = fma(l2d(x), m2, a2_0);
        }
      }
      {  // instruction 3
        const i64 x = c3[r];
        { const int gsel = valid ? grp : -1;
          a0_1 = fma(l2d(x), m0, a0_1);
          a1_1 = fma(l2d(x), m1, a1_1);
          a2_1 = fma(l2d(x), m2, a2_1);
        }
      }
      {  // instruction 4
        const i64 x = c4[r];
        { const int gsel = valid ? grp : -1;
          a0_2 = fma(l2d(x), m0, a0_2);
          a1_2 = fma(l2d(x), m1, a1_2);
          a2_2 = fma(l2d(x), m2, a2_2);
        }
      }
      {  // instruction 5
        const i64 x = d2l(l2d(p.consts[1]) - l2d(c4[r]));
        t0 = x;
      }
      {  // instruction 6
        const i64 x = d2l(l2d(c3[r]) * l2d(t0));
        t1 = x;
        { const int gsel = valid ? grp : -1;
          a0_3 = fma(l2d(x), m0, a0_3);
          a1_3 = fma(l2d(x), m1, a1_3);
          a2_3 = fma(l2d(x), m2, a2_3);
        }
      }
      {  // instruction 7
        const i64 x = d2l(l2d(p.consts[1]) + l2d(c5[r]));
        t0 = x;
      }
      {  // instruction 8
        const i64 x = d2l(l2d(t1) * l2d(t0));
        { const int gsel = valid ? grp : -1;
          a0_4 = fma(l2d(x), m0, a0_4);
          a1_4 = fma(l2d(x), m1, a1_4);
          a2_4 = fma(l2d(x), m2, a2_4);
        }
      }
      {  // instruction 9
        { const int gsel = valid ? grp : -1;
          n0_5 += m0;
          n1_5 += m1;
          n2_5 += m2;
        }
      }
    }
    __syncwarp();
    if (k + NSTAGES < ntiles_w) issue_tile(p, stages, full, stage, gw + (u64)(k + NSTAGES) * nw, lane);
    if (++stage == NSTAGES) {
      stage = 0;
      parity ^= 1u;
    }
  }
  if (bad) atomicOr(p.err, 1);  // MSC_DEVERR_DIV_ZERO
  a0_5 += (i64)n0_5 * p.consts[2];
  a1_5 += (i64)n1_5 * p.consts[2];
  a2_5 += (i64)n2_5 * p.consts[2];
  { const i64 v = warp_fold<0>(d2l(a0_0)); if (lane == 0) red[warp * (NG * STRIDE) + 0] = v; }
  { const i64 v = warp_fold<0>(d2l(a0_1)); if (lane == 0) red[warp * (NG * STRIDE) + 1] = v; }
  { const i64 v = warp_fold<0>(d2l(a0_2)); if (lane == 0) red[warp * (NG * STRIDE) + 2] = v; }
  { const i64 v = warp_fold<0>(d2l(a0_3)); if (lane == 0) red[warp * (NG * STRIDE) + 3] = v; }
  { const i64 v = warp_fold<0>(d2l(a0_4)); if (lane == 0) red[warp * (NG * STRIDE) + 4] = v; }
  { const i64 v = warp_fold<1>(a0_5); if (lane == 0) red[warp * (NG * STRIDE) + 5] = v; }
  { const i64 v = warp_fold<0>(d2l(a1_0)); if (lane == 0) red[warp * (NG * STRIDE) + 6] = v; }
  { const i64 v = warp_fold<0>(d2l(a1_1)); if (lane == 0) red[warp * (NG * STRIDE) + 7] = v; }
  { const i64 v = warp_fold<0>(d2l(a1_2)); if (lane == 0) red[warp * (NG * STRIDE) + 8] = v; }
  { const i64 v = warp_fold<0>(d2l(a1_3)); if (lane == 0) red[warp * (NG * STRIDE) + 9] = v; }
  { const i64 v = warp_fold<0>(d2l(a1_4)); if (lane == 0) red[warp * (NG * STRIDE) + 10] = v; }
  { const i64 v = warp_fold<1>(a1_5); if (lane == 0) red[warp * (NG * STRIDE) + 11] = v; }
  { const i64 v = warp_fold<0>(d2l(a2_0)); if (lane == 0) red[warp * (NG * STRIDE) + 12] = v; }
  { const i64 v = warp_fold<0>(d2l(a2_1)); if (lane == 0) red[warp * (NG * STRIDE) + 13] = v; }
  { const i64 v = warp_fold<0>(d2l(a2_2)); if (lane == 0) red[warp * (NG * STRIDE) + 14] = v; }
  { const i64 v = warp_fold<0>(d2l(a2_3)); if (lane == 0) red[warp * (NG * STRIDE) + 15] = v; }
  { const i64 v = warp_fold<0>(d2l(a2_4)); if (lane == 0) red[warp * (NG * STRIDE) + 16] = v; }
  { const i64 v = warp_fold<1>(a2_5); if (lane == 0) red[warp * (NG * STRIDE) + 17] = v; }
  __syncthreads();
  for (int cell = tid; cell < NG * STRIDE; cell += NT) {
    const int kind = KIND[cell % STRIDE];
    i64 v = red[cell];
#pragma unroll
    for (int w = 1; w < NW; ++w) v = agg_combine_k(kind, v, red[w * (NG * STRIDE) + cell]);
    if (v != INIT[cell % STRIDE]) atomic_fold(kind, p.dense_out + cell, v);
  }
}
