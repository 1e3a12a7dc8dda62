// scan.cu -- the fused scan -> filter -> project -> aggregate kernel and its host entry points.
//
// One persistent kernel walks the relation tile by tile.  Per tile, thread 0 bulk-copies every
// directly scanned column segment global->shared with cp.async.bulk (TMA 1-D, mbarrier
// complete_tx) into a multi-stage ring, so HBM latency is hidden by the ring and not by the
// interpreter.  Every thread owns R consecutive rows and evaluates the expression program for them
// in registers: the program is a postfix stack machine whose stack depth is static per
// instruction, so `s[depth][row]` is always indexed with compile-time constants (the depth switch
// below) and lives in registers.
//
// Replaces, in one pass and without materialising intermediates (reference file:line):
//   FilterTask.execute      src/mini_spark/tasks.py:167-177   (templates/plan.zig:130-147, task_utils.zig:9-51)
//   ProjectTask.execute     src/mini_spark/tasks.py:79-84     (templates/plan.zig:113-125)
//   AggregateTask.execute   src/mini_spark/tasks.py:270-310   (templates/plan.zig:150-253)
//   Col.execute_row         src/mini_spark/sql.py:262-266
#include "common.cuh"

namespace {

constexpr int NT = 128;  // threads per CTA
constexpr int MAX_STAGES = 8;
constexpr int SMEM_HEADER = 128;  // mbarriers [0,64) + block-scan scratch [64,128)

enum Mode { MODE_DENSE = 0, MODE_HASH = 1, MODE_COUNT = 2, MODE_PROJECT = 3 };

struct StagedCol {
  const unsigned char* base;
  uint32_t width;
  uint32_t smem_off;
  int phys;
  int _pad;
};

struct ScanParams {
  uint64_t nrows;
  uint32_t ntiles;
  uint32_t nstages;
  uint32_t stage_bytes;
  uint32_t nstaged;
  uint32_t ntemps;
  StagedCol staged[MSC_VM_MAX_STAGED];
  const void* gather[MSC_VM_MAX_GATHER];
  int gather_phys[MSC_VM_MAX_GATHER];
  const void* luts[MSC_VM_MAX_LUTS];
  uint32_t code[MSC_VM_MAX_CODE];
  long long consts[MSC_VM_MAX_CONSTS];
  int* err;
  // dense aggregation: naggs includes the hidden per-group row counter (last slot)
  int ngroups;
  int naggs;
  long long agg_init[MSC_VM_MAX_AGGS + 1];
  int agg_kind[MSC_VM_MAX_AGGS + 1];
  unsigned long long* dense_out;  // [ngroups][naggs]
  // hash aggregation
  unsigned long long* hkeys;
  unsigned long long* haccs;  // [naggs][capacity]
  uint64_t hcap;              // power of two
  // count / project
  uint32_t* tile_counts;
  const uint64_t* tile_offsets;  // nullptr: no filter, output position = row
  void* out[MSC_VM_MAX_OUT];
  int out_phys[MSC_VM_MAX_OUT];
};

constexpr unsigned long long HASH_EMPTY = 0x8000000000000000ULL;

// ------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ double l2d(long long v) { return __longlong_as_double(v); }
__device__ __forceinline__ long long d2l(double v) { return __double_as_longlong(v); }

// ------------------------------------------------------------------------------------------------
// scalar op semantics that follow Python (the oracle is PythonExecutionEngine, sql.py:262-266)
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ long long py_floordiv_i(long long a, long long b) {
  if (b == 0) return 0;
  long long q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}
__device__ __noinline__ long long py_mod_i(long long a, long long b) {
  if (b == 0) return 0;
  long long m = a % b;
  if (m != 0 && ((m < 0) != (b < 0))) m += b;
  return m;
}
// CPython float_divmod (Objects/floatobject.c): floor division and modulo of doubles
__device__ __noinline__ void py_divmod_f(double vx, double wx, double* fd, double* md) {
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

template <int R>
__device__ __forceinline__ void flag_zero_divisor(const long long (&b)[R], bool is_float, uint32_t vmask, int* err) {
  bool bad = false;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const bool z = is_float ? (l2d(b[r]) == 0.0) : (b[r] == 0);
    bad |= z && ((vmask >> r) & 1u);
  }
  if (bad) atomicOr(err, MSC_DEVERR_DIV_ZERO);
}

// a = a <op> b for R rows; all indices static after inlining
template <int R>
__device__ __forceinline__ void binop(int op, long long (&a)[R], const long long (&b)[R], uint32_t vmask, int* err) {
  switch (op) {
#define F_ARITH(OP, EXPR)                                  \
  case OP: {                                               \
    _Pragma("unroll") for (int r = 0; r < R; ++r) {        \
      const double x = l2d(a[r]), y = l2d(b[r]);           \
      a[r] = d2l(EXPR);                                    \
    }                                                      \
  } break;
#define I_ARITH(OP, EXPR)                                  \
  case OP: {                                               \
    _Pragma("unroll") for (int r = 0; r < R; ++r) {        \
      const long long x = a[r], y = b[r];                  \
      a[r] = (EXPR);                                       \
    }                                                      \
  } break;
    F_ARITH(MSC_OP_ADD_F, x + y)
    F_ARITH(MSC_OP_SUB_F, x - y)
    F_ARITH(MSC_OP_MUL_F, x * y)
    case MSC_OP_DIV_F: {
      flag_zero_divisor<R>(b, true, vmask, err);
#pragma unroll
      for (int r = 0; r < R; ++r) a[r] = d2l(l2d(b[r]) == 0.0 ? 0.0 : l2d(a[r]) / l2d(b[r]));
    } break;
    case MSC_OP_FLOORDIV_F:
    case MSC_OP_MOD_F: {
      flag_zero_divisor<R>(b, true, vmask, err);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        double fd, md;
        py_divmod_f(l2d(a[r]), l2d(b[r]), &fd, &md);
        a[r] = d2l(op == MSC_OP_FLOORDIV_F ? fd : md);
      }
    } break;
    I_ARITH(MSC_OP_ADD_I, x + y)
    I_ARITH(MSC_OP_SUB_I, x - y)
    I_ARITH(MSC_OP_MUL_I, x * y)
    case MSC_OP_FLOORDIV_I: {
      flag_zero_divisor<R>(b, false, vmask, err);
#pragma unroll
      for (int r = 0; r < R; ++r) a[r] = py_floordiv_i(a[r], b[r]);
    } break;
    case MSC_OP_MOD_I: {
      flag_zero_divisor<R>(b, false, vmask, err);
#pragma unroll
      for (int r = 0; r < R; ++r) a[r] = py_mod_i(a[r], b[r]);
    } break;
    default: break;
  }
#undef F_ARITH
#undef I_ARITH
}

// comparisons and boolean ops: result is i64 0/1
template <int R>
__device__ __forceinline__ void cmpop(int op, long long (&a)[R], const long long (&b)[R]) {
  switch (op) {
#define F_CMP(OP, REL)                                                            \
  case OP: {                                                                      \
    _Pragma("unroll") for (int r = 0; r < R; ++r) a[r] = (l2d(a[r]) REL l2d(b[r])) ? 1 : 0; \
  } break;
#define I_CMP(OP, REL)                                                \
  case OP: {                                                          \
    _Pragma("unroll") for (int r = 0; r < R; ++r) a[r] = (a[r] REL b[r]) ? 1 : 0; \
  } break;
    F_CMP(MSC_OP_LT_F, <)
    F_CMP(MSC_OP_LE_F, <=)
    F_CMP(MSC_OP_GT_F, >)
    F_CMP(MSC_OP_GE_F, >=)
    F_CMP(MSC_OP_EQ_F, ==)
    F_CMP(MSC_OP_NE_F, !=)
    I_CMP(MSC_OP_LT_I, <)
    I_CMP(MSC_OP_LE_I, <=)
    I_CMP(MSC_OP_GT_I, >)
    I_CMP(MSC_OP_GE_I, >=)
    I_CMP(MSC_OP_EQ_I, ==)
    I_CMP(MSC_OP_NE_I, !=)
    case MSC_OP_AND: {
#pragma unroll
      for (int r = 0; r < R; ++r) a[r] = a[r] & b[r];
    } break;
    case MSC_OP_OR: {
#pragma unroll
      for (int r = 0; r < R; ++r) a[r] = a[r] | b[r];
    } break;
    default: break;
  }
#undef F_CMP
#undef I_CMP
}

// ------------------------------------------------------------------------------------------------
// column loads
// ------------------------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ void lds_words(const unsigned char* ptr, uint32_t (&w)[W / 4]) {
  if constexpr (W == 4) {
    w[0] = *reinterpret_cast<const uint32_t*>(ptr);
  } else if constexpr (W == 8) {
    const uint2 v = *reinterpret_cast<const uint2*>(ptr);
    w[0] = v.x;
    w[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < W / 16; ++i) {
      const uint4 v = reinterpret_cast<const uint4*>(ptr)[i];
      w[4 * i + 0] = v.x;
      w[4 * i + 1] = v.y;
      w[4 * i + 2] = v.z;
      w[4 * i + 3] = v.w;
    }
  }
}

// Load R consecutive rows of a staged column of physical type PHYS into 64-bit slots.
template <int R, int PHYS>
__device__ __forceinline__ void load_staged(const unsigned char* col_smem, int tid, long long (&dst)[R]) {
  constexpr int WIDTH = (PHYS == MSC_P_U8) ? 1 : (PHYS == MSC_P_U16) ? 2 : (PHYS == MSC_P_I64 || PHYS == MSC_P_F64) ? 8 : 4;
  constexpr int W = WIDTH * R;
  uint32_t w[W / 4];
  lds_words<W>(col_smem + tid * W, w);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if constexpr (PHYS == MSC_P_U8) dst[r] = (w[r / 4] >> (8 * (r % 4))) & 0xffu;
    else if constexpr (PHYS == MSC_P_U16) dst[r] = (w[r / 2] >> (16 * (r % 2))) & 0xffffu;
    else if constexpr (PHYS == MSC_P_U32) dst[r] = static_cast<long long>(w[r]);
    else if constexpr (PHYS == MSC_P_I32) dst[r] = static_cast<long long>(static_cast<int>(w[r]));
    else if constexpr (PHYS == MSC_P_F32) dst[r] = d2l(static_cast<double>(__uint_as_float(w[r])));
    else dst[r] = static_cast<long long>((static_cast<unsigned long long>(w[2 * r + 1]) << 32) | w[2 * r]);
  }
}

// Load through an index vector (staged u32 column): dst[r] = column[index[r]].
template <int R, int PHYS>
__device__ __forceinline__ void load_gather(const unsigned char* idx_smem, const void* col, int tid, uint32_t vmask,
                                            long long (&dst)[R]) {
  uint32_t idx[R];
  lds_words<4 * R>(idx_smem + tid * 4 * R, idx);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    long long v = 0;
    if ((vmask >> r) & 1u) {
      const uint32_t i = idx[r];
      if constexpr (PHYS == MSC_P_U8) v = __ldg(reinterpret_cast<const uint8_t*>(col) + i);
      else if constexpr (PHYS == MSC_P_U16) v = __ldg(reinterpret_cast<const uint16_t*>(col) + i);
      else if constexpr (PHYS == MSC_P_U32) v = __ldg(reinterpret_cast<const uint32_t*>(col) + i);
      else if constexpr (PHYS == MSC_P_I32) v = __ldg(reinterpret_cast<const int*>(col) + i);
      else if constexpr (PHYS == MSC_P_F32) v = d2l(static_cast<double>(__ldg(reinterpret_cast<const float*>(col) + i)));
      else v = __ldg(reinterpret_cast<const long long*>(col) + i);
    }
    dst[r] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// aggregation helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long agg_combine(int kind, long long cur, long long v) {
  switch (kind) {
    case MSC_AGG_SUM_F: return d2l(l2d(cur) + l2d(v));
    case MSC_AGG_SUM_I: return cur + v;
    case MSC_AGG_MIN_F: return (l2d(v) < l2d(cur)) ? v : cur;
    case MSC_AGG_MAX_F: return (l2d(v) > l2d(cur)) ? v : cur;
    case MSC_AGG_MIN_I: return (v < cur) ? v : cur;
    default: return (v > cur) ? v : cur;  // MSC_AGG_MAX_I
  }
}

__device__ __forceinline__ void atomic_fold(int kind, unsigned long long* addr, long long v) {
  switch (kind) {
    case MSC_AGG_SUM_F: atomicAdd(reinterpret_cast<double*>(addr), l2d(v)); break;
    case MSC_AGG_SUM_I: atomicAdd(addr, static_cast<unsigned long long>(v)); break;
    case MSC_AGG_MIN_I: atomicMin(reinterpret_cast<long long*>(addr), v); break;
    case MSC_AGG_MAX_I: atomicMax(reinterpret_cast<long long*>(addr), v); break;
    default: {  // f64 min / max: CAS loop
      unsigned long long old = *addr;
      while (true) {
        const long long merged = agg_combine(kind, static_cast<long long>(old), v);
        if (static_cast<unsigned long long>(merged) == old) break;
        const unsigned long long prev = atomicCAS(addr, old, static_cast<unsigned long long>(merged));
        if (prev == old) break;
        old = prev;
      }
    }
  }
}

template <int R, int KIND>
__device__ __forceinline__ void agg_hash(unsigned long long* haccs, uint64_t hcap, int a, const int (&grp)[R],
                                         const long long (&v)[R]) {
  unsigned long long* base = haccs + static_cast<uint64_t>(a) * hcap;
  // R consecutive rows of one thread often share a key (clustered tables, e.g. lineitem by
  // orderkey): fold each run in registers and issue one atomic per run.
  long long run = v[0];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    constexpr int dummy = 0;
    (void)dummy;
    const int nxt = (r + 1 < R) ? r + 1 : r;
    const bool same_next = (r + 1 < R) && grp[r] >= 0 && grp[nxt] == grp[r];
    if (same_next) {
      run = agg_combine(KIND, run, v[nxt]);
    } else {
      if (grp[r] >= 0) atomic_fold(KIND, base + grp[r], run);
      run = v[nxt];
    }
  }
}

__device__ __forceinline__ int hash_find_or_insert(unsigned long long* keys, uint64_t cap, long long key, int* err) {
  const uint64_t mask = cap - 1;
  uint64_t pos = msc_mix64(static_cast<uint64_t>(key)) & mask;
  const unsigned long long k = static_cast<unsigned long long>(key);
  for (uint64_t probe = 0; probe < cap; ++probe) {
    unsigned long long cur = keys[pos];
    if (cur == k) return static_cast<int>(pos);
    if (cur == HASH_EMPTY) {
      const unsigned long long prev = atomicCAS(keys + pos, HASH_EMPTY, k);
      if (prev == HASH_EMPTY || prev == k) return static_cast<int>(pos);
    }
    pos = (pos + 1) & mask;
  }
  atomicOr(err, MSC_DEVERR_TABLE_FULL);
  return -1;
}

// exclusive prefix sum of one u32 per thread across the CTA; returns the CTA total via *total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* scratch, int tid, uint32_t* total) {
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  __syncthreads();  // scratch may still be read by the previous use
  if (lane == 31) scratch[warp] = inc;
  __syncthreads();
  uint32_t warp_base = 0, sum = 0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) {
    const uint32_t s = scratch[w];
    if (w < warp) warp_base += s;
    sum += s;
  }
  *total = sum;
  return warp_base + inc - v;
}

// ------------------------------------------------------------------------------------------------
// instruction pieces: fetch operands -> compute -> store.  All per-row arrays are indexed with
// unrolled compile-time constants, so they are registers that live for one instruction only.
// ------------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void to_f64(long long (&v)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = d2l(static_cast<double>(v[r]));
}

template <int R>
__device__ __forceinline__ void fetch_staged(const ScanParams& p, const unsigned char* sbase, int idx, int tid, long long (&v)[R]) {
  const unsigned char* col = sbase + p.staged[idx].smem_off;
  switch (p.staged[idx].phys) {
    case MSC_P_U8: load_staged<R, MSC_P_U8>(col, tid, v); break;
    case MSC_P_U16: load_staged<R, MSC_P_U16>(col, tid, v); break;
    case MSC_P_U32: load_staged<R, MSC_P_U32>(col, tid, v); break;
    case MSC_P_I32: load_staged<R, MSC_P_I32>(col, tid, v); break;
    case MSC_P_F32: load_staged<R, MSC_P_F32>(col, tid, v); break;
    default: load_staged<R, MSC_P_I64>(col, tid, v); break;  // I64 and F64: raw 64-bit pattern
  }
}

template <int R>
__device__ __forceinline__ void fetch_gather(const ScanParams& p, const unsigned char* sbase, int idx, int tid, uint32_t vmask,
                                             long long (&v)[R]) {
  const unsigned char* ix = sbase + p.staged[idx >> 6].smem_off;
  const void* col = p.gather[idx & 63];
  switch (p.gather_phys[idx & 63]) {
    case MSC_P_U8: load_gather<R, MSC_P_U8>(ix, col, tid, vmask, v); break;
    case MSC_P_U16: load_gather<R, MSC_P_U16>(ix, col, tid, vmask, v); break;
    case MSC_P_U32: load_gather<R, MSC_P_U32>(ix, col, tid, vmask, v); break;
    case MSC_P_I32: load_gather<R, MSC_P_I32>(ix, col, tid, vmask, v); break;
    case MSC_P_F32: load_gather<R, MSC_P_F32>(ix, col, tid, vmask, v); break;
    default: load_gather<R, MSC_P_I64>(ix, col, tid, vmask, v); break;
  }
}

template <int R>
__device__ __forceinline__ void fetch_temp(const long long* temps, int idx, int tid, long long (&v)[R]) {
  const long long* t = temps + (idx * R) * NT + tid;
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = t[r * NT];
}

// generic operand fetch: the i2f variants are separate switch targets (a flag test would be
// if-converted into R predicated 64-bit conversions on every fetch)
template <int R>
__device__ __forceinline__ void fetch(const ScanParams& p, const unsigned char* sbase, const long long* temps,
                                      uint32_t operand, int tid, uint32_t vmask, long long (&v)[R]) {
  const int idx = operand & 0xfff;
  switch ((operand >> 12) & 15) {
    case MSC_SRC_TEMP: fetch_temp<R>(temps, idx, tid, v); break;
    case MSC_SRC_TEMP | MSC_SRC_I2F: fetch_temp<R>(temps, idx, tid, v); to_f64<R>(v); break;
    case MSC_SRC_STAGED: fetch_staged<R>(p, sbase, idx, tid, v); break;
    case MSC_SRC_STAGED | MSC_SRC_I2F: fetch_staged<R>(p, sbase, idx, tid, v); to_f64<R>(v); break;
    case MSC_SRC_GATHER: fetch_gather<R>(p, sbase, idx, tid, vmask, v); break;
    case MSC_SRC_GATHER | MSC_SRC_I2F: fetch_gather<R>(p, sbase, idx, tid, vmask, v); to_f64<R>(v); break;
    case MSC_SRC_CONST: {
      const long long c = p.consts[idx];
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = c;
    } break;
    default: {
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = 0;
    } break;
  }
}

template <int R, class TLut>
__device__ __forceinline__ void op_lut(long long (&x)[R], const TLut* lut, uint32_t vmask) {
#pragma unroll
  for (int r = 0; r < R; ++r) x[r] = ((vmask >> r) & 1u) ? static_cast<long long>(__ldg(lut + x[r])) : 0;
}

template <int R>
__device__ __forceinline__ void store_temp(long long* temps, int idx, int tid, const long long (&v)[R]) {
  long long* t = temps + (idx * R) * NT + tid;
#pragma unroll
  for (int r = 0; r < R; ++r) t[r * NT] = v[r];
}

// dense aggregation: per-thread shared-memory accumulators acc[(g * naggs + a) * NT + tid]
template <int R>
__device__ __forceinline__ void agg_dense(long long* acc, int naggs, int a, int kind, int tid, const int (&grp)[R],
                                          const long long (&v)[R]) {
  switch (kind) {
#define DENSE_CASE(KIND)                                       \
  case KIND: {                                                 \
    _Pragma("unroll") for (int r = 0; r < R; ++r) {            \
      long long* q = acc + (grp[r] * naggs + a) * NT + tid;    \
      *q = agg_combine(KIND, *q, v[r]);                        \
    }                                                          \
  } break;
    DENSE_CASE(MSC_AGG_SUM_F)
    DENSE_CASE(MSC_AGG_SUM_I)
    DENSE_CASE(MSC_AGG_MIN_F)
    DENSE_CASE(MSC_AGG_MAX_F)
    DENSE_CASE(MSC_AGG_MIN_I)
    default: {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        long long* q = acc + (grp[r] * naggs + a) * NT + tid;
        *q = agg_combine(MSC_AGG_MAX_I, *q, v[r]);
      }
    } break;
#undef DENSE_CASE
  }
}

template <int R>
__device__ __forceinline__ void agg_hash_any(unsigned long long* haccs, uint64_t hcap, int a, int kind, const int (&grp)[R],
                                             const long long (&v)[R]) {
  switch (kind) {
    case MSC_AGG_SUM_F: agg_hash<R, MSC_AGG_SUM_F>(haccs, hcap, a, grp, v); break;
    case MSC_AGG_SUM_I: agg_hash<R, MSC_AGG_SUM_I>(haccs, hcap, a, grp, v); break;
    case MSC_AGG_MIN_F: agg_hash<R, MSC_AGG_MIN_F>(haccs, hcap, a, grp, v); break;
    case MSC_AGG_MAX_F: agg_hash<R, MSC_AGG_MAX_F>(haccs, hcap, a, grp, v); break;
    case MSC_AGG_MIN_I: agg_hash<R, MSC_AGG_MIN_I>(haccs, hcap, a, grp, v); break;
    default: agg_hash<R, MSC_AGG_MAX_I>(haccs, hcap, a, grp, v); break;
  }
}

template <int R>
__device__ __forceinline__ void group_dense(const long long (&x)[R], uint32_t vmask, int ngroups, int naggs, long long* acc,
                                            int tid, int (&grp)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int g = ngroups;  // rows that failed the filter fold into a trash group that is never exported
    if ((vmask >> r) & 1u) {
      const long long code = x[r];
      g = (code >= 0 && code < ngroups) ? static_cast<int>(code) : ngroups;
    }
    grp[r] = g;
    acc[(g * naggs + (naggs - 1)) * NT + tid] += 1;  // hidden per-group row counter
  }
}

template <int R>
__device__ __forceinline__ void group_hash(const long long (&x)[R], uint32_t vmask, unsigned long long* hkeys, uint64_t hcap,
                                           int* err, int (&grp)[R]) {
  long long prev_key = 0;
  int prev_slot = -1;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int slot = -1;
    if ((vmask >> r) & 1u) {
      long long key = x[r];
      if (key == static_cast<long long>(HASH_EMPTY)) key = 0;  // -0.0 groups with +0.0, like a Python dict
      slot = (prev_slot >= 0 && key == prev_key) ? prev_slot : hash_find_or_insert(hkeys, hcap, key, err);
      prev_key = key;
      prev_slot = slot;
    }
    grp[r] = slot;
  }
}

template <int R, class TOut>
__device__ __forceinline__ void store_out(const long long (&x)[R], TOut* out, uint64_t pos, uint32_t vmask) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if ((vmask >> r) & 1u) {
      out[pos] = static_cast<TOut>(x[r]);
      ++pos;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fast shapes: handlers specialised at C++ compile time on (operand kinds, op, destination).  The
// host picks the shape id (`fast` field of w0, include/minispark_cuda.h MSC_FAST_*); one jump-table
// dispatch replaces the generic path's four switches and its register merges.
// ------------------------------------------------------------------------------------------------
struct FastCtx {
  const ScanParams& p;
  const unsigned char* sbase;
  long long* temps;
  long long* acc;
  int tid;
};

template <int R, int FK>
__device__ __forceinline__ void ffetch(const FastCtx& c, int idx, long long (&v)[R]) {
  if constexpr (FK == MSC_FK_TEMP) {
    fetch_temp<R>(c.temps, idx, c.tid, v);
  } else if constexpr (FK == MSC_FK_CONST) {
    const long long k = c.p.consts[idx];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = k;
  } else {
    const unsigned char* col = c.sbase + c.p.staged[idx].smem_off;
    if constexpr (FK == MSC_FK_F32) load_staged<R, MSC_P_F32>(col, c.tid, v);
    else if constexpr (FK == MSC_FK_F64 || FK == MSC_FK_I64) load_staged<R, MSC_P_I64>(col, c.tid, v);
    else if constexpr (FK == MSC_FK_I32) load_staged<R, MSC_P_I32>(col, c.tid, v);
    else if constexpr (FK == MSC_FK_I32F) {
      load_staged<R, MSC_P_I32>(col, c.tid, v);
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = d2l(static_cast<double>(static_cast<int>(v[r])));
    } else if constexpr (FK == MSC_FK_U8) load_staged<R, MSC_P_U8>(col, c.tid, v);
    else if constexpr (FK == MSC_FK_U16) load_staged<R, MSC_P_U16>(col, c.tid, v);
    else load_staged<R, MSC_P_U32>(col, c.tid, v);
  }
}

template <int R, int MODE, int KIND>
__device__ __forceinline__ void fast_agg(const FastCtx& c, int slot, const int (&grp)[R], const long long (&v)[R]) {
  if constexpr (MODE == MODE_DENSE) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      long long* q = c.acc + (grp[r] * c.p.naggs + slot) * NT + c.tid;
      *q = agg_combine(KIND, *q, v[r]);
    }
  } else if constexpr (MODE == MODE_HASH) {
    agg_hash<R, KIND>(c.p.haccs, c.p.hcap, slot, grp, v);
  }
}

// dst <- A (+|-|*) B on f64; DK: 0 TEMP, 1 AGG(SUM_F), 2 AGG(SUM_F) + tee TEMP
template <int R, int MODE, int OPI, int AK, int BK, int DK>
__device__ __forceinline__ void fast_arith(const FastCtx& c, uint32_t w0, uint32_t w1, const int (&grp)[R]) {
  long long a[R], b[R];
  ffetch<R, AK>(c, w1 & 0xfff, a);
  ffetch<R, BK>(c, (w1 >> 16) & 0xfff, b);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const double x = l2d(a[r]), y = l2d(b[r]);
    a[r] = d2l(OPI == 0 ? x + y : (OPI == 1 ? x - y : x * y));
  }
  const int dst = (w0 >> 13) & 0x7f;
  if constexpr (DK == 0) store_temp<R>(c.temps, dst, c.tid, a);
  if constexpr (DK == 2) store_temp<R>(c.temps, static_cast<int>((w0 >> 9) & 0xf) - 1, c.tid, a);
  if constexpr (DK >= 1) fast_agg<R, MODE, MSC_AGG_SUM_F>(c, dst, grp, a);
}

template <int R, int MODE, int KIND, int FK>
__device__ __forceinline__ void fast_aggmov(const FastCtx& c, uint32_t w0, uint32_t w1, const int (&grp)[R]) {
  long long a[R];
  ffetch<R, FK>(c, w1 & 0xfff, a);
  fast_agg<R, MODE, KIND>(c, (w0 >> 13) & 0x7f, grp, a);
}

template <int R, int CMPI, int FK>
__device__ __forceinline__ void fast_cmp_filter(const FastCtx& c, uint32_t w1, uint32_t& vmask) {
  long long a[R];
  ffetch<R, FK>(c, w1 & 0xfff, a);
  const long long k = c.p.consts[(w1 >> 16) & 0xfff];
  constexpr bool is_f = FK == MSC_FK_F32 || FK == MSC_FK_F64 || FK == MSC_FK_I32F;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    bool t;
    if constexpr (is_f) {
      const double x = l2d(a[r]), y = l2d(k);
      t = CMPI == 0 ? x < y : CMPI == 1 ? x <= y : CMPI == 2 ? x > y : CMPI == 3 ? x >= y : CMPI == 4 ? x == y : x != y;
    } else {
      const long long x = a[r];
      t = CMPI == 0 ? x < k : CMPI == 1 ? x <= k : CMPI == 2 ? x > k : CMPI == 3 ? x >= k : CMPI == 4 ? x == k : x != k;
    }
    if (!t) vmask &= ~(1u << r);
  }
}

template <int R, int MODE, int FK>
__device__ __forceinline__ void fast_group(const FastCtx& c, uint32_t w1, uint32_t vmask, int (&grp)[R]) {
  long long a[R];
  ffetch<R, FK>(c, w1 & 0xfff, a);
  if constexpr (MODE == MODE_DENSE) group_dense<R>(a, vmask, c.p.ngroups, c.p.naggs, c.acc, c.tid, grp);
  else if constexpr (MODE == MODE_HASH) group_hash<R>(a, vmask, c.p.hkeys, c.p.hcap, c.p.err, grp);
}

template <int R, int FK, int U32OUT>
__device__ __forceinline__ void fast_out(const FastCtx& c, uint32_t w0, uint32_t w1, uint64_t out_pos, uint32_t vmask) {
  long long a[R];
  ffetch<R, FK>(c, w1 & 0xfff, a);
  void* out = c.p.out[(w0 >> 13) & 0x7f];
  if constexpr (U32OUT) store_out<R, uint32_t>(a, reinterpret_cast<uint32_t*>(out), out_pos, vmask);
  else store_out<R, long long>(a, reinterpret_cast<long long*>(out), out_pos, vmask);
}

// returns false when the id is not a compiled shape (the caller then runs the generic path)
template <int R, int MODE>
__device__ __forceinline__ bool run_fast(const FastCtx& c, int fast, uint32_t w0, uint32_t w1, uint32_t& vmask, int (&grp)[R],
                                         uint64_t out_pos) {
  constexpr bool AGG = MODE == MODE_DENSE || MODE == MODE_HASH;
  switch (fast) {
#define ARITH_ID(OPI, AK, BK, DK) (MSC_FAST_ARITH + (((OPI) * 5 + (AK)) * 5 + (BK)) * 3 + (DK))
#define ARITH_CASE(OPI, AK, BK, DK)                                          \
  case ARITH_ID(OPI, AK, BK, DK):                                            \
    if constexpr (AGG || DK == 0) fast_arith<R, MODE, OPI, AK, BK, DK>(c, w0, w1, grp); \
    return true;
#define ARITH_DK(OPI, AK, BK) ARITH_CASE(OPI, AK, BK, 0) ARITH_CASE(OPI, AK, BK, 1) ARITH_CASE(OPI, AK, BK, 2)
#define ARITH_BK(OPI, AK) ARITH_DK(OPI, AK, 0) ARITH_DK(OPI, AK, 1) ARITH_DK(OPI, AK, 2) ARITH_DK(OPI, AK, 3) ARITH_DK(OPI, AK, 4)
#define ARITH_AK(OPI) ARITH_BK(OPI, 0) ARITH_BK(OPI, 1) ARITH_BK(OPI, 2) ARITH_BK(OPI, 3) ARITH_BK(OPI, 4)
    ARITH_AK(0)
    ARITH_AK(1)
    ARITH_AK(2)
#undef ARITH_AK
#undef ARITH_BK
#undef ARITH_DK
#undef ARITH_CASE
#undef ARITH_ID
#define AGGMOV_CASE(KIND, FK)                                                          \
  case MSC_FAST_AGGMOV + (KIND) * 10 + (FK):                                           \
    if constexpr (AGG) fast_aggmov<R, MODE, KIND, FK>(c, w0, w1, grp);                 \
    return true;
    AGGMOV_CASE(MSC_AGG_SUM_F, MSC_FK_F32) AGGMOV_CASE(MSC_AGG_SUM_F, MSC_FK_F64) AGGMOV_CASE(MSC_AGG_SUM_F, MSC_FK_TEMP)
    AGGMOV_CASE(MSC_AGG_SUM_F, MSC_FK_I32F)
    AGGMOV_CASE(MSC_AGG_SUM_I, MSC_FK_CONST) AGGMOV_CASE(MSC_AGG_SUM_I, MSC_FK_I32) AGGMOV_CASE(MSC_AGG_SUM_I, MSC_FK_I64)
    AGGMOV_CASE(MSC_AGG_SUM_I, MSC_FK_TEMP)
    AGGMOV_CASE(MSC_AGG_MIN_F, MSC_FK_F32) AGGMOV_CASE(MSC_AGG_MIN_F, MSC_FK_F64) AGGMOV_CASE(MSC_AGG_MIN_F, MSC_FK_TEMP)
    AGGMOV_CASE(MSC_AGG_MAX_F, MSC_FK_F32) AGGMOV_CASE(MSC_AGG_MAX_F, MSC_FK_F64) AGGMOV_CASE(MSC_AGG_MAX_F, MSC_FK_TEMP)
    AGGMOV_CASE(MSC_AGG_MIN_I, MSC_FK_I32) AGGMOV_CASE(MSC_AGG_MIN_I, MSC_FK_I64) AGGMOV_CASE(MSC_AGG_MIN_I, MSC_FK_TEMP)
    AGGMOV_CASE(MSC_AGG_MAX_I, MSC_FK_I32) AGGMOV_CASE(MSC_AGG_MAX_I, MSC_FK_I64) AGGMOV_CASE(MSC_AGG_MAX_I, MSC_FK_TEMP)
#undef AGGMOV_CASE
#define CMP_CASE(CMPI, FK) \
  case MSC_FAST_CMP + (CMPI) * 10 + (FK): fast_cmp_filter<R, CMPI, FK>(c, w1, vmask); return true;
#define CMP_ALL(CMPI)                                                                                              \
  CMP_CASE(CMPI, MSC_FK_F32) CMP_CASE(CMPI, MSC_FK_F64) CMP_CASE(CMPI, MSC_FK_I32F) CMP_CASE(CMPI, MSC_FK_I32)     \
  CMP_CASE(CMPI, MSC_FK_I64) CMP_CASE(CMPI, MSC_FK_U8) CMP_CASE(CMPI, MSC_FK_U16) CMP_CASE(CMPI, MSC_FK_U32)
    CMP_ALL(0) CMP_ALL(1) CMP_ALL(2) CMP_ALL(3) CMP_ALL(4) CMP_ALL(5)
#undef CMP_ALL
#undef CMP_CASE
#define GROUP_CASE(FK)                                                   \
  case MSC_FAST_GROUP + (FK):                                            \
    if constexpr (AGG) fast_group<R, MODE, FK>(c, w1, vmask, grp);       \
    return true;
    GROUP_CASE(MSC_FK_U8) GROUP_CASE(MSC_FK_U16) GROUP_CASE(MSC_FK_U32) GROUP_CASE(MSC_FK_I32) GROUP_CASE(MSC_FK_I64)
    GROUP_CASE(MSC_FK_TEMP)
#undef GROUP_CASE
#define OUT_CASE(FK, U32OUT)                                                          \
  case MSC_FAST_OUT + (FK) * 2 + (U32OUT):                                            \
    if constexpr (MODE == MODE_PROJECT) fast_out<R, FK, U32OUT>(c, w0, w1, out_pos, vmask); \
    return true;
    OUT_CASE(MSC_FK_F32, 0) OUT_CASE(MSC_FK_F64, 0) OUT_CASE(MSC_FK_TEMP, 0) OUT_CASE(MSC_FK_I32, 0) OUT_CASE(MSC_FK_I64, 0)
    OUT_CASE(MSC_FK_I32F, 0) OUT_CASE(MSC_FK_U8, 1) OUT_CASE(MSC_FK_U16, 1) OUT_CASE(MSC_FK_U32, 1) OUT_CASE(MSC_FK_TEMP, 1)
#undef OUT_CASE
    default: return false;
  }
}

template <int R>
__device__ __forceinline__ void issue_tile(const ScanParams& p, unsigned char* stages, uint64_t* full, uint32_t k) {
  constexpr uint32_t TILE = NT * R;
  const uint32_t stage = k % p.nstages;
  const uint64_t tile = blockIdx.x + static_cast<uint64_t>(k) * gridDim.x;
  unsigned char* sbase = stages + static_cast<size_t>(stage) * p.stage_bytes;
  uint32_t total = 0;
  for (uint32_t c = 0; c < p.nstaged; ++c) total += p.staged[c].width * TILE;
  mbar_expect_tx(&full[stage], total);
  for (uint32_t c = 0; c < p.nstaged; ++c) {
    const uint32_t bytes = p.staged[c].width * TILE;
    bulk_g2s(sbase + p.staged[c].smem_off, p.staged[c].base + tile * bytes, bytes, &full[stage]);
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int R, int MODE>
__global__ void __launch_bounds__(NT) scan_kernel(const __grid_constant__ ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint32_t* scratch = reinterpret_cast<uint32_t*>(smem + 64);
  unsigned char* stages = smem + SMEM_HEADER;
  long long* temps = reinterpret_cast<long long*>(stages + static_cast<size_t>(p.nstages) * p.stage_bytes);
  long long* acc = temps + static_cast<size_t>(p.ntemps) * R * NT;
  constexpr int TILE = NT * R;
  const int tid = threadIdx.x;

  if (tid == 0) {
    for (uint32_t st = 0; st < p.nstages; ++st) mbar_init(&full[st], 1);
    mbar_fence_init();
  }
  if constexpr (MODE == MODE_DENSE) {
    const int cells = (p.ngroups + 1) * p.naggs;
    for (int c = 0; c < cells; ++c) acc[c * NT + tid] = p.agg_init[c % p.naggs];
  }
  __syncthreads();

  const uint32_t ntiles_cta = (p.ntiles > blockIdx.x) ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
    const uint32_t pre = ntiles_cta < p.nstages ? ntiles_cta : p.nstages;
    for (uint32_t k = 0; k < pre; ++k) issue_tile<R>(p, stages, full, k);
  }

  for (uint32_t k = 0; k < ntiles_cta; ++k) {
    const uint32_t stage = k % p.nstages;
    const uint32_t parity = (k / p.nstages) & 1u;
    const uint64_t tile = blockIdx.x + static_cast<uint64_t>(k) * gridDim.x;
    const unsigned char* sbase = stages + static_cast<size_t>(stage) * p.stage_bytes;
    while (!mbar_try_wait(&full[stage], parity)) {
    }
    const uint64_t row0 = tile * TILE + static_cast<uint64_t>(tid) * R;
    uint32_t vmask = 0;
    int grp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r < p.nrows) vmask |= 1u << r;
      grp[r] = (MODE == MODE_DENSE) ? p.ngroups : -1;
    }
    uint64_t out_pos = row0;  // project: output position of this thread's first surviving row

    const FastCtx fc{p, sbase, temps, acc, tid};
    for (int pc = 0;; pc += 2) {
      const uint32_t w0 = p.code[pc];
      const int op = w0 & 0x3f;
      if (op == MSC_OP_END) break;
      const uint32_t w1 = p.code[pc + 1];
      const int fast = w0 >> 20;
      if (fast != 0 && run_fast<R, MODE>(fc, fast, w0, w1, vmask, grp, out_pos)) continue;
      if (op == MSC_OP_RANK) {
        if constexpr (MODE == MODE_COUNT) {
          uint32_t total;
          (void)block_exclusive_scan(__popc(vmask), scratch, tid, &total);
          if (tid == 0) p.tile_counts[tile] = total;
        } else if constexpr (MODE == MODE_PROJECT) {
          if (p.tile_offsets != nullptr) {
            uint32_t total;
            const uint32_t before = block_exclusive_scan(__popc(vmask), scratch, tid, &total);
            out_pos = p.tile_offsets[tile] + before;
          }
        }
        continue;
      }
      long long a[R];
      fetch<R>(p, sbase, temps, w1 & 0xffffu, tid, vmask, a);
      if (op >= MSC_OP_ADD_F && op <= MSC_OP_OR) {
        long long b[R];
        fetch<R>(p, sbase, temps, w1 >> 16, tid, vmask, b);
        if (op <= MSC_OP_MOD_I) binop<R>(op, a, b, vmask, p.err);
        else cmpop<R>(op, a, b);
      } else if (op == MSC_OP_LUT8) {
        op_lut<R, uint8_t>(a, reinterpret_cast<const uint8_t*>(p.luts[(w1 >> 16) & 0xfff]), vmask);
      } else if (op == MSC_OP_LUT32) {
        op_lut<R, uint32_t>(a, reinterpret_cast<const uint32_t*>(p.luts[(w1 >> 16) & 0xfff]), vmask);
      }
      // ---- store -------------------------------------------------------------------------------
      const int tee = (w0 >> 9) & 0xf;
      if (tee) store_temp<R>(temps, tee - 1, tid, a);
      const int dst = (w0 >> 13) & 0x7f;
      switch ((w0 >> 6) & 7) {
        case MSC_DST_TEMP: store_temp<R>(temps, dst, tid, a); break;
        case MSC_DST_FILTER: {
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (a[r] == 0) vmask &= ~(1u << r);
        } break;
        case MSC_DST_GROUP:
          if constexpr (MODE == MODE_DENSE) group_dense<R>(a, vmask, p.ngroups, p.naggs, acc, tid, grp);
          else if constexpr (MODE == MODE_HASH) group_hash<R>(a, vmask, p.hkeys, p.hcap, p.err, grp);
          break;
        case MSC_DST_AGG:
          if constexpr (MODE == MODE_DENSE) agg_dense<R>(acc, p.naggs, dst, p.agg_kind[dst], tid, grp, a);
          else if constexpr (MODE == MODE_HASH) agg_hash_any<R>(p.haccs, p.hcap, dst, p.agg_kind[dst], grp, a);
          break;
        case MSC_DST_OUT:
          if constexpr (MODE == MODE_PROJECT) {
            if (p.out_phys[dst] == MSC_P_U32) store_out<R, uint32_t>(a, reinterpret_cast<uint32_t*>(p.out[dst]), out_pos, vmask);
            else store_out<R, long long>(a, reinterpret_cast<long long*>(p.out[dst]), out_pos, vmask);
          }
          break;
        default: break;
      }
    }

    __syncthreads();  // every thread is done reading this stage
    if (tid == 0 && k + p.nstages < ntiles_cta) issue_tile<R>(p, stages, full, k + p.nstages);
  }

  if constexpr (MODE == MODE_DENSE) {
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    const int cells = p.ngroups * p.naggs;  // the trash group is not exported
    for (int c = warp; c < cells; c += NT / 32) {
      const int kind = p.agg_kind[c % p.naggs];
      const long long* base = acc + c * NT;
      long long v = base[lane];
#pragma unroll
      for (int j = 1; j < NT / 32; ++j) v = agg_combine(kind, v, base[lane + 32 * j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = agg_combine(kind, v, __shfl_xor_sync(0xffffffffu, v, o));
      if (lane == 0 && v != p.agg_init[c % p.naggs]) atomic_fold(kind, p.dense_out + c, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
__global__ void fill_u64_kernel(unsigned long long* p, unsigned long long v, uint64_t n) {
  for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}

__global__ void dense_init_kernel(unsigned long long* out, int ngroups, int naggs, const long long* init) {
  const int n = ngroups * naggs;
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = static_cast<unsigned long long>(init[i % naggs]);
}

// compact the dense table: one output row per group whose hidden row counter is > 0
__global__ void dense_finalize_kernel(const unsigned long long* table, int ngroups, int naggs_total, uint32_t* out_key,
                                      unsigned long long* const* out_acc, unsigned long long* out_n) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long n = 0;
  for (int g = 0; g < ngroups; ++g) {
    if (table[g * naggs_total + naggs_total - 1] == 0) continue;
    out_key[n] = g;
    for (int a = 0; a < naggs_total - 1; ++a) out_acc[a][n] = table[g * naggs_total + a];
    ++n;
  }
  *out_n = n;
}

constexpr int HTILE = 1024;  // slots per block in the hash-table compaction
__global__ void hash_count_kernel(const unsigned long long* keys, uint64_t cap, uint32_t* tile_counts) {
  __shared__ uint32_t wsum[8];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * HTILE;
  uint32_t c = 0;
  for (int i = threadIdx.x; i < HTILE; i += blockDim.x) c += (base + i < cap && keys[base + i] != HASH_EMPTY);
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t tot = 0;
    for (int w = 0; w < blockDim.x / 32; ++w) tot += wsum[w];
    tile_counts[blockIdx.x] = tot;
  }
}

__global__ void hash_emit_kernel(const unsigned long long* keys, const unsigned long long* accs, uint64_t cap,
                                 int naggs, const uint64_t* tile_offsets, long long* out_key,
                                 unsigned long long* const* out_acc) {
  // 256 threads, HTILE slots: each thread owns 4 consecutive slots so output order is slot order
  __shared__ uint32_t scratch[8];
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * HTILE + threadIdx.x * 4;
  uint32_t c = 0;
  unsigned long long k[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    k[i] = (base + i < cap) ? keys[base + i] : HASH_EMPTY;
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
    for (int a = 0; a < naggs; ++a) out_acc[a][pos] = accs[static_cast<uint64_t>(a) * cap + base + i];
    ++pos;
  }
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
struct LaunchPlan {
  ScanParams p;
  int R;
  size_t smem;
  int grid;
};

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
    case MSC_SRC_LUT:
      if (!allow_lut) return ctx->fail(MSC_ERR_ARG, "operand: LUT reference outside a LUT instruction");
      return idx < sd->nluts ? MSC_OK : ctx->fail(MSC_ERR_ARG, "operand: bad LUT");
    default: return ctx->fail(MSC_ERR_ARG, "operand: unknown kind");
  }
}

int validate_program(msc_ctx* ctx, const msc_scan_desc* sd, int mode, int naggs, int nout) {
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
    const bool lut = op == MSC_OP_LUT8 || op == MSC_OP_LUT32;
    MSC_TRY(validate_operand(ctx, sd, w1 & 0xffffu, false));
    MSC_TRY(validate_operand(ctx, sd, w1 >> 16, lut));
    if (lut && ((w1 >> 28) & 7) != MSC_SRC_LUT) return ctx->fail(MSC_ERR_ARG, "LUT instruction needs a LUT operand");
    const int tee = (w0 >> 9) & 0xf, dkind = (w0 >> 6) & 7, dst = (w0 >> 13) & 0x7f;
    if (tee > sd->ntemps) return ctx->fail(MSC_ERR_ARG, "tee: bad temporary");
    switch (dkind) {
      case MSC_DST_TEMP:
        if (dst >= sd->ntemps) return ctx->fail(MSC_ERR_ARG, "destination: bad temporary");
        break;
      case MSC_DST_FILTER: case MSC_DST_NONE: break;
      case MSC_DST_GROUP:
        if (mode != MODE_DENSE && mode != MODE_HASH) return ctx->fail(MSC_ERR_ARG, "GROUP outside an aggregate scan");
        break;
      case MSC_DST_AGG:
        if (mode != MODE_DENSE && mode != MODE_HASH) return ctx->fail(MSC_ERR_ARG, "AGG outside an aggregate scan");
        if (dst >= naggs) return ctx->fail(MSC_ERR_ARG, "AGG: bad accumulator");
        break;
      case MSC_DST_OUT:
        if (dst >= nout) return ctx->fail(MSC_ERR_ARG, "OUT: bad column");
        break;
      default: return ctx->fail(MSC_ERR_ARG, "unknown destination kind");
    }
  }
  if (!ended) return ctx->fail(MSC_ERR_ARG, "program has no END");
  return MSC_OK;
}

// Build kernel params + launch geometry.  extra_smem = bytes needed after the stage ring and temporaries.
int plan_launch(msc_ctx* ctx, const msc_scan_desc* sd, int R, size_t extra_smem, LaunchPlan* lp) {
  ScanParams& p = lp->p;
  memset(&p, 0, sizeof(p));
  const uint32_t tile = NT * R;
  p.nrows = sd->nrows;
  p.ntiles = static_cast<uint32_t>((sd->nrows + tile - 1) / tile);
  p.nstaged = sd->nstaged;
  p.ntemps = sd->ntemps;
  uint32_t off = 0;
  for (int c = 0; c < sd->nstaged; ++c) {
    const size_t w = msc_phys_width(sd->staged[c].phys);
    if (w == 0 || sd->staged[c].data == nullptr) return ctx->fail(MSC_ERR_ARG, "bad staged column");
    if ((reinterpret_cast<uintptr_t>(sd->staged[c].data) & 15) != 0) return ctx->fail(MSC_ERR_ARG, "staged column not 16B aligned");
    p.staged[c].base = static_cast<const unsigned char*>(sd->staged[c].data);
    p.staged[c].width = static_cast<uint32_t>(w);
    p.staged[c].smem_off = off;
    p.staged[c].phys = sd->staged[c].phys;
    off += static_cast<uint32_t>(msc_round_up(w * tile, 128));
  }
  p.stage_bytes = off ? off : 128;
  for (int c = 0; c < sd->ngather; ++c) {
    if (msc_phys_width(sd->gather[c].phys) == 0 || sd->gather[c].data == nullptr) return ctx->fail(MSC_ERR_ARG, "bad gather column");
    p.gather[c] = sd->gather[c].data;
    p.gather_phys[c] = sd->gather[c].phys;
  }
  for (int c = 0; c < sd->nluts; ++c) p.luts[c] = sd->luts[c];
  memcpy(p.code, sd->code, sizeof(uint32_t) * sd->ncode);
  if (sd->ncode < MSC_VM_MAX_CODE) p.code[sd->ncode] = MSC_OP_END;
  memcpy(p.consts, sd->consts, sizeof(int64_t) * sd->nconsts);
  p.err = ctx->d_err;
  const size_t temps_bytes = static_cast<size_t>(sd->ntemps) * tile * sizeof(long long);
  const size_t fixed = SMEM_HEADER + temps_bytes + extra_smem;
  // ring depth: target 3 CTAs per SM (about 72 KB each) and >= 2 stages; MSC_SCAN_STAGES overrides
  static const int forced = getenv("MSC_SCAN_STAGES") ? atoi(getenv("MSC_SCAN_STAGES")) : 0;
  const size_t budget = 72 * 1024;
  uint32_t ns = budget > fixed ? static_cast<uint32_t>((budget - fixed) / p.stage_bytes) : 0;
  if (ns > 4) ns = 4;
  if (ns < 2) ns = 2;
  if (forced >= 1 && forced <= MAX_STAGES) ns = static_cast<uint32_t>(forced);
  p.nstages = ns;
  lp->R = R;
  lp->smem = fixed + static_cast<size_t>(ns) * p.stage_bytes;
  if (lp->smem > 227 * 1024) return ctx->fail(MSC_ERR_ARG, "scan needs more shared memory than an SM has");
  lp->grid = 0;
  return MSC_OK;
}

template <int R, int MODE>
int launch_scan(msc_ctx* ctx, LaunchPlan* lp) {
  auto kern = scan_kernel<R, MODE>;
  MSC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(lp->smem)));
  int occ = 0;
  MSC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, lp->smem));
  if (occ < 1) return ctx->fail(MSC_ERR_ARG, "scan kernel does not fit on an SM");
  int grid = ctx->sm_count * occ;
  if (static_cast<uint32_t>(grid) > lp->p.ntiles) grid = static_cast<int>(lp->p.ntiles);
  if (grid < 1) grid = 1;
  lp->grid = grid;
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_s0, ctx->stream));
  kern<<<grid, NT, lp->smem, ctx->stream>>>(lp->p);
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_s1, ctx->stream));
  ctx->stats.launches += 1;
  ctx->stats.last_scan_grid = grid;
  ctx->stats.last_scan_stages = static_cast<int32_t>(lp->p.nstages);
  ctx->stats.last_scan_smem = static_cast<int32_t>(lp->smem);
  ctx->stats.last_scan_rows_per_thread = R;
  MSC_CUDA(ctx, cudaGetLastError());
  return MSC_OK;
}

template <int MODE>
int launch_scan_r(msc_ctx* ctx, LaunchPlan* lp) {
  if (lp->R == 8) return launch_scan<8, MODE>(ctx, lp);
  return launch_scan<4, MODE>(ctx, lp);
}

int pick_rows_per_thread() {
  static int r = -1;
  if (r < 0) {
    const char* e = getenv("MSC_SCAN_R");
    r = (e && atoi(e) == 8) ? 8 : 4;
  }
  return r;
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

}  // namespace

int msc_exclusive_scan_u8_u64(msc_ctx* ctx, const uint8_t* in, uint64_t* out, uint64_t n) {
  return exclusive_scan_impl<uint8_t>(ctx, in, out, n);
}
int msc_exclusive_scan_u32_u64(msc_ctx* ctx, const uint32_t* in, uint64_t* out, uint64_t n) {
  return exclusive_scan_impl<uint32_t>(ctx, in, out, n);
}

// =================================================================================================
extern "C" int msc_scan_aggregate(msc_ctx* ctx, const msc_scan_desc* sd, int32_t ngroups, const int32_t* agg_kinds,
                                  int32_t naggs, uint64_t hash_capacity_hint, msc_rel** out) {
  if (!ctx || !sd || !out || naggs < 0 || naggs > MSC_VM_MAX_AGGS) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  const bool dense = ngroups > 0;
  MSC_TRY(validate_program(ctx, sd, dense ? MODE_DENSE : MODE_HASH, naggs, 0));
  long long init[MSC_VM_MAX_AGGS + 1];
  int kinds[MSC_VM_MAX_AGGS + 1];
  for (int a = 0; a < naggs; ++a) {
    kinds[a] = agg_kinds[a];
    switch (agg_kinds[a]) {
      case MSC_AGG_SUM_F: init[a] = __builtin_bit_cast(long long, 0.0); break;
      case MSC_AGG_SUM_I: init[a] = 0; break;
      // MIN/MAX are seeded with the reference's MAX_INT / MIN_INT sentinels (tasks.py:303-310, constants.py:14-15)
      case MSC_AGG_MIN_F: init[a] = __builtin_bit_cast(long long, 2147483647.0); break;
      case MSC_AGG_MAX_F: init[a] = __builtin_bit_cast(long long, -2147483648.0); break;
      case MSC_AGG_MIN_I: init[a] = 2147483647LL; break;
      case MSC_AGG_MAX_I: init[a] = -2147483648LL; break;
      default: return ctx->fail(MSC_ERR_ARG, "bad aggregate kind");
    }
  }
  const int R = pick_rows_per_thread();
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));

  if (dense) {
    const int ntot = naggs + 1;  // + hidden row counter
    kinds[naggs] = MSC_AGG_SUM_I;
    init[naggs] = 0;
    const size_t acc_bytes = static_cast<size_t>(ngroups + 1) * ntot * NT * sizeof(long long);
    if (acc_bytes > 96 * 1024) return ctx->fail(MSC_ERR_ARG, "dense aggregate: groups x aggregates too large; use hash mode");
    LaunchPlan lp;
    MSC_TRY(plan_launch(ctx, sd, R, acc_bytes, &lp));
    lp.p.ngroups = ngroups;
    lp.p.naggs = ntot;
    memcpy(lp.p.agg_init, init, sizeof(long long) * ntot);
    memcpy(lp.p.agg_kind, kinds, sizeof(int) * ntot);
    // global table, initialised with the identities
    DevTmp table(ctx), d_init(ctx), d_n(ctx), d_ptrs(ctx);
    MSC_TRY(table.alloc(sizeof(unsigned long long) * ngroups * ntot));
    MSC_TRY(d_init.alloc(sizeof(long long) * ntot));
    MSC_TRY(d_n.alloc(sizeof(unsigned long long)));
    MSC_CUDA(ctx, cudaMemcpyAsync(d_init.p, init, sizeof(long long) * ntot, cudaMemcpyHostToDevice, ctx->stream));
    dense_init_kernel<<<1, 256, 0, ctx->stream>>>(table.as<unsigned long long>(), ngroups, ntot, d_init.as<long long>());
    ctx->stats.launches += 1;
    lp.p.dense_out = table.as<unsigned long long>();
    if (sd->nrows > 0) MSC_TRY(launch_scan_r<MODE_DENSE>(ctx, &lp));
    // compact present groups into the output relation
    msc_rel* rel = new_rel(ctx, 0);
    int rc = add_col(ctx, rel, MSC_P_U32, ngroups);
    for (int a = 0; rc == MSC_OK && a < naggs; ++a)
      rc = add_col(ctx, rel, (kinds[a] == MSC_AGG_SUM_F || kinds[a] == MSC_AGG_MIN_F || kinds[a] == MSC_AGG_MAX_F) ? MSC_P_F64 : MSC_P_I64, ngroups);
    if (rc != MSC_OK) {
      msc_rel_free(rel);
      return rc;
    }
    std::vector<unsigned long long*> ptrs;
    for (int a = 0; a < naggs; ++a) ptrs.push_back(static_cast<unsigned long long*>(rel->cols[1 + a].data));
    MSC_TRY(d_ptrs.alloc(sizeof(void*) * (naggs + 1)));
    if (naggs) MSC_CUDA(ctx, cudaMemcpyAsync(d_ptrs.p, ptrs.data(), sizeof(void*) * naggs, cudaMemcpyHostToDevice, ctx->stream));
    dense_finalize_kernel<<<1, 32, 0, ctx->stream>>>(table.as<unsigned long long>(), ngroups, ntot,
                                                    static_cast<uint32_t*>(rel->cols[0].data),
                                                    d_ptrs.as<unsigned long long*>(), d_n.as<unsigned long long>());
    ctx->stats.launches += 1;
    MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
    unsigned long long n = 0;
    MSC_CUDA(ctx, cudaMemcpyAsync(&n, d_n.p, sizeof(n), cudaMemcpyDeviceToHost, ctx->stream));
    MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rel->nrows = n;
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
    ctx->stats.last_kernel_ms = ms;
  if (sd->nrows > 0 && cudaEventElapsedTime(&ms, ctx->ev_s0, ctx->ev_s1) == cudaSuccess) ctx->stats.last_scan_ms = ms;
    int drc = msc_check_device_error(ctx);
    if (drc != MSC_OK) {
      msc_rel_free(rel);
      return drc;
    }
    *out = rel;
    return MSC_OK;
  }

  // ---- hash mode ----
  uint64_t want = hash_capacity_hint ? hash_capacity_hint : sd->nrows;
  if (want < 16) want = 16;
  uint64_t cap = 64;
  while (cap < want * 2) cap <<= 1;
  if (cap > (1ULL << 31)) return ctx->fail(MSC_ERR_ARG, "hash aggregate: more than 2^30 groups per GPU is not supported");
  LaunchPlan lp;
  MSC_TRY(plan_launch(ctx, sd, R, 0, &lp));
  lp.p.naggs = naggs;
  memcpy(lp.p.agg_kind, kinds, sizeof(int) * naggs);
  DevTmp keys(ctx), accs(ctx), counts(ctx), offsets(ctx), d_ptrs(ctx);
  MSC_TRY(keys.alloc(cap * sizeof(unsigned long long)));
  MSC_TRY(accs.alloc(cap * sizeof(unsigned long long) * (naggs ? naggs : 1)));
  const int fill_grid = ctx->sm_count * 8;
  fill_u64_kernel<<<fill_grid, 256, 0, ctx->stream>>>(keys.as<unsigned long long>(), HASH_EMPTY, cap);
  for (int a = 0; a < naggs; ++a)
    fill_u64_kernel<<<fill_grid, 256, 0, ctx->stream>>>(accs.as<unsigned long long>() + static_cast<uint64_t>(a) * cap,
                                                        static_cast<unsigned long long>(init[a]), cap);
  ctx->stats.launches += 1 + naggs;
  lp.p.hkeys = keys.as<unsigned long long>();
  lp.p.haccs = accs.as<unsigned long long>();
  lp.p.hcap = cap;
  if (sd->nrows > 0) MSC_TRY(launch_scan_r<MODE_HASH>(ctx, &lp));
  // compaction of occupied slots
  const uint64_t nht = (cap + HTILE - 1) / HTILE;
  MSC_TRY(counts.alloc(nht * sizeof(uint32_t)));
  MSC_TRY(offsets.alloc((nht + 1) * sizeof(uint64_t)));
  hash_count_kernel<<<static_cast<unsigned>(nht), 256, 0, ctx->stream>>>(keys.as<unsigned long long>(), cap, counts.as<uint32_t>());
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
  hash_emit_kernel<<<static_cast<unsigned>(nht), 256, 0, ctx->stream>>>(keys.as<unsigned long long>(), accs.as<unsigned long long>(), cap, naggs,
                                                                      offsets.as<uint64_t>(), static_cast<long long*>(rel->cols[0].data),
                                                                      d_ptrs.as<unsigned long long*>());
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

extern "C" int msc_scan_project(msc_ctx* ctx, const msc_scan_desc* sd, const int32_t* out_phys, int32_t nout, msc_rel** out) {
  if (!ctx || !sd || !out || nout < 0 || nout > MSC_VM_MAX_OUT) return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  MSC_TRY(validate_program(ctx, sd, MODE_PROJECT, 0, nout));
  for (int i = 0; i < nout; ++i)
    if (out_phys[i] != MSC_P_I64 && out_phys[i] != MSC_P_F64 && out_phys[i] != MSC_P_U32)
      return ctx->fail(MSC_ERR_ARG, "project output must be I64, F64 or U32");
  const int R = pick_rows_per_thread();
  bool has_filter = false;
  int rank_pc = -1;
  for (int pc = 0; pc + 1 < sd->ncode; pc += 2) {
    const int op = sd->code[pc] & 0x3f;
    if (op == MSC_OP_END) break;
    if (((sd->code[pc] >> 6) & 7) == MSC_DST_FILTER && op != MSC_OP_RANK) has_filter = true;
    if (op == MSC_OP_RANK) rank_pc = pc;
  }
  if (has_filter && rank_pc < 0) return ctx->fail(MSC_ERR_ARG, "filtered projection needs a RANK instruction");
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  LaunchPlan lp;
  MSC_TRY(plan_launch(ctx, sd, R, 0, &lp));
  uint64_t nout_rows = sd->nrows;
  DevTmp counts(ctx), offsets(ctx);
  if (has_filter && sd->nrows > 0) {
    // pass 1: rows surviving per tile (program truncated after RANK)
    LaunchPlan cp = lp;
    cp.p.code[rank_pc + 2] = MSC_OP_END;
    MSC_TRY(counts.alloc(sizeof(uint32_t) * cp.p.ntiles));
    MSC_TRY(offsets.alloc(sizeof(uint64_t) * (cp.p.ntiles + 1)));
    cp.p.tile_counts = counts.as<uint32_t>();
    MSC_TRY(launch_scan_r<MODE_COUNT>(ctx, &cp));
    MSC_TRY(msc_exclusive_scan_u32_u64(ctx, counts.as<uint32_t>(), offsets.as<uint64_t>(), cp.p.ntiles));
    MSC_CUDA(ctx, cudaMemcpyAsync(&nout_rows, offsets.as<uint64_t>() + cp.p.ntiles, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int drc = msc_check_device_error(ctx);
    if (drc != MSC_OK) return drc;
    lp.p.tile_offsets = offsets.as<uint64_t>();
  }
  msc_rel* rel = new_rel(ctx, nout_rows);
  for (int i = 0; i < nout; ++i) {
    int rc = add_col(ctx, rel, out_phys[i], nout_rows);
    if (rc != MSC_OK) {
      msc_rel_free(rel);
      return rc;
    }
    lp.p.out[i] = rel->cols[i].data;
    lp.p.out_phys[i] = out_phys[i];
  }
  if (sd->nrows > 0 && nout_rows > 0 && nout > 0) {
    int rc = launch_scan_r<MODE_PROJECT>(ctx, &lp);
    if (rc != MSC_OK) {
      msc_rel_free(rel);
      return rc;
    }
  }
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b);
  ctx->stats.last_kernel_ms = ms;
  if (sd->nrows > 0 && cudaEventElapsedTime(&ms, ctx->ev_s0, ctx->ev_s1) == cudaSuccess) ctx->stats.last_scan_ms = ms;
  int drc = msc_check_device_error(ctx);
  if (drc != MSC_OK) {
    msc_rel_free(rel);
    return drc;
  }
  *out = rel;
  return MSC_OK;
}
