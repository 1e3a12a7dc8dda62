// scan_kernel.cuh -- the fused scan -> filter -> project -> aggregate kernel (device side).
//
// One persistent kernel walks the relation in warp tiles of 32 x R rows.  Every warp runs its own
// pipeline: lane 0 bulk-copies the tile's column segments global->shared with cp.async.bulk (TMA
// 1-D, mbarrier complete_tx) into a private multi-stage ring, so HBM latency is hidden by the ring
// and warps never wait for each other.  Each lane owns R consecutive rows and evaluates the
// three-address expression program (include/minispark_cuda.h) for them.  Temporaries are per-lane
// slots in shared memory, so no interpreter state lives in registers between instructions; hot
// instruction shapes run through handlers specialised at C++ compile time (run_fast), everything
// else through the generic fetch / compute / store path.
//
// Replaces, in one pass and without materialising intermediates (reference file:line):
//   FilterTask.execute      src/mini_spark/tasks.py:167-177   (templates/plan.zig:130-147, task_utils.zig:9-51)
//   ProjectTask.execute     src/mini_spark/tasks.py:79-84     (templates/plan.zig:113-125)
//   AggregateTask.execute   src/mini_spark/tasks.py:270-310   (templates/plan.zig:150-253)
//   Col.execute_row         src/mini_spark/sql.py:262-266
#pragma once
#include "common.cuh"
#include "join_table.cuh"

namespace mscan {

constexpr int NT = 128;  // threads per CTA (4 independent warp pipelines)
constexpr int NW = NT / 32;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_HEADER = NW * MAX_STAGES * 8;  // mbarriers: full[warp][stage]

// MODE_RUNS: GROUP BY a sorted key column -- every run of equal keys is a group, its output row is its run number
// MODE_BUILD: the build half of a hash join as a scan -- rows that pass the filters insert (GROUP value = key, row number)
// into a compact join table (join_table.cuh) without being materialised
enum Mode { MODE_DENSE = 0, MODE_HASH = 1, MODE_COUNT = 2, MODE_PROJECT = 3, MODE_RUNS = 4, MODE_BUILD = 5 };

struct StagedCol {
  const unsigned char* base;
  uint32_t width;
  uint32_t smem_off;    // offset inside one warp stage
  int phys;
  uint32_t tile_bytes;  // width * rows per warp tile: bytes of one bulk copy
};

struct ScanParams {
  uint64_t nrows;
  uint32_t ntiles;       // warp tiles of 32*R rows
  uint32_t nstages;
  uint32_t stage_bytes;  // one warp stage (all staged columns)
  uint32_t warp_bytes;   // nstages * stage_bytes + temporaries of one warp
  uint32_t nstaged;
  uint32_t ntemps;
  uint32_t tile_tx_bytes;  // sum of tile_bytes: what one warp tile's mbarrier expects
  uint32_t _pad0;
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
  // one slot = [key, accumulator 0 .. naggs-1] padded to 1 << hshift words: a row touches ONE 32-byte sector when it
  // has up to 3 accumulators (separate key / accumulator arrays cost one random sector each)
  unsigned long long* htbl;
  uint32_t hshift;
  uint64_t hcap;  // slots, a power of two
  // CTA-local pre-aggregation in front of htbl (the north_star's "smem pre-aggregation"): lcap slots (a power of two, 0 = off)
  // of [key] then [naggs accumulators] in shared memory.  A row whose key finds a local slot costs shared-memory atomics
  // only; a key the local table has no room for goes to htbl directly; the CTA folds its slots into htbl when it is done.
  // Few groups make every global accumulator a hot spot (50 groups x 7 sums at sf10: 18 ms of same-address atomics).
  uint32_t lcap;
  uint32_t _pad1;
  // optimistic sizing: [0] counts the keys inserted into htbl, [1] != 0 once there are more than hlimit of them -- the
  // scan then stops early and the host repeats it with a larger table (nullptr: htbl is large enough for any input)
  unsigned long long* hstate;
  unsigned long long hlimit;
  // count / project
  uint32_t* tile_counts;
  const uint64_t* tile_offsets;  // nullptr: no filter, output position = row
  void* out[MSC_VM_MAX_OUT];
  int out_phys[MSC_VM_MAX_OUT];
  const unsigned long long* nrows_dev;  // when set: the row count is the device's (nrows is an upper bound)
  // MODE_RUNS: staged slot of the key column; tile_offsets[] = runs that start before each warp tile; out[0] = key
  // column of the result (i64), out[1 + a] = accumulator column a, already holding its identity
  int run_key_col;
  // MODE_BUILD: the compact join table being filled (header, presence bitmap, 8-byte slots)
  MscJoinTableHeader* jheader;
  uint32_t* jbitmap;
  unsigned long long* jslots;
};

struct LaunchPlan {
  ScanParams p;
  int R;
  size_t smem;
  int grid;
  bool timed = true;  // record ev_s0 / ev_s1 around the launch (off for the follow-up scans of a pending chain)
};

template <int R, int MODE>
int launch_scan(msc_ctx* ctx, LaunchPlan* lp);

constexpr unsigned long long HASH_EMPTY = 0x8000000000000000ULL;
constexpr int LOCAL_PROBES = 4;  // slots a key may sit away from its home in a CTA-local table

#if defined(__CUDACC__) && !defined(MSCAN_DECL_ONLY)
// ------------------------------------------------------------------------------------------------
// mbarrier / bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
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
static __device__ __noinline__ long long py_floordiv_i(long long a, long long b) {
  if (b == 0) return 0;
  long long q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}
static __device__ __noinline__ long long py_mod_i(long long a, long long b) {
  if (b == 0) return 0;
  long long m = a % b;
  if (m != 0 && ((m < 0) != (b < 0))) m += b;
  return m;
}
// CPython float_divmod (Objects/floatobject.c): floor division and modulo of doubles
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

// a = a <op> b for R rows (generic path)
template <int R>
__device__ __forceinline__ void binop(int op, long long (&a)[R], const long long (&b)[R], uint32_t vmask, int* err) {
  switch (op) {
#define F_ARITH(OP, EXPR)                             \
  case OP: {                                          \
    _Pragma("unroll") for (int r = 0; r < R; ++r) {   \
      const double x = l2d(a[r]), y = l2d(b[r]);      \
      a[r] = d2l(EXPR);                               \
    }                                                 \
  } break;
#define I_ARITH(OP, EXPR)                             \
  case OP: {                                          \
    _Pragma("unroll") for (int r = 0; r < R; ++r) {   \
      const long long x = a[r], y = b[r];             \
      a[r] = (EXPR);                                  \
    }                                                 \
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

// comparisons and boolean ops: result is i64 0/1 (generic path)
template <int R>
__device__ __forceinline__ void cmpop(int op, long long (&a)[R], const long long (&b)[R]) {
  switch (op) {
#define F_CMP(OP, REL)                                                                        \
  case OP: {                                                                                  \
    _Pragma("unroll") for (int r = 0; r < R; ++r) a[r] = (l2d(a[r]) REL l2d(b[r])) ? 1 : 0;   \
  } break;
#define I_CMP(OP, REL)                                                              \
  case OP: {                                                                        \
    _Pragma("unroll") for (int r = 0; r < R; ++r) a[r] = (a[r] REL b[r]) ? 1 : 0;   \
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
// column loads.  A lane's R rows are consecutive, so a column segment is read with the widest
// shared-memory vector loads.
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

template <int R, int PHYS>
__device__ __forceinline__ void load_staged(const unsigned char* col_smem, int lane, long long (&dst)[R]) {
  constexpr int WIDTH = (PHYS == MSC_P_U8) ? 1 : (PHYS == MSC_P_U16) ? 2 : (PHYS == MSC_P_I64 || PHYS == MSC_P_F64) ? 8 : 4;
  constexpr int W = WIDTH * R;
  uint32_t w[W / 4];
  lds_words<W>(col_smem + lane * W, w);
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
__device__ __forceinline__ void load_gather(const unsigned char* idx_smem, const void* col, int lane, uint32_t vmask,
                                            long long (&dst)[R]) {
  uint32_t idx[R];
  lds_words<4 * R>(idx_smem + lane * 4 * R, idx);
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

// hash mode: one atomic per run of equal keys among a lane's R consecutive rows (clustered tables,
// e.g. lineitem by orderkey, fold in registers first)
template <int R, int KIND>
__device__ __forceinline__ void agg_hash(unsigned long long* htbl, uint32_t hshift, int a, const int (&grp)[R],
                                         const long long (&v)[R]) {
  unsigned long long* base = htbl + 1 + a;
  long long run = v[0];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int nxt = (r + 1 < R) ? r + 1 : r;
    const bool same_next = (r + 1 < R) && grp[r] >= 0 && grp[nxt] == grp[r];
    if (same_next) {
      run = agg_combine(KIND, run, v[nxt]);
    } else {
      if (grp[r] >= 0) atomic_fold(KIND, base + (static_cast<uint64_t>(grp[r]) << hshift), run);
      run = v[nxt];
    }
  }
}

// the same fold on a shared-memory cell (state space spelled out: a generic atomic on a shared address is far slower)
__device__ __forceinline__ void atomic_fold_shared(int kind, unsigned long long* cell, long long v) {
  const uint32_t addr = smem_u32(cell);
  switch (kind) {
    case MSC_AGG_SUM_F: asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(addr), "d"(l2d(v)) : "memory"); break;
    case MSC_AGG_SUM_I: asm volatile("red.shared.add.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); break;
    case MSC_AGG_MIN_I: asm volatile("red.shared.min.s64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); break;
    case MSC_AGG_MAX_I: asm volatile("red.shared.max.s64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); break;
    default: {  // f64 min / max: CAS loop
      unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(cell);
      while (true) {
        const long long merged = agg_combine(kind, static_cast<long long>(old), v);
        if (static_cast<unsigned long long>(merged) == old) break;
        unsigned long long prev;
        asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(prev) : "r"(addr), "l"(old), "l"(merged) : "memory");
        if (prev == old) break;
        old = prev;
      }
    }
  }
}

// hash mode behind a CTA-local table: grp >= 0 is a LOCAL slot, grp <= -2 the global slot -2 - grp, -1 no group
template <int R, int KIND>
__device__ __forceinline__ void agg_lhash(unsigned long long* lcells, int naggs, unsigned long long* htbl, uint32_t hshift, int a,
                                          const int (&grp)[R], const long long (&v)[R]) {
  long long run = v[0];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int nxt = (r + 1 < R) ? r + 1 : r;
    const bool same_next = (r + 1 < R) && grp[r] != -1 && grp[nxt] == grp[r];
    if (same_next) {
      run = agg_combine(KIND, run, v[nxt]);
    } else {
      if (grp[r] >= 0) atomic_fold_shared(KIND, lcells + grp[r] * naggs + a, run);
      else if (grp[r] != -1) atomic_fold(KIND, htbl + 1 + a + (static_cast<uint64_t>(-2 - grp[r]) << hshift), run);
      run = v[nxt];
    }
  }
}

// exclusive prefix sum of one u32 per lane across the warp; total in *total
__device__ __forceinline__ uint32_t warp_exclusive_scan(uint32_t v, int lane, uint32_t* total) {
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  *total = __shfl_sync(0xffffffffu, inc, 31);
  return inc - v;
}

// ------------------------------------------------------------------------------------------------
// execution context of one warp tile
// ------------------------------------------------------------------------------------------------
struct Ctx {
  const ScanParams& p;
  const unsigned char* sbase;  // this warp's current stage
  long long* temps;            // this warp's temporaries: [(idx * R + r) * 32 + lane]
  long long* acc;              // CTA accumulators: [(group * naggs + slot) * NT + tid]
  int lane;
  int tid;
  uint64_t tile;  // warp tile being evaluated
};

template <int R>
__device__ __forceinline__ void fetch_temp(const Ctx& c, int idx, long long (&v)[R]) {
  const long long* t = c.temps + (idx * R) * 32 + c.lane;
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = t[r * 32];
}
template <int R>
__device__ __forceinline__ void store_temp(const Ctx& c, int idx, const long long (&v)[R]) {
  long long* t = c.temps + (idx * R) * 32 + c.lane;
#pragma unroll
  for (int r = 0; r < R; ++r) t[r * 32] = v[r];
}
template <int R>
__device__ __forceinline__ void to_f64(long long (&v)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = d2l(static_cast<double>(v[r]));
}

template <int R>
__device__ __forceinline__ void fetch_staged(const Ctx& c, int idx, long long (&v)[R]) {
  const unsigned char* col = c.sbase + c.p.staged[idx].smem_off;
  switch (c.p.staged[idx].phys) {
    case MSC_P_U8: load_staged<R, MSC_P_U8>(col, c.lane, v); break;
    case MSC_P_U16: load_staged<R, MSC_P_U16>(col, c.lane, v); break;
    case MSC_P_U32: load_staged<R, MSC_P_U32>(col, c.lane, v); break;
    case MSC_P_I32: load_staged<R, MSC_P_I32>(col, c.lane, v); break;
    case MSC_P_F32: load_staged<R, MSC_P_F32>(col, c.lane, v); break;
    default: load_staged<R, MSC_P_I64>(col, c.lane, v); break;  // I64 and F64: raw 64-bit pattern
  }
}

template <int R>
__device__ __forceinline__ void fetch_gather(const Ctx& c, int idx, uint32_t vmask, long long (&v)[R]) {
  const unsigned char* ix = c.sbase + c.p.staged[idx >> 6].smem_off;
  const void* col = c.p.gather[idx & 63];
  switch (c.p.gather_phys[idx & 63]) {
    case MSC_P_U8: load_gather<R, MSC_P_U8>(ix, col, c.lane, vmask, v); break;
    case MSC_P_U16: load_gather<R, MSC_P_U16>(ix, col, c.lane, vmask, v); break;
    case MSC_P_U32: load_gather<R, MSC_P_U32>(ix, col, c.lane, vmask, v); break;
    case MSC_P_I32: load_gather<R, MSC_P_I32>(ix, col, c.lane, vmask, v); break;
    case MSC_P_F32: load_gather<R, MSC_P_F32>(ix, col, c.lane, vmask, v); break;
    default: load_gather<R, MSC_P_I64>(ix, col, c.lane, vmask, v); break;
  }
}

// Load through per-row indices held in a temporary (the build-side row a PROBE found): dst[r] = column[temp[r]].
template <int R>
__device__ __forceinline__ void fetch_gather_temp(const Ctx& c, int idx, uint32_t vmask, long long (&v)[R]) {
  long long ix[R];
  fetch_temp<R>(c, idx >> 6, ix);
  const void* col = c.p.gather[idx & 63];
  const int phys = c.p.gather_phys[idx & 63];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    long long x = 0;
    if (((vmask >> r) & 1u) && ix[r] >= 0) {
      const uint64_t i = static_cast<uint64_t>(ix[r]);
      switch (phys) {
        case MSC_P_U8: x = __ldg(reinterpret_cast<const uint8_t*>(col) + i); break;
        case MSC_P_U16: x = __ldg(reinterpret_cast<const uint16_t*>(col) + i); break;
        case MSC_P_U32: x = __ldg(reinterpret_cast<const uint32_t*>(col) + i); break;
        case MSC_P_I32: x = __ldg(reinterpret_cast<const int*>(col) + i); break;
        case MSC_P_F32: x = d2l(static_cast<double>(__ldg(reinterpret_cast<const float*>(col) + i))); break;
        default: x = __ldg(reinterpret_cast<const long long*>(col) + i); break;
      }
    }
    v[r] = x;
  }
}

// MSC_OP_PROBE: x[r] <- build-side row of key x[r] in a join table (msc_join_build), or -1.  The lane's R lookups advance
// together, one table sector per row and step, so a tile waits for its longest probe sequence and not for R in turn.
template <int R>
__device__ __forceinline__ void op_probe(long long (&x)[R], const void* table, uint32_t vmask) {
  const MscJoinTableHeader* h = static_cast<const MscJoinTableHeader*>(table);
  const uint64_t mask = h->cap - 1;
  uint64_t pos[R];
  bool pend[R];
  if (h->slot_bytes == 8) {  // compact: u64 = build row << 32 | (u32)key
    const unsigned long long* slots = reinterpret_cast<const unsigned long long*>(msc_join_slots(h));
    uint32_t key[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      pend[r] = ((vmask >> r) & 1u) && msc_join_key_is_32bit(x[r]);
      key[r] = static_cast<uint32_t>(x[r]);
      pos[r] = msc_fmix32(key[r]) & mask;
      x[r] = -1;
    }
    while (true) {
      unsigned long long raw[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (pend[r]) raw[r] = __ldg(slots + pos[r]);
      bool more = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (!pend[r]) continue;
        if (raw[r] == MSC_J_EMPTY8) {
          pend[r] = false;
        } else if (static_cast<uint32_t>(raw[r]) == key[r]) {
          x[r] = static_cast<long long>(raw[r] >> 32);
          pend[r] = false;
        } else {
          pos[r] = (pos[r] + 1) & mask;
          more = true;
        }
      }
      if (!more) break;
    }
    return;
  }
  const uint4* slots = reinterpret_cast<const uint4*>(msc_join_slots(h));
  unsigned long long key[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    key[r] = msc_join_norm_key(x[r]);
    pend[r] = (vmask >> r) & 1u;
    pos[r] = msc_mix64(key[r]) & mask;
    x[r] = -1;
  }
  while (true) {
    uint4 raw[R];
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (pend[r]) raw[r] = __ldg(slots + pos[r]);
    bool more = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (!pend[r]) continue;
      const unsigned long long cur = (static_cast<unsigned long long>(raw[r].y) << 32) | raw[r].x;
      if (cur == key[r]) {
        x[r] = static_cast<long long>(raw[r].z);
        pend[r] = false;
      } else if (cur == MSC_J_EMPTY) {
        pend[r] = false;
      } else {
        pos[r] = (pos[r] + 1) & mask;
        more = true;
      }
    }
    if (!more) break;
  }
}

// generic operand fetch: the i2f variants are separate switch targets (a flag test would be
// if-converted into R predicated 64-bit conversions on every fetch)
template <int R>
__device__ __forceinline__ void fetch(const Ctx& c, uint32_t operand, uint32_t vmask, long long (&v)[R]) {
  const int idx = operand & 0xfff;
  switch ((operand >> 12) & 15) {
    case MSC_SRC_TEMP: fetch_temp<R>(c, idx, v); break;
    case MSC_SRC_TEMP | MSC_SRC_I2F: fetch_temp<R>(c, idx, v); to_f64<R>(v); break;
    case MSC_SRC_STAGED: fetch_staged<R>(c, idx, v); break;
    case MSC_SRC_STAGED | MSC_SRC_I2F: fetch_staged<R>(c, idx, v); to_f64<R>(v); break;
    case MSC_SRC_GATHER: fetch_gather<R>(c, idx, vmask, v); break;
    case MSC_SRC_GATHER | MSC_SRC_I2F: fetch_gather<R>(c, idx, vmask, v); to_f64<R>(v); break;
    case MSC_SRC_GATHER_T: fetch_gather_temp<R>(c, idx, vmask, v); break;
    case MSC_SRC_GATHER_T | MSC_SRC_I2F: fetch_gather_temp<R>(c, idx, vmask, v); to_f64<R>(v); break;
    case MSC_SRC_CONST: {
      const long long k = c.p.consts[idx];
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = k;
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

template <int R, int KIND>
__device__ __forceinline__ void agg_dense_k(const Ctx& c, int slot, const int (&grp)[R], const long long (&v)[R]) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    long long* q = c.acc + (grp[r] * c.p.naggs + slot) * NT + c.tid;
    *q = agg_combine(KIND, *q, v[r]);
  }
}

// local table of a hash scan, at the CTA's accumulator area: keys[lcap], cells[lcap][naggs], then two counters
__device__ __forceinline__ unsigned long long* local_keys(const Ctx& c) { return reinterpret_cast<unsigned long long*>(c.acc); }
__device__ __forceinline__ unsigned long long* local_cells(const Ctx& c) { return reinterpret_cast<unsigned long long*>(c.acc) + c.p.lcap; }
__device__ __forceinline__ uint32_t* local_flags(const Ctx& c) {  // [0] inserts that found no room, [1] != 0: stop trying
  return reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned long long*>(c.acc) + static_cast<size_t>(c.p.lcap) * (1 + c.p.naggs));
}

template <int R, int KIND>
__device__ __forceinline__ void agg_hash_any(const Ctx& c, int a, const int (&grp)[R], const long long (&v)[R]) {
  if (c.p.lcap) agg_lhash<R, KIND>(local_cells(c), c.p.naggs, c.p.htbl, c.p.hshift, a, grp, v);
  else agg_hash<R, KIND>(c.p.htbl, c.p.hshift, a, grp, v);
}

template <int R, int MODE>
__device__ __forceinline__ void agg_any(const Ctx& c, int slot, int kind, const int (&grp)[R], const long long (&v)[R]) {
  switch (kind) {
#define AGG_ANY_CASE(KIND)                                                                          \
  case KIND:                                                                                        \
    if constexpr (MODE == MODE_DENSE) agg_dense_k<R, KIND>(c, slot, grp, v);                        \
    else if constexpr (MODE == MODE_HASH) agg_hash_any<R, KIND>(c, slot, grp, v);                    \
    else if constexpr (MODE == MODE_RUNS) agg_hash<R, KIND>(static_cast<unsigned long long*>(c.p.out[1 + slot]) - 1, 0, 0, grp, v); \
    break;
    AGG_ANY_CASE(MSC_AGG_SUM_F)
    AGG_ANY_CASE(MSC_AGG_SUM_I)
    AGG_ANY_CASE(MSC_AGG_MIN_F)
    AGG_ANY_CASE(MSC_AGG_MAX_F)
    AGG_ANY_CASE(MSC_AGG_MIN_I)
    default:
      if constexpr (MODE == MODE_DENSE) agg_dense_k<R, MSC_AGG_MAX_I>(c, slot, grp, v);
      else if constexpr (MODE == MODE_HASH) agg_hash_any<R, MSC_AGG_MAX_I>(c, slot, grp, v);
      else if constexpr (MODE == MODE_RUNS) agg_hash<R, MSC_AGG_MAX_I>(static_cast<unsigned long long*>(c.p.out[1 + slot]) - 1, 0, 0, grp, v);
      break;
#undef AGG_ANY_CASE
  }
}

template <int R, int MODE>
__device__ __forceinline__ void set_group(const Ctx& c, const long long (&x)[R], uint32_t vmask, int (&grp)[R]) {
  if constexpr (MODE == MODE_DENSE) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int g = c.p.ngroups;  // rows that failed the filter fold into a trash group that is never exported
      if ((vmask >> r) & 1u) {
        const long long code = x[r];
        g = (code >= 0 && code < c.p.ngroups) ? static_cast<int>(code) : c.p.ngroups;
      }
      grp[r] = g;
      c.acc[(g * c.p.naggs + (c.p.naggs - 1)) * NT + c.tid] += 1;  // hidden per-group row counter
    }
  } else if constexpr (MODE == MODE_HASH) {
    // A lane's rows are consecutive, so equal adjacent keys (clustered tables) share one lookup.  Find-or-insert is ONE
    // atomicCAS(slot, EMPTY, key) per probe step -- the old value says "inserted", "found" or "someone else's" -- and the
    // steps of a lane's R rows are issued together: R independent atomics in flight, then R checks, then the next step
    // for the rows that collided.  A tile therefore waits for as many round trips to the table as its longest probe
    // sequence, not for load -> CAS -> walk of every row in turn (prof_hash: 13 long-scoreboard stall cycles per issued
    // instruction, 3.2 ms for sf10 GROUP BY l_orderkey).
    const uint64_t mask = c.p.hcap - 1;
    unsigned long long key[R];
    uint64_t pos[R];
    bool head[R], pend[R];
    int slot[R];
    const bool local = c.p.lcap != 0;
    const bool local_open = local && *reinterpret_cast<volatile uint32_t*>(local_flags(c) + 1) == 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      key[r] = static_cast<unsigned long long>(x[r]);
      if (key[r] == HASH_EMPTY) key[r] = 0;  // -0.0 groups with +0.0, like a Python dict
      const bool valid = (vmask >> r) & 1u;
      const bool prev_valid = r > 0 && ((vmask >> (r > 0 ? r - 1 : 0)) & 1u);
      head[r] = valid && !(prev_valid && key[r] == key[r > 0 ? r - 1 : 0]);
      pend[r] = head[r];
      const uint64_t h = msc_mix64(key[r]);
      pos[r] = h & mask;
      slot[r] = -1;
      if (local_open && head[r]) {
        // the CTA's own table first: a read finds a key that is already there, one CAS claims an empty slot
        unsigned long long* lkeys = local_keys(c);
        const uint32_t lmask = c.p.lcap - 1;
        uint32_t lp = static_cast<uint32_t>(h >> 32) & lmask;
        for (int step = 0; step < LOCAL_PROBES; ++step, lp = (lp + 1) & lmask) {
          unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(lkeys + lp);
          if (cur == HASH_EMPTY) asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(cur) : "r"(smem_u32(lkeys + lp)), "l"(HASH_EMPTY), "l"(key[r]) : "memory");
          if (cur == HASH_EMPTY || cur == key[r]) {
            slot[r] = static_cast<int>(lp);
            pend[r] = false;
            break;
          }
        }
        // no room near its home slot: this row goes to the global table; a CTA that keeps failing stops trying (more groups
        // than local slots: the local probes would only cost time)
        if (pend[r] && atomicAdd(local_flags(c), 1u) > 4u * c.p.lcap) *reinterpret_cast<volatile uint32_t*>(local_flags(c) + 1) = 1u;
      }
    }
    bool any = false;
#pragma unroll
    for (int r = 0; r < R; ++r) any |= pend[r];
    if (any) {
      for (uint64_t step = 0;; ++step) {
        unsigned long long got[R];
#pragma unroll
        for (int r = 0; r < R; ++r) got[r] = pend[r] ? atomicCAS(c.p.htbl + (pos[r] << c.p.hshift), HASH_EMPTY, key[r]) : 0ull;
        bool more = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (!pend[r]) continue;
          if (got[r] == HASH_EMPTY || got[r] == key[r]) {
            slot[r] = local ? -2 - static_cast<int>(pos[r]) : static_cast<int>(pos[r]);
            pend[r] = false;
            if (got[r] == HASH_EMPTY && c.p.hstate != nullptr && atomicAdd(c.p.hstate, 1ull) >= c.p.hlimit)
              *reinterpret_cast<volatile unsigned long long*>(c.p.hstate + 1) = 1ull;  // more groups than this table was sized for
          } else {
            pos[r] = (pos[r] + 1) & mask;
            more = true;
          }
        }
        if (!more) break;
        if (step >= 32 && c.p.hstate != nullptr && *reinterpret_cast<volatile unsigned long long*>(c.p.hstate + 1) != 0) break;  // (the scan is being abandoned)
        if (step >= c.p.hcap) {  // every slot belongs to another key
          atomicOr(c.p.err, MSC_DEVERR_TABLE_FULL);
          break;
        }
      }
    }
    int prev_slot = -1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int sl = -1;
      if ((vmask >> r) & 1u) {
        sl = head[r] ? slot[r] : prev_slot;
        prev_slot = sl;
      }
      grp[r] = sl;
    }
  } else if constexpr (MODE == MODE_BUILD) {
    // key -> row number into the compact table (msc_join_build's format: 8-byte slots row << 32 | (u32)key, presence bitmap).
    // A key outside the 32-bit range or a second row with the same key only raises its flag: the host then builds the
    // general way.
    const uint64_t row0 = c.tile * (32ull * R) + static_cast<uint64_t>(c.lane) * R;
    const uint64_t mask = c.p.hcap - 1, bmask = c.p.jheader->bitmap_bits - 1;
    uint32_t k32[R];
    uint64_t pos[R];
    bool pend[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      grp[r] = -1;
      pend[r] = (vmask >> r) & 1u;
      if (pend[r] && !msc_join_key_is_32bit(x[r])) {
        c.p.jheader->wide_keys = 1;
        pend[r] = false;
      }
      k32[r] = static_cast<uint32_t>(x[r]);
      pos[r] = msc_fmix32(k32[r]) & mask;
      if (pend[r]) {
        const uint64_t bit = msc_fmix32(k32[r] ^ 0x9e3779b9u) & bmask;
        atomicOr(&c.p.jbitmap[bit >> 5], 1u << (bit & 31));
      }
    }
    while (true) {  // the lane's R inserts advance together: R atomics in flight per step (cf. the hash aggregate above)
      unsigned long long prev[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (pend[r]) prev[r] = atomicCAS(&c.p.jslots[pos[r]], MSC_J_EMPTY8, ((row0 + r) << 32) | k32[r]);
      bool more = false;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (!pend[r]) continue;
        if (prev[r] == MSC_J_EMPTY8) {
          pend[r] = false;
        } else if (static_cast<uint32_t>(prev[r]) == k32[r]) {
          c.p.jheader->duplicates = 1;
          pend[r] = false;
        } else {
          pos[r] = (pos[r] + 1) & mask;
          more = true;
        }
      }
      if (!more) break;
    }
  } else if constexpr (MODE == MODE_RUNS) {
    // Sorted key column (the host checked): a row's group is the number of runs that started at or before it, minus one.
    // Runs before the tile come from a prefix sum over per-tile counts (run_heads_kernel in scan.cu), runs before the
    // lane from a warp scan; a row starts a run when its key differs from the previous ROW's, which for a lane's first
    // row is the previous lane's last (a shuffle) and for the tile's first row the column element before the tile.
    constexpr int WT = 32 * R;
    const uint64_t row0 = c.tile * WT + static_cast<uint64_t>(c.lane) * R;
    long long prev = __shfl_up_sync(0xffffffffu, x[R - 1], 1);
    if (c.lane == 0 && row0 > 0) {
      const StagedCol& kc = c.p.staged[c.p.run_key_col];
      const unsigned char* e = kc.base + (row0 - 1) * kc.width;
      switch (kc.phys) {
        case MSC_P_U8: prev = *e; break;
        case MSC_P_U16: prev = *reinterpret_cast<const uint16_t*>(e); break;
        case MSC_P_U32: prev = *reinterpret_cast<const uint32_t*>(e); break;
        case MSC_P_I32: prev = *reinterpret_cast<const int*>(e); break;
        default: prev = *reinterpret_cast<const long long*>(e); break;
      }
    }
    bool head[R];
    uint32_t mine = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const bool valid = (vmask >> r) & 1u;
      head[r] = valid && (row0 + r == 0 || x[r] != (r > 0 ? x[r > 0 ? r - 1 : 0] : prev));
      mine += head[r];
    }
    uint32_t total;
    uint64_t run = c.p.tile_offsets[c.tile] + warp_exclusive_scan(mine, c.lane, &total);  // runs started before this lane's rows
    long long* out_key = static_cast<long long*>(c.p.out[0]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (head[r]) out_key[run++] = x[r];
      grp[r] = ((vmask >> r) & 1u) ? static_cast<int>(run) - 1 : -1;  // the run this row belongs to
    }
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
// host picks the shape id (`fast` field of w0, include/minispark_cuda.h MSC_FAST_*) and rewrites
// staged operand indices into (shared-memory offset / 16), so a handler needs no table lookups.
// The ids are dense, so the switch below compiles to a single jump table.
// ------------------------------------------------------------------------------------------------
template <int R, int FK>
__device__ __forceinline__ void ffetch(const Ctx& c, int idx, long long (&v)[R]) {
  if constexpr (FK == MSC_FK_TEMP) {
    fetch_temp<R>(c, idx, v);
  } else if constexpr (FK == MSC_FK_CONST) {
    const long long k = c.p.consts[idx];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = k;
  } else {
    const unsigned char* col = c.sbase + (idx << 4);
    if constexpr (FK == MSC_FK_F32) load_staged<R, MSC_P_F32>(col, c.lane, v);
    else if constexpr (FK == MSC_FK_F64 || FK == MSC_FK_I64) load_staged<R, MSC_P_I64>(col, c.lane, v);
    else if constexpr (FK == MSC_FK_I32) load_staged<R, MSC_P_I32>(col, c.lane, v);
    else if constexpr (FK == MSC_FK_I32F) {
      load_staged<R, MSC_P_I32>(col, c.lane, v);
#pragma unroll
      for (int r = 0; r < R; ++r) v[r] = d2l(static_cast<double>(static_cast<int>(v[r])));
    } else if constexpr (FK == MSC_FK_U8) load_staged<R, MSC_P_U8>(col, c.lane, v);
    else if constexpr (FK == MSC_FK_U16) load_staged<R, MSC_P_U16>(col, c.lane, v);
    else load_staged<R, MSC_P_U32>(col, c.lane, v);
  }
}

template <int R, int MODE, int KIND>
__device__ __forceinline__ void fast_agg(const Ctx& c, int slot, const int (&grp)[R], const long long (&v)[R]) {
  if constexpr (MODE == MODE_DENSE) agg_dense_k<R, KIND>(c, slot, grp, v);
  else if constexpr (MODE == MODE_HASH) agg_hash_any<R, KIND>(c, slot, grp, v);
  // a run's accumulator is element [run] of its output column: agg_hash with one-word slots and no key word
  else if constexpr (MODE == MODE_RUNS) agg_hash<R, KIND>(static_cast<unsigned long long*>(c.p.out[1 + slot]) - 1, 0, 0, grp, v);
}

// dst <- A (+|-|*) B on f64; DK: 0 TEMP, 1 AGG(SUM_F), 2 AGG(SUM_F) + tee TEMP
template <int R, int MODE, int OPI, int AK, int BK, int DK>
__device__ __forceinline__ void fast_arith(const Ctx& c, uint32_t w0, uint32_t w1, const int (&grp)[R]) {
  long long a[R], b[R];
  ffetch<R, AK>(c, w1 & 0xfff, a);
  ffetch<R, BK>(c, (w1 >> 16) & 0xfff, b);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const double x = l2d(a[r]), y = l2d(b[r]);
    a[r] = d2l(OPI == 0 ? x + y : (OPI == 1 ? x - y : x * y));
  }
  const int dst = (w0 >> 13) & 0x7f;
  if constexpr (DK == 0) store_temp<R>(c, dst, a);
  if constexpr (DK == 2) store_temp<R>(c, static_cast<int>((w0 >> 9) & 0xf) - 1, a);
  if constexpr (DK >= 1) fast_agg<R, MODE, MSC_AGG_SUM_F>(c, dst, grp, a);
}

template <int R, int MODE, int KIND, int FK>
__device__ __forceinline__ void fast_aggmov(const Ctx& c, uint32_t w0, uint32_t w1, const int (&grp)[R]) {
  long long a[R];
  ffetch<R, FK>(c, w1 & 0xfff, a);
  fast_agg<R, MODE, KIND>(c, (w0 >> 13) & 0x7f, grp, a);
}

template <int R, int CMPI, int FK>
__device__ __forceinline__ void fast_cmp_filter(const Ctx& c, uint32_t w1, uint32_t& vmask) {
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
__device__ __forceinline__ void fast_group(const Ctx& c, uint32_t w1, uint32_t vmask, int (&grp)[R]) {
  long long a[R];
  ffetch<R, FK>(c, w1 & 0xfff, a);
  set_group<R, MODE>(c, a, vmask, grp);
}

template <int R, int FK, int U32OUT>
__device__ __forceinline__ void fast_out(const Ctx& c, uint32_t w0, uint32_t w1, uint64_t out_pos, uint32_t vmask) {
  long long a[R];
  ffetch<R, FK>(c, w1 & 0xfff, a);
  void* out = c.p.out[(w0 >> 13) & 0x7f];
  if constexpr (U32OUT) store_out<R, uint32_t>(a, reinterpret_cast<uint32_t*>(out), out_pos, vmask);
  else store_out<R, long long>(a, reinterpret_cast<long long*>(out), out_pos, vmask);
}

__host__ __device__ constexpr bool fk_is_float(int fk) { return fk == MSC_FK_F32 || fk == MSC_FK_F64 || fk == MSC_FK_I32F; }
__host__ __device__ constexpr bool fk_is_int_col(int fk) { return fk == MSC_FK_I32 || fk == MSC_FK_I64; }
__host__ __device__ constexpr bool fk_is_code(int fk) { return fk == MSC_FK_U8 || fk == MSC_FK_U16 || fk == MSC_FK_U32; }
__host__ __device__ constexpr bool aggmov_valid(int kind, int fk) {
  const bool fkind = kind == MSC_AGG_SUM_F || kind == MSC_AGG_MIN_F || kind == MSC_AGG_MAX_F;
  if (fk == MSC_FK_TEMP) return true;
  if (fkind) return fk_is_float(fk);
  return fk_is_int_col(fk) || (fk == MSC_FK_CONST && kind == MSC_AGG_SUM_I);
}
__host__ __device__ constexpr bool cmp_valid(int fk) { return fk != MSC_FK_TEMP && fk != MSC_FK_CONST; }
__host__ __device__ constexpr bool group_valid(int fk) { return fk == MSC_FK_TEMP || fk_is_int_col(fk) || fk_is_code(fk); }
__host__ __device__ constexpr bool out_valid(int fk, int u32) {
  if (fk == MSC_FK_CONST) return false;
  return u32 ? (fk_is_code(fk) || fk == MSC_FK_TEMP) : !fk_is_code(fk);
}

// returns false when the id is not a compiled shape (the caller then runs the generic path)
template <int R, int MODE>
__device__ __forceinline__ bool run_fast(const Ctx& c, int fast, uint32_t w0, uint32_t w1, uint32_t& vmask, int (&grp)[R],
                                         uint64_t out_pos) {
  constexpr bool AGG = MODE == MODE_DENSE || MODE == MODE_HASH || MODE == MODE_RUNS;  // (MODE_BUILD takes its GROUP through the generic path)
  switch (fast) {
#define ARITH_CASE(OPI, AK, BK, DK)                                                      \
  case MSC_FAST_ARITH + (((OPI) * 5 + (AK)) * 5 + (BK)) * 3 + (DK):                      \
    if constexpr (AGG || DK == 0) {                                                      \
      fast_arith<R, MODE, OPI, AK, BK, DK>(c, w0, w1, grp);                              \
      return true;                                                                       \
    } else {                                                                             \
      return false;                                                                      \
    }
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
#define FK_ALL(M, X) M(X, 0) M(X, 1) M(X, 2) M(X, 3) M(X, 4) M(X, 5) M(X, 6) M(X, 7) M(X, 8) M(X, 9)
#define AGGMOV_CASE(KIND, FK)                                     \
  case MSC_FAST_AGGMOV + (KIND) * 10 + (FK):                      \
    if constexpr (AGG && aggmov_valid(KIND, FK)) {                \
      fast_aggmov<R, MODE, KIND, FK>(c, w0, w1, grp);             \
      return true;                                                \
    } else {                                                      \
      return false;                                               \
    }
    FK_ALL(AGGMOV_CASE, 0) FK_ALL(AGGMOV_CASE, 1) FK_ALL(AGGMOV_CASE, 2) FK_ALL(AGGMOV_CASE, 3) FK_ALL(AGGMOV_CASE, 4)
    FK_ALL(AGGMOV_CASE, 5)
#undef AGGMOV_CASE
#define CMP_CASE(CMPI, FK)                                        \
  case MSC_FAST_CMP + (CMPI) * 10 + (FK):                         \
    if constexpr (cmp_valid(FK)) {                                \
      fast_cmp_filter<R, CMPI, FK>(c, w1, vmask);                 \
      return true;                                                \
    } else {                                                      \
      return false;                                               \
    }
    FK_ALL(CMP_CASE, 0) FK_ALL(CMP_CASE, 1) FK_ALL(CMP_CASE, 2) FK_ALL(CMP_CASE, 3) FK_ALL(CMP_CASE, 4) FK_ALL(CMP_CASE, 5)
#undef CMP_CASE
#define GROUP_CASE(UNUSED, FK)                                    \
  case MSC_FAST_GROUP + (FK):                                     \
    if constexpr ((AGG || MODE == MODE_BUILD) && group_valid(FK)) { \
      fast_group<R, MODE, FK>(c, w1, vmask, grp);                 \
      return true;                                                \
    } else {                                                      \
      return false;                                               \
    }
    FK_ALL(GROUP_CASE, 0)
#undef GROUP_CASE
#define OUT_CASE(U32OUT, FK)                                      \
  case MSC_FAST_OUT + (FK) * 2 + (U32OUT):                        \
    if constexpr (MODE == MODE_PROJECT && out_valid(FK, U32OUT)) {\
      fast_out<R, FK, U32OUT>(c, w0, w1, out_pos, vmask);         \
      return true;                                                \
    } else {                                                      \
      return false;                                               \
    }
    FK_ALL(OUT_CASE, 0) FK_ALL(OUT_CASE, 1)
#undef OUT_CASE
#undef FK_ALL
    default: return false;
  }
}

// Start the bulk copies of warp tile `tile` into ring slot `stage`: lane 0 arms the mbarrier with the
// tile's byte count, then lane c starts column c's copy (one cp.async.bulk per referenced column).
__device__ __forceinline__ void issue_tile(const ScanParams& p, unsigned char* stages, uint64_t* full, uint32_t stage,
                                           uint64_t tile, int lane) {
  if (lane == 0) mbar_expect_tx(&full[stage], p.tile_tx_bytes);
  __syncwarp();
  if (lane < static_cast<int>(p.nstaged)) {
    const StagedCol& c = p.staged[lane];
    bulk_g2s(stages + stage * p.stage_bytes + c.smem_off, c.base + tile * c.tile_bytes, c.tile_bytes, &full[stage]);
  }
}

// validity bits of a lane's R consecutive rows starting at row0
template <int R>
__device__ __forceinline__ uint32_t row_mask(uint64_t row0, uint64_t nrows) {
  if (row0 + R <= nrows) return (1u << R) - 1u;
  uint32_t m = 0;
#pragma unroll
  for (int r = 0; r < R; ++r)
    if (row0 + r < nrows) m |= 1u << r;
  return m;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int R, int MODE>
__global__ void __launch_bounds__(NT) scan_kernel(const __grid_constant__ ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr uint32_t WT = 32 * R;  // rows per warp tile
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem) + warp * MAX_STAGES;
  unsigned char* stages = smem + SMEM_HEADER + static_cast<size_t>(warp) * p.warp_bytes;
  long long* temps = reinterpret_cast<long long*>(stages + static_cast<size_t>(p.nstages) * p.stage_bytes);
  long long* acc = reinterpret_cast<long long*>(smem + SMEM_HEADER + static_cast<size_t>(NW) * p.warp_bytes);

  if (lane == 0) {
    for (uint32_t st = 0; st < p.nstages; ++st) mbar_init(&full[st], 1);
    mbar_fence_init();
  }
  if constexpr (MODE == MODE_DENSE) {
    const int cells = (p.ngroups + 1) * p.naggs;
    for (int c = 0; c < cells; ++c) acc[c * NT + tid] = p.agg_init[c % p.naggs];
  }
  if constexpr (MODE == MODE_HASH) {
    if (p.lcap) {
      unsigned long long* lkeys = reinterpret_cast<unsigned long long*>(acc);
      for (uint32_t i = tid; i < p.lcap; i += NT) lkeys[i] = HASH_EMPTY;
      for (uint32_t i = tid; i < p.lcap * p.naggs; i += NT) lkeys[p.lcap + i] = static_cast<unsigned long long>(p.agg_init[i % p.naggs]);
      if (tid < 2) reinterpret_cast<uint32_t*>(lkeys + static_cast<size_t>(p.lcap) * (1 + p.naggs))[tid] = 0;
    }
  }
  __syncthreads();

  const uint32_t gw = blockIdx.x * NW + warp;  // global warp id
  const uint32_t nw = gridDim.x * NW;
  uint32_t ntiles_w = (p.ntiles > gw) ? (p.ntiles - gw + nw - 1) / nw : 0;
  {
    const uint32_t pre = ntiles_w < p.nstages ? ntiles_w : p.nstages;
    for (uint32_t k = 0; k < pre; ++k) issue_tile(p, stages, full, k, gw + static_cast<uint64_t>(k) * nw, lane);
  }
  const uint64_t nrows = p.nrows_dev ? *p.nrows_dev : p.nrows;  // tiles past the device's count are fully masked

  uint32_t stage = 0, parity = 0;
  for (uint32_t k = 0; k < ntiles_w; ++k) {
    const uint64_t tile = gw + static_cast<uint64_t>(k) * nw;
    const unsigned char* sbase = stages + static_cast<size_t>(stage) * p.stage_bytes;
    while (!mbar_try_wait(&full[stage], parity)) {
    }
    if constexpr (MODE == MODE_HASH) {
      // the table turned out too small for this input: take what is already on its way into shared memory and leave
      if (p.hstate != nullptr && *reinterpret_cast<volatile unsigned long long*>(p.hstate + 1) != 0 && ntiles_w > k + p.nstages) ntiles_w = k + p.nstages;
    }
    const uint64_t row0 = tile * WT + static_cast<uint64_t>(lane) * R;
    uint32_t vmask = row_mask<R>(row0, nrows);
    int grp[R];
#pragma unroll
    for (int r = 0; r < R; ++r) grp[r] = (MODE == MODE_DENSE) ? p.ngroups : -1;
    uint64_t out_pos = row0;  // project: output position of this lane's first surviving row
    const Ctx c{p, sbase, temps, acc, lane, tid, tile};

    for (int pc = 0;; pc += 2) {
      const uint32_t w0 = p.code[pc];
      const int op = w0 & 0x3f;
      if (op == MSC_OP_END) break;
      const uint32_t w1 = p.code[pc + 1];
      const int fast = w0 >> 20;
      if (fast != 0 && run_fast<R, MODE>(c, fast, w0, w1, vmask, grp, out_pos)) continue;
      if (op == MSC_OP_RANK) {
        if constexpr (MODE == MODE_COUNT) {
          uint32_t total;
          (void)warp_exclusive_scan(__popc(vmask), lane, &total);
          if (lane == 0) p.tile_counts[tile] = total;
          break;  // everything after RANK computes output columns: nothing the count pass needs
        } else if constexpr (MODE == MODE_PROJECT) {
          if (p.tile_offsets != nullptr) {
            uint32_t total;
            out_pos = p.tile_offsets[tile] + warp_exclusive_scan(__popc(vmask), lane, &total);
          }
        }
        continue;
      }
      long long a[R];
      fetch<R>(c, w1 & 0xffffu, vmask, a);
      if (op >= MSC_OP_ADD_F && op <= MSC_OP_OR) {
        long long b[R];
        fetch<R>(c, w1 >> 16, vmask, b);
        if (op <= MSC_OP_MOD_I) binop<R>(op, a, b, vmask, p.err);
        else cmpop<R>(op, a, b);
      } else if (op == MSC_OP_LUT8) {
        op_lut<R, uint8_t>(a, reinterpret_cast<const uint8_t*>(p.luts[(w1 >> 16) & 0xfff]), vmask);
      } else if (op == MSC_OP_LUT32) {
        op_lut<R, uint32_t>(a, reinterpret_cast<const uint32_t*>(p.luts[(w1 >> 16) & 0xfff]), vmask);
      } else if (op == MSC_OP_PROBE) {
        op_probe<R>(a, p.luts[(w1 >> 16) & 0x7ff], vmask);  // (bit 11 = MSC_PROBE_COMPACT: a promise to the kernel generator)
      }
      const int tee = (w0 >> 9) & 0xf;
      if (tee) store_temp<R>(c, tee - 1, a);
      const int dst = (w0 >> 13) & 0x7f;
      switch ((w0 >> 6) & 7) {
        case MSC_DST_TEMP: store_temp<R>(c, dst, a); break;
        case MSC_DST_FILTER: {
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (a[r] == 0) vmask &= ~(1u << r);
        } break;
        case MSC_DST_GROUP: set_group<R, MODE>(c, a, vmask, grp); break;
        case MSC_DST_AGG: agg_any<R, MODE>(c, dst, p.agg_kind[dst], grp, a); break;
        case MSC_DST_OUT:
          if constexpr (MODE == MODE_PROJECT) {
            if (p.out_phys[dst] == MSC_P_U32) store_out<R, uint32_t>(a, reinterpret_cast<uint32_t*>(p.out[dst]), out_pos, vmask);
            else store_out<R, long long>(a, reinterpret_cast<long long*>(p.out[dst]), out_pos, vmask);
          }
          break;
        default: break;
      }
    }

    __syncwarp();  // every lane is done reading this stage
    if (k + p.nstages < ntiles_w) issue_tile(p, stages, full, stage, gw + static_cast<uint64_t>(k + p.nstages) * nw, lane);
    if (++stage == p.nstages) {
      stage = 0;
      parity ^= 1u;
    }
  }

  if constexpr (MODE == MODE_HASH) {
    if (p.lcap) {
      // fold the CTA's table into the global one: a key costs one find-or-insert and one atomic per accumulator that moved
      __syncthreads();
      const unsigned long long* lkeys = reinterpret_cast<const unsigned long long*>(acc);
      const unsigned long long* lcells = lkeys + p.lcap;
      const bool abandoned = p.hstate != nullptr && *reinterpret_cast<volatile unsigned long long*>(p.hstate + 1) != 0;
      const uint64_t mask = p.hcap - 1;
      for (uint32_t s = tid; s < p.lcap && !abandoned; s += NT) {
        const unsigned long long key = lkeys[s];
        if (key == HASH_EMPTY) continue;
        uint64_t pos = msc_mix64(key) & mask;
        bool placed = false;
        for (uint64_t step = 0; step <= p.hcap; ++step, pos = (pos + 1) & mask) {
          const unsigned long long got = atomicCAS(p.htbl + (pos << p.hshift), HASH_EMPTY, key);
          if (got == HASH_EMPTY || got == key) {
            placed = true;
            if (got == HASH_EMPTY && p.hstate != nullptr && atomicAdd(p.hstate, 1ull) >= p.hlimit)
              *reinterpret_cast<volatile unsigned long long*>(p.hstate + 1) = 1ull;
            break;
          }
          if (step >= 32 && p.hstate != nullptr && *reinterpret_cast<volatile unsigned long long*>(p.hstate + 1) != 0) break;
        }
        if (!placed) {
          if (p.hstate == nullptr) atomicOr(p.err, MSC_DEVERR_TABLE_FULL);
          continue;
        }
        for (int a = 0; a < p.naggs; ++a) {
          const long long v = static_cast<long long>(lcells[static_cast<size_t>(s) * p.naggs + a]);
          if (v != p.agg_init[a]) atomic_fold(p.agg_kind[a], p.htbl + (pos << p.hshift) + 1 + a, v);
        }
      }
    }
  }
  if constexpr (MODE == MODE_DENSE) {
    __syncthreads();
    const int cells = p.ngroups * p.naggs;  // the trash group is not exported
    for (int cell = warp; cell < cells; cell += NW) {
      const int kind = p.agg_kind[cell % p.naggs];
      const long long* base = acc + cell * NT;
      long long v = base[lane];
#pragma unroll
      for (int j = 1; j < NW; ++j) v = agg_combine(kind, v, base[lane + 32 * j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v = agg_combine(kind, v, __shfl_xor_sync(0xffffffffu, v, o));
      if (lane == 0 && v != p.agg_init[cell % p.naggs]) atomic_fold(kind, p.dense_out + cell, v);
    }
  }
}

// host launcher, instantiated once per (R, MODE) in its own translation unit (scan_inst_*.cu)
template <int R, int MODE>
int launch_scan(msc_ctx* ctx, LaunchPlan* lp) {
  auto kern = scan_kernel<R, MODE>;
  // attribute + occupancy query cost ~10 us per launch, which shows on sub-millisecond scans: remember the last answer
  static size_t cached_smem = ~static_cast<size_t>(0);
  static int cached_occ = 0, cached_dev = -1;
  if (cached_smem != lp->smem || cached_dev != ctx->device) {
    MSC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(lp->smem)));
    MSC_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached_occ, kern, NT, lp->smem));
    cached_smem = lp->smem;
    cached_dev = ctx->device;
  }
  const int occ = cached_occ;
  if (occ < 1) return ctx->fail(MSC_ERR_ARG, "scan kernel does not fit on an SM");
  uint64_t grid = static_cast<uint64_t>(ctx->sm_count) * occ;
  const uint64_t need = (lp->p.ntiles + NW - 1) / NW;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  lp->grid = static_cast<int>(grid);
  if (lp->timed) MSC_CUDA(ctx, cudaEventRecord(ctx->ev_s0, ctx->stream));
  kern<<<lp->grid, NT, lp->smem, ctx->stream>>>(lp->p);
  ctx->stats.launches += 1;
  if (lp->timed) {
    MSC_CUDA(ctx, cudaEventRecord(ctx->ev_s1, ctx->stream));
    ctx->stats.last_scan_grid = lp->grid;
    ctx->stats.last_scan_stages = static_cast<int32_t>(lp->p.nstages);
    ctx->stats.last_scan_smem = static_cast<int32_t>(lp->smem);
    ctx->stats.last_scan_rows_per_thread = R;
    ctx->stats.last_scan_kind = MSC_SCAN_KIND_VM;
    ctx->stats.last_scan_regs = 0;
  }
  MSC_CUDA(ctx, cudaGetLastError());
  return MSC_OK;
}
#endif  // __CUDACC__ && !MSCAN_DECL_ONLY

}  // namespace mscan
