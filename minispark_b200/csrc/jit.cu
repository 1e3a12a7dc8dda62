// jit.cu -- query-specialised dense aggregate scan: source generation, NVRTC compilation, kernel cache, launch.
//
// The reference's native engine renders one program per query from templates/plan.zig (src/mini_spark/codegen.py),
// compiles it and caches the executable (execution.py:139-160).  This is the same idea on the GPU: the query's
// three-address expression program (include/minispark_cuda.h) is translated into straight-line CUDA C++ for a
// lane's 8 rows, appended to the hand-written kernel frame of jit_prelude.inc (persistent grid, warp-private
// cp.async.bulk rings), compiled for sm_100a with NVRTC and kept in a per-process cache keyed by the source text
// (plus an optional on-disk cubin cache).  Group accumulators live in REGISTERS for the whole kernel -- one
// predicated add per (group, aggregate) and row -- so there is no dispatch, no mask arithmetic and no shared-memory
// accumulator traffic; the interpreters (scan_regvm_impl.cuh, scan_kernel.cuh) remain for one-shot queries, whose
// scan is shorter than a compile, and for shapes outside this generator (many groups).
//
// libnvrtc and libcuda are dlopen()ed on first use: the library still loads (and every other entry point works)
// where they are absent, and msc_jit_* then fail with a message.
#include <dlfcn.h>
#include <sys/stat.h>

#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>

#include "jit.h"

namespace mscan {
namespace {

const char* const kPrelude =
#include "jit_prelude.inc"
    ;

struct JitParams {  // must match struct JitParams in jit_prelude.inc
  unsigned long long nrows;
  const unsigned long long* nrows_dev;
  uint32_t ntiles;
  uint32_t _pad;
  const unsigned char* col[24];
  const void* gather[16];
  const void* luts[8];
  long long consts[32];
  unsigned long long* dense_out;
  int* err;
  uint32_t* tile_counts;
  const unsigned long long* tile_offsets;
  void* out[24];
  long long fconsts[32];
  unsigned long long* fmeta;
  uint32_t* ticket;
  unsigned long long* mailbox[8];
  const int* inv;
  unsigned long long epoch;
  int rank, world, nlocal, slot_cells, nglobal, _pad3;
};
static_assert(MSC_VM_MAX_STAGED == 24 && MSC_VM_MAX_GATHER == 16 && MSC_VM_MAX_LUTS == 8 && MSC_VM_MAX_CONSTS == 32 && MSC_VM_MAX_OUT == 24,
              "JitParams layout");

// ---- NVRTC + driver API through dlopen ------------------------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
typedef struct CUmod_st* CUmodule;
typedef struct CUfunc_st* CUfunction;
typedef struct CUstream_st* CUstream;

struct Api {
  void *nvrtc = nullptr, *cuda = nullptr;
  int (*nvrtcCreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*nvrtcCompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*nvrtcGetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*nvrtcGetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*nvrtcGetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*nvrtcGetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*nvrtcDestroyProgram)(nvrtcProgram*) = nullptr;
  int (*nvrtcVersion)(int*, int*) = nullptr;
  int (*cuModuleLoadData)(CUmodule*, const void*) = nullptr;
  int (*cuModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  int (*cuFuncSetAttribute)(CUfunction, int, int) = nullptr;
  int (*cuFuncGetAttribute)(int*, int, CUfunction) = nullptr;
  int (*cuOccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
  int (*cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  int (*cuGetErrorString)(int, const char**) = nullptr;
};

Api& api() {
  static Api a;
  return a;
}

template <class F>
bool sym(void* lib, const char* name, F* out, std::string* why) {
  *out = reinterpret_cast<F>(dlsym(lib, name));
  if (!*out) *why = std::string("missing symbol ") + name;
  return *out != nullptr;
}

bool load_nvrtc(std::string* why) {
  Api& a = api();
  if (a.nvrtc) return true;
  if (getenv("MSC_JIT_NO_NVRTC")) {  // tests: behave like a machine without the compiler library
    *why = "libnvrtc.so.12 not found (MSC_JIT_NO_NVRTC is set)";
    return false;
  }
  const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
  for (const char* n : names)
    if ((a.nvrtc = dlopen(n, RTLD_NOW | RTLD_LOCAL)) != nullptr) break;
  if (!a.nvrtc) {
    *why = "libnvrtc.so.12 not found (needed to specialise scan kernels)";
    return false;
  }
  return sym(a.nvrtc, "nvrtcCreateProgram", &a.nvrtcCreateProgram, why) && sym(a.nvrtc, "nvrtcCompileProgram", &a.nvrtcCompileProgram, why) &&
         sym(a.nvrtc, "nvrtcGetProgramLogSize", &a.nvrtcGetProgramLogSize, why) && sym(a.nvrtc, "nvrtcGetProgramLog", &a.nvrtcGetProgramLog, why) &&
         sym(a.nvrtc, "nvrtcGetCUBINSize", &a.nvrtcGetCUBINSize, why) && sym(a.nvrtc, "nvrtcGetCUBIN", &a.nvrtcGetCUBIN, why) &&
         sym(a.nvrtc, "nvrtcDestroyProgram", &a.nvrtcDestroyProgram, why) && sym(a.nvrtc, "nvrtcVersion", &a.nvrtcVersion, why);
}

bool load_driver(std::string* why) {
  Api& a = api();
  if (a.cuda) return true;
  a.cuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
  if (!a.cuda) {
    *why = "libcuda.so.1 not found";
    return false;
  }
  return sym(a.cuda, "cuModuleLoadData", &a.cuModuleLoadData, why) && sym(a.cuda, "cuModuleGetFunction", &a.cuModuleGetFunction, why) &&
         sym(a.cuda, "cuFuncSetAttribute", &a.cuFuncSetAttribute, why) && sym(a.cuda, "cuFuncGetAttribute", &a.cuFuncGetAttribute, why) &&
         sym(a.cuda, "cuOccupancyMaxActiveBlocksPerMultiprocessor", &a.cuOccupancyMaxActiveBlocksPerMultiprocessor, why) &&
         sym(a.cuda, "cuLaunchKernel", &a.cuLaunchKernel, why) && sym(a.cuda, "cuGetErrorString", &a.cuGetErrorString, why);
}

constexpr int CU_FUNC_ATTRIBUTE_NUM_REGS = 4;
constexpr int CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES = 8;

// ---- source generation ----------------------------------------------------------------------------------------
struct Layout {  // of one warp stage, as plan_launch lays it out
  uint32_t off[MSC_VM_MAX_STAGED];
  uint32_t tile_bytes[MSC_VM_MAX_STAGED];
  uint32_t stage_bytes = 0, tx_bytes = 0;
};

Layout stage_layout(const msc_scan_desc* sd) {
  Layout l;
  uint32_t off = 0;
  for (int c = 0; c < sd->nstaged; ++c) {
    const uint32_t w = static_cast<uint32_t>(msc_phys_width(sd->staged[c].phys));
    l.off[c] = off;
    l.tile_bytes[c] = w * 256;
    l.tx_bytes += w * 256;
    off += static_cast<uint32_t>(msc_round_up(w * 256, 128));
  }
  l.stage_bytes = off ? off : 128;
  return l;
}

const char* ld_fn(int phys) {
  switch (phys) {
    case MSC_P_U8: return "ld_u8";
    case MSC_P_U16: return "ld_u16";
    case MSC_P_U32: return "ld_u32";
    case MSC_P_I32: return "ld_i32";
    case MSC_P_F32: return "ld_f32";
    default: return "ld_64";
  }
}

bool is_float_kind(int k) { return k == MSC_AGG_SUM_F || k == MSC_AGG_MIN_F || k == MSC_AGG_MAX_F; }

// a scan that probes a join table keeps the table in L2 by streaming its columns through with an evict-first policy
bool program_probes(const msc_scan_desc* sd) {
  for (int pc = 0; pc + 1 < sd->ncode; pc += 2) {
    const int op = sd->code[pc] & 0x3f;
    if (op == MSC_OP_END) break;
    if (op == MSC_OP_PROBE) return true;
  }
  return false;
}

struct Gen {
  const msc_scan_desc* sd;
  int ngroups, naggs, stride;  // stride = naggs, or naggs + 1 with the hidden per-group row counter in slot naggs
  const int* kinds;            // [stride]
  const long long* init;       // [stride]
  int nstages;
  bool masked;  // SUM_F / COUNT as acc = fma(v, m, acc) with one-hot f64 masks m read from a shared-memory table
  // Accumulators in shared memory, one private copy per thread ([cell][thread]): a row updates the `stride` cells of ITS group
  // (load, add, store) instead of every group's register accumulator under a predicate.  Past ~32 cells the register form
  // issues groups x accumulators FP64 instructions per row and is bound by that pipe, not by HBM (sf10, 7 groups x 6
  // accumulators: 0.30 ms for 0.78 GB).
  bool smem_cells = false;
  // fused finish (optional): the final projection over the aggregate's groups, run by the last CTA of the scan
  const msc_scan_desc* fin = nullptr;
  const int32_t* fin_cols = nullptr;  // staged slot of `fin` -> column of the compacted relation (0 = group id, 1 + s = accumulator s)
  const int32_t* fin_phys = nullptr;  // physical types of its outputs
  int fin_nout = 0, count_slot = 0;
  bool fin_peer = false;  // merge the partial tables of all ranks over NVLink peer memory before the projection
  bool in_finish = false;
  std::ostringstream o;
  std::string why;
  bool counted[MSC_VM_MAX_AGGS + 1] = {};  // slot is "SUM_I of a constant": counted in a u32 per tile, folded at the tile's end
  int count_const[MSC_VM_MAX_AGGS + 1] = {};  // index of that constant in p.consts (-1: the hidden row counter, times 1)
  std::string count_mul(int s) { return count_const[s] < 0 ? std::string("1ll") : "p.consts[" + std::to_string(count_const[s]) + "]"; }

  std::string operand(uint32_t opnd, bool* ok) {
    const int kind = (opnd >> 12) & 7, idx = opnd & 0xfff;
    const bool i2f = ((opnd >> 12) & MSC_SRC_I2F) != 0;
    std::string s;
    if (in_finish && (kind == MSC_SRC_GATHER || kind == MSC_SRC_GATHER_T || kind == MSC_SRC_LUT)) {
      *ok = false;
      return "0ll";
    }
    switch (kind) {
      case MSC_SRC_TEMP: s = temp(idx); break;
      case MSC_SRC_STAGED:
        if (tail_mode) s = "ldrow<" + std::to_string(sd->staged[idx].phys) + ">(sb + " + std::to_string(tail_lay.off[idx]) + ", row)";
        else s = in_finish ? "f" + std::to_string(idx) : "c" + std::to_string(idx) + "[r]";
        break;
      case MSC_SRC_CONST: s = (in_finish ? "p.fconsts[" : "p.consts[") + std::to_string(idx) + "]"; break;
      case MSC_SRC_GATHER:
        s = "gather_at<" + std::to_string(sd->gather[idx & 63].phys) + ">(p.gather[" + std::to_string(idx & 63) + "], c" +
            std::to_string(idx >> 6) + "[r], valid)";
        break;
      case MSC_SRC_GATHER_T:  // a build-side column of a fused join, read through the row its PROBE matched
        s = "gather_at<" + std::to_string(sd->gather[idx & 63].phys) + ">(p.gather[" + std::to_string(idx & 63) + "], " + temp(idx >> 6) +
            ", valid && " + temp(idx >> 6) + " >= 0)";
        break;
      case MSC_SRC_NONE: s = "0ll"; break;
      default: *ok = false; return "0ll";
    }
    if (i2f) s = "d2l((double)(" + s + "))";
    return s;
  }

  // the value of one instruction as an i64 expression of a and b (raw 64-bit operands)
  std::string compute(int op, const std::string& a, const std::string& b, uint32_t opnd_b, bool* ok) {
    auto f2 = [&](const char* sym) { return "d2l(l2d(" + a + ") " + sym + " l2d(" + b + "))"; };
    auto i2 = [&](const char* sym) { return "((" + a + ") " + sym + " (" + b + "))"; };
    auto cf = [&](const char* sym) { return "((l2d(" + a + ") " + sym + " l2d(" + b + ")) ? 1ll : 0ll)"; };
    auto ci = [&](const char* sym) { return "(((" + a + ") " + sym + " (" + b + ")) ? 1ll : 0ll)"; };
    switch (op) {
      case MSC_OP_MOV: return a;
      case MSC_OP_ADD_F: return f2("+");
      case MSC_OP_SUB_F: return f2("-");
      case MSC_OP_MUL_F: return f2("*");
      case MSC_OP_DIV_F: return "(l2d(" + b + ") == 0.0 ? d2l(0.0) : d2l(l2d(" + a + ") / l2d(" + b + ")))";
      case MSC_OP_FLOORDIV_F: return "py_floordiv_f(" + a + ", " + b + ")";
      case MSC_OP_MOD_F: return "py_mod_f(" + a + ", " + b + ")";
      case MSC_OP_ADD_I: return i2("+");
      case MSC_OP_SUB_I: return i2("-");
      case MSC_OP_MUL_I: return i2("*");
      case MSC_OP_FLOORDIV_I: return "py_floordiv_i(" + a + ", " + b + ")";
      case MSC_OP_MOD_I: return "py_mod_i(" + a + ", " + b + ")";
      case MSC_OP_LT_F: return cf("<");
      case MSC_OP_LE_F: return cf("<=");
      case MSC_OP_GT_F: return cf(">");
      case MSC_OP_GE_F: return cf(">=");
      case MSC_OP_EQ_F: return cf("==");
      case MSC_OP_NE_F: return cf("!=");
      case MSC_OP_LT_I: return ci("<");
      case MSC_OP_LE_I: return ci("<=");
      case MSC_OP_GT_I: return ci(">");
      case MSC_OP_GE_I: return ci(">=");
      case MSC_OP_EQ_I: return ci("==");
      case MSC_OP_NE_I: return ci("!=");
      case MSC_OP_AND: return i2("&");
      case MSC_OP_OR: return i2("|");
      case MSC_OP_LUT8:
        return "(valid ? (i64)__ldg(reinterpret_cast<const unsigned char*>(p.luts[" + std::to_string(opnd_b & 0xfff) + "]) + (" + a + ")) : 0ll)";
      case MSC_OP_LUT32:
        return "(valid ? (i64)__ldg(reinterpret_cast<const u32*>(p.luts[" + std::to_string(opnd_b & 0xfff) + "]) + (" + a + ")) : 0ll)";
      case MSC_OP_PROBE:
        if (in_finish) {
          *ok = false;
          return "0ll";
        }
        return std::string("join_probe<") + ((opnd_b & MSC_PROBE_COMPACT) ? "true" : "false") + ">(p.luts[" + std::to_string(opnd_b & 0x7ff) + "], " + a +
               ", valid, keep_policy)";
      default: *ok = false; return "0ll";
    }
  }

  bool temp_arrays = false;  // project scans run the program in two row loops, so temporaries are arrays over the rows
  bool tail_mode = false;    // the compacted tail of a probing scan: one surviving row per lane, staged values read by row index
  Layout tail_lay;
  std::string temp(int idx) {
    if (tail_mode) return "w" + std::to_string(idx);
    return (in_finish ? "ft" : "t") + std::to_string(idx) + (temp_arrays ? "[r]" : "");
  }
  std::string acc(int g, int s) { return "a" + std::to_string(g) + "_" + std::to_string(s); }
  std::string cnt(int g, int s) { return "n" + std::to_string(g) + "_" + std::to_string(s); }

  // accumulators are typed: double for the float kinds, i64 otherwise
  void emit_agg(int slot, const std::string& x) {
    if (smem_cells) {
      // (cellrow: the row's group, or the trash group NG for rows without one -- straight-line code, no divergence)
      o << "        { i64* cell = cellrow + " << slot << " * NT; ";
      if (counted[slot]) o << "*cell += " << count_mul(slot) << ";";
      else if (kinds[slot] == MSC_AGG_SUM_F) o << "*cell = d2l(l2d(*cell) + l2d(" << x << "));";
      else if (kinds[slot] == MSC_AGG_SUM_I) o << "*cell += " << x << ";";
      else o << "*cell = agg_combine<" << kinds[slot] << ">(*cell, " << x << ");";
      o << " }\n";
      return;
    }
    o << "        { const int gsel = valid ? grp : -1;\n";
    for (int g = 0; g < ngroups; ++g) {
      o << "          ";
      const std::string a = acc(g, slot), gs = std::to_string(g);
      if (counted[slot] && masked) o << cnt(g, slot) << " += m" << gs << ";\n";
      else if (counted[slot]) o << "inc_if<" << gs << ">(" << cnt(g, slot) << ", gsel);\n";
      else if (kinds[slot] == MSC_AGG_SUM_F && masked) o << a << " = fma(l2d(" << x << "), m" << gs << ", " << a << ");\n";
      else if (kinds[slot] == MSC_AGG_SUM_F) o << "addf_if<" << gs << ">(" << a << ", l2d(" << x << "), gsel);\n";
      else if (kinds[slot] == MSC_AGG_SUM_I) o << "addi_if<" << gs << ">(" << a << ", " << x << ", gsel);\n";
      else if (is_float_kind(kinds[slot])) o << "if (gsel == " << gs << ") " << a << " = l2d(agg_combine<" << kinds[slot] << ">(d2l(" << a << "), " << x << "));\n";
      else o << "if (gsel == " << gs << ") " << a << " = agg_combine<" << kinds[slot] << ">(" << a << ", " << x << ");\n";
    }
    o << "        }\n";
  }

  // shared text: constants of the stage layout + the tile issue function
  void emit_layout(const Layout& lay) {
    o << "constexpr u32 STAGE_BYTES = " << lay.stage_bytes << ", TX_BYTES = " << lay.tx_bytes << ", NSTAGED = " << sd->nstaged << ";\n";
    o << "__device__ const u32 COL_OFF[" << std::max(1, sd->nstaged) << "] = {";
    for (int c = 0; c < sd->nstaged; ++c) o << (c ? ", " : "") << lay.off[c];
    if (!sd->nstaged) o << "0";
    o << "};\n__device__ const u32 COL_BYTES[" << std::max(1, sd->nstaged) << "] = {";
    for (int c = 0; c < sd->nstaged; ++c) o << (c ? ", " : "") << lay.tile_bytes[c];
    if (!sd->nstaged) o << "0";
    o << "};\n";
    o << R"(__device__ __forceinline__ void issue_tile(const JitParams& p, unsigned char* stages, u64* full, u32 stage, u64 tile, int lane) {
  if (lane == 0) mbar_expect_tx(&full[stage], TX_BYTES);
  __syncwarp();
  if (lane < (int)NSTAGED) bulk_g2s(stages + stage * STAGE_BYTES + COL_OFF[lane], p.col[lane] + tile * COL_BYTES[lane], COL_BYTES[lane], &full[stage]);
}
)";
  }

  // one instruction of the row program inside a row loop (temporaries through temp()); returns false on an unsupported one.
  // OUT destinations go to o<dst>[r]; FILTER narrows `valid`.
  bool emit_row_instruction(int pc, bool allow_out) {
    const uint32_t w0 = sd->code[pc], w1 = sd->code[pc + 1];
    const int op = w0 & 0x3f;
    const int dkind = (w0 >> 6) & 7, tee = (w0 >> 9) & 0xf, dst = (w0 >> 13) & 0x7f;
    const uint32_t oa = w1 & 0xffffu, ob = w1 >> 16;
    const bool lut = op == MSC_OP_LUT8 || op == MSC_OP_LUT32 || op == MSC_OP_PROBE;
    bool ok = true;
    const std::string a = operand(oa, &ok), b = lut ? std::string("0ll") : operand(ob, &ok);
    if (dkind == MSC_DST_OUT && !allow_out) return true;  // a count scan only needs the filters
    o << "      {  // instruction " << pc / 2 << "\n";
    if (op == MSC_OP_DIV_F || op == MSC_OP_FLOORDIV_F || op == MSC_OP_MOD_F) o << "        bad |= valid && (l2d(" << b << ") == 0.0);\n";
    if (op == MSC_OP_FLOORDIV_I || op == MSC_OP_MOD_I) o << "        bad |= valid && ((" << b << ") == 0);\n";
    o << "        const i64 x = " << compute(op, a, b, ob, &ok) << ";\n";
    if (tee) o << "        " << temp(tee - 1) << " = x;\n";
    switch (dkind) {
      case MSC_DST_TEMP: o << "        " << temp(dst) << " = x;\n"; break;
      case MSC_DST_FILTER: o << "        valid = valid && (x != 0);\n"; break;
      case MSC_DST_OUT: o << "        o" << dst << "[r] = x;\n"; break;
      case MSC_DST_NONE: break;
      default: why = "destination kind outside a project scan"; return false;
    }
    o << "      }\n";
    if (!ok) why = "operand or opcode outside the generator";
    return ok;
  }

  // COUNT (count_only) and PROJECT scans: FilterTask / ProjectTask (tasks.py:79-84, 167-177) with stable compaction.
  // The program is [filters..] RANK [outputs..] (lowering.py compile_project); without a filter there is no RANK and
  // a row's output position is its row number.
  bool generate_project(bool count_only, const int32_t* out_phys, int nout) {
    const Layout lay = stage_layout(sd);
    temp_arrays = true;
    int rank_pc = -1, end_pc = 0;
    bool has_filter = false;
    for (int pc = 0; pc + 1 < sd->ncode; pc += 2) {
      const int op = sd->code[pc] & 0x3f;
      if (op == MSC_OP_END) break;
      end_pc = pc + 2;
      if (op == MSC_OP_RANK) rank_pc = pc;
      const int dk = (sd->code[pc] >> 6) & 7;
      if (dk == MSC_DST_FILTER) {
        has_filter = true;
        if (rank_pc >= 0) {
          why = "filter after RANK";
          return false;
        }
      }
      if (dk == MSC_DST_GROUP || dk == MSC_DST_AGG) {
        why = "aggregate destination in a project scan";
        return false;
      }
    }
    if (has_filter && rank_pc < 0) {
      why = "filtered projection needs a RANK instruction";
      return false;
    }
    if (count_only && !has_filter) {
      why = "count scan without a filter";
      return false;
    }
    o << (program_probes(sd) ? "#define MSC_STREAM_EVICT_FIRST 1\n" : "") << "#define MINCTAS " << min_ctas(JIT_MIN_CTAS) << "\n" << kPrelude;
    o << "constexpr int NSTAGES = " << nstages << ";\n";
    emit_layout(lay);
    o << R"(
extern "C" __global__ void __launch_bounds__(NT, MINCTAS) msc_jit_scan(const __grid_constant__ JitParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u64* full = reinterpret_cast<u64*>(smem) + warp * 8;
  unsigned char* stages = smem + SMEM_HEADER + warp * (NSTAGES * STAGE_BYTES);
  if (lane == 0) {
    for (u32 st = 0; st < NSTAGES; ++st) mbar_init(&full[st], 1);
    mbar_fence_init();
  }
  __syncwarp();
  const u32 gw = blockIdx.x * NW + warp, nw = gridDim.x * NW;
  const u32 ntiles_w = (p.ntiles > gw) ? (p.ntiles - gw + nw - 1) / nw : 0;
  {
    const u32 pre = ntiles_w < NSTAGES ? ntiles_w : NSTAGES;
    for (u32 k = 0; k < pre; ++k) issue_tile(p, stages, full, k, gw + (u64)k * nw, lane);
  }
  const u64 nrows = p.nrows_dev ? *p.nrows_dev : p.nrows;
  const u64 keep_policy = l2_keep_policy();  // (join-table reads ask L2 to keep their lines)
  bool bad = false;
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
)";
    for (int c = 0; c < sd->nstaged; ++c)
      o << "    i64 c" << c << "[R]; " << ld_fn(sd->staged[c].phys) << "(sb + " << lay.off[c] << ", lane, c" << c << ");\n";
    for (int t = 0; t < sd->ntemps; ++t) o << "    i64 t" << t << "[R];\n";
    if (!count_only)
      for (int c = 0; c < nout; ++c) o << "    i64 o" << c << "[R];\n";
    // phase 1: the filters
    const int phase1_end = rank_pc >= 0 ? rank_pc : 0;
    o << "    u32 keep = 0;\n#pragma unroll\n    for (int r = 0; r < R; ++r) {\n      bool valid = (vmask >> r) & 1u;\n";
    for (int t = 0; t < sd->ntemps; ++t) o << "      t" << t << "[r] = 0;\n";
    for (int pc = 0; pc < phase1_end; pc += 2)
      if (!emit_row_instruction(pc, false)) return false;
    o << "      keep |= (valid ? 1u : 0u) << r;\n    }\n";
    if (count_only) {
      o << R"(    {
      u32 n = __popc(keep);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
      if (lane == 0) p.tile_counts[tile] = n;
    }
)";
    } else {
      // phase 2: the outputs of the surviving rows (rows that were filtered out compute harmless values; loads through
      // gathers / lookup tables look at `valid`)
      o << "#pragma unroll\n    for (int r = 0; r < R; ++r) {\n      const bool valid = (keep >> r) & 1u;\n";
      for (int pc = rank_pc >= 0 ? rank_pc + 2 : 0; pc < end_pc; pc += 2)
        if (!emit_row_instruction(pc, true)) return false;
      o << "    }\n";
      // stable output positions: rows 4*lane..4*lane+3 of each 128-row half, halves in order
      if (has_filter) {
        o << R"(    u64 pos0, pos1;
    {
      const u32 c = __popc(keep & 0xfu) | (__popc(keep >> 4) << 16);  // survivors of this lane in half 0 | half 1
      u32 inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
      }
      const u32 total = __shfl_sync(0xffffffffu, inc, 31), excl = inc - c;
      const u64 base = p.tile_offsets[tile];
      pos0 = base + (excl & 0xffffu);
      pos1 = base + (total & 0xffffu) + (excl >> 16);
    }
)";
      } else {
        o << "    const u64 pos0 = tile_row0 + 4 * lane, pos1 = pos0 + 128;\n";
      }
      for (int c = 0; c < nout; ++c) {
        const bool u32out = out_phys[c] == MSC_P_U32;
        const std::string ty = u32out ? "u32" : "i64";
        o << "    {\n      " << ty << "* out = reinterpret_cast<" << ty << "*>(p.out[" << c << "]);\n";
        if (!has_filter) {  // whole tiles go out as 128-bit stores
          o << "      if (vmask == 0xffu) {\n";
          if (u32out) {
            o << "        *reinterpret_cast<uint4*>(out + pos0) = make_uint4((u32)o" << c << "[0], (u32)o" << c << "[1], (u32)o" << c << "[2], (u32)o" << c << "[3]);\n";
            o << "        *reinterpret_cast<uint4*>(out + pos1) = make_uint4((u32)o" << c << "[4], (u32)o" << c << "[5], (u32)o" << c << "[6], (u32)o" << c << "[7]);\n";
          } else {
            for (int h = 0; h < 2; ++h)
              for (int j = 0; j < 2; ++j)
                o << "        *reinterpret_cast<longlong2*>(out + pos" << h << " + " << 2 * j << ") = make_longlong2(o" << c << "[" << 4 * h + 2 * j << "], o" << c
                  << "[" << 4 * h + 2 * j + 1 << "]);\n";
          }
          o << "      } else {\n";
        } else {
          o << "      {\n";
        }
        o << "        u64 q0 = pos0, q1 = pos1;\n#pragma unroll\n        for (int r = 0; r < 4; ++r) {\n";
        o << "          if ((keep >> r) & 1u) out[" << (has_filter ? "q0++" : "q0 + r") << "] = (" << ty << ")o" << c << "[r];\n";
        o << "          if ((keep >> (r + 4)) & 1u) out[" << (has_filter ? "q1++" : "q1 + r") << "] = (" << ty << ")o" << c << "[r + 4];\n";
        o << "        }\n      }\n    }\n";
      }
    }
    o << R"(    __syncwarp();
    if (k + NSTAGES < ntiles_w) issue_tile(p, stages, full, stage, gw + (u64)(k + NSTAGES) * nw, lane);
    if (++stage == NSTAGES) {
      stage = 0;
      parity ^= 1u;
    }
  }
  if (bad) atomicOr(p.err, 1);  // MSC_DEVERR_DIV_ZERO
}
)";
    return true;
  }

  // resident CTAs per SM the register allocator is asked to fit (experiments: MSC_JIT_MINCTAS overrides)
  static int min_ctas(int chosen) {
    const char* e = getenv("MSC_JIT_MINCTAS");
    const int v = e ? atoi(e) : 0;
    return (v >= 1 && v <= 16) ? v : chosen;
  }

  // GROUP BY a sorted integer key column (scan.cu checked: no descents, no filter): every run of equal keys is a group and
  // its run number is its output row (scan_kernel.cuh MODE_RUNS is the interpreted twin).
  bool generate_runs(int key_col) {
    const Layout lay = stage_layout(sd);
    temp_arrays = false;
    // (six resident CTAs, <= 80 registers: since the kernel stores before it adds it is bound by its own instruction stream
    // and fixed-latency waits, which more warps hide -- 0.354 ms with four CTAs per SM, 0.331 with five, 0.320 with six)
    o << (program_probes(sd) ? "#define MSC_STREAM_EVICT_FIRST 1\n" : "") << "#define MINCTAS " << min_ctas(6) << "\n" << kPrelude;
    o << "constexpr int NSTAGES = " << nstages << ";\n";
    emit_layout(lay);
    o << R"(
extern "C" __global__ void __launch_bounds__(NT, MINCTAS) msc_jit_runs(const __grid_constant__ JitParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u64* full = reinterpret_cast<u64*>(smem) + warp * 8;
  unsigned char* stages = smem + SMEM_HEADER + warp * (NSTAGES * STAGE_BYTES);
  if (lane == 0) {
    for (u32 st = 0; st < NSTAGES; ++st) mbar_init(&full[st], 1);
    mbar_fence_init();
  }
  __syncwarp();
  const u32 gw = blockIdx.x * NW + warp, nw = gridDim.x * NW;
  const u32 ntiles_w = (p.ntiles > gw) ? (p.ntiles - gw + nw - 1) / nw : 0;
  {
    const u32 pre = ntiles_w < NSTAGES ? ntiles_w : NSTAGES;
    for (u32 k = 0; k < pre; ++k) issue_tile(p, stages, full, k, gw + (u64)k * nw, lane);
  }
  const u64 nrows = p.nrows;
  bool bad = false;
  i64* out_key = reinterpret_cast<i64*>(p.out[0]);
  u32 stage = 0, parity = 0;
)";
    // Two values of a tile live in global memory, not in its staged columns: the number of runs that start before it and
    // the key of the row before its first.  Both are fetched one tile AHEAD (prof_runs: the two loads, issued where they
    // were needed, held 23 % of the kernel's stall samples).
    const int kphys = sd->staged[key_col].phys;
    const char* kty = kphys == MSC_P_U8 ? "unsigned char" : kphys == MSC_P_U16 ? "unsigned short" : kphys == MSC_P_U32 ? "u32"
                      : kphys == MSC_P_I32 ? "int" : "i64";
    o << "  const " << kty << "* key_col = reinterpret_cast<const " << kty << "*>(p.col[" << key_col << "]);\n";
    // (next_before keeps the column's own type: widening it where it is LOADED made the warp wait for the load right there --
    // prof_runs, 22 % of the stall samples on that one shift)
    o << "  u64 next_base = 0;\n  " << kty << " next_before = 0;\n";
    o << R"(  if (ntiles_w > 0) {
    next_base = p.tile_offsets[gw];
    if (lane == 0 && gw > 0) next_before = key_col[(u64)gw * WT - 1];
  }
  for (u32 k = 0; k < ntiles_w; ++k) {
    const u64 tile = gw + (u64)k * nw;
    const u64 base = next_base;  // runs that start before this tile
    const i64 before = (i64)next_before;
    if (k + 1 < ntiles_w) {
      next_base = p.tile_offsets[tile + nw];
      if (lane == 0) next_before = key_col[(tile + nw) * WT - 1];
    }
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
)";
    for (int c = 0; c < sd->nstaged; ++c)
      o << "    i64 c" << c << "[R]; " << ld_fn(sd->staged[c].phys) << "(sb + " << lay.off[c] << ", lane, c" << c << ");\n";
    o << "    i64 key[R];\n";
    bool used[MSC_VM_MAX_AGGS] = {};
    for (int pc = 0; pc + 1 < sd->ncode; pc += 2) {
      const uint32_t w0 = sd->code[pc];
      if ((w0 & 0x3f) == MSC_OP_END) break;
      if (((w0 >> 6) & 7) == MSC_DST_AGG) used[(w0 >> 13) & 0x7f] = true;
    }
    for (int a = 0; a < naggs; ++a)
      if (used[a]) o << "    i64 v" << a << "[R];\n";
    o << "#pragma unroll\n    for (int r = 0; r < R; ++r) {\n      const bool valid = (vmask >> r) & 1u;\n";
    for (int t = 0; t < sd->ntemps; ++t) o << "      i64 t" << t << " = 0;\n";
    bool ok = true, grouped = false;
    for (int pc = 0; pc + 1 < sd->ncode; pc += 2) {
      const uint32_t w0 = sd->code[pc], w1 = sd->code[pc + 1];
      const int op = w0 & 0x3f;
      if (op == MSC_OP_END) break;
      const int dkind = (w0 >> 6) & 7, tee = (w0 >> 9) & 0xf, dst = (w0 >> 13) & 0x7f;
      const uint32_t oa = w1 & 0xffffu, ob = w1 >> 16;
      const bool lut = op == MSC_OP_LUT8 || op == MSC_OP_LUT32 || op == MSC_OP_PROBE;
      if (op == MSC_OP_RANK || dkind == MSC_DST_FILTER || dkind == MSC_DST_OUT) {
        why = "filter / projection instruction in a streaming aggregate";
        return false;
      }
      const std::string a = operand(oa, &ok), b = lut ? std::string("0ll") : operand(ob, &ok);
      o << "      {\n";
      if (op == MSC_OP_DIV_F || op == MSC_OP_FLOORDIV_F || op == MSC_OP_MOD_F) o << "        bad |= valid && (l2d(" << b << ") == 0.0);\n";
      if (op == MSC_OP_FLOORDIV_I || op == MSC_OP_MOD_I) o << "        bad |= valid && ((" << b << ") == 0);\n";
      o << "        const i64 x = " << compute(op, a, b, ob, &ok) << ";\n";
      if (tee) o << "        " << temp(tee - 1) << " = x;\n";
      switch (dkind) {
        case MSC_DST_TEMP: o << "        " << temp(dst) << " = x;\n"; break;
        case MSC_DST_GROUP: o << "        key[r] = x;\n"; grouped = true; break;
        case MSC_DST_AGG: o << "        v" << dst << "[r] = x;\n"; break;
        case MSC_DST_NONE: break;
        default: why = "destination kind outside an aggregate scan"; return false;
      }
      o << "      }\n";
    }
    o << "    }\n";
    if (!ok || !grouped) {
      why = ok ? "program has no GROUP" : "operand or opcode outside the generator";
      return false;
    }
    o << "    // a row starts a run when its key differs from the previous ROW's: rows 4*lane..4*lane+3 of each 128-row half\n";
    o << "    i64 prev0 = __shfl_up_sync(0xffffffffu, key[3], 1), prev1 = __shfl_up_sync(0xffffffffu, key[7], 1);\n";
    o << "    const i64 row127 = __shfl_sync(0xffffffffu, key[3], 31);\n";
    o << "    if (lane == 0) {\n      prev1 = row127;\n      prev0 = tile_row0 > 0 ? before : ~key[0];\n    }\n";
    o << R"(    bool h[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const i64 left = (r == 0) ? prev0 : (r == 4) ? prev1 : key[r > 0 ? r - 1 : 0];
      h[r] = ((vmask >> r) & 1u) && key[r] != left;
    }
    int idx[R];
    {
      const u32 c = (u32)(h[0] + h[1] + h[2] + h[3]) | ((u32)(h[4] + h[5] + h[6] + h[7]) << 16);  // runs starting in this lane: half 0 | half 1
      u32 inc = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
      }
      const u32 total = __shfl_sync(0xffffffffu, inc, 31), excl = inc - c;
      u64 run0 = base + (excl & 0xffffu), run1 = base + (total & 0xffffu) + (excl >> 16);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (h[r]) out_key[run0] = key[r];
        run0 += h[r];
        idx[r] = ((vmask >> r) & 1u) ? (int)run0 - 1 : -1;
        if (h[r + 4]) out_key[run1] = key[r + 4];
        run1 += h[r + 4];
        idx[r + 4] = ((vmask >> (r + 4)) & 1u) ? (int)run1 - 1 : -1;
      }
    }
)";
    // phase 1: the runs that start in a lane's segments are stored; phase 2, behind the warp's barrier: the leading rows of a
    // segment are added to the run they continue (fold_store_heads in the prelude)
    for (int a = 0; a < naggs; ++a)
      if (used[a])
        o << "    const i64 lead0_" << a << " = fold_store_heads<" << kinds[a] << ">(reinterpret_cast<u64*>(p.out[" << 1 + a << "]), h, idx, v" << a
          << ", vmask & 0xfu), lead1_" << a << " = fold_store_heads<" << kinds[a] << ">(reinterpret_cast<u64*>(p.out[" << 1 + a << "]), h + 4, idx + 4, v" << a
          << " + 4, (vmask >> 4) & 0xfu);\n";
    o << "    __syncwarp();\n    {\n      const long long open_run = (long long)base - 1;  // the run still open when this tile begins: its part goes to the tile's carry cell\n"
      << "      const bool l0 = (vmask & 1u) && !h[0], l1 = ((vmask >> 4) & 1u) && !h[4];\n";
    for (int a = 0; a < naggs; ++a)
      if (used[a]) {
        const std::string col = "reinterpret_cast<u64*>(p.out[" + std::to_string(1 + a) + "])", carry = "(p.dense_out + " + std::to_string(a) + "ull * ((p.ntiles + 1u) & ~1u) + tile)";  // (rows of even length: 16-byte aligned)
        o << "      if (l0) atomic_fold(" << kinds[a] << ", idx[0] == open_run ? " << carry << " : " << col << " + idx[0], lead0_" << a << ");\n"
          << "      if (l1) atomic_fold(" << kinds[a] << ", idx[4] == open_run ? " << carry << " : " << col << " + idx[4], lead1_" << a << ");\n";
      }
    o << "    }\n";
    o << R"(    __syncwarp();
    if (k + NSTAGES < ntiles_w) issue_tile(p, stages, full, stage, gw + (u64)(k + NSTAGES) * nw, lane);
    if (++stage == NSTAGES) {
      stage = 0;
      parity ^= 1u;
    }
  }
  if (bad) atomicOr(p.err, 1);  // MSC_DEVERR_DIV_ZERO
}
)";
    return true;
  }

  bool generate() {
    const Layout lay = stage_layout(sd);
    // which accumulators are SUM_I of a constant (COUNT, plan.py:190-204 / sql.py:463-464)?
    int ninstr = 0;
    for (int pc = 0; pc + 1 < sd->ncode; pc += 2, ++ninstr) {
      const uint32_t w0 = sd->code[pc], w1 = sd->code[pc + 1];
      const int op = w0 & 0x3f;
      if (op == MSC_OP_END) break;
      const int dkind = (w0 >> 6) & 7, dst = (w0 >> 13) & 0x7f, tee = (w0 >> 9) & 0xf;
      const uint32_t a = w1 & 0xffffu;
      if (dkind == MSC_DST_AGG && op == MSC_OP_MOV && tee == 0 && ((a >> 12) & 15) == MSC_SRC_CONST && kinds[dst] == MSC_AGG_SUM_I) {
        counted[dst] = true;
        count_const[dst] = static_cast<int>(a & 0xfff);
      }
    }
    for (int pc = 0, seen[MSC_VM_MAX_AGGS + 1] = {}; pc + 1 < sd->ncode; pc += 2) {  // a slot written twice is not a plain count
      const uint32_t w0 = sd->code[pc];
      if ((w0 & 0x3f) == MSC_OP_END) break;
      if (((w0 >> 6) & 7) == MSC_DST_AGG && ++seen[(w0 >> 13) & 0x7f] > 1) counted[(w0 >> 13) & 0x7f] = false;
    }
    if (stride > naggs) {  // hidden per-group row counter
      counted[naggs] = true;
      count_const[naggs] = -1;
    }

    // (a probing scan waits on table reads, not on arithmetic: with few cells a fifth resident CTA hides more of that latency --
    // config 5: 0.573 -> 0.513 ms)
    // registers: 2 per accumulator cell; up to 32 cells fit 4 CTAs of 128 threads per SM (128 registers), more need 3 (168)
    o << (program_probes(sd) ? "#define MSC_STREAM_EVICT_FIRST 1\n" : "") << "#define MINCTAS " << min_ctas(program_probes(sd) && ngroups * stride <= 16 ? 5 : (smem_cells || ngroups * stride <= JIT_REG_CELLS_4CTAS) ? 4 : 3) << "\n" << kPrelude;
    if (masked) {  // the masks are fixed at GROUP: a later filter would not reach them
      bool grouped_seen = false;
      for (int pc = 0; pc + 1 < sd->ncode; pc += 2) {
        const uint32_t w0 = sd->code[pc];
        if ((w0 & 0x3f) == MSC_OP_END) break;
        const int dk = (w0 >> 6) & 7;
        if (dk == MSC_DST_GROUP) grouped_seen = true;
        if (dk == MSC_DST_FILTER && grouped_seen) {
          why = "filter after GROUP (masked variant)";
          return false;
        }
      }
    }
    o << "constexpr int NG = " << ngroups << ", NGP = " << (ngroups + 1) / 2 * 2 << ", STRIDE = " << stride << ", NSTAGES = " << nstages << ";\n";
    o << "constexpr u32 STAGE_BYTES = " << lay.stage_bytes << ", TX_BYTES = " << lay.tx_bytes << ", NSTAGED = " << sd->nstaged << ";\n";
    o << "__device__ const u32 COL_OFF[" << std::max(1, sd->nstaged) << "] = {";
    for (int c = 0; c < sd->nstaged; ++c) o << (c ? ", " : "") << lay.off[c];
    if (!sd->nstaged) o << "0";
    o << "};\n__device__ const u32 COL_BYTES[" << std::max(1, sd->nstaged) << "] = {";
    for (int c = 0; c < sd->nstaged; ++c) o << (c ? ", " : "") << lay.tile_bytes[c];
    if (!sd->nstaged) o << "0";
    o << "};\n__device__ const int KIND[STRIDE] = {";
    for (int s = 0; s < stride; ++s) o << (s ? ", " : "") << kinds[s];
    o << "};\n__device__ const i64 INIT[STRIDE] = {";
    for (int s = 0; s < stride; ++s) o << (s ? ", " : "") << init[s] << "ll";
    o << "};\n\n";
    o << R"(__device__ __forceinline__ void issue_tile(const JitParams& p, unsigned char* stages, u64* full, u32 stage, u64 tile, int lane) {
  if (lane == 0) mbar_expect_tx(&full[stage], TX_BYTES);
  __syncwarp();
  if (lane < (int)NSTAGED) bulk_g2s(stages + stage * STAGE_BYTES + COL_OFF[lane], p.col[lane] + tile * COL_BYTES[lane], COL_BYTES[lane], &full[stage]);
}

extern "C" __global__ void __launch_bounds__(NT, MINCTAS) msc_jit_dense(const __grid_constant__ JitParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  u64* full = reinterpret_cast<u64*>(smem) + warp * 8;
  unsigned char* stages = smem + SMEM_HEADER + warp * (NSTAGES * STAGE_BYTES);
)" << (smem_cells ? R"(  i64* cellbase = reinterpret_cast<i64*>(smem + SMEM_HEADER + NW * (NSTAGES * STAGE_BYTES));  // [(NG + 1) * STRIDE][NT], group NG = trash
  i64* cells = cellbase + tid;  // this thread's copy of cell c: cells[c * NT] (a warp's 32 copies are 32 consecutive words: no bank conflicts)
  u32* queue = reinterpret_cast<u32*>(cellbase + (NG + 1) * STRIDE * NT) + warp * (2 * WT);  // (probing scans only: survivors of a tile)
  if (lane == 0) {
    for (u32 st = 0; st < NSTAGES; ++st) mbar_init(&full[st], 1);
    mbar_fence_init();
  }
  for (int c = 0; c < (NG + 1) * STRIDE; ++c) cells[c * NT] = INIT[c % STRIDE];
  __syncthreads();
)" : R"(  i64* red = reinterpret_cast<i64*>(smem + SMEM_HEADER + NW * (NSTAGES * STAGE_BYTES));  // [NW][NG * STRIDE]
  // one-hot f64 masks: row e = 0 is "no group" (filtered out / past the end / code out of range), row g + 1 selects group g
  double* mlut = reinterpret_cast<double*>(red + NW * NG * STRIDE);  // [NG + 1][NGP]
  u32* queue = reinterpret_cast<u32*>(&mlut[(NG + 1) * NGP]) + warp * (2 * WT);  // (probing scans only: survivors of a tile)
  if (lane == 0) {
    for (u32 st = 0; st < NSTAGES; ++st) mbar_init(&full[st], 1);
    mbar_fence_init();
  }
  for (int i = tid; i < (NG + 1) * NGP; i += NT) mlut[i] = (i / NGP >= 1 && i % NGP == i / NGP - 1) ? 1.0 : 0.0;
  __syncthreads();
)") << R"(
  const u32 gw = blockIdx.x * NW + warp, nw = gridDim.x * NW;
  const u32 ntiles_w = (p.ntiles > gw) ? (p.ntiles - gw + nw - 1) / nw : 0;
  {
    const u32 pre = ntiles_w < NSTAGES ? ntiles_w : NSTAGES;
    for (u32 k = 0; k < pre; ++k) issue_tile(p, stages, full, k, gw + (u64)k * nw, lane);
  }
  const u64 nrows = p.nrows_dev ? *p.nrows_dev : p.nrows;
  const u64 keep_policy = l2_keep_policy();  // (join-table reads ask L2 to keep their lines)
  bool bad = false;
)";
    for (int g = 0; g < ngroups && !smem_cells; ++g)
      for (int s = 0; s < stride; ++s) {
        if (is_float_kind(kinds[s])) o << "  double " << acc(g, s) << " = l2d(INIT[" << s << "]);";
        else o << "  i64 " << acc(g, s) << " = INIT[" << s << "];";
        if (counted[s]) o << (masked ? " double " : " u32 ") << cnt(g, s) << " = 0;";
        o << "\n";
      }
    o << R"(  u32 stage = 0, parity = 0;
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
)";
    // Rows past the end of the relation exist only in the last tile.  When the group key is a staged column read as it
    // is, that tile overwrites the key of its invalid rows with -1 ("no group") and the row loop needs no validity
    // bits at all; programs that dereference row values (gathers, lookup tables) keep the per-row bit.
    int key_col = -1;
    bool derefs = false;
    for (int pc = 0; pc + 1 < sd->ncode; pc += 2) {
      const uint32_t w0 = sd->code[pc], w1 = sd->code[pc + 1];
      const int op = w0 & 0x3f;
      if (op == MSC_OP_END) break;
      const uint32_t oa = w1 & 0xffffu, ob = w1 >> 16;
      if (op == MSC_OP_LUT8 || op == MSC_OP_LUT32 || op == MSC_OP_PROBE || ((oa >> 12) & 7) == MSC_SRC_GATHER || ((ob >> 12) & 7) == MSC_SRC_GATHER ||
          ((oa >> 12) & 7) == MSC_SRC_GATHER_T || ((ob >> 12) & 7) == MSC_SRC_GATHER_T)
        derefs = true;
      if (((w0 >> 6) & 7) == MSC_DST_GROUP && op == MSC_OP_MOV && ((oa >> 12) & 15) == MSC_SRC_STAGED) key_col = oa & 0xfff;
    }
    const bool valid_bits = derefs || key_col < 0;
    for (int c = 0; c < sd->nstaged; ++c)
      o << "    i64 c" << c << "[R]; " << ld_fn(sd->staged[c].phys) << "(sb + " << lay.off[c] << ", lane, c" << c << ");\n";
    if (!valid_bits) {
      o << "    if (vmask != 0xffu) {\n#pragma unroll\n      for (int r = 0; r < R; ++r)\n        if (!((vmask >> r) & 1u)) c" << key_col
        << "[r] = -1;\n    }\n";
    }
    // A join probe (MSC_OP_PROBE) is a dependent random read: run the program in TWO row loops around the first one --
    // loop 1 evaluates what precedes it and ISSUES the table read of every row of the lane, loop 2 resolves the probes and
    // runs the rest -- so that a lane's 8 lookups are in flight together instead of one after the other.
    int probe_pc = -1;
    for (int pc = 0; pc + 1 < sd->ncode; pc += 2) {
      const int op = sd->code[pc] & 0x3f, dkind = (sd->code[pc] >> 6) & 7;
      if (op == MSC_OP_END) break;
      if (op == MSC_OP_PROBE) {
        if (dkind == MSC_DST_TEMP && ((sd->code[pc] >> 9) & 0xf) == 0) probe_pc = pc;
        break;
      }
      if (dkind != MSC_DST_TEMP && dkind != MSC_DST_FILTER) break;  // only filters and temporaries may precede a split
    }
    temp_arrays = probe_pc >= 0;
    bool ok = true, grouped = false;
    if (probe_pc >= 0) {
      for (int t = 0; t < sd->ntemps; ++t) o << "    i64 t" << t << "[R];\n";
      o << "    JoinProbe<" << (((sd->code[probe_pc + 1] >> 16) & MSC_PROBE_COMPACT) ? "true" : "false") << "> pq[R]; u32 vm = 0;\n#pragma unroll\n    for (int r = 0; r < R; ++r) {\n      bool valid = (vmask >> r) & 1u;\n";
      for (int t = 0; t < sd->ntemps; ++t) o << "      t" << t << "[r] = 0;\n";
      for (int pc = 0; pc < probe_pc; pc += 2) {
        const uint32_t w0 = sd->code[pc], w1 = sd->code[pc + 1];
        const int op = w0 & 0x3f, dkind = (w0 >> 6) & 7, tee = (w0 >> 9) & 0xf, dst = (w0 >> 13) & 0x7f;
        const uint32_t oa = w1 & 0xffffu, ob = w1 >> 16;
        const bool lut = op == MSC_OP_LUT8 || op == MSC_OP_LUT32;
        const std::string a = operand(oa, &ok), b = lut ? std::string("0ll") : operand(ob, &ok);
        o << "      {  // instruction " << pc / 2 << "\n";
        if (op == MSC_OP_DIV_F || op == MSC_OP_FLOORDIV_F || op == MSC_OP_MOD_F) o << "        bad |= valid && (l2d(" << b << ") == 0.0);\n";
        if (op == MSC_OP_FLOORDIV_I || op == MSC_OP_MOD_I) o << "        bad |= valid && ((" << b << ") == 0);\n";
        o << "        const i64 x = " << compute(op, a, b, ob, &ok) << ";\n";
        if (tee) o << "        " << temp(tee - 1) << " = x;\n";
        if (dkind == MSC_DST_TEMP) o << "        " << temp(dst) << " = x;\n";
        else o << "        valid = valid && (x != 0);\n";
        o << "      }\n";
      }
      {
        const uint32_t w1 = sd->code[probe_pc + 1];
        const std::string a = operand(w1 & 0xffffu, &ok);
        o << "      join_probe_issue(p.luts[" << ((w1 >> 16) & 0x7ff) << "], " << a << ", valid, pq[r], keep_policy);  // instruction "
          << probe_pc / 2 << ", first half\n";
      }
      o << "      vm |= (valid ? 1u : 0u) << r;\n    }\n";
    }
    // Compacted tail: when the instruction after the probe is its match filter (FILTER <- probe >= 0) and nothing later reads a
    // temporary computed before the probe, the rows that found a partner are queued per warp (tile row, build row) and the rest
    // of the program -- gathers, GROUP, aggregates -- runs once per SURVIVOR, a lane each, instead of once per scanned row.
    // A selective join (config 5: 4 % of the probe side's rows survive) then pays the expensive half for those rows only.
    bool compact_tail = false;
    int probe_dst = 0;
    if (probe_pc >= 0) {
      probe_dst = (sd->code[probe_pc] >> 13) & 0x7f;
      const uint32_t n0 = sd->code[probe_pc + 2], n1 = sd->code[probe_pc + 3];
      const bool match_filter = (n0 & 0x3f) == MSC_OP_GE_I && ((n0 >> 6) & 7) == MSC_DST_FILTER && ((n0 >> 9) & 0xf) == 0 &&
                                (n1 & 0xffffu) == static_cast<uint32_t>((MSC_SRC_TEMP << 12) | probe_dst) && ((n1 >> 28) & 7) == MSC_SRC_CONST &&
                                sd->consts[(n1 >> 16) & 0xfff] == 0;
      uint32_t early = 0;  // temporaries written before the probe
      for (int pc = 0; pc < probe_pc; pc += 2) {
        const uint32_t w0 = sd->code[pc];
        if (((w0 >> 6) & 7) == MSC_DST_TEMP) early |= 1u << ((w0 >> 13) & 0x7f);
        if ((w0 >> 9) & 0xf) early |= 1u << (((w0 >> 9) & 0xf) - 1);
      }
      bool reads_early = false;
      for (int pc = probe_pc + 4; pc + 1 < sd->ncode; pc += 2) {
        const uint32_t w0 = sd->code[pc], w1 = sd->code[pc + 1];
        if ((w0 & 0x3f) == MSC_OP_END) break;
        for (int side = 0; side < 2; ++side) {
          const uint32_t opnd = (w1 >> (16 * side)) & 0xffffu;
          const int kind = (opnd >> 12) & 7, idx = opnd & 0xfff;
          if (kind == MSC_SRC_TEMP && ((early >> idx) & 1u) && idx != probe_dst) reads_early = true;
          if (kind == MSC_SRC_GATHER_T && (idx >> 6) != probe_dst && ((early >> (idx >> 6)) & 1u)) reads_early = true;
        }
        const int dk = (w0 >> 6) & 7;  // a later instruction that rewrites an early temporary makes it "late" from there on: keep it simple
        if (dk == MSC_DST_TEMP && ((early >> ((w0 >> 13) & 0x7f)) & 1u)) reads_early = true;
      }
      static const bool compaction_enabled = !(getenv("MSC_JIT_COMPACT_TAIL") && atoi(getenv("MSC_JIT_COMPACT_TAIL")) == 0);
      compact_tail = match_filter && !reads_early && compaction_enabled && fin == nullptr;
    }
    if (compact_tail) {
      const uint32_t w1 = sd->code[probe_pc + 1];
      o << "    u32 hits = 0;\n#pragma unroll\n    for (int r = 0; r < R; ++r) {\n      const bool valid = (vm >> r) & 1u;\n"
        << "      t" << probe_dst << "[r] = join_probe_resolve(p.luts[" << ((w1 >> 16) & 0x7ff) << "], pq[r], valid, keep_policy);  // instruction " << probe_pc / 2
        << ", second half\n      if (valid && t" << probe_dst << "[r] >= 0) hits |= 1u << r;  // instruction " << probe_pc / 2 + 1 << "\n    }\n";
      o << R"(    u32 survivors;
    {
      u32 at = warp_exclusive_scan(__popc(hits), lane, &survivors);
#pragma unroll
      for (int r = 0; r < R; ++r)
        if ((hits >> r) & 1u) {
          queue[2 * at] = (r / 4) * 128 + 4 * lane + (r % 4);
)" << "          queue[2 * at + 1] = static_cast<u32>(t" << probe_dst << R"([r]);
          ++at;
        }
    }
    __syncwarp();
    for (u32 e = lane; e < survivors; e += 32) {
      const u32 row = queue[2 * e];
      bool valid = true;
      int grp = -1;
)";
      if (smem_cells) o << "      i64* cellrow = cells + NG * (STRIDE * NT);\n";
      if (masked) {
        o << "      double";
        for (int g = 0; g < (ngroups + 1) / 2 * 2; ++g) o << (g ? ", m" : " m") << g << " = 0.0";
        o << ";\n";
      }
      for (int t = 0; t < sd->ntemps; ++t) o << "      i64 w" << t << " = " << (t == probe_dst ? "queue[2 * e + 1]" : "0") << ";\n";
      tail_mode = true;
      tail_lay = lay;
    } else {
    o << "#pragma unroll\n    for (int r = 0; r < R; ++r) {\n      bool valid = "
      << (probe_pc >= 0 ? "(vm >> r) & 1u" : (valid_bits ? "(vmask >> r) & 1u" : "true")) << ";\n      int grp = -1;\n";
    if (smem_cells) o << "      i64* cellrow = cells + NG * (STRIDE * NT);\n";
    if (masked) {
      o << "      double";
      for (int g = 0; g < (ngroups + 1) / 2 * 2; ++g) o << (g ? ", m" : " m") << g << " = 0.0";
      o << ";\n";
    }
    if (probe_pc < 0) {
      for (int t = 0; t < sd->ntemps; ++t) o << "      i64 t" << t << " = 0;\n";
    } else {
      const uint32_t w0 = sd->code[probe_pc], w1 = sd->code[probe_pc + 1];
      o << "      " << temp((w0 >> 13) & 0x7f) << " = join_probe_resolve(p.luts[" << ((w1 >> 16) & 0x7ff) << "], pq[r], valid, keep_policy);  // instruction "
        << probe_pc / 2 << ", second half\n";
    }
    }
    for (int pc = probe_pc >= 0 ? probe_pc + (compact_tail ? 4 : 2) : 0; pc + 1 < sd->ncode; pc += 2) {
      const uint32_t w0 = sd->code[pc], w1 = sd->code[pc + 1];
      const int op = w0 & 0x3f;
      if (op == MSC_OP_END) break;
      if (op == MSC_OP_RANK) {
        why = "RANK in an aggregate scan";
        return false;
      }
      const int dkind = (w0 >> 6) & 7, tee = (w0 >> 9) & 0xf, dst = (w0 >> 13) & 0x7f;
      const uint32_t oa = w1 & 0xffffu, ob = w1 >> 16;
      const bool lut = op == MSC_OP_LUT8 || op == MSC_OP_LUT32 || op == MSC_OP_PROBE;
      const std::string a = operand(oa, &ok), b = lut ? std::string("0ll") : operand(ob, &ok);
      o << "      {  // instruction " << pc / 2 << "\n";
      if (op == MSC_OP_DIV_F || op == MSC_OP_FLOORDIV_F || op == MSC_OP_MOD_F)
        o << "        bad |= valid && (l2d(" << b << ") == 0.0);\n";
      if (op == MSC_OP_FLOORDIV_I || op == MSC_OP_MOD_I) o << "        bad |= valid && ((" << b << ") == 0);\n";
      const bool plain_count = dkind == MSC_DST_AGG && counted[dst];
      if (!plain_count) o << "        const i64 x = " << compute(op, a, b, ob, &ok) << ";\n";
      if (tee) o << "        " << temp(tee - 1) << " = x;\n";
      switch (dkind) {
        case MSC_DST_TEMP: o << "        " << temp(dst) << " = x;\n"; break;
        case MSC_DST_FILTER:
          o << "        valid = valid && (x != 0);\n";
          if (smem_cells && grouped) o << "        cellrow = cells + ((valid && grp >= 0) ? grp : NG) * (STRIDE * NT);\n";
          break;
        case MSC_DST_GROUP:
          o << "        grp = (x >= 0 && x < NG) ? (int)x : -1;\n";
          if (smem_cells) o << "        cellrow = cells + ((valid && grp >= 0) ? grp : NG) * (STRIDE * NT);\n";
          if (masked) {
            o << "        { const double2* mrow = reinterpret_cast<const double2*>(mlut + (valid ? grp + 1 : 0) * NGP);\n";
            for (int g = 0; g < (ngroups + 1) / 2 * 2; g += 2)
              o << "          { const double2 mm = mrow[" << g / 2 << "]; m" << g << " = mm.x; m" << g + 1 << " = mm.y; }\n";
            o << "        }\n";
          }
          if (stride > naggs) emit_agg(naggs, "");
          grouped = true;
          break;
        case MSC_DST_AGG:
          if (!grouped) {
            why = "aggregate before GROUP";
            return false;
          }
          emit_agg(dst, "x");
          break;
        case MSC_DST_NONE: break;
        default: why = "destination kind outside an aggregate scan"; return false;
      }
      o << "      }\n";
    }
    if (!ok) {
      why = "operand or opcode outside the generator";
      return false;
    }
    if (!grouped) {
      why = "program has no GROUP";
      return false;
    }
    o << "    }\n";
    if (compact_tail) o << "    __syncwarp();  // the queue is rewritten by the next tile\n";
    tail_mode = false;
    // fold the tile's u32 row counts into their i64 accumulators (8 rows per lane and tile: no overflow)
    for (int g = 0; g < ngroups && !smem_cells; ++g)
      for (int s = 0; s < stride; ++s)
        if (counted[s] && !masked) o << "    " << acc(g, s) << " += (i64)" << cnt(g, s) << " * " << count_mul(s) << "; " << cnt(g, s) << " = 0;\n";
    o << R"(    __syncwarp();
    if (k + NSTAGES < ntiles_w) issue_tile(p, stages, full, stage, gw + (u64)(k + NSTAGES) * nw, lane);
    if (++stage == NSTAGES) {
      stage = 0;
      parity ^= 1u;
    }
  }
  if (bad) atomicOr(p.err, 1);  // MSC_DEVERR_DIV_ZERO
)";
    if (smem_cells) {
      // every thread's copy of a cell, folded by one warp: 4 x 32 consecutive words, a shuffle tree, one atomic per CTA and cell
      o << R"(  __syncthreads();
  for (int cell = warp; cell < NG * STRIDE; cell += NW) {
    const int kind = KIND[cell % STRIDE];
    const i64* copies = cellbase + cell * NT;
    i64 v = copies[lane];
#pragma unroll
    for (int w = 1; w < NW; ++w) v = agg_combine_k(kind, v, copies[lane + 32 * w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = agg_combine_k(kind, v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0 && v != INIT[cell % STRIDE]) atomic_fold(kind, p.dense_out + cell, v);
  }
)";
    } else {
    if (masked)  // row counts were summed as f64 (exact below 2^53)
      for (int g = 0; g < ngroups; ++g)
        for (int s = 0; s < stride; ++s)
          if (counted[s]) o << "  " << acc(g, s) << " += (i64)" << cnt(g, s) << " * " << count_mul(s) << ";\n";
    for (int g = 0; g < ngroups; ++g)
      for (int s = 0; s < stride; ++s) {
        const std::string raw = is_float_kind(kinds[s]) ? "d2l(" + acc(g, s) + ")" : acc(g, s);
        o << "  { const i64 v = warp_fold<" << kinds[s] << ">(" << raw << "); if (lane == 0) red[warp * (NG * STRIDE) + " << g * stride + s << "] = v; }\n";
      }
    o << R"(  __syncthreads();
  for (int cell = tid; cell < NG * STRIDE; cell += NT) {
    const int kind = KIND[cell % STRIDE];
    i64 v = red[cell];
#pragma unroll
    for (int w = 1; w < NW; ++w) v = agg_combine_k(kind, v, red[w * (NG * STRIDE) + cell]);
    if (v != INIT[cell % STRIDE]) atomic_fold(kind, p.dense_out + cell, v);
  }
)";
    }
    if (fin != nullptr && !emit_finish()) return false;
    o << "}\n";
    return true;
  }

  // The last CTA to fold its accumulators into the table also finishes the query: lane g of its first warp takes group
  // g -- compaction of the groups that received rows (dense_finalize_kernel) and the final projection over them (AVG =
  // SUM / COUNT, HAVING, output order: what msc_scan_project would run as a second scan) -- and leaves the row count,
  // the non-finite flag and the device error word in p.fmeta.  No launch after the scan: prepared queries need it
  // (bench/step_probe.py: the three follow-up launches and their gaps cost ~50 us of a 0.44 ms step).
  bool emit_finish() {
    if (ngroups > 32) {
      why = "fused finish handles at most 32 groups";
      return false;
    }
    in_finish = true;
    temp_arrays = false;
    o << R"(  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last && warp == 0) {
    __threadfence();
)";
    if (!fin_peer) {
      o << R"(    const int g = lane;
    const bool in_range = g < NG;
    const u64* cell = p.dense_out + (in_range ? g : 0) * STRIDE;
    i64 a[STRIDE];
#pragma unroll
    for (int s = 0; s < STRIDE; ++s) a[s] = static_cast<i64>(__ldcg(cell + s));
    if (in_range) {  // leave the identities behind: the next pass of a prepared query needs no initialisation launch
#pragma unroll
      for (int s = 0; s < STRIDE; ++s) p.dense_out[g * STRIDE + s] = static_cast<u64>(INIT[s]);
    }
)";
    } else {
      // the exchange, inside the same kernel: push this rank's table into every rank's mailbox (peer stores over NVLink),
      // publish the epoch, wait for the W epochs of the own mailbox, fold the W tables in rank order
      o << R"(    const u32 set = static_cast<u32>(p.epoch & 1u);
    const u64 flag_off = static_cast<u64>(set) * p.world, slot_off = 2ull * p.world + (static_cast<u64>(set) * p.world) * p.slot_cells;
    for (int peer = 0; peer < p.world; ++peer) {
      u64* dst = p.mailbox[peer] + slot_off + static_cast<u64>(p.rank) * p.slot_cells;
      for (int i = lane; i < p.nlocal * STRIDE; i += 32) dst[i] = __ldcg(p.dense_out + i);
    }
    __syncwarp();
    for (int i = lane; i < p.nlocal * STRIDE; i += 32) p.dense_out[i] = static_cast<u64>(INIT[i % STRIDE]);  // identities for the next pass
    __threadfence_system();
    __syncwarp();
    if (lane < p.world) st_release_sys(p.mailbox[lane] + flag_off + p.rank, p.epoch);
    bool late = false;
    if (lane < p.world) {
      const u64* flag = p.mailbox[p.rank] + flag_off + lane;
      const long long t0 = clock64();
      while (ld_acquire_sys(flag) != p.epoch) {
        if (clock64() - t0 > 20000000000ll) {  // ~10 s: a peer never launched this pass
          late = true;
          break;
        }
        __nanosleep(200);
      }
    }
    if (__any_sync(0xffffffffu, late) && lane == 0) atomicOr(p.err, 32);  // MSC_DEVERR_PEER_TIMEOUT
    __syncwarp();
    const int g = lane;  // merged (global) group
    const bool in_range = g < p.nglobal;
    i64 a[STRIDE];
#pragma unroll
    for (int s = 0; s < STRIDE; ++s) a[s] = INIT[s];
    for (int r = 0; r < p.world; ++r) {
      const int lg = in_range ? p.inv[r * 32 + g] : -1;
      if (lg < 0) continue;
      const u64* cell = p.mailbox[p.rank] + slot_off + static_cast<u64>(r) * p.slot_cells + static_cast<u64>(lg) * STRIDE;
#pragma unroll
      for (int s = 0; s < STRIDE; ++s) a[s] = agg_combine_k(KIND[s], a[s], static_cast<i64>(ld_volatile(cell + s)));
    }
)";
    }
    o << "    bool valid = in_range && a[" << count_slot << "] != 0;\n    bool nonfinite = false, bad = false;\n";
    for (int s = 0; s < naggs; ++s)
      if (kinds[s] == MSC_AGG_SUM_F) o << "    nonfinite |= in_range && !isfinite(l2d(a[" << s << "]));\n";
    for (int c = 0; c < fin->nstaged; ++c) {
      if (fin_cols[c] < 0 || fin_cols[c] > naggs) {
        why = "final projection refers to a column the aggregate does not have";
        return false;
      }
      o << "    const i64 f" << c << " = " << (fin_cols[c] == 0 ? std::string("(i64)g") : "a[" + std::to_string(fin_cols[c] - 1) + "]") << ";\n";
    }
    for (int t = 0; t < fin->ntemps; ++t) o << "    i64 ft" << t << " = 0;\n";
    for (int c = 0; c < fin_nout; ++c) o << "    i64 fo" << c << " = 0;\n";
    bool ok = true;
    for (int pc = 0; pc + 1 < fin->ncode; pc += 2) {
      const uint32_t w0 = fin->code[pc], w1 = fin->code[pc + 1];
      const int op = w0 & 0x3f;
      if (op == MSC_OP_END) break;
      if (op == MSC_OP_RANK) continue;  // positions come from the ballot below
      if (op == MSC_OP_LUT8 || op == MSC_OP_LUT32) {
        why = "lookup table in the final projection";
        return false;
      }
      const int dkind = (w0 >> 6) & 7, tee = (w0 >> 9) & 0xf, dst = (w0 >> 13) & 0x7f;
      const uint32_t oa = w1 & 0xffffu, ob = w1 >> 16;
      const std::string x = compute(op, operand(oa, &ok), operand(ob, &ok), ob, &ok);
      o << "    {\n";
      if (op == MSC_OP_DIV_F || op == MSC_OP_FLOORDIV_F || op == MSC_OP_MOD_F) o << "      bad |= valid && (l2d(" << operand(ob, &ok) << ") == 0.0);\n";
      if (op == MSC_OP_FLOORDIV_I || op == MSC_OP_MOD_I) o << "      bad |= valid && ((" << operand(ob, &ok) << ") == 0);\n";
      o << "      const i64 x = " << x << ";\n";
      if (tee) o << "      " << temp(tee - 1) << " = x;\n";
      switch (dkind) {
        case MSC_DST_TEMP: o << "      " << temp(dst) << " = x;\n"; break;
        case MSC_DST_FILTER: o << "      valid = valid && (x != 0);\n"; break;
        case MSC_DST_OUT:
          if (dst >= fin_nout) {
            why = "final projection writes a column it does not have";
            return false;
          }
          o << "      fo" << dst << " = x;\n";
          break;
        case MSC_DST_NONE: break;
        default: why = "destination kind outside a final projection"; return false;
      }
      o << "    }\n";
    }
    if (!ok) {
      why = "operand or opcode outside the generator (final projection)";
      return false;
    }
    o << "    const u32 keep = __ballot_sync(0xffffffffu, valid);\n    const int pos = __popc(keep & ((1u << lane) - 1u));\n    if (valid) {\n";
    for (int c = 0; c < fin_nout; ++c) {
      const std::string ty = fin_phys[c] == MSC_P_U32 ? "u32" : "i64";
      o << "      reinterpret_cast<" << ty << "*>(p.out[" << c << "])[pos] = (" << ty << ")fo" << c << ";\n";
    }
    o << R"(    }
    const bool any_nonfinite = __any_sync(0xffffffffu, nonfinite), any_bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
      p.fmeta[0] = __popc(keep);
      p.fmeta[1] = any_nonfinite ? 1 : 0;
      p.fmeta[2] = static_cast<u64>(*reinterpret_cast<volatile int*>(p.err) | (any_bad ? 1 : 0));
      *p.err = 0;
      *p.ticket = 0;
    }
  }
)";
    in_finish = false;
    return true;
  }
};

// ---- kernel cache ------------------------------------------------------------------------------------------------
struct Kernel {
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  std::vector<char> cubin;
  int regs = 0, occ = 0, nstages = 0;
  size_t smem = 0;
};

std::unordered_map<std::string, std::unique_ptr<Kernel>>& cache() {
  static std::unordered_map<std::string, std::unique_ptr<Kernel>> c;
  return c;
}

uint64_t fnv1a(const std::string& s) {
  uint64_t h = 1469598103934665603ull;
  for (unsigned char ch : s) {
    h ^= ch;
    h *= 1099511628211ull;
  }
  return h;
}

std::string disk_cache_path(const std::string& source) {
  const char* dir = getenv("MSC_JIT_CACHE");
  if (!dir || !*dir) return "";
  int major = 0, minor = 0;
  api().nvrtcVersion(&major, &minor);
  char name[96];
  snprintf(name, sizeof(name), "/msc_jit_%016llx_%zu_nvrtc%d.%d_sm100a.cubin", static_cast<unsigned long long>(fnv1a(source)), source.size(), major, minor);
  return std::string(dir) + name;
}

int compile(const std::string& source, std::vector<char>* cubin, std::string* err) {
  if (!load_nvrtc(err)) return MSC_ERR_ARG;
  const std::string path = disk_cache_path(source);
  if (!path.empty()) {
    std::ifstream f(path, std::ios::binary);
    if (f) {
      cubin->assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
      if (!cubin->empty()) return MSC_OK;
    }
  }
  Api& a = api();
  nvrtcProgram prog = nullptr;
  // MSC_JIT_DUMP_DIR=<dir>: keep the generated source as a file and compile it under that name, so that profilers
  // (ncu --import-source) and cuobjdump -lineinfo can show it
  std::string name = "msc_jit_dense.cu";
  if (const char* dump = getenv("MSC_JIT_DUMP_DIR")) {
    mkdir(dump, 0755);
    char file[64];
    snprintf(file, sizeof(file), "/msc_jit_%016llx.cu", static_cast<unsigned long long>(fnv1a(source)));
    name = std::string(dump) + file;
    std::ofstream f(name);
    f << source;
  }
  if (a.nvrtcCreateProgram(&prog, source.c_str(), name.c_str(), 0, nullptr, nullptr) != 0) {
    *err = "nvrtcCreateProgram failed";
    return MSC_ERR_ARG;
  }
  // --fmad=false: Python rounds every operation (sql.py:262-266); a contracted a * b + c would not
  const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device", "--fmad=false"};
  const int rc = a.nvrtcCompileProgram(prog, 5, opts);
  if (rc != 0) {
    size_t n = 0;
    a.nvrtcGetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) a.nvrtcGetProgramLog(prog, &log[0]);
    a.nvrtcDestroyProgram(&prog);
    if (getenv("MSC_JIT_DUMP")) fprintf(stderr, "%s\n", source.c_str());
    *err = "compilation failed: " + log.substr(0, 1500);
    return MSC_ERR_ARG;
  }
  size_t n = 0;
  a.nvrtcGetCUBINSize(prog, &n);
  cubin->resize(n);
  a.nvrtcGetCUBIN(prog, cubin->data());
  a.nvrtcDestroyProgram(&prog);
  if (!path.empty()) {
    mkdir(getenv("MSC_JIT_CACHE"), 0755);
    const std::string tmp = path + ".tmp" + std::to_string(getpid());
    std::ofstream f(tmp, std::ios::binary);
    f.write(cubin->data(), static_cast<std::streamsize>(cubin->size()));
    f.close();
    rename(tmp.c_str(), path.c_str());
  }
  return MSC_OK;
}

int cu_fail(msc_ctx* ctx, const char* what, int rc) {
  const char* s = nullptr;
  if (api().cuGetErrorString) api().cuGetErrorString(rc, &s);
  return ctx->fail(MSC_ERR_CUDA, std::string("jit: ") + what + " -> " + (s ? s : "error"));
}

// generated source for this query, or "" with ctx->err set
int generate(const msc_scan_desc* sd, int ngroups, int naggs, int stride, const int* kinds, const long long* init, int nstages, bool masked,
             std::string* source, std::string* err, const JitFinish* fin = nullptr) {
  Gen g{sd, ngroups, naggs, stride, kinds, init, nstages, masked};
  g.smem_cells = jit_dense_cells_in_smem(ngroups, stride);
  if (g.smem_cells && masked) {
    *err = "the mask-table form keeps its accumulators in registers";
    return MSC_ERR_ARG;
  }
  if (fin) {
    g.fin = fin->scan;
    g.fin_cols = fin->cols;
    g.fin_phys = fin->out_phys;
    g.fin_nout = fin->nout;
    g.count_slot = fin->count_slot;
    g.fin_peer = fin->peer != nullptr;
  }
  if (!g.generate()) {
    *err = g.why;
    return MSC_ERR_ARG;
  }
  *source = g.o.str();
  return MSC_OK;
}

}  // namespace

bool jit_dense_cells_in_smem(int ngroups, int stride) {
  static const int from = getenv("MSC_JIT_SMEM_CELLS_FROM") ? atoi(getenv("MSC_JIT_SMEM_CELLS_FROM")) : JIT_REG_CELLS_4CTAS + 1;
  return ngroups * stride >= from;
}

bool jit_dense_supported(const msc_scan_desc* sd, int ngroups, int stride) {
  // register accumulators: groups x accumulators x 2 registers must leave room for the rows in flight; shared-memory cells: a
  // private copy per thread, 1 KB per cell and CTA
  const int most = jit_dense_cells_in_smem(ngroups, stride) ? JIT_MAX_SMEM_CELLS : JIT_MAX_REG_CELLS;
  return ngroups >= 1 && ngroups * stride <= most && sd->nstaged >= 1 && sd->nstaged <= 24;
}

// masked: try the mask-table variant first, fall back to the exact one when the program does not allow it
int generate_either(const msc_scan_desc* sd, int ngroups, int naggs, int stride, const int* kinds, const long long* init, bool* masked,
                    std::string* source, std::string* err, const JitFinish* fin = nullptr) {
  if (*masked && generate(sd, ngroups, naggs, stride, kinds, init, 2, true, source, err, fin) == MSC_OK) return MSC_OK;
  *masked = false;
  return generate(sd, ngroups, naggs, stride, kinds, init, 2, false, source, err, fin);
}

int jit_dense_source(const msc_scan_desc* sd, int ngroups, int naggs, int stride, const int* kinds, const long long* init, bool masked,
                     std::string* source, std::string* err, const JitFinish* fin) {
  return generate_either(sd, ngroups, naggs, stride, kinds, init, &masked, source, err, fin);
}

int jit_compile_source(const std::string& source, std::vector<char>* cubin, std::string* err) { return compile(source, cubin, err); }

// Everything the generated source depends on, as bytes: a launch finds its kernel through this key without building
// the source text again (constants are kernel parameters, so a different date literal reuses the kernel).
std::string shape_key(msc_ctx* ctx, const msc_scan_desc* sd, int ngroups, int naggs, int stride, const int* kinds, const long long* init, bool masked) {
  std::string k;
  auto put = [&](const void* p, size_t n) { k.append(static_cast<const char*>(p), n); };
  const int head[8] = {ctx->device, masked ? 1 : 0, ngroups, naggs, stride, sd->nstaged, sd->ngather, sd->ntemps};
  put(head, sizeof(head));
  put(kinds, sizeof(int) * stride);
  put(init, sizeof(long long) * stride);
  for (int c = 0; c < sd->nstaged; ++c) put(&sd->staged[c].phys, sizeof(int32_t));
  for (int c = 0; c < sd->ngather; ++c) put(&sd->gather[c].phys, sizeof(int32_t));
  int n = 0;
  while (n + 1 < sd->ncode && (sd->code[n] & 0x3f) != MSC_OP_END) n += 2;
  put(sd->code, sizeof(uint32_t) * n);
  return k;
}

struct ShapeEntry {
  Kernel* kernel;
  bool masked;  // the variant that was generated (a program may refuse the masked form)
};
std::unordered_map<std::string, ShapeEntry>& shapes() {
  static std::unordered_map<std::string, ShapeEntry> m;
  return m;
}

bool jit_dense_cached(msc_ctx* ctx, const msc_scan_desc* sd, int ngroups, int naggs, int stride, const int* kinds, const long long* init,
                      bool masked) {
  return shapes().count(shape_key(ctx, sd, ngroups, naggs, stride, kinds, init, masked)) != 0;
}

namespace {

// compile (or find) the kernel of `source`, load it into this device's context and size its launch
int load_kernel(msc_ctx* ctx, const std::string& source, const char* fn_name, size_t smem, Kernel** out) {
  std::string why;
  const std::string key = std::to_string(ctx->device) + "#" + source;
  auto it = cache().find(key);
  if (it == cache().end()) {
    if (!load_driver(&why)) return ctx->fail(MSC_ERR_CUDA, "jit: " + why);
    auto k = std::make_unique<Kernel>();
    const auto t0 = std::chrono::steady_clock::now();
    if (compile(source, &k->cubin, &why) != MSC_OK) return ctx->fail(MSC_ERR_ARG, "jit: " + why);
    ctx->stats.last_jit_compile_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    Api& a = api();
    int rc = a.cuModuleLoadData(&k->mod, k->cubin.data());
    if (rc != 0) return cu_fail(ctx, "cuModuleLoadData", rc);
    rc = a.cuModuleGetFunction(&k->fn, k->mod, fn_name);
    if (rc != 0) return cu_fail(ctx, "cuModuleGetFunction", rc);
    k->nstages = 2;
    k->smem = smem;
    if (k->smem > 227 * 1024) return ctx->fail(MSC_ERR_ARG, "jit: scan needs more shared memory than an SM has");
    rc = a.cuFuncSetAttribute(k->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, static_cast<int>(k->smem));
    if (rc != 0) return cu_fail(ctx, "cuFuncSetAttribute", rc);
    a.cuFuncGetAttribute(&k->regs, CU_FUNC_ATTRIBUTE_NUM_REGS, k->fn);
    rc = a.cuOccupancyMaxActiveBlocksPerMultiprocessor(&k->occ, k->fn, NT, k->smem);
    if (rc != 0 || k->occ < 1) return cu_fail(ctx, "cuOccupancyMaxActiveBlocksPerMultiprocessor", rc);
    it = cache().emplace(key, std::move(k)).first;
    ctx->stats.jit_compiles += 1;
  }
  *out = it->second.get();
  return MSC_OK;
}

int fill_params(msc_ctx* ctx, const msc_scan_desc* sd, JitParams* p) {
  memset(p, 0, sizeof(*p));
  p->nrows = sd->nrows;
  p->nrows_dev = reinterpret_cast<const unsigned long long*>(sd->nrows_dev);
  p->ntiles = static_cast<uint32_t>((sd->nrows + 255) / 256);
  for (int c = 0; c < sd->nstaged; ++c) {
    if (!sd->staged[c].data || (reinterpret_cast<uintptr_t>(sd->staged[c].data) & 15) != 0) return ctx->fail(MSC_ERR_ARG, "staged column not 16B aligned");
    p->col[c] = static_cast<const unsigned char*>(sd->staged[c].data);
  }
  for (int c = 0; c < sd->ngather; ++c) p->gather[c] = sd->gather[c].data;
  for (int c = 0; c < sd->nluts; ++c) p->luts[c] = sd->luts[c];
  memcpy(p->consts, sd->consts, sizeof(int64_t) * sd->nconsts);
  p->err = ctx->d_err;
  return MSC_OK;
}

int launch(msc_ctx* ctx, Kernel& k, JitParams* p, bool timed) {
  uint64_t grid = static_cast<uint64_t>(ctx->sm_count) * k.occ;
  const uint64_t need = (p->ntiles + NW - 1) / NW;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  void* args[] = {p};
  if (timed) MSC_CUDA(ctx, cudaEventRecord(ctx->ev_s0, ctx->stream));
  const int rc = api().cuLaunchKernel(k.fn, static_cast<unsigned>(grid), 1, 1, NT, 1, 1, static_cast<unsigned>(k.smem),
                                      reinterpret_cast<CUstream>(ctx->stream), args, nullptr);
  if (rc != 0) return cu_fail(ctx, "cuLaunchKernel", rc);
  ctx->stats.launches += 1;
  if (timed) {
    MSC_CUDA(ctx, cudaEventRecord(ctx->ev_s1, ctx->stream));
    ctx->stats.last_scan_grid = static_cast<int32_t>(grid);
    ctx->stats.last_scan_stages = k.nstages;
    ctx->stats.last_scan_smem = static_cast<int32_t>(k.smem);
    ctx->stats.last_scan_rows_per_thread = 8;
    ctx->stats.last_scan_kind = MSC_SCAN_KIND_JIT;
    ctx->stats.last_scan_regs = k.regs;
  }
  return MSC_OK;
}

std::string project_key(msc_ctx* ctx, const msc_scan_desc* sd, bool count_only, const int32_t* out_phys, int nout) {
  const int kinds[1] = {count_only ? -2 : -3};  // keeps these keys apart from the dense ones
  const long long init[1] = {nout};
  std::string k = shape_key(ctx, sd, 0, 0, 1, kinds, init, false);
  if (!count_only) k.append(reinterpret_cast<const char*>(out_phys), sizeof(int32_t) * nout);
  return k;
}

int project_source(const msc_scan_desc* sd, bool count_only, const int32_t* out_phys, int nout, std::string* source, std::string* err) {
  Gen g{sd, 0, 0, 0, nullptr, nullptr, 2, false};
  if (!g.generate_project(count_only, out_phys, nout)) {
    *err = g.why;
    return MSC_ERR_ARG;
  }
  *source = g.o.str();
  return MSC_OK;
}

}  // namespace

int jit_dense_launch(msc_ctx* ctx, const msc_scan_desc* sd, int ngroups, int naggs, int stride, const int* kinds, const long long* init,
                     unsigned long long* table, bool timed, bool* masked, const JitFinish* fin) {
  std::string skey = shape_key(ctx, sd, ngroups, naggs, stride, kinds, init, *masked);
  if (fin) {  // the final projection is part of the kernel text
    const int head[4] = {fin->peer ? -8 : -7, fin->nout, fin->count_slot, fin->scan->nstaged};
    skey.append(reinterpret_cast<const char*>(head), sizeof(head));
    skey.append(reinterpret_cast<const char*>(fin->cols), sizeof(int32_t) * fin->scan->nstaged);
    skey.append(reinterpret_cast<const char*>(fin->out_phys), sizeof(int32_t) * fin->nout);
    int n = 0;
    while (n + 1 < fin->scan->ncode && (fin->scan->code[n] & 0x3f) != MSC_OP_END) n += 2;
    skey.append(reinterpret_cast<const char*>(fin->scan->code), sizeof(uint32_t) * n);
  }
  auto sit = shapes().find(skey);
  if (sit == shapes().end()) {
    std::string source, why;
    if (generate_either(sd, ngroups, naggs, stride, kinds, init, masked, &source, &why, fin) != MSC_OK)
      return ctx->fail(MSC_ERR_ARG, "jit: " + why);
    const Layout lay = stage_layout(sd);
    const size_t acc_bytes = jit_dense_cells_in_smem(ngroups, stride)
                                 ? static_cast<size_t>(ngroups + 1) * stride * NT * 8
                                 : static_cast<size_t>(NW) * ngroups * stride * 8 + static_cast<size_t>(ngroups + 1) * ((ngroups + 1) / 2 * 2) * 8;
    const size_t smem = 4 * 8 * 8 + static_cast<size_t>(NW) * 2 * lay.stage_bytes + acc_bytes + (program_probes(sd) ? static_cast<size_t>(NW) * 2 * 256 * 4 : 0);
    Kernel* k = nullptr;
    MSC_TRY(load_kernel(ctx, source, "msc_jit_dense", smem, &k));
    sit = shapes().emplace(skey, ShapeEntry{k, *masked}).first;
  }
  *masked = sit->second.masked;
  Kernel& k = *sit->second.kernel;
  if (fin && (fin->compile_only || (fin->peer && fin->peer->compile_only))) return MSC_OK;
  if (sd->nrows == 0 && !(fin && fin->peer)) return MSC_OK;  // (a rank without rows still takes part in the exchange)
  JitParams p;
  MSC_TRY(fill_params(ctx, sd, &p));
  p.dense_out = table;
  if (fin && fin->peer) {
    const msc_peer_spec& ps = *fin->peer;
    for (int r = 0; r < ps.world; ++r) p.mailbox[r] = static_cast<unsigned long long*>(ps.mailbox[r]);
    p.inv = ps.inv;
    p.epoch = ps.epoch;
    p.rank = ps.rank;
    p.world = ps.world;
    p.nlocal = ps.nlocal;
    p.slot_cells = ps.gmax * stride;
    p.nglobal = ps.nglobal;
  }
  if (fin) {
    memcpy(p.fconsts, fin->scan->consts, sizeof(int64_t) * fin->scan->nconsts);
    for (int c = 0; c < fin->nout; ++c) p.out[c] = fin->outs[c];
    p.fmeta = fin->meta;
    p.ticket = fin->ticket;
  }
  return launch(ctx, k, &p, timed);
}

namespace {
std::string runs_key(msc_ctx* ctx, const msc_scan_desc* sd, int naggs, const int* kinds, int key_col) {
  const long long init[1] = {key_col};
  int k2[MSC_VM_MAX_AGGS + 1];
  k2[0] = -4;  // keeps these keys apart from the dense / project ones
  for (int a = 0; a < naggs; ++a) k2[1 + a] = kinds[a];
  (void)init;
  std::string k = shape_key(ctx, sd, 0, naggs, 1, k2, init, false);
  k.append(reinterpret_cast<const char*>(k2), sizeof(int) * (1 + naggs));
  return k;
}
int runs_kernel(msc_ctx* ctx, const msc_scan_desc* sd, int naggs, const int* kinds, int key_col, Kernel** out) {
  const std::string skey = runs_key(ctx, sd, naggs, kinds, key_col);
  auto sit = shapes().find(skey);
  if (sit == shapes().end()) {
    Gen g{sd, 0, naggs, naggs, kinds, nullptr, 2, false};
    if (!g.generate_runs(key_col)) return ctx->fail(MSC_ERR_ARG, "jit: " + g.why);
    const Layout lay = stage_layout(sd);
    Kernel* k = nullptr;
    MSC_TRY(load_kernel(ctx, g.o.str(), "msc_jit_runs", 4 * 8 * 8 + static_cast<size_t>(NW) * 2 * lay.stage_bytes, &k));
    sit = shapes().emplace(skey, ShapeEntry{k, false}).first;
  }
  *out = sit->second.kernel;
  return MSC_OK;
}
}  // namespace

bool jit_runs_cached(msc_ctx* ctx, const msc_scan_desc* sd, int naggs, const int* kinds, int key_col) {
  return shapes().count(runs_key(ctx, sd, naggs, kinds, key_col)) != 0;
}

int jit_runs_source(const msc_scan_desc* sd, int naggs, const int* kinds, int key_col, std::string* source, std::string* err) {
  Gen g{sd, 0, naggs, naggs, kinds, nullptr, 2, false};
  if (!g.generate_runs(key_col)) {
    *err = g.why;
    return MSC_ERR_ARG;
  }
  *source = g.o.str();
  return MSC_OK;
}

int jit_runs_launch(msc_ctx* ctx, const msc_scan_desc* sd, int naggs, const int* kinds, int key_col, const uint64_t* tile_offsets, void* const* outs,
                    unsigned long long* carry, bool timed) {
  Kernel* k = nullptr;
  MSC_TRY(runs_kernel(ctx, sd, naggs, kinds, key_col, &k));
  JitParams p;
  MSC_TRY(fill_params(ctx, sd, &p));
  p.tile_offsets = reinterpret_cast<const unsigned long long*>(tile_offsets);
  p.dense_out = carry;  // [accumulator][tile]: what a tile adds to the run that was open when it began (holding the identities)
  for (int c = 0; c < 1 + naggs; ++c) p.out[c] = outs[c];
  return launch(ctx, *k, &p, timed);
}

int jit_project_source(const msc_scan_desc* sd, bool count_only, const int32_t* out_phys, int nout, std::string* source, std::string* err) {
  return project_source(sd, count_only, out_phys, nout, source, err);
}

bool jit_project_cached(msc_ctx* ctx, const msc_scan_desc* sd, const int32_t* out_phys, int nout) {
  return shapes().count(project_key(ctx, sd, false, out_phys, nout)) != 0;
}

namespace {
int project_kernel(msc_ctx* ctx, const msc_scan_desc* sd, bool count_only, const int32_t* out_phys, int nout, Kernel** out) {
  const std::string skey = project_key(ctx, sd, count_only, out_phys, nout);
  auto sit = shapes().find(skey);
  if (sit == shapes().end()) {
    std::string source, why;
    if (project_source(sd, count_only, out_phys, nout, &source, &why) != MSC_OK) return ctx->fail(MSC_ERR_ARG, "jit: " + why);
    const Layout lay = stage_layout(sd);
    Kernel* k = nullptr;
    MSC_TRY(load_kernel(ctx, source, "msc_jit_scan", 4 * 8 * 8 + static_cast<size_t>(NW) * 2 * lay.stage_bytes, &k));
    sit = shapes().emplace(skey, ShapeEntry{k, false}).first;
  }
  *out = sit->second.kernel;
  return MSC_OK;
}
}  // namespace

int jit_project_compile(msc_ctx* ctx, const msc_scan_desc* sd, bool with_count_pass, const int32_t* out_phys, int nout) {
  Kernel* k = nullptr;
  if (with_count_pass) MSC_TRY(project_kernel(ctx, sd, true, out_phys, nout, &k));
  return project_kernel(ctx, sd, false, out_phys, nout, &k);
}

int jit_project_launch(msc_ctx* ctx, const msc_scan_desc* sd, bool count_only, const int32_t* out_phys, int nout, uint32_t* tile_counts,
                       const uint64_t* tile_offsets, void* const* outs, bool timed) {
  Kernel* kp = nullptr;
  MSC_TRY(project_kernel(ctx, sd, count_only, out_phys, nout, &kp));
  Kernel& k = *kp;
  if (sd->nrows == 0) return MSC_OK;
  JitParams p;
  MSC_TRY(fill_params(ctx, sd, &p));
  p.tile_counts = tile_counts;
  p.tile_offsets = reinterpret_cast<const unsigned long long*>(tile_offsets);
  for (int c = 0; c < nout && !count_only; ++c) p.out[c] = outs[c];
  return launch(ctx, k, &p, timed);
}

}  // namespace mscan
