// jit.h -- query-specialised dense aggregate scan (jit.cu): declarations for scan.cu / core.cu.
#pragma once
#include <string>
#include <vector>

#include "scan_kernel.cuh"

namespace mscan {

constexpr int JIT_MAX_REG_CELLS = 56;    // groups x accumulators kept in registers by the specialised kernel
constexpr int JIT_REG_CELLS_4CTAS = 32;  // up to here 128 registers per thread suffice (4 CTAs per SM), beyond: 168 (3 CTAs)
constexpr int JIT_MAX_SMEM_CELLS = 96;   // groups x accumulators of the form that keeps them in shared memory, a copy per thread (1 KB per cell and CTA)
constexpr int JIT_MIN_CTAS = 4;        // __launch_bounds__(128, 4): at most 128 registers per thread

// Optional fused finish of a dense aggregate scan: the last CTA compacts the groups and runs the final projection
// (jit.cu emit_finish).  `scan` supplies the program and constants only.
struct JitFinish {
  const msc_scan_desc* scan;
  const int32_t* cols;      // staged slot of `scan` -> column of the compacted relation (0 = group id, 1 + s = accumulator s)
  const int32_t* out_phys;  // [nout]
  int nout;
  int count_slot;
  void* const* outs;             // [nout] device columns of the final relation (ngroups rows each)
  unsigned long long* meta;      // device: {rows, non-finite flag, error word}
  uint32_t* ticket;              // device counter, zero between launches
  const msc_peer_spec* peer = nullptr;  // merge all ranks' tables over NVLink peer memory first (msc_dense_fused_peer)
  bool compile_only = false;            // generate, compile and load the kernel, do not launch it
};

// can this scan run on a specialised kernel at all (cheap checks; the generator may still refuse a program)?
bool jit_dense_supported(const msc_scan_desc* sd, int ngroups, int stride);
bool jit_dense_cells_in_smem(int ngroups, int stride);  // which of the two accumulator forms a table of this size gets
// CUDA C++ source of the specialised kernel (no device needed)
// masked: SUM_F / COUNT by fma with one-hot f64 masks (1 instruction per group and aggregate instead of 3; a non-finite
// input leaks into the other groups as NaN, so the caller must check the sums and rerun unmasked -- as for the masked
// regvm variants, gen_regvm.py); programs that do not allow it get the exact form anyway
int jit_dense_source(const msc_scan_desc* sd, int ngroups, int naggs, int stride, const int* kinds, const long long* init, bool masked,
                     std::string* source, std::string* err, const JitFinish* fin = nullptr);
// NVRTC: source -> sm_100a cubin (no device needed)
int jit_compile_source(const std::string& source, std::vector<char>* cubin, std::string* err);
// is the kernel for this scan already compiled and loaded in this process?
bool jit_dense_cached(msc_ctx* ctx, const msc_scan_desc* sd, int ngroups, int naggs, int stride, const int* kinds, const long long* init,
                      bool masked);
// compile (or take from the cache) and launch into `table` ([ngroups][stride], already holding the identities)
int jit_dense_launch(msc_ctx* ctx, const msc_scan_desc* sd, int ngroups, int naggs, int stride, const int* kinds, const long long* init,
                     unsigned long long* table, bool timed, bool* masked, const JitFinish* fin = nullptr);  // *masked in: allowed, out: used


// ---- filter / project scans (MODE_COUNT and MODE_PROJECT of scan_kernel.cuh), warp tiles of 256 rows ----------------
// count_only: write every tile's surviving row count to tile_counts[]; otherwise write the output columns, at
// tile_offsets[tile] + rank for filtered scans (stable) or at the row number for unfiltered ones
int jit_project_source(const msc_scan_desc* sd, bool count_only, const int32_t* out_phys, int nout, std::string* source, std::string* err);
bool jit_project_cached(msc_ctx* ctx, const msc_scan_desc* sd, const int32_t* out_phys, int nout);
// compile (or find) both passes' kernels without launching: lets the caller fall back to the interpreter as a whole
int jit_project_compile(msc_ctx* ctx, const msc_scan_desc* sd, bool with_count_pass, const int32_t* out_phys, int nout);
int jit_project_launch(msc_ctx* ctx, const msc_scan_desc* sd, bool count_only, const int32_t* out_phys, int nout, uint32_t* tile_counts,
                       const uint64_t* tile_offsets, void* const* outs, bool timed);


// ---- streaming aggregate over the runs of a sorted key (MODE_RUNS of scan_kernel.cuh), 256-row tiles ------------------
// outs[0] = key column of the result (i64), outs[1 + a] = accumulator column a holding its identity; tile_offsets[t] = runs
// that start before tile t
bool jit_runs_cached(msc_ctx* ctx, const msc_scan_desc* sd, int naggs, const int* kinds, int key_col);
int jit_runs_source(const msc_scan_desc* sd, int naggs, const int* kinds, int key_col, std::string* source, std::string* err);
int jit_runs_launch(msc_ctx* ctx, const msc_scan_desc* sd, int naggs, const int* kinds, int key_col, const uint64_t* tile_offsets, void* const* outs,
                    unsigned long long* carry, bool timed);

}  // namespace mscan
