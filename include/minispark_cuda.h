/* minispark_cuda.h -- C-ABI of libminispark_cuda.so, the B200 (sm_100a) operator library behind
 * `CudaExecutionEngine`.
 *
 * The reference (david-westreicher/minispark) has no FFI: its native path is a code-generated Zig
 * executable driven over stdin/stdout (src/mini_spark/jobs.py:45-79, zig-src/src/job.zig:11-51,
 * src/mini_spark/execution.py:198-219).  This header therefore defines the boundary a maintainer
 * would bind instead: one entry point per *operator* of the reference's hot path, each citing the
 * reference code it replaces.  Everything is `extern "C"`, plain pointers and sizes, opaque
 * handles, `int` return (0 = ok, <0 = error; text via msc_last_error).  No exceptions cross the
 * boundary.  Host buffers are caller-owned, device buffers are library-owned unless stated.
 *
 * Thread model: one host thread per msc_ctx; a ctx owns one device, one compute stream and two
 * copy streams.  Multi-GPU = one process (rank) per GPU, each with its own ctx; the exchange step
 * between ranks moves the buffers produced by msc_partition() (torch.distributed / NCCL).
 */
#ifndef MINISPARK_CUDA_H
#define MINISPARK_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSC_ABI_VERSION 1

#if defined(__GNUC__)
#define MSC_API __attribute__((visibility("default")))
#else
#define MSC_API
#endif

/* ---- error codes ------------------------------------------------------------------------ */
#define MSC_OK 0
#define MSC_ERR_CUDA (-1)      /* a CUDA runtime call failed */
#define MSC_ERR_IO (-2)        /* file open/read/write failed or malformed BlockFile */
#define MSC_ERR_ARG (-3)       /* invalid argument / program */
#define MSC_ERR_DIV_ZERO (-4)  /* a row divided by zero (reference: ZeroDivisionError in sql.py:262-266) */
#define MSC_ERR_OVERFLOW (-5)  /* INT result does not fit i32 on write (reference: io.py:87-90) */
#define MSC_ERR_COLLISION (-6) /* unresolved 64-bit string-hash collision in a dictionary */
#define MSC_ERR_STRLEN (-7)    /* string longer than 255 bytes (BlockFile limit, io.py:42-44) */
#define MSC_ERR_PEER (-8)      /* a peer GPU did not deliver its part of a fused exchange in time */

/* ---- logical column types: the BlockFile schema ordinals (constants.py:18-23) ------------- */
#define MSC_T_INTEGER 0
#define MSC_T_STRING 1
#define MSC_T_FLOAT 2
#define MSC_T_TIMESTAMP 3

/* ---- physical (device) element types ---------------------------------------------------- */
#define MSC_P_U8 0   /* dictionary code, <=256 entries / bool flags */
#define MSC_P_U16 1  /* dictionary code, <=65536 entries */
#define MSC_P_U32 2  /* dictionary code / row index */
#define MSC_P_I32 3  /* INTEGER as stored on disk */
#define MSC_P_I64 4  /* TIMESTAMP (microseconds); INTEGER in the wide layout; computed INTEGER */
#define MSC_P_F32 5  /* FLOAT as stored on disk */
#define MSC_P_F64 6  /* FLOAT in the wide layout; computed FLOAT */

/* ---- device layouts for msc_table_load --------------------------------------------------- */
#define MSC_LAYOUT_NATIVE 0 /* disk widths: i32 / f32 / i64 / narrowest dictionary code */
#define MSC_LAYOUT_WIDE 1   /* BASELINE.json north_star widths: i64 / f64 / i64 / u32 code */

/* ---- expression program ----------------------------------------------------------------- *
 * A scan evaluates a short three-address program for every row.  It replaces the reference's
 * per-row tree interpreter (`Col.execute_row`, sql.py:127-128,192-194,262-266,371-372) and the
 * generated Zig condition/projection functions (templates/plan.zig:62-75,80-103).
 *
 * One instruction is two u32 words:
 *   w0 = op | dst_kind << 6 | tee << 9 | dst_index << 13 | fast << 20
 *   w1 = operand A | operand B << 16        operand = index | src_kind << 12 | i2f << 15
 * result = op(A, B); it goes to `dst` and, when tee != 0, also to temporary tee-1.
 * Values are 64-bit: i64 (INTEGER, TIMESTAMP, booleans, dictionary codes) or f64 (FLOAT).
 * Temporaries are per-thread slots in shared memory, so no interpreter state lives in registers
 * between instructions.  `fast` != 0 names a pre-compiled specialisation of exactly this
 * instruction shape (operand kinds / physical types / destination baked in at C++ compile time,
 * see MSC_FAST_*); the kernel dispatches it with one jump and falls back to the generic
 * fetch/compute/store path when fast == 0.  Python parses these names (minispark_b200/native.py). */
enum msc_opcode {
  MSC_OP_END = 0,
  MSC_OP_MOV = 1, /* result = A */
  MSC_OP_ADD_F = 2, MSC_OP_SUB_F = 3, MSC_OP_MUL_F = 4, MSC_OP_DIV_F = 5, MSC_OP_FLOORDIV_F = 6, MSC_OP_MOD_F = 7,
  MSC_OP_ADD_I = 8, MSC_OP_SUB_I = 9, MSC_OP_MUL_I = 10, MSC_OP_FLOORDIV_I = 11, MSC_OP_MOD_I = 12,
  MSC_OP_LT_F = 13, MSC_OP_LE_F = 14, MSC_OP_GT_F = 15, MSC_OP_GE_F = 16, MSC_OP_EQ_F = 17, MSC_OP_NE_F = 18,
  MSC_OP_LT_I = 19, MSC_OP_LE_I = 20, MSC_OP_GT_I = 21, MSC_OP_GE_I = 22, MSC_OP_EQ_I = 23, MSC_OP_NE_I = 24,
  MSC_OP_AND = 25, MSC_OP_OR = 26,
  MSC_OP_LUT8 = 27,  /* result = luts[B.index][A]  (u8 table: LIKE over dictionary codes) */
  MSC_OP_LUT32 = 28, /* result = luts[B.index][A]  (u32 table: code translation between dictionaries) */
  MSC_OP_RANK = 29,  /* no operands: fix each surviving row's stable output position (project scans) */
  MSC_OP_PROBE = 30, /* result = build-side row of key A in the join table luts[B.index] (msc_join_build), or -1: the probe
                      * half of a hash join evaluated per scanned row (BroadcastHashJoinTask, tasks.py:219-240), for build
                      * sides without duplicate keys */
  MSC_OP__COUNT = 31
};

/* operand (source) kinds */
#define MSC_SRC_NONE 0
#define MSC_SRC_TEMP 1   /* temporary `index` */
#define MSC_SRC_STAGED 2 /* staged column `index` of the current row (converted from its physical type) */
#define MSC_SRC_CONST 3  /* consts[index] */
#define MSC_SRC_GATHER 4 /* gather column (index & 63) read through staged index vector (index >> 6) */
#define MSC_SRC_LUT 5    /* luts[index]: only as operand B of LUT8 / LUT32 / PROBE */
#define MSC_SRC_GATHER_T 6 /* gather column (index & 63) read through the row index held in TEMPORARY (index >> 6): the
                            * build-side columns of a row that MSC_OP_PROBE matched (negative index: 0) */
#define MSC_SRC_I2F 8    /* flag: convert the fetched i64 to f64 (INT->FLOAT coercion, sql.py:277-290) */

/* destination kinds */
#define MSC_DST_TEMP 0   /* temporary dst_index */
#define MSC_DST_FILTER 1 /* row stays valid only if result != 0 (FilterTask, tasks.py:167-177) */
#define MSC_DST_GROUP 2  /* result is the dense group id, or the 64-bit group key in hash mode */
#define MSC_DST_AGG 3    /* fold result into accumulator dst_index of the row's group (tasks.py:293-310) */
#define MSC_DST_OUT 4    /* write result to output column dst_index at the row's output position */
#define MSC_DST_NONE 5

/* ---- fast shapes: `fast` field of w0 ------------------------------------------------------ *
 * source kinds of a specialised handler */
#define MSC_FK_F32 0   /* staged f32 column widened to f64 */
#define MSC_FK_F64 1   /* staged f64 column */
#define MSC_FK_TEMP 2  /* temporary */
#define MSC_FK_CONST 3 /* constant */
#define MSC_FK_I32F 4  /* staged i32 column converted to f64 */
#define MSC_FK_I32 5   /* staged i32 column as i64 */
#define MSC_FK_I64 6   /* staged i64 column */
#define MSC_FK_U8 7    /* staged u8 dictionary code */
#define MSC_FK_U16 8
#define MSC_FK_U32 9
#define MSC_FK__COUNT 10
/* families: id = base + formula
 *   ARITH  1 + ((opi*5 + ak)*5 + bk)*3 + dk   opi: 0 ADD_F 1 SUB_F 2 MUL_F; ak,bk in F32,F64,TEMP,CONST,I32F;
 *                                             dk: 0 -> TEMP, 1 -> AGG (SUM_F), 2 -> AGG + tee TEMP
 *   AGGMOV 226 + agg_kind*10 + fk             accumulator <- source
 *   CMP    286 + cmpi*10 + fk                 FILTER <- source <cmp> CONST; cmpi: LT LE GT GE EQ NE
 *   GROUP  346 + fk                           GROUP <- source
 *   OUT    356 + fk*2 + (out is U32)          OUT column <- source
 * The ids are dense (1..375) so the kernel's dispatch is a single jump table.  For a fast
 * instruction the library rewrites STAGED operand indices into shared-memory offsets / 16. */
#define MSC_FAST_ARITH 1
#define MSC_FAST_AGGMOV 226
#define MSC_FAST_CMP 286
#define MSC_FAST_GROUP 346
#define MSC_FAST_OUT 356

#define MSC_VM_MAX_TEMPS 8
#define MSC_VM_MAX_CODE 192  /* u32 words = 96 instructions */
#define MSC_VM_MAX_CODE2 128 /* regvm instructions */
#define MSC_VM_MAX_CONSTS 32
#define MSC_VM_MAX_STAGED 24 /* directly scanned columns (incl. index vectors); <= 32: one lane issues one column's copy */
#define MSC_VM_MAX_GATHER 16 /* columns read through an index vector */
#define MSC_VM_MAX_LUTS 8
#define MSC_VM_MAX_AGGS 16
#define MSC_VM_MAX_OUT 24

/* aggregate kinds for msc_scan_aggregate */
#define MSC_AGG_SUM_F 0
#define MSC_AGG_SUM_I 1
#define MSC_AGG_MIN_F 2
#define MSC_AGG_MAX_F 3
#define MSC_AGG_MIN_I 4
#define MSC_AGG_MAX_I 5

typedef struct msc_ctx msc_ctx;
typedef struct msc_table msc_table; /* an opened BlockFile (path or host memory image) */
typedef struct msc_rel msc_rel;     /* device-resident relation: N rows x columns */
typedef struct msc_dict msc_dict;   /* device-resident string dictionary of one STRING column */

/* A column bound to a scan: device pointer + physical type. */
typedef struct msc_colbind {
  const void* data;
  int32_t phys;
  int32_t _pad;
} msc_colbind;

/* One fused scan over `nrows` rows.  staged[] columns are read row-aligned (bulk async copies
 * into shared memory, tile by tile); gather[] columns are read through staged index vectors. */
typedef struct msc_scan_desc {
  uint64_t nrows;
  int32_t nstaged;
  int32_t ngather;
  msc_colbind staged[MSC_VM_MAX_STAGED];
  msc_colbind gather[MSC_VM_MAX_GATHER];
  int32_t ncode;   /* u32 words, two per instruction */
  int32_t nconsts;
  uint32_t code[MSC_VM_MAX_CODE];
  int64_t consts[MSC_VM_MAX_CONSTS];
  int32_t nluts;
  int32_t ntemps;  /* temporaries the program uses */
  const void* luts[MSC_VM_MAX_LUTS];
  /* Optional second encoding of the SAME query for the register-resident interpreter (dense
   * aggregate scans only; minispark_b200/csrc/gen_regvm.py, regvm_handlers.h): one u32 per
   * instruction = handler | a1 << 8 | a2 << 16, column operands given as staged slots.  When present
   * and valid the library runs it instead of `code`; ncode2 = 0 means "not available". */
  int32_t ncode2;
  int32_t count_slot2; /* accumulator that code2 increments once per surviving row (its COUNT), or -1 */
  uint32_t code2[MSC_VM_MAX_CODE2];
  /* Optional: the row count lives on the device (a relation still pending, see msc_rel_settle).  `nrows` is then
   * an upper bound used for launch geometry and allocation; msc_scan_project returns a pending relation without
   * waiting for the device. */
  const uint64_t* nrows_dev;
  /* != 0: run this scan on a kernel specialised for its program (compiled with NVRTC on first use, see MSC_DENSE_JIT);
   * 0: use such a kernel only when this process has already compiled it.  msc_scan_project reads it; the dense
   * aggregate entry points take MSC_DENSE_JIT in their flags instead. */
  int32_t want_jit;
  /* != 0: the staged columns belong to a loaded table and do not change while it is loaded; the library may then keep what it
   * derives from them (the run index of a GROUP BY key column: run count, sortedness, runs before every tile) for later scans
   * of the same column.  0 for relations that are rewritten in place (exchange buffers). */
  int32_t table_columns;
} msc_scan_desc;

typedef struct msc_stats {
  double last_kernel_ms;    /* device time of the last scan/aggregate launch sequence (CUDA events) */
  double last_ingest_ms;    /* wall time of the last msc_table_load */
  uint64_t last_ingest_bytes; /* bytes copied host->device by the last msc_table_load */
  uint64_t launches;        /* kernels launched by this ctx since creation */
  uint64_t device_bytes;    /* bytes currently allocated by this ctx */
  double last_scan_ms;      /* device time of the last fused scan kernel alone (CUDA events on its stream) */
  int32_t last_scan_grid;   /* CTAs of that launch */
  int32_t last_scan_stages; /* shared-memory ring depth */
  int32_t last_scan_smem;   /* dynamic shared memory bytes per CTA */
  int32_t last_scan_rows_per_thread;
  int32_t last_scan_kind;   /* which kernel ran the last fused scan: MSC_SCAN_KIND_* */
  int32_t last_scan_regs;   /* registers per thread of a specialised kernel (0 otherwise) */
  uint64_t jit_compiles;    /* specialised kernels compiled by this ctx since creation */
  double last_jit_compile_ms; /* host time of the last NVRTC compilation (or on-disk cache hit) */
  int32_t last_agg_runs;    /* != 0: the last msc_scan_aggregate streamed over the runs of a sorted key, so its result rows
                             * ascend by key (what a range-partitioned shuffle of partial results builds on) */
  int32_t last_hash_local_slots; /* hash aggregate: slots of the CTA-local pre-aggregation table of the last scan (0: none) */
  int32_t last_hash_attempts;    /* hash aggregate: scans it took to find a table large enough (1 unless the group count was unknown) */
  int32_t last_run_index_hit;    /* != 0: the last hash aggregate found the run index of its key column (a table column seen before) */
} msc_stats;
#define MSC_SCAN_KIND_VM 0    /* C++ three-address interpreter (scan_kernel.cuh) */
#define MSC_SCAN_KIND_REGVM 1 /* register-resident PTX interpreter (scan_regvm_impl.cuh) */
#define MSC_SCAN_KIND_JIT 2   /* query-specialised kernel compiled at run time (jit.cu) */
#define MSC_SCAN_KIND_RUNS 3  /* the C++ interpreter as a streaming aggregate over the runs of a sorted key (MODE_RUNS) */

/* ---- context ---------------------------------------------------------------------------- */
MSC_API int msc_abi_version(void);
MSC_API int msc_create(int device, msc_ctx** out);
MSC_API void msc_destroy(msc_ctx* ctx);
MSC_API const char* msc_last_error(msc_ctx* ctx);
MSC_API int msc_sync(msc_ctx* ctx);
MSC_API int msc_get_stats(msc_ctx* ctx, msc_stats* out);
/* the compute stream (a cudaStream_t) every call of this context is ordered on: lets the host language order its own
 * work -- e.g. the NCCL collective between msc_scan_dense_table(ASYNC) and msc_dense_merge_compact -- without host waits */
MSC_API int msc_stream_handle(msc_ctx* ctx, void** stream);
/* device time of a region of library calls: start records an event on the compute stream, stop records another,
 * waits for it and returns the milliseconds in between (what bench.py brackets its K steps with) */
MSC_API int msc_timer_start(msc_ctx* ctx);
MSC_API int msc_timer_stop(msc_ctx* ctx, double* ms);
/* pinned host memory for callers that stage BlockFile images themselves */
MSC_API int msc_host_alloc(msc_ctx* ctx, size_t nbytes, void** out);
MSC_API int msc_host_free(msc_ctx* ctx, void* p);
/* raw device scratch for the host side (exchange buffers) */
MSC_API int msc_dev_alloc(msc_ctx* ctx, size_t nbytes, void** out);
MSC_API int msc_dev_free(msc_ctx* ctx, void* p);
MSC_API int msc_memcpy_d2h(msc_ctx* ctx, void* host_dst, const void* dev_src, size_t nbytes);
MSC_API int msc_memcpy_h2d(msc_ctx* ctx, void* dev_dst, const void* host_src, size_t nbytes);

/* ---- BlockFile ingest: replaces BlockFile._deserialize_block (io.py:112-163), LoadTableBlockTask
 * (tasks.py:117-121) and zig ColumnData.readColumn / LoadTableBlockProducer
 * (block_file.zig:225-268, tasks.zig:212-222).  Only the requested columns are read. -------- */
MSC_API int msc_table_open(msc_ctx* ctx, const char* path, msc_table** out);
MSC_API int msc_table_open_mem(msc_ctx* ctx, const void* image, size_t nbytes, msc_table** out);
MSC_API void msc_table_close(msc_table* t);
MSC_API int msc_table_info(msc_table* t, int32_t* ncols, int32_t* nblocks, uint64_t* nrows);
MSC_API int msc_table_col_info(msc_table* t, int32_t col, int32_t* type, char* name, int32_t name_cap);
MSC_API int msc_table_block_rows(msc_table* t, int32_t block, uint32_t* rows);
/* Load `ncols` columns of `nblocks` blocks (concatenated in the order given) into a new relation.
 * STRING columns are dictionary-encoded on the device; dicts[i] is in/out: pass NULL to create a
 * dictionary, or an existing one to keep encoding against it.  Non-string slots are ignored. */
MSC_API int msc_table_load(msc_ctx* ctx, msc_table* t, const int32_t* cols, int32_t ncols,
                   const int32_t* blocks, int32_t nblocks, int32_t layout,
                   msc_dict** dicts, msc_rel** out);

/* ---- relations -------------------------------------------------------------------------- */
MSC_API int msc_rel_info(msc_rel* r, uint64_t* nrows, int32_t* ncols);
MSC_API int msc_rel_col(msc_rel* r, int32_t col, void** dev_ptr, int32_t* phys);
MSC_API void msc_rel_free(msc_rel* r);
/* Pending relations.  The *_async entry points and msc_scan_project with scan->nrows_dev enqueue their work and
 * return a relation whose row count is still on the device (msc_rel_info reports the upper bound it was allocated
 * for).  msc_rel_nrows_dev gives that device word for the next scan's nrows_dev; msc_rel_settle waits ONCE for all
 * listed relations, fixes their row counts, reports the device error word and whether an aggregate among them
 * produced a non-finite SUM.  A chain scan-aggregate -> final projection thus costs one host wait, not three. */
MSC_API int msc_rel_nrows_dev(msc_rel* r, const uint64_t** nrows_dev);
MSC_API int msc_rel_settle(msc_ctx* ctx, msc_rel* const* rels, int32_t nrels, int32_t* nonfinite);
/* pointers and physical types of all columns in one call (cols[ncols], caller-owned) */
MSC_API int msc_rel_cols(msc_rel* r, msc_colbind* cols, int32_t ncols);
/* new relation of `nrows` rows with zero-initialised, tile-padded columns of the given physical types
 * (filled by the caller, e.g. with rows received from other ranks) */
MSC_API int msc_rel_alloc(msc_ctx* ctx, uint64_t nrows, const int32_t* phys, int32_t ncols, msc_rel** out);
/* wrap caller-owned device memory (e.g. exchange receive buffers) as a relation; not freed */
MSC_API int msc_rel_wrap(msc_ctx* ctx, uint64_t nrows, const msc_colbind* cols, int32_t ncols, msc_rel** out);

/* ---- fused scan -> filter -> project -> aggregate: replaces FilterTask.execute
 * (tasks.py:167-177), ProjectTask.execute (tasks.py:79-84), AggregateTask.execute
 * (tasks.py:270-310) and the generated Zig consumers (templates/plan.zig:113-253). ----------- */
/* Dense mode (ngroups > 0): MSC_DST_GROUP receives a group id in [0, ngroups).  Hash mode
 * (ngroups == 0): it receives an arbitrary 64-bit key; `hash_capacity_hint` bounds the distinct keys
 * (0 = unknown: the table starts small and the scan is repeated with a larger one when it fills up;
 * with MSC_HASH_HINT_SOFT set the rest is a guess -- e.g. what the same scan produced last time -- that
 * sizes the first table but is not relied upon).  Rows first meet a CTA-local table in shared memory
 * (the reference's per-worker dict, tasks.py:347-375) that is folded into the global one at the end.
 * Output relation: column 0 = group id (U32) or key (I64), columns 1..naggs =
 * accumulators (I64 / F64), one row per group that received at least one row. */
#define MSC_HASH_HINT_SOFT 9223372036854775808ull
MSC_API int msc_scan_aggregate(msc_ctx* ctx, const msc_scan_desc* scan, int32_t ngroups,
                       const int32_t* agg_kinds, int32_t naggs, uint64_t hash_capacity_hint,
                       msc_rel** out);
/* The dense mode in pieces, for callers that put an exchange between the scan and the result (multi-GPU: every
 * rank scans its row-blocks into a table, the tables are all-gathered over NVLink, merged, compacted -- the
 * reference's pre-aggregate -> shuffle -> final aggregate, plan.py:190-199).  A table is [ngroups][stride] 64-bit
 * cells (I64, or F64 bit patterns); msc_dense_layout gives `stride` (naggs, or naggs + 1 when the library keeps
 * its own row counter) and the slot that tells whether a group received rows. */
MSC_API int msc_dense_layout(msc_ctx* ctx, const msc_scan_desc* scan, const int32_t* agg_kinds, int32_t naggs,
                     int32_t* stride, int32_t* count_slot);
/* identities + fused scan into `table` (device, ngroups * stride cells).  flags: MSC_DENSE_EXACT = never use the
 * register-reduction kernels (which cannot be trusted when a SUM comes out non-finite; without this flag the call
 * checks and reruns by itself), MSC_DENSE_ASYNC = only enqueue on the compute stream: no check, no host wait --
 * the caller looks at `nonfinite` of msc_dense_merge_compact and repeats the pass with MSC_DENSE_EXACT if set. */
#define MSC_DENSE_EXACT 1
#define MSC_DENSE_ASYNC 2
/* MSC_DENSE_JIT = run the scan on a kernel specialised for this query (jit.cu): the three-address program is turned
 * into straight-line CUDA C++ and compiled for sm_100a with NVRTC on first use (0.1-0.3 s; cached in the process and,
 * with MSC_JIT_CACHE=<dir>, on disk), the way the reference's ThreadEngine compiles one Zig program per query
 * (execution.py:139-160).  Without the flag a scan uses an already compiled kernel when there is one and the
 * interpreters otherwise.  Exact arithmetic (no masked reduction), so MSC_DENSE_EXACT is implied.  Scans the generator
 * cannot express (groups x accumulators > 32) ignore the flag. */
#define MSC_DENSE_JIT 4
MSC_API int msc_scan_dense_table(msc_ctx* ctx, const msc_scan_desc* scan, int32_t ngroups, const int32_t* agg_kinds,
                         int32_t naggs, void* table, int32_t flags);
/* fold `world` tables of [gmax][stride] cells (device, rank-major) into out_table[ngroups_out][stride];
 * perm[r * gmax + g] (host) = output group of rank r's group g, or < 0.  Folds in rank order. */
MSC_API int msc_dense_merge(msc_ctx* ctx, const void* tables, int32_t world, int32_t gmax, int32_t stride,
                    const int32_t* agg_kinds, int32_t naggs, const int32_t* perm, int32_t ngroups_out, void* out_table);
/* merge (perm_dev: the permutation already on the device) + compaction in one stream-ordered sequence with a single
 * host wait at the end; scratch_table: ngroups_out * stride cells of device memory */
MSC_API int msc_dense_merge_compact(msc_ctx* ctx, const void* tables, int32_t world, int32_t gmax, int32_t stride,
                            const int32_t* agg_kinds, int32_t naggs, const int32_t* perm_dev, int32_t ngroups_out,
                            int32_t count_slot, void* scratch_table, msc_rel** out, int32_t* nonfinite);
/* the same without the host wait: returns a pending relation (msc_rel_settle) */
MSC_API int msc_dense_merge_compact_async(msc_ctx* ctx, const void* tables, int32_t world, int32_t gmax, int32_t stride,
                                  const int32_t* agg_kinds, int32_t naggs, const int32_t* perm_dev, int32_t ngroups_out,
                                  int32_t count_slot, void* scratch_table, msc_rel** out);
MSC_API int msc_dense_compact_async(msc_ctx* ctx, const void* table, int32_t ngroups, int32_t stride, const int32_t* agg_kinds,
                            int32_t naggs, int32_t count_slot, msc_rel** out);
/* The specialised kernel's CUDA C++ source for a dense aggregate scan (column pointers are not looked at) and its
 * compilation, as two device-free steps: lets a host inspect / test / pre-compile what MSC_DENSE_JIT would run.
 * msc_jit_dense_source writes at most `cap` bytes (NUL-terminated) and the full length to *len; msc_jit_compile returns
 * the sm_100a cubin the same way and NVRTC's log (or the reason no compiler is available) in `log`.  masked != 0 asks
 * for the variant MSC_DENSE_JIT runs first (SUM / COUNT through one-hot f64 masks, one fma per group and aggregate;
 * the exact variant is what MSC_DENSE_EXACT, or a non-finite sum, falls back to). */
MSC_API int msc_jit_dense_source(const msc_scan_desc* scan, int32_t ngroups, const int32_t* agg_kinds, int32_t naggs, int32_t masked, char* buf,
                         size_t cap, size_t* len);
/* ... and with the final projection fused into the kernel's last CTA (what msc_dense_fused runs) */
MSC_API int msc_jit_dense_fused_source(const msc_scan_desc* scan, int32_t ngroups, const int32_t* agg_kinds, int32_t naggs, int32_t masked,
                               const msc_scan_desc* final_scan, const int32_t* final_cols, const int32_t* out_phys, int32_t nout,
                               int32_t peer /* != 0: the cross-rank variant of msc_dense_fused_peer */, char* buf, size_t cap, size_t* len);
/* ... for the streaming aggregate over the runs of sorted key column `key_col` (a staged slot; what msc_scan_aggregate runs in
 * hash mode when it finds the key column sorted and the scan asks for a specialised kernel) */
MSC_API int msc_jit_runs_source(const msc_scan_desc* scan, const int32_t* agg_kinds, int32_t naggs, int32_t key_col, char* buf, size_t cap,
                        size_t* len);
/* the same for a filter / project scan: count_only != 0 gives the first pass (surviving rows per 256-row tile), else
 * the pass that writes the output columns at their stable positions */
MSC_API int msc_jit_project_source(const msc_scan_desc* scan, int32_t count_only, const int32_t* out_phys, int32_t nout, char* buf, size_t cap,
                           size_t* len);
MSC_API int msc_jit_compile(const char* source, void* cubin, size_t cap, size_t* len, char* log, size_t log_cap);
/* one pass of a prepared dense aggregate in one call: msc_scan_dense_table(flags | ASYNC) -> msc_dense_compact_async ->
 * msc_scan_project of `final_scan` (its staged slot s is bound to column final_cols[s] of the compacted relation, its
 * row count to that relation's device count) -> msc_rel_settle of both.  *nonfinite != 0: repeat with MSC_DENSE_EXACT. */
MSC_API int msc_dense_chain(msc_ctx* ctx, const msc_scan_desc* scan, int32_t ngroups, const int32_t* agg_kinds, int32_t naggs, void* table,
                    int32_t stride, int32_t count_slot, int32_t flags, msc_scan_desc* final_scan, const int32_t* final_cols,
                    const int32_t* out_phys, int32_t nout, msc_rel** raw, msc_rel** final_rel, int32_t* nonfinite);
/* the same pass with no launch after the scan: the specialised kernel's last CTA compacts the groups and evaluates
 * the final projection.  Needs MSC_DENSE_JIT in `flags`; *final_rel comes back NULL (and MSC_OK) when this query cannot
 * be fused -- no specialised kernel, lookup tables in the final projection, more than 32 groups, no rows -- and the
 * caller uses msc_dense_chain instead. */
MSC_API int msc_dense_fused(msc_ctx* ctx, const msc_scan_desc* scan, int32_t ngroups, const int32_t* agg_kinds, int32_t naggs, void* table,
                    int32_t flags, const msc_scan_desc* final_scan, const int32_t* final_cols, const int32_t* out_phys, int32_t nout,
                    msc_rel** final_rel, int32_t* nonfinite);
/* ---- fused scan + cross-GPU merge over NVLink peer memory -----------------------------------------------------------
 * One process per GPU.  Every rank owns a small "mailbox" in device memory that its peers can write (CUDA IPC):
 * msc_peer_alloc creates it and returns the 64-byte handle to send to the other ranks (the host language moves the
 * handles, e.g. torch.distributed.all_gather_object), msc_peer_open maps a peer's mailbox.  Mailbox layout for a world
 * of W ranks and tables of C = gmax * stride cells: u64 flag[2][W], then u64 slot[2][W][C]
 * (MSC_PEER_MAILBOX_BYTES).  It replaces the shuffle files between the pre-aggregate and the final aggregate
 * (plan.py:94-118, 190-199) for low-cardinality GROUP BY.
 *
 * msc_dense_fused_peer is msc_dense_fused across ranks, still ONE kernel per rank and pass: the scan's last CTA stores
 * this rank's table into slot[epoch & 1][rank] of every mailbox (peer stores over NVLink), publishes flag = epoch,
 * waits for the W flags of its own mailbox, folds the W tables in rank order through `inv` (inv[r * 32 + G] = rank
 * r's local group of global group G, or -1) and evaluates the final projection over the `nglobal` <= 32 merged groups.
 * Every rank ends with the same bits.  `epoch` must grow by one per pass on all ranks (two slot sets: a rank can be at
 * most one pass ahead of its slowest peer).  compile_only != 0 compiles the kernel and returns (all ranks should agree
 * that it compiled before the first real pass; a rank that never launches makes its peers time out after ~10 s with
 * MSC_ERR_PEER).  *final_rel == NULL with MSC_OK: this query cannot be fused (as for msc_dense_fused). */
#define MSC_PEER_MAX_WORLD 8
#define MSC_PEER_MAILBOX_BYTES(world, cells) (sizeof(uint64_t) * (2 * (size_t)(world) + 2 * (size_t)(world) * (size_t)(cells)))
MSC_API int msc_peer_alloc(msc_ctx* ctx, size_t nbytes, void** dev_ptr, void* handle64);
MSC_API int msc_peer_open(msc_ctx* ctx, const void* handle64, void** peer_ptr);
MSC_API int msc_peer_close(msc_ctx* ctx, void* peer_ptr);
MSC_API int msc_peer_free(msc_ctx* ctx, void* dev_ptr);
typedef struct msc_peer_spec {
  void* mailbox[MSC_PEER_MAX_WORLD]; /* [rank] = this rank's own mailbox, the others as mapped by msc_peer_open */
  const int32_t* inv;                /* device, [world][32] */
  uint64_t epoch;
  int32_t rank, world;
  int32_t nlocal;  /* groups of this rank's table */
  int32_t gmax;    /* groups a slot has room for (max over ranks) */
  int32_t nglobal; /* merged groups (<= 32) */
  int32_t compile_only;
} msc_peer_spec;
MSC_API int msc_dense_fused_peer(msc_ctx* ctx, const msc_scan_desc* scan, const int32_t* agg_kinds, int32_t naggs, void* table, int32_t flags,
                         const msc_scan_desc* final_scan, const int32_t* final_cols, const int32_t* out_phys, int32_t nout,
                         const msc_peer_spec* peer, msc_rel** final_rel, int32_t* nonfinite);
/* ---- prepared dense aggregate ---------------------------------------------------------------------------------------
 * msc_dense_fused / msc_dense_fused_peer with everything that does not change between passes done once: programs
 * validated and copied, accumulator table and result relations allocated, kernel compiled.  A pass is then ONE kernel
 * launch (the kernel's finish leaves the table's identities behind for the next pass), one 24-byte read and one host
 * wait.  *out == NULL with MSC_OK: this query cannot be fused (see msc_dense_fused).  The result relation handed out by
 * msc_prepared_run / msc_prepared_wait belongs to the prepared object and is overwritten MSC_PREPARED_RING passes later;
 * `peer` may be NULL (one GPU), `epoch` is ignored then. */
typedef struct msc_prepared msc_prepared;
MSC_API int msc_prepared_create(msc_ctx* ctx, const msc_scan_desc* scan, int32_t ngroups, const int32_t* agg_kinds, int32_t naggs,
                        const msc_scan_desc* final_scan, const int32_t* final_cols, const int32_t* out_phys, int32_t nout,
                        const msc_peer_spec* peer, msc_prepared** out);
MSC_API int msc_prepared_run(msc_prepared* p, int32_t flags, uint64_t epoch, msc_rel** result, uint64_t* nrows, int32_t* nonfinite);
/* The same in two halves, so that the host can run ahead of the device: msc_prepared_enqueue launches a pass (kernel +
 * the 24-byte read of its row count into pinned memory) and returns; msc_prepared_wait collects the OLDEST pass still in
 * flight.  At most MSC_PREPARED_RING passes may be in flight; a collected result stays valid until that many later passes
 * have been enqueued.  Back-to-back passes keep the device busy across the host's per-pass work (launch latency, the
 * wait, the caller's own code) -- and with several ranks keep every rank's next kernel queued behind its current one, so
 * that the in-kernel exchange waits for the slowest GPU, not for the slowest host. */
#define MSC_PREPARED_RING 4
MSC_API int msc_prepared_enqueue(msc_prepared* p, int32_t flags, uint64_t epoch);
MSC_API int msc_prepared_wait(msc_prepared* p, msc_rel** result, uint64_t* nrows, int32_t* nonfinite);
MSC_API void msc_prepared_free(msc_prepared* p);
/* table -> relation: group id (U32) + the first naggs accumulators of every group whose count_slot is non-zero */
MSC_API int msc_dense_compact(msc_ctx* ctx, const void* table, int32_t ngroups, int32_t stride, const int32_t* agg_kinds,
                      int32_t naggs, int32_t count_slot, msc_rel** out);
/* Filter + project with stable compaction (output keeps input order, tasks.py:177).  Output
 * column i is written by the instructions whose destination is MSC_DST_OUT i; out_phys[i] in {I64,F64,U32}. */
MSC_API int msc_scan_project(msc_ctx* ctx, const msc_scan_desc* scan, const int32_t* out_phys,
                     int32_t nout, msc_rel** out);

/* ---- string dictionaries (STRING = u8-length-prefixed bytes on disk, io.py:100-104) ------- */
MSC_API int msc_dict_create(msc_ctx* ctx, msc_dict** out);
MSC_API void msc_dict_free(msc_dict* d);
MSC_API int msc_dict_size(msc_dict* d, uint32_t* nentries, uint64_t* nbytes);
/* code of a host string; insert=1 adds it when missing, else *code = -1 when missing */
MSC_API int msc_dict_lookup(msc_ctx* ctx, msc_dict* d, const char* s, size_t len, int32_t insert, int64_t* code);
/* u8 LUT over the dictionary entries: 1 where the entry matches the LIKE pattern
 * (LikeColumn, sql.py:178-179,192-194; zig-regex call at templates/plan.zig:66-68) */
MSC_API int msc_dict_like(msc_ctx* ctx, msc_dict* d, const char* pattern, size_t len, void** lut_dev);
/* u32 LUT translating codes of `from` into codes of `to` (insert=1 extends `to`;
 * otherwise missing entries map to 0xFFFFFFFF) */
MSC_API int msc_dict_translate(msc_ctx* ctx, msc_dict* from, msc_dict* to, int32_t insert, void** lut_dev);
/* copy the dictionary to the host: lens[nentries] (u32) and concatenated bytes */
MSC_API int msc_dict_export(msc_ctx* ctx, msc_dict* d, uint32_t* lens, uint8_t* bytes);
/* fill an EMPTY dictionary from the host in one go: entry i (code i) is the next lens[i] bytes of `bytes`; the
 * entries must be distinct.  Codes follow the given order, so ranks that load the same list agree on every code
 * (multi-GPU joins and merges unify string columns this way; the reference compares the strings themselves,
 * tasks.py:201-240). */
MSC_API int msc_dict_load(msc_ctx* ctx, msc_dict* d, const uint32_t* lens, const uint8_t* bytes, uint32_t nentries);
/* STRING '+' STRING (sql.py:331-333, zig concatStrings utils.zig:118-131): per-row concatenation
 * of `nparts` parts; part i is a code column (codes[i] != NULL, with its dictionary) or a literal.
 * Result: new U32 code column (1-column relation) over dictionary `out_dict`. */
typedef struct msc_concat_part {
  msc_colbind codes;      /* data == NULL -> literal */
  msc_dict* dict;
  const char* literal;
  uint64_t literal_len;
} msc_concat_part;
MSC_API int msc_str_concat(msc_ctx* ctx, const msc_concat_part* parts, int32_t nparts, uint64_t nrows,
                   msc_dict* out_dict, msc_rel** out);

/* ---- hash join: replaces BroadcastHashJoinTask.generate_chunks (tasks.py:201-240) and zig
 * JoinProducer (tasks.zig:21-196).  Keys are 64-bit (INTEGER/TIMESTAMP value, FLOAT bits or a
 * code in a shared dictionary).  Output: 2-column relation of U32 row indices (left, right),
 * one row per matching pair, right-row major like the reference. ---------------------------- */
MSC_API int msc_hash_join(msc_ctx* ctx, const int64_t* left_keys, uint64_t nleft,
                  const int64_t* right_keys, uint64_t nright, msc_rel** out_pairs);

/* The build half alone: key -> build row, for scans that carry the probe half in their row program (MSC_OP_PROBE with the
 * table as a LUT operand, build-side columns read through MSC_SRC_GATHER_T).  *out_table is a relation that owns the table
 * (free it with msc_rel_free; its one column's data pointer is what goes into msc_scan_desc.luts[]).  *unique == 0: some
 * key occurs more than once on the build side -- such a join must use msc_hash_join, which emits every pair. */
MSC_API int msc_join_build(msc_ctx* ctx, const int64_t* keys, uint64_t nkeys, msc_rel** out_table, int32_t* unique, int32_t* slot_bytes);
/* *slot_bytes = 8 (compact table: every key is a sign-extended 32-bit value) or 16.  A PROBE instruction may promise the
 * compact format to the kernel generator by setting MSC_PROBE_COMPACT in its LUT operand's index (the specialised kernel
 * then carries only that format's code and state); the interpreter reads the format from the table's header either way. */
#define MSC_PROBE_COMPACT 2048

/* msc_join_build without materialising the build side: ONE scan over its base rows whose program holds the side's filters
 * (then a RANK, as in msc_scan_project) and GROUP <- key; the rows that pass insert (key, their base row number) into a
 * compact table (a count pass sizes it when the scan filters).  Build-side columns are then read from the BASE columns
 * through the probe's match.  *usable == 0 (and *out_table == NULL): a key repeats or does not fit 32 bits -- build the
 * general way (materialise, msc_join_build / msc_hash_join).  *nkeys = rows that passed the filters. */
MSC_API int msc_scan_join_build(msc_ctx* ctx, const msc_scan_desc* scan, msc_rel** out_table, int32_t* usable, uint64_t* nkeys);

/* ---- shuffle partitioning: replaces WriteToShufflePartitions.write (tasks.py:347-375) and zig
 * fill_buckets (task_utils.zig:53-98).  Rows are routed by hash(key) % nparts; every column is
 * scattered into partition-contiguous order, STABLE (rows of a partition keep their input order, as the reference's
 * per-bucket appends do).  counts[nparts] (host) receives rows per partition.  Output relation has the same columns,
 * permuted.  When the rows are partition-contiguous already (one partition; a sorted key routed by range) nothing is
 * copied: the result SHARES the input's columns, and `rel` must stay alive as long as `*out` is used. -------------- */
MSC_API int msc_partition(msc_ctx* ctx, msc_rel* rel, int32_t key_col, int32_t nparts, uint64_t* counts,
                  msc_rel** out);
/* the same by key RANGE: partition p receives the rows with lower_bounds[p] <= key < lower_bounds[p + 1] (signed 64-bit
 * compare; lower_bounds[0] is ignored = -infinity; the bounds must not decrease).  A sorted input stays sorted inside
 * every partition -- what lets the final aggregate after a shuffle of sorted partial results stream over runs instead of
 * probing a hash table. */
MSC_API int msc_partition_range(msc_ctx* ctx, msc_rel* rel, int32_t key_col, int32_t nparts, const int64_t* lower_bounds,
                        uint64_t* counts, msc_rel** out);

/* ---- row exchange between ranks over NVLink peer memory: replaces the shuffle FILES between stages --
 * WriteToShufflePartitions.write (tasks.py:347-375) on the sending side, LoadShuffleFilesTask (tasks.py:144-150,
 * plan.py:94-118) on the receiving side.  One process per GPU; a rank's rows are partitioned on its GPU (msc_partition)
 * and every partition-contiguous column segment is copied by ONE push kernel straight into the receiving rank's buffer
 * (plain coalesced stores into CUDA-IPC peer memory): no collective library call, no staging copy.
 *
 * Set-up (collective; the host language moves the 64-byte handles, e.g. torch.distributed.all_gather_object):
 *   msc_shuffle_create -> handle of this rank's control block;  msc_shuffle_attach(all ranks' handles, rank-major)
 *   msc_shuffle_slot_alloc(slot, bytes) -> handle of a receive buffer;  msc_shuffle_slot_attach(slot, all handles)
 *   (to grow a slot: every rank msc_shuffle_slot_detach, a host barrier, then alloc + attach again)
 * One exchange (`epoch` = 1, 2, 3, ... the same on every rank):
 *   msc_shuffle_begin   partitions `rel` on column key_col by hash(key) % world (key_col < 0: every row goes to every
 *                       rank), publishes this rank's row of the rows[src][dst] matrix to all ranks, waits for theirs and
 *                       returns the matrix (world x world, src-major) and need_bytes[dst] = slot bytes rank dst needs.
 *                       All ranks see the same numbers, so they can size their slots without talking again.
 *   msc_shuffle_finish  pushes this rank's segments into the receivers' slot `slot`, tells them, waits (on the device)
 *                       for everybody's rows and returns a relation that WRAPS this rank's slot (columns in the order
 *                       and physical types of `rel`, rows ordered by sending rank, then input order).  Stream-ordered:
 *                       the host does not wait.  The relation is valid until the slot is used by another exchange.
 *   msc_shuffle_wait    optional: host wait + device error word (MSC_ERR_PEER when a rank never arrived) + the device
 *                       milliseconds of the last finish.
 * msc_shuffle_allgather: `nbytes` (<= MSC_SHUFFLE_TABLE_BYTES, a multiple of 8) from every rank into dst[world][nbytes]
 * through the control blocks, one small kernel, no host wait (its own `epoch` sequence 1, 2, ...): the partial tables of a
 * low-cardinality GROUP BY (plan.py:190-199) in one-shot queries. */
#define MSC_SHUFFLE_TABLE_BYTES 16384
typedef struct msc_shuffle msc_shuffle;
MSC_API int msc_shuffle_create(msc_ctx* ctx, int32_t rank, int32_t world, msc_shuffle** out, void* handle64);
MSC_API int msc_shuffle_attach(msc_shuffle* sh, const void* handles /* world x 64 bytes */);
MSC_API int msc_shuffle_slot_alloc(msc_shuffle* sh, int32_t slot, size_t nbytes, void* handle64);
MSC_API int msc_shuffle_slot_attach(msc_shuffle* sh, int32_t slot, const void* handles /* world x 64 bytes */);
MSC_API int msc_shuffle_slot_detach(msc_shuffle* sh, int32_t slot);
MSC_API int msc_shuffle_slot_bytes(msc_shuffle* sh, int32_t slot, size_t* nbytes);
MSC_API int msc_shuffle_begin(msc_shuffle* sh, msc_rel* rel, int32_t key_col, uint64_t epoch, uint64_t* matrix, uint64_t* need_bytes);
/* msc_shuffle_begin with rank r receiving the keys in [lower_bounds[r], lower_bounds[r + 1]) (msc_partition_range) */
MSC_API int msc_shuffle_begin_range(msc_shuffle* sh, msc_rel* rel, int32_t key_col, const int64_t* lower_bounds, uint64_t epoch,
                            uint64_t* matrix, uint64_t* need_bytes);
MSC_API int msc_shuffle_finish(msc_shuffle* sh, int32_t slot, uint64_t epoch, msc_rel** out);
MSC_API int msc_shuffle_wait(msc_shuffle* sh, double* ms);
MSC_API int msc_shuffle_allgather(msc_shuffle* sh, const void* src_dev, size_t nbytes, uint64_t epoch, void* dst_dev);
MSC_API void msc_shuffle_free(msc_shuffle* sh);

/* ---- results: replaces WriteToLocalFileTask.write (tasks.py:399-410) / zig BlockFile.appendData
 * (block_file.zig:413-456).  INTEGER narrows to i32 (error on overflow), FLOAT to f32. ------- */
MSC_API int msc_rel_copy_column(msc_ctx* ctx, msc_rel* r, int32_t col, void* host_dst, size_t cap_bytes);
/* Rows `rows[0..nreq)` (nreq <= 4) of a relation as raw 64-bit values, out[nreq][ncols] (integers sign / zero extended, F32
 * widened to F64 bits), and the fold of one row of partial aggregates into row `row`: column i takes values[i] by
 * kinds[i] = MSC_AGG_* (kinds[i] < 0: untouched).  Together they are the final aggregate of a shuffle whose partial results
 * ascend by key on every rank and follow each other in rank order: only the group that straddles two ranks has a partner,
 * so its row is read, sent (msc_shuffle_allgather) and folded instead of re-aggregating everything (plan.py:190-199). */
MSC_API int msc_rel_read_rows(msc_ctx* ctx, msc_rel* r, const uint64_t* rows, int32_t nreq, int64_t* out);
MSC_API int msc_rel_fold_row(msc_ctx* ctx, msc_rel* r, uint64_t row, const int64_t* values, const int32_t* kinds);
typedef struct msc_out_col {
  const char* name;
  int32_t type;      /* MSC_T_* */
  int32_t rel_col;   /* column of the relation */
  msc_dict* dict;    /* for MSC_T_STRING */
} msc_out_col;
MSC_API int msc_write_blockfile(msc_ctx* ctx, msc_rel* r, const msc_out_col* cols, int32_t ncols,
                        const char* path, uint32_t rows_per_block);

#ifdef __cplusplus
}
#endif
#endif /* MINISPARK_CUDA_H */
