// shuffle.cu -- the row exchange between ranks over NVLink peer memory (one process per GPU, CUDA IPC).
//
// Replaces the reference's shuffle files: WriteToShufflePartitions.write appends every row to `hash(key) % n` bucket
// files (src/mini_spark/tasks.py:347-375, zig task_utils.zig:53-98) and LoadShuffleFilesTask streams a partition's
// files back in (tasks.py:144-150, plan.py:94-118).  Here a rank partitions its rows on the device (msc_partition,
// stable), every rank learns the whole rows[src][dst] matrix through a control block its peers write into, and a push
// kernel copies each partition-contiguous column segment straight into the receiver's buffer -- plain stores to peer
// memory, coalesced, no collective library and no staging copy.  Flags with release / acquire semantics at system scope
// order the phases, as in the fused low-cardinality exchange (jit.cu emit_finish):
//
//   begin   partition -> publish my row of the matrix to every peer -> wait for all rows -> matrix to the host
//           (the host sizes the receive buffers from it: every rank sees the same matrix, so all decide alike)
//   finish  push kernel (my segments -> the receivers' slots) -> publish "data of epoch e is there" -> wait for all
//           -> a relation that wraps my slot.  Everything in `finish` is stream-ordered; the host does not wait.
//
// A rank's row of the matrix for epoch e doubles as the permission to write into its slot: it is published after all of
// the rank's earlier device work (the kernels that read the slot's previous contents) has been enqueued before it on the
// same stream, and senders push only after they have seen every rank's row.  Two flag / matrix sets (epoch parity) are
// enough because a rank can be at most one exchange ahead of its slowest peer.
#include "common.cuh"

namespace {

constexpr int SH_MAXW = MSC_PEER_MAX_WORLD;
constexpr size_t SH_TABLE_BYTES = MSC_SHUFFLE_TABLE_BYTES;

struct ShuffleCtrl {
  unsigned long long flag_counts[2][SH_MAXW];  // [parity][src] = epoch once src's row of the matrix is in
  unsigned long long flag_data[2][SH_MAXW];    // [parity][src] = epoch once src's rows are in my slot
  unsigned long long flag_table[2][SH_MAXW];   // the small all-gather (msc_shuffle_allgather)
  unsigned long long matrix[2][SH_MAXW][SH_MAXW];  // [parity][src][dst] rows
  unsigned long long tables[2][SH_MAXW][SH_TABLE_BYTES / 8];
};

struct PeerCtrl {
  ShuffleCtrl* p[SH_MAXW];
};

struct Slot {
  void* own = nullptr;
  size_t bytes = 0;
  void* peer[SH_MAXW] = {nullptr};
};

struct Segment {
  const unsigned char* src;
  unsigned char* dst;
  unsigned long long bytes;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* addr, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* addr) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* addr) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
  return v;
}

// lanes < world wait until flag[lane] == epoch; returns true (to every lane) when a peer did not show up in ~30 s
__device__ bool wait_flags(const unsigned long long* flags, int world, unsigned long long epoch) {
  bool late = false;
  const int lane = threadIdx.x & 31;
  if (lane < world) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(flags + lane) != epoch) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 30000000000ull) {
        late = true;
        break;
      }
      __nanosleep(500);
    }
  }
  return __any_sync(0xffffffffu, late);
}

// one warp: my row of the matrix -> every rank's control block, then wait for all rows and copy the matrix out
__global__ void shuffle_counts_kernel(PeerCtrl peers, int rank, int world, unsigned long long epoch, const unsigned long long* counts,
                                      unsigned long long* matrix_out, int* err) {
  const int lane = threadIdx.x, par = static_cast<int>(epoch & 1u);
  for (int peer = 0; peer < world; ++peer)
    if (lane < world) peers.p[peer]->matrix[par][rank][lane] = counts[lane];
  __threadfence_system();
  __syncwarp();
  if (lane < world) st_release_sys(&peers.p[lane]->flag_counts[par][rank], epoch);
  if (wait_flags(peers.p[rank]->flag_counts[par], world, epoch) && lane == 0) atomicOr(err, MSC_DEVERR_PEER_TIMEOUT);
  __syncwarp();
  for (int i = lane; i < world * world; i += 32)
    matrix_out[i] = ld_relaxed_sys(&peers.p[rank]->matrix[par][i / world][i % world]);
}

// every segment (one column of the rows going to one rank) copied with the widest unit both ends are aligned for
__global__ void shuffle_push_kernel(const Segment* segs, int nsegs) {
  const uint64_t tid = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  const uint64_t nthreads = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  for (int s = 0; s < nsegs; ++s) {
    const Segment g = segs[s];
    const uint64_t a = reinterpret_cast<uint64_t>(g.src) | reinterpret_cast<uint64_t>(g.dst) | g.bytes;
    if ((a & 15) == 0) {
      const uint4* src = reinterpret_cast<const uint4*>(g.src);
      uint4* dst = reinterpret_cast<uint4*>(g.dst);
      for (uint64_t i = tid; i < g.bytes / 16; i += nthreads) dst[i] = src[i];
    } else if ((a & 7) == 0) {
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(g.src);
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(g.dst);
      for (uint64_t i = tid; i < g.bytes / 8; i += nthreads) dst[i] = src[i];
    } else if ((a & 3) == 0) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(g.src);
      uint32_t* dst = reinterpret_cast<uint32_t*>(g.dst);
      for (uint64_t i = tid; i < g.bytes / 4; i += nthreads) dst[i] = src[i];
    } else {
      for (uint64_t i = tid; i < g.bytes; i += nthreads) g.dst[i] = g.src[i];
    }
  }
}

// one warp, after the push kernel on the same stream: tell every rank my rows are in its slot, wait for everyone's
__global__ void shuffle_done_kernel(PeerCtrl peers, int rank, int world, unsigned long long epoch, int* err) {
  const int lane = threadIdx.x, par = static_cast<int>(epoch & 1u);
  __threadfence_system();
  if (lane < world) st_release_sys(&peers.p[lane]->flag_data[par][rank], epoch);
  if (wait_flags(peers.p[rank]->flag_data[par], world, epoch) && lane == 0) atomicOr(err, MSC_DEVERR_PEER_TIMEOUT);
}

// small fixed-size all-gather through the control blocks (partial aggregate tables): one CTA, no host wait
__global__ void shuffle_allgather_kernel(PeerCtrl peers, int rank, int world, unsigned long long epoch, const unsigned long long* src,
                                         int words, unsigned long long* dst, int* err) {
  const int par = static_cast<int>(epoch & 1u);
  for (int peer = 0; peer < world; ++peer) {
    unsigned long long* to = peers.p[peer]->tables[par][rank];
    for (int i = threadIdx.x; i < words; i += blockDim.x) to[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < 32) {
    if (threadIdx.x < world) st_release_sys(&peers.p[threadIdx.x]->flag_table[par][rank], epoch);
    if (wait_flags(peers.p[rank]->flag_table[par], world, epoch) && threadIdx.x == 0) atomicOr(err, MSC_DEVERR_PEER_TIMEOUT);
  }
  __syncthreads();
  for (int r = 0; r < world; ++r) {
    const unsigned long long* from = peers.p[rank]->tables[par][r];
    for (int i = threadIdx.x; i < words; i += blockDim.x) dst[static_cast<size_t>(r) * words + i] = ld_relaxed_sys(from + i);
  }
}

// byte offset of column c inside a slot that receives `total` rows of columns with the given widths: every column keeps
// the tile padding the scan kernels rely on (common.cuh MSC_ROW_PAD), and starts 256-byte aligned
size_t column_offset(const std::vector<size_t>& widths, uint64_t total, size_t c) {
  size_t off = 0;
  for (size_t i = 0; i < c; ++i) off += static_cast<size_t>(msc_round_up(total ? total : 1, MSC_ROW_PAD)) * widths[i] + 256;
  return off;
}

}  // namespace

struct msc_shuffle {
  msc_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  ShuffleCtrl* own = nullptr;
  PeerCtrl peers{};
  std::vector<Slot> slots;
  // between begin and finish
  msc_rel* staged = nullptr;   // partition-contiguous rows (owned), or the caller's relation for a broadcast
  bool staged_owned = false;
  bool broadcast = false;
  uint64_t bounds[SH_MAXW + 1] = {0};
  uint64_t matrix[SH_MAXW * SH_MAXW] = {0};
  uint64_t epoch = 0;
  unsigned long long* d_small = nullptr;  // device: counts[SH_MAXW] then matrix[SH_MAXW * SH_MAXW]
  Segment* h_segs = nullptr;              // pinned: the push kernel's segment list (rewritten only after begin's host wait)
  Segment* d_segs = nullptr;
  static constexpr size_t kMaxSegs = SH_MAXW * 64;
};

extern "C" int msc_shuffle_create(msc_ctx* ctx, int32_t rank, int32_t world, msc_shuffle** out, void* handle64) {
  if (!ctx || !out || !handle64 || world < 1 || world > SH_MAXW || rank < 0 || rank >= world)
    return ctx ? ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  msc_shuffle* sh = new msc_shuffle();
  sh->ctx = ctx;
  sh->rank = rank;
  sh->world = world;
  void* p = nullptr;
  int rc = msc_peer_alloc(ctx, sizeof(ShuffleCtrl), &p, handle64);
  if (rc == MSC_OK && cudaMalloc(reinterpret_cast<void**>(&sh->d_small), sizeof(unsigned long long) * (SH_MAXW + SH_MAXW * SH_MAXW)) != cudaSuccess)
    rc = ctx->fail(MSC_ERR_CUDA, "cudaMalloc failed");
  if (rc == MSC_OK && (cudaHostAlloc(reinterpret_cast<void**>(&sh->h_segs), sizeof(Segment) * msc_shuffle::kMaxSegs, cudaHostAllocDefault) != cudaSuccess ||
                       cudaMalloc(reinterpret_cast<void**>(&sh->d_segs), sizeof(Segment) * msc_shuffle::kMaxSegs) != cudaSuccess))
    rc = ctx->fail(MSC_ERR_CUDA, "allocation of the segment list failed");
  if (rc != MSC_OK) {
    if (p) cudaFree(p);
    if (sh->d_small) cudaFree(sh->d_small);
    if (sh->h_segs) cudaFreeHost(sh->h_segs);
    if (sh->d_segs) cudaFree(sh->d_segs);
    delete sh;
    return rc;
  }
  sh->own = static_cast<ShuffleCtrl*>(p);
  sh->peers.p[rank] = sh->own;
  *out = sh;
  return MSC_OK;
}

extern "C" int msc_shuffle_attach(msc_shuffle* sh, const void* handles) {
  if (!sh || !handles) return MSC_ERR_ARG;
  for (int r = 0; r < sh->world; ++r) {
    if (r == sh->rank) continue;
    void* p = nullptr;
    MSC_TRY(msc_peer_open(sh->ctx, static_cast<const char*>(handles) + 64 * r, &p));
    sh->peers.p[r] = static_cast<ShuffleCtrl*>(p);
  }
  return MSC_OK;
}

extern "C" int msc_shuffle_slot_detach(msc_shuffle* sh, int32_t slot) {
  if (!sh || slot < 0) return MSC_ERR_ARG;
  if (slot >= static_cast<int32_t>(sh->slots.size())) return MSC_OK;
  Slot& s = sh->slots[slot];
  for (int r = 0; r < sh->world; ++r)
    if (r != sh->rank && s.peer[r]) {
      MSC_TRY(msc_peer_close(sh->ctx, s.peer[r]));
      s.peer[r] = nullptr;
    }
  return MSC_OK;
}

extern "C" int msc_shuffle_slot_alloc(msc_shuffle* sh, int32_t slot, size_t nbytes, void* handle64) {
  if (!sh || slot < 0 || slot >= 64 || !handle64 || nbytes == 0) return sh ? sh->ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  msc_ctx* ctx = sh->ctx;
  if (slot >= static_cast<int32_t>(sh->slots.size())) sh->slots.resize(slot + 1);
  Slot& s = sh->slots[slot];
  for (int r = 0; r < sh->world; ++r)
    if (r != sh->rank && s.peer[r]) return ctx->fail(MSC_ERR_ARG, "slot is still attached to its peers (msc_shuffle_slot_detach first)");
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (s.own) MSC_CUDA(ctx, cudaFree(s.own));
  s.own = nullptr;
  s.bytes = 0;
  void* p = nullptr;
  MSC_TRY(msc_peer_alloc(ctx, nbytes, &p, handle64));
  s.own = p;
  s.bytes = nbytes;
  s.peer[sh->rank] = p;
  return MSC_OK;
}

extern "C" int msc_shuffle_slot_attach(msc_shuffle* sh, int32_t slot, const void* handles) {
  if (!sh || !handles || slot < 0 || slot >= static_cast<int32_t>(sh->slots.size())) return MSC_ERR_ARG;
  Slot& s = sh->slots[slot];
  for (int r = 0; r < sh->world; ++r) {
    if (r == sh->rank) continue;
    MSC_TRY(msc_peer_open(sh->ctx, static_cast<const char*>(handles) + 64 * r, &s.peer[r]));
  }
  return MSC_OK;
}

extern "C" int msc_shuffle_slot_bytes(msc_shuffle* sh, int32_t slot, size_t* nbytes) {
  if (!sh || !nbytes || slot < 0) return MSC_ERR_ARG;
  *nbytes = slot < static_cast<int32_t>(sh->slots.size()) ? sh->slots[slot].bytes : 0;
  return MSC_OK;
}

static void drop_staged(msc_shuffle* sh) {
  if (sh->staged && sh->staged_owned) msc_rel_free(sh->staged);
  sh->staged = nullptr;
  sh->staged_owned = false;
}

static int shuffle_begin(msc_shuffle* sh, msc_rel* rel, int32_t key_col, const int64_t* lower_bounds, uint64_t epoch, uint64_t* matrix,
                         uint64_t* need_bytes);

extern "C" int msc_shuffle_begin(msc_shuffle* sh, msc_rel* rel, int32_t key_col, uint64_t epoch, uint64_t* matrix, uint64_t* need_bytes) {
  return shuffle_begin(sh, rel, key_col, nullptr, epoch, matrix, need_bytes);
}

extern "C" int msc_shuffle_begin_range(msc_shuffle* sh, msc_rel* rel, int32_t key_col, const int64_t* lower_bounds, uint64_t epoch,
                                       uint64_t* matrix, uint64_t* need_bytes) {
  if (!lower_bounds || key_col < 0) return sh ? sh->ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  return shuffle_begin(sh, rel, key_col, lower_bounds, epoch, matrix, need_bytes);
}

static int shuffle_begin(msc_shuffle* sh, msc_rel* rel, int32_t key_col, const int64_t* lower_bounds, uint64_t epoch, uint64_t* matrix,
                         uint64_t* need_bytes) {
  if (!sh || !rel || !matrix || !need_bytes || epoch == 0) return sh ? sh->ctx->fail(MSC_ERR_ARG, "bad arguments") : MSC_ERR_ARG;
  msc_ctx* ctx = sh->ctx;
  const int W = sh->world;
  for (int r = 0; r < W; ++r)
    if (!sh->peers.p[r]) return ctx->fail(MSC_ERR_ARG, "shuffle is not attached to its peers");
  drop_staged(sh);
  uint64_t counts[SH_MAXW] = {0};
  sh->broadcast = key_col < 0;
  if (sh->broadcast) {  // every row to every rank (small partial tables: the all-gather of plan.py:190-199's final aggregate)
    sh->staged = rel;
    sh->staged_owned = false;
    for (int d = 0; d < W; ++d) counts[d] = rel->nrows;
  } else {
    msc_rel* part = nullptr;
    if (lower_bounds) MSC_TRY(msc_partition_range(ctx, rel, key_col, W, lower_bounds, counts, &part));
    else MSC_TRY(msc_partition(ctx, rel, key_col, W, counts, &part));
    sh->staged = part;
    sh->staged_owned = true;
    sh->bounds[0] = 0;
    for (int d = 0; d < W; ++d) sh->bounds[d + 1] = sh->bounds[d] + counts[d];
  }
  sh->epoch = epoch;
  unsigned long long h_counts[SH_MAXW];
  for (int d = 0; d < SH_MAXW; ++d) h_counts[d] = d < W ? counts[d] : 0;
  MSC_CUDA(ctx, cudaMemcpyAsync(sh->d_small, h_counts, sizeof(h_counts), cudaMemcpyHostToDevice, ctx->stream));
  shuffle_counts_kernel<<<1, 32, 0, ctx->stream>>>(sh->peers, sh->rank, W, epoch, sh->d_small, sh->d_small + SH_MAXW, ctx->d_err);
  ctx->stats.launches += 1;
  unsigned long long h_matrix[SH_MAXW * SH_MAXW];
  MSC_CUDA(ctx, cudaMemcpyAsync(h_matrix, sh->d_small + SH_MAXW, sizeof(unsigned long long) * W * W, cudaMemcpyDeviceToHost, ctx->stream));
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  MSC_CUDA(ctx, cudaGetLastError());
  int rc = msc_check_device_error(ctx);
  if (rc != MSC_OK) {
    drop_staged(sh);
    return rc;
  }
  std::vector<size_t> widths;
  for (auto& c : sh->staged->cols) widths.push_back(msc_phys_width(c.phys));
  for (int i = 0; i < W * W; ++i) sh->matrix[i] = matrix[i] = h_matrix[i];
  for (int d = 0; d < W; ++d) {
    uint64_t total = 0;
    for (int s = 0; s < W; ++s) total += h_matrix[s * W + d];
    need_bytes[d] = column_offset(widths, total, widths.size());
  }
  return MSC_OK;
}

extern "C" int msc_shuffle_finish(msc_shuffle* sh, int32_t slot, uint64_t epoch, msc_rel** out) {
  if (!sh || !out || !sh->staged || epoch != sh->epoch || slot < 0 || slot >= static_cast<int32_t>(sh->slots.size()))
    return sh ? sh->ctx->fail(MSC_ERR_ARG, "msc_shuffle_finish without a matching msc_shuffle_begin / slot") : MSC_ERR_ARG;
  msc_ctx* ctx = sh->ctx;
  const int W = sh->world, me = sh->rank;
  Slot& s = sh->slots[slot];
  std::vector<size_t> widths;
  for (auto& c : sh->staged->cols) widths.push_back(msc_phys_width(c.phys));
  const size_t ncols = widths.size();
  uint64_t total[SH_MAXW] = {0};
  for (int d = 0; d < W; ++d) {
    for (int src = 0; src < W; ++src) total[d] += sh->matrix[src * W + d];
    if (!s.peer[d] || column_offset(widths, total[d], ncols) > s.bytes) {
      drop_staged(sh);
      return ctx->fail(MSC_ERR_ARG, "exchange slot missing or too small for the rows it receives");
    }
  }
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
  std::vector<Segment> segs;
  for (int d = 0; d < W; ++d) {
    const uint64_t n = sh->matrix[me * W + d];
    if (n == 0) continue;
    uint64_t before = 0;  // rows of the ranks below me come first in the receiver's columns
    for (int src = 0; src < me; ++src) before += sh->matrix[src * W + d];
    const uint64_t first = sh->broadcast ? 0 : sh->bounds[d];
    for (size_t c = 0; c < ncols; ++c) {
      Segment g;
      g.src = static_cast<const unsigned char*>(sh->staged->cols[c].data) + first * widths[c];
      g.dst = static_cast<unsigned char*>(s.peer[d]) + column_offset(widths, total[d], c) + before * widths[c];
      g.bytes = n * widths[c];
      segs.push_back(g);
    }
  }
  // my own columns' padding must be defined: the scans bulk-copy whole tiles past the last row (rows there are masked)
  for (size_t c = 0; c < ncols; ++c) {
    const size_t lo = column_offset(widths, total[me], c) + total[me] * widths[c];
    const size_t hi = c + 1 < ncols ? column_offset(widths, total[me], c + 1) : column_offset(widths, total[me], ncols);
    MSC_CUDA(ctx, cudaMemsetAsync(static_cast<char*>(s.own) + lo, 0, hi - lo, ctx->stream));
  }
  if (segs.size() > msc_shuffle::kMaxSegs) {
    drop_staged(sh);
    return ctx->fail(MSC_ERR_ARG, "too many columns in one exchange");
  }
  if (!segs.empty()) {
    memcpy(sh->h_segs, segs.data(), sizeof(Segment) * segs.size());
    MSC_CUDA(ctx, cudaMemcpyAsync(sh->d_segs, sh->h_segs, sizeof(Segment) * segs.size(), cudaMemcpyHostToDevice, ctx->stream));
    shuffle_push_kernel<<<ctx->sm_count * 4, 512, 0, ctx->stream>>>(sh->d_segs, static_cast<int>(segs.size()));
    ctx->stats.launches += 1;
  }
  shuffle_done_kernel<<<1, 32, 0, ctx->stream>>>(sh->peers, me, W, epoch, ctx->d_err);
  ctx->stats.launches += 1;
  MSC_CUDA(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
  MSC_CUDA(ctx, cudaGetLastError());
  msc_rel* rel = new msc_rel();
  rel->ctx = ctx;
  rel->nrows = total[me];
  for (size_t c = 0; c < ncols; ++c) {
    msc_col col;
    col.data = static_cast<char*>(s.own) + column_offset(widths, total[me], c);
    col.phys = sh->staged->cols[c].phys;
    col.owned = false;
    rel->cols.push_back(col);
  }
  drop_staged(sh);  // (stream-ordered free: the push kernel ahead of it still reads the rows)
  *out = rel;
  return MSC_OK;
}

extern "C" int msc_shuffle_wait(msc_shuffle* sh, double* ms) {
  if (!sh) return MSC_ERR_ARG;
  msc_ctx* ctx = sh->ctx;
  MSC_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float f = 0;
  if (cudaEventElapsedTime(&f, ctx->ev_a, ctx->ev_b) == cudaSuccess) {
    ctx->stats.last_kernel_ms = f;
    if (ms) *ms = f;
  }
  return msc_check_device_error(ctx);
}

extern "C" int msc_shuffle_allgather(msc_shuffle* sh, const void* src_dev, size_t nbytes, uint64_t epoch, void* dst_dev) {
  if (!sh || !src_dev || !dst_dev || epoch == 0 || nbytes == 0 || nbytes % 8 != 0 || nbytes > SH_TABLE_BYTES)
    return sh ? sh->ctx->fail(MSC_ERR_ARG, "bad arguments (table larger than MSC_SHUFFLE_TABLE_BYTES?)") : MSC_ERR_ARG;
  msc_ctx* ctx = sh->ctx;
  for (int r = 0; r < sh->world; ++r)
    if (!sh->peers.p[r]) return ctx->fail(MSC_ERR_ARG, "shuffle is not attached to its peers");
  shuffle_allgather_kernel<<<1, 256, 0, ctx->stream>>>(sh->peers, sh->rank, sh->world, epoch, static_cast<const unsigned long long*>(src_dev),
                                                       static_cast<int>(nbytes / 8), static_cast<unsigned long long*>(dst_dev), ctx->d_err);
  ctx->stats.launches += 1;
  MSC_CUDA(ctx, cudaGetLastError());
  return MSC_OK;
}

extern "C" void msc_shuffle_free(msc_shuffle* sh) {
  if (!sh) return;
  msc_ctx* ctx = sh->ctx;
  cudaStreamSynchronize(ctx->stream);
  drop_staged(sh);
  for (size_t i = 0; i < sh->slots.size(); ++i) {
    msc_shuffle_slot_detach(sh, static_cast<int32_t>(i));
    if (sh->slots[i].own) cudaFree(sh->slots[i].own);
  }
  for (int r = 0; r < sh->world; ++r)
    if (r != sh->rank && sh->peers.p[r]) cudaIpcCloseMemHandle(sh->peers.p[r]);
  if (sh->own) cudaFree(sh->own);
  if (sh->d_small) cudaFree(sh->d_small);
  if (sh->h_segs) cudaFreeHost(sh->h_segs);
  if (sh->d_segs) cudaFree(sh->d_segs);
  delete sh;
}
