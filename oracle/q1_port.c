/* q1_port.c -- CPU restatement of the reference's native path for TPC-H Q1.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: nothing under minispark_b200/ links or executes this.
 * Used by tests (full-size parity check of the CUDA engine) and by bench.py (cpu_baseline and the
 * `--impl reference` arm).  The reference's Zig ThreadEngine cannot be built here (no zig, its
 * zig-regex dependency needs the network), so this follows the steps of its generated stage
 * program, with the arithmetic of the results oracle (PythonExecutionEngine: f64):
 *   job per row-block on a worker pool            src/mini_spark/plan.py:90-93, execution.py:126-157
 *   read + decode every column of the block       zig-src/src/block_file.zig:225-268,297-306 (no pruning)
 *   condition -> materialising filter of all cols templates/plan.zig:62-75,130-147; zig-src/src/task_utils.zig:9-51
 *   hash aggregate keyed by the group string      templates/plan.zig:170-251 / src/mini_spark/tasks.py:270-310
 *   per-block partials -> final merge, AVG=SUM/COUNT  src/mini_spark/plan.py:190-203
 * Query (examples/benchmark.py:51-68): GROUP BY l_returnflag WHERE l_shipdate <= cutoff;
 * SUM(qty), SUM(price), SUM(price*(1-disc)), SUM(price*(1-disc)*(1+tax)), AVG(qty), AVG(price), AVG(disc), COUNT.
 *
 * usage: q1_port <blockfile> <threads> <max_blocks|0> <wire 0|1>
 * prints one JSON object: rows, seconds, groups[{key, sums..., count}].
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#define MAX_COLS 32
#define MAX_GROUPS 64
#define NSUM 5 /* qty, price, disc_price, charge, disc */

typedef struct {
  char key[256];
  int key_len;
  double sum[NSUM];
  long long count;
} group_t;

typedef struct {
  group_t g[MAX_GROUPS];
  int n;
} table_t;

typedef struct {
  const uint8_t* base;
  size_t size;
  int ncols;
  int types[MAX_COLS];
  char names[MAX_COLS][256];
  uint32_t nblocks;
  const uint64_t* starts;
  int c_qty, c_price, c_disc, c_tax, c_flag, c_ship;
  int64_t cutoff_us;
  int wire;
  /* work queue */
  pthread_mutex_t mu;
  uint32_t next_block, max_blocks;
  table_t merged;
  uint64_t rows;
} job_ctx;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static group_t* find_group(table_t* t, const char* key, int len) {
  for (int i = 0; i < t->n; ++i)
    if (t->g[i].key_len == len && memcmp(t->g[i].key, key, len) == 0) return &t->g[i];
  if (t->n == MAX_GROUPS) {
    fprintf(stderr, "too many groups\n");
    exit(2);
  }
  group_t* g = &t->g[t->n++];
  memset(g, 0, sizeof(*g));
  memcpy(g->key, key, len);
  g->key_len = len;
  return g;
}

/* one ScanJob: decode all columns, filter (materialising every column), aggregate */
static void run_block(job_ctx* J, uint32_t b, table_t* out, uint64_t* rows_out) {
  const uint8_t* p = J->base + J->starts[b];
  uint32_t rows;
  memcpy(&rows, p, 4);
  p += 4;
  /* decode: copy each column payload into its own typed array (LoadTableBlockProducer) */
  void* col[MAX_COLS];
  uint8_t* slen[MAX_COLS];
  uint32_t* soff[MAX_COLS];
  for (int c = 0; c < J->ncols; ++c) {
    uint64_t nbytes;
    memcpy(&nbytes, p, 8);
    p += 8;
    col[c] = malloc(nbytes ? nbytes : 1);
    memcpy(col[c], p, nbytes);
    slen[c] = NULL;
    soff[c] = NULL;
    if (J->types[c] == 1) { /* STRING: u8 lengths then bytes -> offsets */
      slen[c] = (uint8_t*)col[c];
      soff[c] = malloc(sizeof(uint32_t) * (rows + 1));
      uint32_t o = rows;
      for (uint32_t r = 0; r < rows; ++r) {
        soff[c][r] = o;
        o += slen[c][r];
      }
      soff[c][rows] = o;
    }
    p += nbytes;
  }
  /* condition */
  const int64_t* ship = (const int64_t*)col[J->c_ship];
  uint32_t* sel = malloc(sizeof(uint32_t) * (rows ? rows : 1));
  uint32_t nsel = 0;
  for (uint32_t r = 0; r < rows; ++r)
    if (ship[r] <= J->cutoff_us) sel[nsel++] = r;
  /* FilterTask: every column is materialised for the surviving rows */
  void* fcol[MAX_COLS];
  uint32_t* fsoff[MAX_COLS];
  for (int c = 0; c < J->ncols; ++c) {
    fsoff[c] = NULL;
    if (J->types[c] == 1) {
      uint64_t bytes = 0;
      for (uint32_t i = 0; i < nsel; ++i) bytes += slen[c][sel[i]];
      uint8_t* dst = malloc(bytes ? bytes : 1);
      fsoff[c] = malloc(sizeof(uint32_t) * (nsel + 1));
      uint32_t o = 0;
      for (uint32_t i = 0; i < nsel; ++i) {
        const uint32_t r = sel[i];
        fsoff[c][i] = o;
        memcpy(dst + o, (uint8_t*)col[c] + soff[c][r], slen[c][r]);
        o += slen[c][r];
      }
      fsoff[c][nsel] = o;
      fcol[c] = dst;
    } else if (J->types[c] == 3) {
      int64_t* dst = malloc(8 * (nsel ? nsel : 1));
      for (uint32_t i = 0; i < nsel; ++i) dst[i] = ((int64_t*)col[c])[sel[i]];
      fcol[c] = dst;
    } else {
      uint32_t* dst = malloc(4 * (nsel ? nsel : 1));
      for (uint32_t i = 0; i < nsel; ++i) dst[i] = ((uint32_t*)col[c])[sel[i]];
      fcol[c] = dst;
    }
  }
  /* aggregate (row order, f64 accumulators: the PythonExecutionEngine arithmetic) */
  table_t t;
  t.n = 0;
  const float* qty = (const float*)fcol[J->c_qty];
  const float* price = (const float*)fcol[J->c_price];
  const float* disc = (const float*)fcol[J->c_disc];
  const float* tax = (const float*)fcol[J->c_tax];
  const uint8_t* flag = (const uint8_t*)fcol[J->c_flag];
  group_t* last = NULL;
  for (uint32_t i = 0; i < nsel; ++i) {
    const char* key = (const char*)flag + fsoff[J->c_flag][i];
    const int len = (int)(fsoff[J->c_flag][i + 1] - fsoff[J->c_flag][i]);
    group_t* g = (last && last->key_len == len && memcmp(last->key, key, len) == 0) ? last : find_group(&t, key, len);
    last = g;
    const double q = qty[i], pr = price[i], d = disc[i], tx = tax[i];
    const double dp = pr * (1.0 - d);
    g->sum[0] += q;
    g->sum[1] += pr;
    g->sum[2] += dp;
    g->sum[3] += dp * (1.0 + tx);
    g->sum[4] += d;
    g->count += 1;
  }
  if (J->wire) /* partials cross a shuffle BlockFile: FLOAT is stored as f32 (io.py:91-94) */
    for (int i = 0; i < t.n; ++i)
      for (int s = 0; s < NSUM; ++s) t.g[i].sum[s] = (double)(float)t.g[i].sum[s];
  *out = t;
  *rows_out = rows;
  for (int c = 0; c < J->ncols; ++c) {
    free(col[c]);
    free(soff[c]);
    free(fcol[c]);
    free(fsoff[c]);
  }
  free(sel);
}

static void* worker(void* arg) {
  job_ctx* J = (job_ctx*)arg;
  for (;;) {
    pthread_mutex_lock(&J->mu);
    const uint32_t b = J->next_block++;
    pthread_mutex_unlock(&J->mu);
    if (b >= J->max_blocks) break;
    table_t t;
    uint64_t rows;
    run_block(J, b, &t, &rows);
    pthread_mutex_lock(&J->mu); /* final aggregate: merge the block's partials */
    for (int i = 0; i < t.n; ++i) {
      group_t* g = find_group(&J->merged, t.g[i].key, t.g[i].key_len);
      for (int s = 0; s < NSUM; ++s) g->sum[s] += t.g[i].sum[s];
      g->count += t.g[i].count;
    }
    J->rows += rows;
    pthread_mutex_unlock(&J->mu);
  }
  return NULL;
}

static int col_index(job_ctx* J, const char* name) {
  for (int c = 0; c < J->ncols; ++c)
    if (strcmp(J->names[c], name) == 0) return c;
  fprintf(stderr, "column %s missing\n", name);
  exit(2);
}

int main(int argc, char** argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s <blockfile> <threads> <max_blocks|0> <wire 0|1>\n", argv[0]);
    return 2;
  }
  const int threads = atoi(argv[2]);
  job_ctx J;
  memset(&J, 0, sizeof(J));
  int fd = open(argv[1], O_RDONLY);
  struct stat st;
  if (fd < 0 || fstat(fd, &st) != 0) {
    perror("open");
    return 2;
  }
  J.size = st.st_size;
  J.base = mmap(NULL, J.size, PROT_READ, MAP_PRIVATE, fd, 0);
  if (J.base == MAP_FAILED) {
    perror("mmap");
    return 2;
  }
  J.ncols = J.base[0];
  size_t pos = 1;
  for (int c = 0; c < J.ncols; ++c) {
    J.types[c] = J.base[pos];
    const int nl = J.base[pos + 1];
    memcpy(J.names[c], J.base + pos + 2, nl);
    J.names[c][nl] = 0;
    pos += 2 + nl;
  }
  memcpy(&J.nblocks, J.base + J.size - 4, 4);
  J.starts = (const uint64_t*)(J.base + J.size - 4 - 8ULL * J.nblocks);
  J.c_qty = col_index(&J, "l_quantity");
  J.c_price = col_index(&J, "l_extendedprice");
  J.c_disc = col_index(&J, "l_discount");
  J.c_tax = col_index(&J, "l_tax");
  J.c_flag = col_index(&J, "l_returnflag");
  J.c_ship = col_index(&J, "l_shipdate");
  J.cutoff_us = 912470400LL * 1000000LL; /* 1998-12-01T00:00:00 UTC */
  J.wire = atoi(argv[4]);
  J.max_blocks = atoi(argv[3]) > 0 && (uint32_t)atoi(argv[3]) < J.nblocks ? (uint32_t)atoi(argv[3]) : J.nblocks;
  pthread_mutex_init(&J.mu, NULL);
  pthread_t th[256];
  const int nt = threads < 1 ? 1 : (threads > 256 ? 256 : threads);
  const double t0 = now_s();
  for (int i = 0; i < nt; ++i) pthread_create(&th[i], NULL, worker, &J);
  for (int i = 0; i < nt; ++i) pthread_join(th[i], NULL);
  const double t1 = now_s();
  printf("{\"rows\": %llu, \"seconds\": %.6f, \"threads\": %d, \"blocks\": %u, \"groups\": [", (unsigned long long)J.rows, t1 - t0, nt, J.max_blocks);
  for (int i = 0; i < J.merged.n; ++i) {
    group_t* g = &J.merged.g[i];
    printf("%s{\"key\": \"%.*s\", \"sum_qty\": %.17g, \"sum_base_price\": %.17g, \"sum_disc_price\": %.17g, \"sum_charge\": %.17g, "
           "\"sum_disc\": %.17g, \"count\": %lld}",
           i ? ", " : "", g->key_len, g->key, g->sum[0], g->sum[1], g->sum[2], g->sum[3], g->sum[4], g->count);
  }
  printf("]}\n");
  return 0;
}
