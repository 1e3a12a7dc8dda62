/* cfg_port.c -- CPU restatement of the reference's path for BASELINE.json configs 4 and 5.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: nothing under minispark_b200/ links or executes this.  bench.py and tests/ run it
 * as the full-size checker of the CUDA engine's `extra.highcard` / `extra.join` results.  It follows the reference's
 * PythonExecutionEngine (the results oracle: f64 arithmetic), stage by stage:
 *
 *   highcard  SELECT l_orderkey, SUM(l_quantity), AVG(l_extendedprice) FROM lineitem GROUP BY l_orderkey
 *     one job per row-block                                   src/mini_spark/plan.py:90-93
 *     per-block pre-aggregation keyed by the group value      src/mini_spark/tasks.py:270-310 (before_shuffle=True)
 *     shuffle on the key, final aggregate re-SUMs partials    src/mini_spark/plan.py:190-199, tasks.py:347-375
 *     AVG = SUM / COUNT projected after the final aggregate   src/mini_spark/plan.py:200-203, sql.py:436-446
 *   join      orders o JOIN lineitem l ON o.o_orderkey = l.l_orderkey
 *             WHERE o.o_orderdate BETWEEN lo AND hi AND l.l_shipmode LIKE '%AIR%'
 *             GROUP BY o.o_orderpriority: COUNT(), SUM(l.l_extendedprice)
 *     build key -> left rows over the whole left side, stream right rows, emit per match   tasks.py:201-240
 *     the filters run on the joined rows (FilterTask above the join)                       tasks.py:167-177
 *     LIKE = anchored regex with % -> .*  (here: the pattern is %<literal>%, a substring test)   sql.py:178-194
 *     BETWEEN is inclusive on both ends (parser.py desugars it to >= and <=)
 *
 * usage: cfg_port highcard <lineitem.bin> <out.bin>
 *          out.bin = u64 n, then n x i64 key (ascending), n x f64 sum(l_quantity), n x f64 sum(l_extendedprice), n x i64 count
 *        cfg_port join <orders.bin> <lineitem.bin> <lo_us> <hi_us> <needle>
 *          prints one JSON object: {"pairs": joined rows before the filters, "groups": [{"key", "count", "sum"}]}
 *        cfg_port groupby <lineitem.bin> <key column>
 *          prints one JSON object: {"rows", "groups": [{"key", "count", "sum_q", "sum_p", "sum_pq", "min_p", "max_p"}]} (keys ascending by bytes)
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#define MAX_COLS 32
#define T_INT 0
#define T_STR 1
#define T_FLOAT 2
#define T_TS 3

typedef struct {
  const uint8_t* base;
  size_t size;
  int ncols;
  int types[MAX_COLS];
  char names[MAX_COLS][256];
  uint32_t nblocks;
  const uint64_t* starts;
} file_t;

typedef struct {
  uint32_t rows;
  const uint8_t* payload[MAX_COLS];
  uint64_t nbytes[MAX_COLS];
} block_t;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static void die(const char* what) {
  fprintf(stderr, "cfg_port: %s\n", what);
  exit(2);
}

/* BlockFile layout: src/mini_spark/io.py:47-60 (schema), :74-109 (blocks), :217-229 (footer) */
static void open_file(file_t* f, const char* path) {
  int fd = open(path, O_RDONLY);
  struct stat st;
  if (fd < 0 || fstat(fd, &st) != 0) die("cannot open input");
  f->size = st.st_size;
  f->base = mmap(NULL, f->size, PROT_READ, MAP_PRIVATE, fd, 0);
  if (f->base == MAP_FAILED) die("mmap failed");
  f->ncols = f->base[0];
  size_t pos = 1;
  for (int c = 0; c < f->ncols; ++c) {
    f->types[c] = f->base[pos];
    const int nl = f->base[pos + 1];
    memcpy(f->names[c], f->base + pos + 2, nl);
    f->names[c][nl] = 0;
    pos += 2 + nl;
  }
  memcpy(&f->nblocks, f->base + f->size - 4, 4);
  f->starts = (const uint64_t*)(f->base + f->size - 4 - 8ULL * f->nblocks);
}

static int col_index(const file_t* f, const char* name) {
  for (int c = 0; c < f->ncols; ++c)
    if (strcmp(f->names[c], name) == 0) return c;
  fprintf(stderr, "cfg_port: column %s missing\n", name);
  exit(2);
}

static void read_block(const file_t* f, uint32_t b, block_t* out) {
  const uint8_t* p = f->base + f->starts[b];
  memcpy(&out->rows, p, 4);
  p += 4;
  for (int c = 0; c < f->ncols; ++c) {
    memcpy(&out->nbytes[c], p, 8);
    p += 8;
    out->payload[c] = p;
    p += out->nbytes[c];
  }
}

static inline uint64_t mix64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}

/* ---- config 4 ------------------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t key;
  double sum_q, sum_p;
  int64_t count; /* 0 = empty slot */
} cell_t;

typedef struct {
  cell_t* cells;
  uint64_t cap, used;
} table_t;

static void table_init(table_t* t, uint64_t min_cap) {
  t->cap = 64;
  while (t->cap < min_cap) t->cap <<= 1;
  t->cells = calloc(t->cap, sizeof(cell_t));
  if (!t->cells) die("out of memory");
  t->used = 0;
}

static cell_t* table_find(table_t* t, int64_t key) {
  uint64_t pos = mix64((uint64_t)key) & (t->cap - 1);
  for (;;) {
    cell_t* c = &t->cells[pos];
    if (c->count == 0) {
      c->key = key;
      t->used++;
      return c;
    }
    if (c->key == key) return c;
    pos = (pos + 1) & (t->cap - 1);
  }
}

static void table_grow(table_t* t) {
  table_t n;
  table_init(&n, t->cap * 2);
  for (uint64_t i = 0; i < t->cap; ++i)
    if (t->cells[i].count) {
      cell_t* c = table_find(&n, t->cells[i].key);
      *c = t->cells[i];
    }
  free(t->cells);
  *t = n;
}

static int cmp_cell(const void* a, const void* b) {
  const int64_t x = ((const cell_t*)a)->key, y = ((const cell_t*)b)->key;
  return x < y ? -1 : x > y;
}

static int run_highcard(const char* path, const char* out_path) {
  file_t f;
  open_file(&f, path);
  const int c_key = col_index(&f, "l_orderkey"), c_q = col_index(&f, "l_quantity"), c_p = col_index(&f, "l_extendedprice");
  if (f.types[c_key] != T_INT || f.types[c_q] != T_FLOAT || f.types[c_p] != T_FLOAT) die("unexpected column types");
  const double t0 = now_s();
  table_t final;
  table_init(&final, 1 << 20);
  uint64_t rows_total = 0;
  for (uint32_t b = 0; b < f.nblocks; ++b) {
    block_t blk;
    read_block(&f, b, &blk);
    const int32_t* key = (const int32_t*)blk.payload[c_key];
    const float* q = (const float*)blk.payload[c_q];
    const float* p = (const float*)blk.payload[c_p];
    /* pre-aggregation of this block (AggregateTask before the shuffle): sums in row order, f64 */
    table_t part;
    table_init(&part, 1 << 16);
    for (uint32_t r = 0; r < blk.rows; ++r) {
      if (part.used * 2 >= part.cap) table_grow(&part);
      cell_t* c = table_find(&part, key[r]);
      c->sum_q += (double)q[r];
      c->sum_p += (double)p[r];
      c->count += 1;
    }
    /* final aggregate over the partial rows (plan.py:199): SUM of sums, SUM of counts */
    for (uint64_t i = 0; i < part.cap; ++i) {
      const cell_t* s = &part.cells[i];
      if (!s->count) continue;
      if (final.used * 2 >= final.cap) table_grow(&final);
      cell_t* c = table_find(&final, s->key);
      c->sum_q += s->sum_q;
      c->sum_p += s->sum_p;
      c->count += s->count;
    }
    free(part.cells);
    rows_total += blk.rows;
  }
  cell_t* out = malloc(sizeof(cell_t) * (final.used ? final.used : 1));
  uint64_t n = 0;
  for (uint64_t i = 0; i < final.cap; ++i)
    if (final.cells[i].count) out[n++] = final.cells[i];
  qsort(out, n, sizeof(cell_t), cmp_cell);
  FILE* fo = fopen(out_path, "wb");
  if (!fo) die("cannot open output");
  fwrite(&n, 8, 1, fo);
  for (uint64_t i = 0; i < n; ++i) fwrite(&out[i].key, 8, 1, fo);
  for (uint64_t i = 0; i < n; ++i) fwrite(&out[i].sum_q, 8, 1, fo);
  for (uint64_t i = 0; i < n; ++i) fwrite(&out[i].sum_p, 8, 1, fo);
  for (uint64_t i = 0; i < n; ++i) fwrite(&out[i].count, 8, 1, fo);
  fclose(fo);
  printf("{\"rows\": %llu, \"groups\": %llu, \"seconds\": %.6f}\n", (unsigned long long)rows_total, (unsigned long long)n, now_s() - t0);
  return 0;
}

/* ---- config 5 ------------------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t key;
  uint32_t head; /* chain of left rows with this key, newest first; 0xFFFFFFFF = empty slot */
} jslot_t;

typedef struct {
  char key[256];
  int len;
  long long count;
  double sum;
} jgroup_t;

static int run_join(const char* orders_path, const char* lineitem_path, int64_t lo_us, int64_t hi_us, const char* needle) {
  file_t fo, fl;
  open_file(&fo, orders_path);
  open_file(&fl, lineitem_path);
  const int o_key = col_index(&fo, "o_orderkey"), o_date = col_index(&fo, "o_orderdate"), o_prio = col_index(&fo, "o_orderpriority");
  const int l_key = col_index(&fl, "l_orderkey"), l_mode = col_index(&fl, "l_shipmode"), l_price = col_index(&fl, "l_extendedprice");
  const double t0 = now_s();
  /* left side: every orders row (the reference filters after the join) */
  uint64_t nleft = 0;
  for (uint32_t b = 0; b < fo.nblocks; ++b) {
    block_t blk;
    read_block(&fo, b, &blk);
    nleft += blk.rows;
  }
  int64_t* okey = malloc(8 * (nleft ? nleft : 1));
  int64_t* odate = malloc(8 * (nleft ? nleft : 1));
  const uint8_t** oprio = malloc(sizeof(uint8_t*) * (nleft ? nleft : 1));
  uint8_t* oprio_len = malloc(nleft ? nleft : 1);
  uint32_t* next = malloc(4 * (nleft ? nleft : 1));
  uint64_t at = 0;
  for (uint32_t b = 0; b < fo.nblocks; ++b) {
    block_t blk;
    read_block(&fo, b, &blk);
    const int32_t* k = (const int32_t*)blk.payload[o_key];
    const int64_t* d = (const int64_t*)blk.payload[o_date];
    const uint8_t* lens = blk.payload[o_prio];
    const uint8_t* bytes = lens + blk.rows;
    for (uint32_t r = 0; r < blk.rows; ++r, ++at) {
      okey[at] = k[r];
      odate[at] = d[r];
      oprio[at] = bytes;
      oprio_len[at] = lens[r];
      bytes += lens[r];
    }
  }
  uint64_t cap = 64;
  while (cap < 2 * nleft) cap <<= 1;
  jslot_t* slots = malloc(sizeof(jslot_t) * cap);
  for (uint64_t i = 0; i < cap; ++i) slots[i].head = 0xFFFFFFFFu;
  for (uint64_t i = 0; i < nleft; ++i) { /* dict key -> [row idx] (tasks.py:213-218) */
    uint64_t pos = mix64((uint64_t)okey[i]) & (cap - 1);
    while (slots[pos].head != 0xFFFFFFFFu && slots[pos].key != okey[i]) pos = (pos + 1) & (cap - 1);
    slots[pos].key = okey[i];
    next[i] = slots[pos].head;
    slots[pos].head = (uint32_t)i;
  }
  /* stream the right side; for each match the filters and the aggregate */
  jgroup_t groups[64];
  int ngroups = 0;
  unsigned long long pairs = 0;
  const size_t needle_len = strlen(needle);
  for (uint32_t b = 0; b < fl.nblocks; ++b) {
    block_t blk;
    read_block(&fl, b, &blk);
    const int32_t* k = (const int32_t*)blk.payload[l_key];
    const float* price = (const float*)blk.payload[l_price];
    const uint8_t* lens = blk.payload[l_mode];
    const uint8_t* bytes = lens + blk.rows;
    for (uint32_t r = 0; r < blk.rows; ++r) {
      const uint8_t* s = bytes;
      const int slen = lens[r];
      bytes += slen;
      uint64_t pos = mix64((uint64_t)(int64_t)k[r]) & (cap - 1);
      while (slots[pos].head != 0xFFFFFFFFu && slots[pos].key != k[r]) pos = (pos + 1) & (cap - 1);
      if (slots[pos].head == 0xFFFFFFFFu) continue;
      const int like = needle_len == 0 || memmem(s, slen, needle, needle_len) != NULL;
      for (uint32_t l = slots[pos].head; l != 0xFFFFFFFFu; l = next[l]) {
        ++pairs;
        if (odate[l] < lo_us || odate[l] > hi_us || !like) continue;
        int g = 0;
        for (; g < ngroups; ++g)
          if (groups[g].len == oprio_len[l] && memcmp(groups[g].key, oprio[l], oprio_len[l]) == 0) break;
        if (g == ngroups) {
          if (ngroups == 64) die("too many groups");
          memcpy(groups[g].key, oprio[l], oprio_len[l]);
          groups[g].len = oprio_len[l];
          groups[g].count = 0;
          groups[g].sum = 0.0;
          ++ngroups;
        }
        groups[g].count += 1;
        groups[g].sum += (double)price[r];
      }
    }
  }
  printf("{\"pairs\": %llu, \"seconds\": %.6f, \"groups\": [", pairs, now_s() - t0);
  for (int g = 0; g < ngroups; ++g)
    printf("%s{\"key\": \"%.*s\", \"count\": %lld, \"sum\": %.17g}", g ? ", " : "", groups[g].len, groups[g].key, groups[g].count, groups[g].sum);
  printf("]}\n");
  return 0;
}

/* ---- a GROUP BY with few groups (bench.py extra.midcard) ------------------------------------------------------------------
 * SELECT key, COUNT(), SUM(l_quantity), SUM(l_extendedprice), SUM(l_extendedprice * l_quantity), MIN(l_extendedprice),
 *        MAX(l_extendedprice) FROM lineitem GROUP BY key        -- key: a STRING, FLOAT or INTEGER column
 * Same stages as highcard above (tasks.py:270-310, 347-375): values widen from the file's f32 to f64 (io.py:91-94), every
 * accumulator is f64 / i64, groups are keyed by the value itself. */
typedef struct {
  uint8_t key[256];
  int len; /* -1 = empty slot */
  long long count;
  double sum_q, sum_p, sum_pq, min_p, max_p;
} gcell_t;

static int cmp_gcell(const void* a, const void* b) {
  const gcell_t *x = a, *y = b;
  const int n = x->len < y->len ? x->len : y->len;
  const int c = memcmp(x->key, y->key, n);
  return c ? c : x->len - y->len;
}

static int run_groupby(const char* path, const char* keyname) {
  const double t0 = now_s();
  file_t f;
  open_file(&f, path);
  const int kc = col_index(&f, keyname), qc = col_index(&f, "l_quantity"), pc = col_index(&f, "l_extendedprice");
  const int ktype = f.types[kc];
  const uint64_t cap = 1u << 16;
  gcell_t* cells = malloc(cap * sizeof(gcell_t));
  if (!cells) die("out of memory");
  for (uint64_t i = 0; i < cap; ++i) cells[i].len = -1;
  uint64_t used = 0, rows = 0;
  for (uint32_t b = 0; b < f.nblocks; ++b) {
    block_t blk;
    read_block(&f, b, &blk);
    const float* q = (const float*)blk.payload[qc];
    const float* price = (const float*)blk.payload[pc];
    const uint8_t* lens = blk.payload[kc];
    const uint8_t* bytes = lens + blk.rows;
    rows += blk.rows;
    for (uint32_t r = 0; r < blk.rows; ++r) {
      const uint8_t* k;
      int klen;
      if (ktype == T_STR) {
        k = bytes;
        klen = lens[r];
        bytes += klen;
      } else {
        klen = ktype == T_TS ? 8 : 4;
        k = blk.payload[kc] + (size_t)klen * r;
      }
      uint64_t h = 1469598103934665603ULL;
      for (int i = 0; i < klen; ++i) h = (h ^ k[i]) * 1099511628211ULL;
      uint64_t pos = mix64(h) & (cap - 1);
      while (cells[pos].len >= 0 && !(cells[pos].len == klen && memcmp(cells[pos].key, k, klen) == 0)) pos = (pos + 1) & (cap - 1);
      gcell_t* c = &cells[pos];
      if (c->len < 0) {
        if (++used > cap / 2) die("groupby: more than 32768 groups");
        memcpy(c->key, k, klen);
        c->len = klen;
        c->count = 0;
        c->sum_q = c->sum_p = c->sum_pq = 0.0;
        c->min_p = 1.0 / 0.0;
        c->max_p = -1.0 / 0.0;
      }
      const double qq = (double)q[r], pp = (double)price[r];
      c->count += 1;
      c->sum_q += qq;
      c->sum_p += pp;
      c->sum_pq += pp * qq;
      if (pp < c->min_p) c->min_p = pp;
      if (pp > c->max_p) c->max_p = pp;
    }
  }
  uint64_t n = 0;
  for (uint64_t i = 0; i < cap; ++i)
    if (cells[i].len >= 0) cells[n++] = cells[i];
  qsort(cells, n, sizeof(gcell_t), cmp_gcell);
  printf("{\"rows\": %llu, \"seconds\": %.6f, \"groups\": [", (unsigned long long)rows, now_s() - t0);
  for (uint64_t i = 0; i < n; ++i) {
    const gcell_t* c = &cells[i];
    printf("%s{\"key\": ", i ? ", " : "");
    if (ktype == T_STR) {
      printf("\"%.*s\"", c->len, (const char*)c->key);
    } else if (ktype == T_FLOAT) {
      float v;
      memcpy(&v, c->key, 4);
      printf("%.17g", (double)v);
    } else if (ktype == T_TS) {
      int64_t v;
      memcpy(&v, c->key, 8);
      printf("%lld", (long long)v);
    } else {
      int32_t v;
      memcpy(&v, c->key, 4);
      printf("%d", v);
    }
    printf(", \"count\": %lld, \"sum_q\": %.17g, \"sum_p\": %.17g, \"sum_pq\": %.17g, \"min_p\": %.17g, \"max_p\": %.17g}", c->count, c->sum_q, c->sum_p,
           c->sum_pq, c->min_p, c->max_p);
  }
  printf("]}\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 4 && strcmp(argv[1], "groupby") == 0) return run_groupby(argv[2], argv[3]);
  if (argc == 4 && strcmp(argv[1], "highcard") == 0) return run_highcard(argv[2], argv[3]);
  if (argc == 7 && strcmp(argv[1], "join") == 0) return run_join(argv[2], argv[3], atoll(argv[4]), atoll(argv[5]), argv[6]);
  fprintf(stderr, "usage: %s highcard <lineitem> <out.bin> | join <orders> <lineitem> <lo_us> <hi_us> <needle> | groupby <lineitem> <key column>\n", argv[0]);
  return 2;
}
