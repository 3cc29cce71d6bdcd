/* Implementation of the MPI stand-in declared in mpi.h (test infrastructure: lets the UNMODIFIED reference run on
 * P > 1 ranks in a container without MPI, so that multi-rank ogs maps / halo lists / operator results can be dumped
 * as golden fixtures).  One PROCESS per rank (started by mpirun_stub.sh with MPISTUB_RANK / MPISTUB_SIZE /
 * MPISTUB_DIR), so every rank has its own globals and its own glibc rand() stream, as under a real launcher.
 *
 * Transport: a message from world rank s to world rank d is one file  DIR/m_<s>_<d>_<seq>  (written under a
 * temporary name, then renamed), seq counting the messages of that ordered pair.  Sends are eager and never block;
 * a receive reads the files of its source in sequence order into an "unexpected" list and matches (context, tag)
 * in arrival order, which gives MPI's non-overtaking rule.  Collectives are built on that with reserved tags.
 * Reductions are evaluated on rank 0 of the communicator in ascending rank order (a defined order, bcast back).
 * Performance is irrelevant: fixtures are small.
 */
#include "mpi.h"

#include <errno.h>
#include <sched.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#define SZ(t) ((size_t)((t)&0xffff))
#define CLS(t) (((t) >> 16) & 0xff)
#define MAXCOMM 512
#define MAXREQ 65536
#define TAG_COLL 0x40000000 /* reserved tag space for collectives */

typedef struct { int used, ctx, size, rank; int* ranks; } comm_s;
typedef struct msg_s { struct msg_s* next; int ctx, tag; long long bytes; char* data; } msg_s;
typedef struct { int used; void* buf; long long bytes; int src_world, tag, ctx; } req_s;

static int g_init = 0, g_size = 1, g_rank = 0, g_ctx_counter = 1;
static char g_dir[1024];
static comm_s g_comm[MAXCOMM];
static msg_s** g_queue;               /* per source world rank: unexpected messages, arrival order */
static unsigned long long *g_sseq, *g_rseq;
static req_s g_req[MAXREQ];

static void die(const char* what) {
  fprintf(stderr, "mpistub[rank %d]: %s\n", g_rank, what);
  abort();
}

static void ensure_init(void) {
  if (g_init) return;
  g_init = 1;
  const char* s = getenv("MPISTUB_SIZE");
  const char* r = getenv("MPISTUB_RANK");
  const char* d = getenv("MPISTUB_DIR");
  g_size = s ? atoi(s) : 1;
  g_rank = r ? atoi(r) : 0;
  if (g_size < 1 || g_rank < 0 || g_rank >= g_size) die("bad MPISTUB_SIZE / MPISTUB_RANK");
  if (g_size > 1 && !d) die("MPISTUB_DIR is not set");
  snprintf(g_dir, sizeof(g_dir), "%s", d ? d : "/tmp");
  g_queue = (msg_s**)calloc((size_t)g_size, sizeof(msg_s*));
  g_sseq = (unsigned long long*)calloc((size_t)g_size, sizeof(unsigned long long));
  g_rseq = (unsigned long long*)calloc((size_t)g_size, sizeof(unsigned long long));
  comm_s* w = &g_comm[MPI_COMM_WORLD];
  w->used = 1; w->ctx = 1; w->size = g_size; w->rank = g_rank;
  w->ranks = (int*)malloc(sizeof(int) * (size_t)g_size);
  for (int i = 0; i < g_size; ++i) w->ranks[i] = i;
}

static comm_s* C(MPI_Comm c) {
  ensure_init();
  if (c <= 0 || c >= MAXCOMM || !g_comm[c].used) die("invalid communicator");
  return &g_comm[c];
}

/* ---------------------------------------------------------------- transport */
static void put(int dst_world, int ctx, int tag, const void* buf, long long bytes) {
  char tmp[1200], fin[1200];
  const unsigned long long seq = g_sseq[dst_world]++;
  snprintf(tmp, sizeof(tmp), "%s/t_%d_%d_%llu", g_dir, g_rank, dst_world, seq);
  snprintf(fin, sizeof(fin), "%s/m_%d_%d_%llu", g_dir, g_rank, dst_world, seq);
  FILE* f = fopen(tmp, "wb");
  if (!f) die("cannot create message file");
  long long hdr[3] = {ctx, tag, bytes};
  if (fwrite(hdr, sizeof(hdr), 1, f) != 1) die("short write");
  if (bytes > 0 && fwrite(buf, 1, (size_t)bytes, f) != (size_t)bytes) die("short write");
  fclose(f);
  if (rename(tmp, fin) != 0) die("rename failed");
}

static void pull(int src_world) { /* next message of src into the unexpected list (blocks) */
  char fin[1200];
  snprintf(fin, sizeof(fin), "%s/m_%d_%d_%llu", g_dir, src_world, g_rank, g_rseq[src_world]);
  FILE* f = NULL;
  const time_t t0 = time(NULL);
  long spins = 0;
  while (!(f = fopen(fin, "rb"))) {
    if (errno != ENOENT) die("cannot open message file");
    if (++spins > 2000) usleep(200); else sched_yield();
    if ((spins & 0xfff) == 0 && time(NULL) - t0 > 300) die("receive timed out (a peer died?)");
  }
  long long hdr[3];
  if (fread(hdr, sizeof(hdr), 1, f) != 1) die("short read");
  msg_s* m = (msg_s*)malloc(sizeof(msg_s));
  m->next = NULL; m->ctx = (int)hdr[0]; m->tag = (int)hdr[1]; m->bytes = hdr[2];
  m->data = (char*)malloc((size_t)(m->bytes > 0 ? m->bytes : 1));
  if (m->bytes > 0 && fread(m->data, 1, (size_t)m->bytes, f) != (size_t)m->bytes) die("short read");
  fclose(f);
  unlink(fin);
  g_rseq[src_world]++;
  msg_s** q = &g_queue[src_world];
  while (*q) q = &(*q)->next;
  *q = m;
}

static void get(int src_world, int ctx, int tag, void* buf, long long maxbytes) {
  for (;;) {
    for (msg_s** q = &g_queue[src_world]; *q; q = &(*q)->next) {
      msg_s* m = *q;
      if (m->ctx == ctx && m->tag == tag) {
        if (m->bytes > maxbytes) die("message truncated");
        if (m->bytes > 0) memcpy(buf, m->data, (size_t)m->bytes);
        *q = m->next;
        free(m->data);
        free(m);
        return;
      }
    }
    pull(src_world);
  }
}

/* ---------------------------------------------------------------- environment */
int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; ensure_init(); return 0; }
int MPI_Finalize(void) { if (g_init && g_size > 1) MPI_Barrier(MPI_COMM_WORLD); return 0; }
int MPI_Comm_rank(MPI_Comm c, int* r) { *r = C(c)->rank; return 0; }
int MPI_Comm_size(MPI_Comm c, int* s) { *s = C(c)->size; return 0; }
int MPI_Get_processor_name(char* name, int* len) { strcpy(name, "localhost"); *len = 9; return 0; }
int MPI_Type_contiguous(int n, MPI_Datatype t, MPI_Datatype* nt) { *nt = (0xff << 16) | (int)(n * SZ(t)); return 0; }
int MPI_Type_commit(MPI_Datatype* t) { (void)t; return 0; }
int MPI_Type_free(MPI_Datatype* t) { (void)t; return 0; }

static MPI_Comm new_comm(int ctx, int size, int rank, int* ranks) {
  for (int i = 2; i < MAXCOMM; ++i)
    if (!g_comm[i].used) {
      g_comm[i].used = 1; g_comm[i].ctx = ctx; g_comm[i].size = size; g_comm[i].rank = rank; g_comm[i].ranks = ranks;
      return i;
    }
  die("too many communicators");
  return 0;
}
/* a context id every member of the parent agrees on and nobody has used: max of the members' counters + 1 */
static int agree_ctx(MPI_Comm parent) {
  int mine = g_ctx_counter, mx = 0;
  MPI_Allreduce(&mine, &mx, 1, MPI_INT, MPI_MAX, parent);
  g_ctx_counter = mx + 1;
  return g_ctx_counter;
}
int MPI_Comm_dup(MPI_Comm c, MPI_Comm* n) {
  comm_s* p = C(c);
  const int ctx = agree_ctx(c);
  int* ranks = (int*)malloc(sizeof(int) * (size_t)p->size);
  memcpy(ranks, p->ranks, sizeof(int) * (size_t)p->size);
  *n = new_comm(ctx, p->size, p->rank, ranks);
  return 0;
}
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* n) {
  comm_s* p = C(c);
  const int ctx = agree_ctx(c);
  int mine[2] = {color, key};
  int* all = (int*)malloc(sizeof(int) * 2 * (size_t)p->size);
  MPI_Allgather(mine, 2, MPI_INT, all, 2, MPI_INT, c);
  int* ranks = (int*)malloc(sizeof(int) * (size_t)p->size);
  int cnt = 0;
  for (int i = 0; i < p->size; ++i)
    if (all[2 * i] == color) ranks[cnt++] = i;  /* parent ranks of my colour */
  for (int i = 1; i < cnt; ++i) {               /* stable insertion sort by key (ties: parent rank) */
    int v = ranks[i], j = i - 1;
    while (j >= 0 && all[2 * ranks[j] + 1] > all[2 * v + 1]) { ranks[j + 1] = ranks[j]; --j; }
    ranks[j + 1] = v;
  }
  int me = -1;
  for (int i = 0; i < cnt; ++i) {
    if (ranks[i] == p->rank) me = i;
    ranks[i] = p->ranks[ranks[i]];  /* to world ranks */
  }
  free(all);
  /* colours are disjoint groups sharing one context id: ranks never talk across groups on it */
  *n = new_comm(ctx, cnt, me, ranks);
  return 0;
}
int MPI_Comm_free(MPI_Comm* c) {
  if (*c > MPI_COMM_WORLD && *c < MAXCOMM && g_comm[*c].used) { free(g_comm[*c].ranks); g_comm[*c].used = 0; }
  *c = MPI_COMM_NULL;
  return 0;
}

/* ---------------------------------------------------------------- point to point */
int MPI_Send(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) {
  comm_s* p = C(c);
  if (d < 0 || d >= p->size) die("MPI_Send: bad destination");
  put(p->ranks[d], p->ctx, tag, b, (long long)n * (long long)SZ(t));
  return 0;
}
int MPI_Recv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status* st) {
  comm_s* p = C(c);
  if (s < 0 || s >= p->size) die("MPI_Recv: bad source");
  get(p->ranks[s], p->ctx, tag, b, (long long)n * (long long)SZ(t));
  if (st) { st->MPI_SOURCE = s; st->MPI_TAG = tag; st->MPI_ERROR = 0; }
  return 0;
}
int MPI_Isend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request* r) {
  MPI_Send(b, n, t, d, tag, c);  /* eager: complete on return */
  *r = MPI_REQUEST_NULL;
  return 0;
}
int MPI_Irecv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request* r) {
  comm_s* p = C(c);
  if (s < 0 || s >= p->size) die("MPI_Irecv: bad source");
  for (int i = 1; i < MAXREQ; ++i)
    if (!g_req[i].used) {
      g_req[i].used = 1; g_req[i].buf = b; g_req[i].bytes = (long long)n * (long long)SZ(t);
      g_req[i].src_world = p->ranks[s]; g_req[i].tag = tag; g_req[i].ctx = p->ctx;
      *r = i;
      return 0;
    }
  die("too many requests");
  return 1;
}
int MPI_Wait(MPI_Request* r, MPI_Status* s) {
  (void)s;
  if (*r > 0 && *r < MAXREQ && g_req[*r].used) {
    req_s* q = &g_req[*r];
    get(q->src_world, q->ctx, q->tag, q->buf, q->bytes);
    q->used = 0;
  }
  *r = MPI_REQUEST_NULL;
  return 0;
}
int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s) {
  (void)s;
  for (int i = 0; i < n; ++i) MPI_Wait(&r[i], MPI_STATUS_IGNORE);
  return 0;
}

/* ---------------------------------------------------------------- collectives */
static void cpy(const void* s, void* r, size_t bytes) {
  if (s != MPI_IN_PLACE && s != r && r && s && bytes) memmove(r, s, bytes);
}
#define RED_LOOP(T)                                                        \
  { T* a = (T*)acc; const T* b = (const T*)in;                             \
    for (int i = 0; i < n; ++i) {                                          \
      switch (op) {                                                        \
        case MPI_MAX: if (b[i] > a[i]) a[i] = b[i]; break;                 \
        case MPI_MIN: if (b[i] < a[i]) a[i] = b[i]; break;                 \
        case MPI_SUM: a[i] = a[i] + b[i]; break;                           \
        case MPI_PROD: a[i] = a[i] * b[i]; break;                          \
        case MPI_LAND: a[i] = (T)((a[i] != 0) && (b[i] != 0)); break;      \
        case MPI_LOR: a[i] = (T)((a[i] != 0) || (b[i] != 0)); break;       \
        case MPI_LXOR: a[i] = (T)((a[i] != 0) != (b[i] != 0)); break;      \
        default: die("unknown reduction op");                              \
      } } }
static void reduce_into(void* acc, const void* in, int n, MPI_Datatype t, MPI_Op op) {
  switch (CLS(t)) {
    case 0: RED_LOOP(char) break;
    case 1: RED_LOOP(int) break;
    case 2: RED_LOOP(long long) break;
    case 3: RED_LOOP(float) break;
    case 4: RED_LOOP(double) break;
    default: die("reduction on an opaque datatype");
  }
}

int MPI_Barrier(MPI_Comm c) {
  comm_s* p = C(c);
  if (p->size == 1) return 0;
  char z = 0;
  if (p->rank == 0) {
    for (int i = 1; i < p->size; ++i) get(p->ranks[i], p->ctx, TAG_COLL + 1, &z, 1);
    for (int i = 1; i < p->size; ++i) put(p->ranks[i], p->ctx, TAG_COLL + 2, &z, 1);
  } else {
    put(p->ranks[0], p->ctx, TAG_COLL + 1, &z, 1);
    get(p->ranks[0], p->ctx, TAG_COLL + 2, &z, 1);
  }
  return 0;
}
int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) {
  comm_s* p = C(c);
  const long long bytes = (long long)n * (long long)SZ(t);
  if (p->size == 1) return 0;
  if (p->rank == root) {
    for (int i = 0; i < p->size; ++i)
      if (i != root) put(p->ranks[i], p->ctx, TAG_COLL + 3, b, bytes);
  } else {
    get(p->ranks[root], p->ctx, TAG_COLL + 3, b, bytes);
  }
  return 0;
}
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
  comm_s* p = C(c);
  const size_t bytes = (size_t)n * SZ(t);
  if (p->size == 1) { cpy(s, r, bytes); return 0; }
  const void* mine = (s == MPI_IN_PLACE) ? r : s;
  if (p->rank != root) { put(p->ranks[root], p->ctx, TAG_COLL + 4, mine, (long long)bytes); return 0; }
  /* ascending rank order: acc = v0 op v1 op v2 ... */
  char* acc = (char*)malloc(bytes ? bytes : 1);
  char* in = (char*)malloc(bytes ? bytes : 1);
  for (int i = 0; i < p->size; ++i) {
    if (i == root) memcpy(in, mine, bytes);
    else get(p->ranks[i], p->ctx, TAG_COLL + 4, in, (long long)bytes);
    if (i == 0) memcpy(acc, in, bytes);
    else reduce_into(acc, in, n, t, op);
  }
  memcpy(r, acc, bytes);
  free(acc); free(in);
  return 0;
}
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  MPI_Reduce(s, r, n, t, op, 0, c);
  return MPI_Bcast(r, n, t, 0, c);
}
int MPI_Iallreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c, MPI_Request* q) {
  *q = MPI_REQUEST_NULL;
  return MPI_Allreduce(s, r, n, t, op, c);
}
int MPI_Scan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { /* inclusive */
  comm_s* p = C(c);
  const size_t bytes = (size_t)n * SZ(t);
  cpy(s, r, bytes);
  if (p->size == 1) return 0;
  if (p->rank > 0) {
    char* prev = (char*)malloc(bytes ? bytes : 1);
    char* mine = (char*)malloc(bytes ? bytes : 1);
    memcpy(mine, r, bytes);
    get(p->ranks[p->rank - 1], p->ctx, TAG_COLL + 5, prev, (long long)bytes);
    memcpy(r, prev, bytes);
    reduce_into(r, mine, n, t, op);  /* prefix(rank-1) op mine */
    free(prev); free(mine);
  }
  if (p->rank + 1 < p->size) put(p->ranks[p->rank + 1], p->ctx, TAG_COLL + 5, r, (long long)bytes);
  return 0;
}
int MPI_Gatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rn, const int* displs, MPI_Datatype rt,
                int root, MPI_Comm c) {
  comm_s* p = C(c);
  if (p->rank != root) { put(p->ranks[root], p->ctx, TAG_COLL + 6, s, (long long)sn * (long long)SZ(st)); return 0; }
  for (int i = 0; i < p->size; ++i) {
    char* dst = (char*)r + (size_t)displs[i] * SZ(rt);
    if (i == root) cpy(s, dst, (size_t)sn * SZ(st));
    else get(p->ranks[i], p->ctx, TAG_COLL + 6, dst, (long long)rn[i] * (long long)SZ(rt));
  }
  return 0;
}
int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  comm_s* p = C(c);
  if (p->rank != root) { put(p->ranks[root], p->ctx, TAG_COLL + 7, s, (long long)sn * (long long)SZ(st)); return 0; }
  for (int i = 0; i < p->size; ++i) {
    char* dst = (char*)r + (size_t)i * (size_t)rn * SZ(rt);
    if (i == root) cpy(s, dst, (size_t)sn * SZ(st));
    else get(p->ranks[i], p->ctx, TAG_COLL + 7, dst, (long long)rn * (long long)SZ(rt));
  }
  return 0;
}
int MPI_Scatter(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  comm_s* p = C(c);
  if (p->rank == root) {
    for (int i = 0; i < p->size; ++i) {
      const char* src = (const char*)s + (size_t)i * (size_t)sn * SZ(st);
      if (i == root) { if (r != MPI_IN_PLACE) cpy(src, r, (size_t)sn * SZ(st)); }
      else put(p->ranks[i], p->ctx, TAG_COLL + 8, src, (long long)sn * (long long)SZ(st));
    }
  } else {
    get(p->ranks[root], p->ctx, TAG_COLL + 8, r, (long long)rn * (long long)SZ(rt));
  }
  return 0;
}
int MPI_Scatterv(const void* s, const int* sn, const int* displs, MPI_Datatype st, void* r, int rn, MPI_Datatype rt,
                 int root, MPI_Comm c) {
  comm_s* p = C(c);
  if (p->rank == root) {
    for (int i = 0; i < p->size; ++i) {
      const char* src = (const char*)s + (size_t)displs[i] * SZ(st);
      if (i == root) { if (r != MPI_IN_PLACE) cpy(src, r, (size_t)sn[i] * SZ(st)); }
      else put(p->ranks[i], p->ctx, TAG_COLL + 9, src, (long long)sn[i] * (long long)SZ(st));
    }
  } else {
    get(p->ranks[root], p->ctx, TAG_COLL + 9, r, (long long)rn * (long long)SZ(rt));
  }
  return 0;
}
int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rn, const int* displs, MPI_Datatype rt,
                   MPI_Comm c) {
  comm_s* p = C(c);
  const void* mine = (s == MPI_IN_PLACE) ? (const void*)((char*)r + (size_t)displs[p->rank] * SZ(rt)) : s;
  const long long mybytes = (s == MPI_IN_PLACE) ? (long long)rn[p->rank] * (long long)SZ(rt) : (long long)sn * (long long)SZ(st);
  for (int i = 0; i < p->size; ++i)
    if (i != p->rank) put(p->ranks[i], p->ctx, TAG_COLL + 10, mine, mybytes);
  for (int i = 0; i < p->size; ++i) {
    char* dst = (char*)r + (size_t)displs[i] * SZ(rt);
    if (i == p->rank) cpy(mine, dst, (size_t)mybytes);
    else get(p->ranks[i], p->ctx, TAG_COLL + 10, dst, (long long)rn[i] * (long long)SZ(rt));
  }
  return 0;
}
int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) {
  comm_s* p = C(c);
  const size_t chunk = (size_t)rn * SZ(rt);
  const void* mine = (s == MPI_IN_PLACE) ? (const void*)((char*)r + (size_t)p->rank * chunk) : s;
  const long long mybytes = (s == MPI_IN_PLACE) ? (long long)chunk : (long long)sn * (long long)SZ(st);
  for (int i = 0; i < p->size; ++i)
    if (i != p->rank) put(p->ranks[i], p->ctx, TAG_COLL + 11, mine, mybytes);
  for (int i = 0; i < p->size; ++i) {
    char* dst = (char*)r + (size_t)i * chunk;
    if (i == p->rank) cpy(mine, dst, (size_t)mybytes);
    else get(p->ranks[i], p->ctx, TAG_COLL + 11, dst, (long long)chunk);
  }
  return 0;
}
int MPI_Alltoall(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) {
  comm_s* p = C(c);
  const size_t sc = (size_t)sn * SZ(st), rc = (size_t)rn * SZ(rt);
  if (s == MPI_IN_PLACE) die("MPI_Alltoall: MPI_IN_PLACE is not supported");
  for (int i = 0; i < p->size; ++i)
    if (i != p->rank) put(p->ranks[i], p->ctx, TAG_COLL + 12, (const char*)s + (size_t)i * sc, (long long)sc);
  for (int i = 0; i < p->size; ++i) {
    if (i == p->rank) cpy((const char*)s + (size_t)i * sc, (char*)r + (size_t)i * rc, sc);
    else get(p->ranks[i], p->ctx, TAG_COLL + 12, (char*)r + (size_t)i * rc, (long long)rc);
  }
  return 0;
}
int MPI_Alltoallv(const void* s, const int* sn, const int* sd, MPI_Datatype st, void* r, const int* rn, const int* rd,
                  MPI_Datatype rt, MPI_Comm c) {
  comm_s* p = C(c);
  if (s == MPI_IN_PLACE) die("MPI_Alltoallv: MPI_IN_PLACE is not supported");
  for (int i = 0; i < p->size; ++i)
    if (i != p->rank)
      put(p->ranks[i], p->ctx, TAG_COLL + 13, (const char*)s + (size_t)sd[i] * SZ(st), (long long)sn[i] * (long long)SZ(st));
  for (int i = 0; i < p->size; ++i) {
    char* dst = (char*)r + (size_t)rd[i] * SZ(rt);
    if (i == p->rank) cpy((const char*)s + (size_t)sd[i] * SZ(st), dst, (size_t)sn[i] * SZ(st));
    else get(p->ranks[i], p->ctx, TAG_COLL + 13, dst, (long long)rn[i] * (long long)SZ(rt));
  }
  return 0;
}
int MPI_Ialltoallv(const void* s, const int* sn, const int* sd, MPI_Datatype st, void* r, const int* rn, const int* rd,
                   MPI_Datatype rt, MPI_Comm c, MPI_Request* q) {
  *q = MPI_REQUEST_NULL;
  return MPI_Alltoallv(s, sn, sd, st, r, rn, rd, rt, c);
}
