/* Single-rank MPI stand-in used ONLY to build the reference (libParanumal) in this
 * container for golden-fixture generation (there is no MPI in the image).
 * Test infrastructure - never linked into the product library.
 *
 * Datatype handle = (class << 16) | size_in_bytes. With one rank every collective is a
 * memcpy honouring MPI_IN_PLACE and displacement 0; point-to-point is never reached
 * (reference ogs picks the size==1 shortcut, libs/ogs/ogsAuto.cpp:153-155).
 */
#ifndef LIBP_B200_MPISTUB_H
#define LIBP_B200_MPISTUB_H
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Datatype;
typedef int MPI_Comm;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_NULL  0
#define MPI_COMM_WORLD 1
#define MPI_IN_PLACE ((void*)-1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_MAX_PROCESSOR_NAME 256

#define MPI_CHAR          ((0<<16)|1)
#define MPI_INT           ((1<<16)|4)
#define MPI_LONG_LONG_INT ((2<<16)|8)
#define MPI_FLOAT         ((3<<16)|4)
#define MPI_DOUBLE        ((4<<16)|8)

#define MPI_MAX 1
#define MPI_MIN 2
#define MPI_SUM 3
#define MPI_PROD 4
#define MPI_LAND 5
#define MPI_LOR 6
#define MPI_LXOR 7

#define MPISTUB_SZ(t) ((size_t)((t)&0xffff))

static inline void mpistub_cpy(const void* s, void* r, int n, MPI_Datatype t) {
  if (s != MPI_IN_PLACE && s != r && r && s) memcpy(r, s, (size_t)n*MPISTUB_SZ(t));
}
static inline int mpistub_p2p(const char* what) {
  fprintf(stderr, "mpistub: %s reached with a single rank\n", what); abort(); return 1;
}

static inline int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_rank(MPI_Comm c, int* r) { (void)c; *r = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm c, int* s) { (void)c; *s = 1; return 0; }
static inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm* n) { *n = c; return 0; }
static inline int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* n) { (void)color; (void)key; *n = c; return 0; }
static inline int MPI_Comm_free(MPI_Comm* c) { *c = MPI_COMM_NULL; return 0; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return 0; }
static inline int MPI_Get_processor_name(char* name, int* len) { strcpy(name, "localhost"); *len = 9; return 0; }
static inline int MPI_Type_contiguous(int n, MPI_Datatype t, MPI_Datatype* nt) { *nt = (int)(n*MPISTUB_SZ(t)); return 0; }
static inline int MPI_Type_commit(MPI_Datatype* t) { (void)t; return 0; }
static inline int MPI_Type_free(MPI_Datatype* t) { (void)t; return 0; }

static inline int MPI_Send(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c) { (void)b;(void)n;(void)t;(void)d;(void)tag;(void)c; return mpistub_p2p("MPI_Send"); }
static inline int MPI_Recv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status* st) { (void)b;(void)n;(void)t;(void)s;(void)tag;(void)c;(void)st; return mpistub_p2p("MPI_Recv"); }
static inline int MPI_Isend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request* r) { (void)b;(void)n;(void)t;(void)d;(void)tag;(void)c;(void)r; return mpistub_p2p("MPI_Isend"); }
static inline int MPI_Irecv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request* r) { (void)b;(void)n;(void)t;(void)s;(void)tag;(void)c;(void)r; return mpistub_p2p("MPI_Irecv"); }
static inline int MPI_Wait(MPI_Request* r, MPI_Status* s) { (void)r; (void)s; return 0; }
static inline int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s) { (void)n; (void)r; (void)s; return 0; }

static inline int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void)b;(void)n;(void)t;(void)root;(void)c; return 0; }
static inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) { (void)op;(void)root;(void)c; mpistub_cpy(s, r, n, t); return 0; }
static inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op;(void)c; mpistub_cpy(s, r, n, t); return 0; }
static inline int MPI_Iallreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c, MPI_Request* q) { (void)op;(void)c; *q = 0; mpistub_cpy(s, r, n, t); return 0; }
static inline int MPI_Scan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) { (void)op;(void)c; mpistub_cpy(s, r, n, t); return 0; }
static inline int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { (void)rn;(void)rt;(void)root;(void)c; mpistub_cpy(s, r, sn, st); return 0; }
static inline int MPI_Gatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rn, const int* displs, MPI_Datatype rt, int root, MPI_Comm c) { (void)rn;(void)root;(void)c; mpistub_cpy(s, (char*)r + (size_t)displs[0]*MPISTUB_SZ(rt), sn, st); return 0; }
static inline int MPI_Scatter(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { (void)rn;(void)rt;(void)root;(void)c; if (r != MPI_IN_PLACE) mpistub_cpy(s, r, sn, st); return 0; }
static inline int MPI_Scatterv(const void* s, const int* sn, const int* displs, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c) { (void)rn;(void)rt;(void)root;(void)c; if (r != MPI_IN_PLACE) mpistub_cpy((const char*)s + (size_t)displs[0]*MPISTUB_SZ(st), r, sn[0], st); return 0; }
static inline int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) { (void)rn;(void)rt;(void)c; mpistub_cpy(s, r, sn, st); return 0; }
static inline int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rn, const int* displs, MPI_Datatype rt, MPI_Comm c) { (void)rn;(void)c; mpistub_cpy(s, (char*)r + (size_t)displs[0]*MPISTUB_SZ(rt), sn, st); return 0; }
static inline int MPI_Alltoall(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c) { (void)rn;(void)rt;(void)c; mpistub_cpy(s, r, sn, st); return 0; }
static inline int MPI_Alltoallv(const void* s, const int* sn, const int* sd, MPI_Datatype st, void* r, const int* rn, const int* rd, MPI_Datatype rt, MPI_Comm c) { (void)rn;(void)c; mpistub_cpy((const char*)s + (size_t)sd[0]*MPISTUB_SZ(st), (char*)r + (size_t)rd[0]*MPISTUB_SZ(rt), sn[0], st); return 0; }
static inline int MPI_Ialltoallv(const void* s, const int* sn, const int* sd, MPI_Datatype st, void* r, const int* rn, const int* rd, MPI_Datatype rt, MPI_Comm c, MPI_Request* q) { *q = 0; return MPI_Alltoallv(s, sn, sd, st, r, rn, rd, rt, c); }

#ifdef __cplusplus
}
#endif
#endif
