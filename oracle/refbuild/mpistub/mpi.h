/* MPI stand-in used ONLY to build and run the reference (libParanumal) in this container for golden-fixture
 * generation (there is no MPI in the image).  Test infrastructure - never linked into the product library.
 *
 * The implementation lives in mpistub.c.  With MPISTUB_SIZE unset (or 1) every collective is a memcpy honouring
 * MPI_IN_PLACE and the displacements of rank 0.  With MPISTUB_SIZE = P > 1 the P ranks are separate PROCESSES started
 * by mpirun_stub.sh (own globals, own glibc rand() stream - exactly as under a real MPI launcher) that exchange
 * messages through files in a shared directory (MPISTUB_DIR); see mpistub.c for the protocol.  Only the ~40 entry
 * points the reference uses (include/comm.hpp, libs/core/comm.cpp) exist.
 *
 * Datatype handle = (class << 16) | size_in_bytes; MPI_Type_contiguous gives class 0xff (opaque bytes).
 */
#ifndef LIBP_B200_MPISTUB_H
#define LIBP_B200_MPISTUB_H
#include <stddef.h>

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
#define MPI_REQUEST_NULL 0
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

int MPI_Init(int* argc, char*** argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm c, int* r);
int MPI_Comm_size(MPI_Comm c, int* s);
int MPI_Comm_dup(MPI_Comm c, MPI_Comm* n);
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm* n);
int MPI_Comm_free(MPI_Comm* c);
int MPI_Barrier(MPI_Comm c);
int MPI_Get_processor_name(char* name, int* len);
int MPI_Type_contiguous(int n, MPI_Datatype t, MPI_Datatype* nt);
int MPI_Type_commit(MPI_Datatype* t);
int MPI_Type_free(MPI_Datatype* t);

int MPI_Send(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c);
int MPI_Recv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Status* st);
int MPI_Isend(const void* b, int n, MPI_Datatype t, int d, int tag, MPI_Comm c, MPI_Request* r);
int MPI_Irecv(void* b, int n, MPI_Datatype t, int s, int tag, MPI_Comm c, MPI_Request* r);
int MPI_Wait(MPI_Request* r, MPI_Status* s);
int MPI_Waitall(int n, MPI_Request* r, MPI_Status* s);

int MPI_Bcast(void* b, int n, MPI_Datatype t, int root, MPI_Comm c);
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c);
int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Iallreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c, MPI_Request* q);
int MPI_Scan(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c);
int MPI_Gather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Gatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rn, const int* displs, MPI_Datatype rt,
                int root, MPI_Comm c);
int MPI_Scatter(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, int root, MPI_Comm c);
int MPI_Scatterv(const void* s, const int* sn, const int* displs, MPI_Datatype st, void* r, int rn, MPI_Datatype rt,
                 int root, MPI_Comm c);
int MPI_Allgather(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c);
int MPI_Allgatherv(const void* s, int sn, MPI_Datatype st, void* r, const int* rn, const int* displs, MPI_Datatype rt,
                   MPI_Comm c);
int MPI_Alltoall(const void* s, int sn, MPI_Datatype st, void* r, int rn, MPI_Datatype rt, MPI_Comm c);
int MPI_Alltoallv(const void* s, const int* sn, const int* sd, MPI_Datatype st, void* r, const int* rn, const int* rd,
                  MPI_Datatype rt, MPI_Comm c);
int MPI_Ialltoallv(const void* s, const int* sn, const int* sd, MPI_Datatype st, void* r, const int* rn, const int* rd,
                   MPI_Datatype rt, MPI_Comm c, MPI_Request* q);

#ifdef __cplusplus
}
#endif
#endif
