/*
 * Single-rank MPI stand-in used ONLY to compile the unmodified CombBLAS reference
 * (read in place from /root/reference) into the parity oracle `oracle/_ref/`.
 *
 * TEST INFRASTRUCTURE. Not part of the product, never linked into libcbgpu.so.
 *
 * World size is 1: collectives copy the caller's own contribution (honouring
 * MPI_IN_PLACE, counts and displacements); point-to-point and one-sided calls abort,
 * they are never reached at P=1 on the SpGEMM path. Datatype handles carry their byte
 * size so that MPI_Type_contiguous(sizeof(tuple), MPI_CHAR) (how the reference ships
 * std::tuple buffers, ParFriends.h:3557, MPIType.h:113) copies the right amount.
 */
#ifndef CBGPU_ORACLE_MPI_SHIM_H
#define CBGPU_ORACLE_MPI_SHIM_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Win;
typedef int MPI_Group;
typedef int MPI_Request;
typedef int MPI_Info;
typedef int MPI_File;
typedef int MPI_Errhandler;
typedef long MPI_Aint;
typedef long long MPI_Offset;
typedef struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; int _count; } MPI_Status;
typedef void(MPI_User_function)(void *, void *, int *, MPI_Datatype *);

#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_COMM_NULL 0
#define MPI_GROUP_NULL 0
#define MPI_IN_PLACE ((void *)1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_REQUEST_NULL 0
#define MPI_INFO_NULL 0
#define MPI_OP_NULL 0
#define MPI_DATATYPE_NULL 0
#define MPI_WIN_NULL 0
#define MPI_FILE_NULL 0
#define MPI_IDENT 0
#define MPI_CONGRUENT 1
#define MPI_SIMILAR 2
#define MPI_UNEQUAL 3
#define MPI_SUCCESS 0
#define MPI_MAX_ERROR_STRING 256
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3
#define MPI_LOCK_EXCLUSIVE 1
#define MPI_LOCK_SHARED 2
#define MPI_MODE_CREATE 1
#define MPI_MODE_RDONLY 2
#define MPI_MODE_WRONLY 4
#define MPI_MODE_RDWR 8
#define MPI_MODE_NOCHECK 16
#define MPI_MODE_NOSTORE 32
#define MPI_MODE_NOPUT 64
#define MPI_MODE_NOPRECEDE 128
#define MPI_MODE_NOSUCCEED 256
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_UNDEFINED (-32766)
#define MPI_ERRORS_RETURN 1
#define MPI_ERRORS_ARE_FATAL 2

/* ops */
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_PROD 4
#define MPI_BAND 5
#define MPI_BOR 6
#define MPI_BXOR 7
#define MPI_LAND 8
#define MPI_LOR 9
#define MPI_LXOR 10
#define MPI_MAXLOC 11
#define MPI_MINLOC 12
#define MPI_REPLACE 13

/* datatype handle = (kind << 20) | byte size */
#define CBSHIM_DT(kind, size) (((kind) << 20) | (size))
#define CBSHIM_DT_SIZE(dt) ((long)((dt)&0xFFFFF))
#define MPI_CHAR CBSHIM_DT(1, 1)
#define MPI_SIGNED_CHAR CBSHIM_DT(2, 1)
#define MPI_UNSIGNED_CHAR CBSHIM_DT(3, 1)
#define MPI_BYTE CBSHIM_DT(4, 1)
#define MPI_SHORT CBSHIM_DT(5, 2)
#define MPI_UNSIGNED_SHORT CBSHIM_DT(6, 2)
#define MPI_INT CBSHIM_DT(7, 4)
#define MPI_UNSIGNED CBSHIM_DT(8, 4)
#define MPI_LONG CBSHIM_DT(9, 8)
#define MPI_UNSIGNED_LONG CBSHIM_DT(10, 8)
#define MPI_LONG_LONG CBSHIM_DT(11, 8)
#define MPI_LONG_LONG_INT CBSHIM_DT(11, 8)
#define MPI_UNSIGNED_LONG_LONG CBSHIM_DT(12, 8)
#define MPI_FLOAT CBSHIM_DT(13, 4)
#define MPI_DOUBLE CBSHIM_DT(14, 8)
#define MPI_LONG_DOUBLE CBSHIM_DT(15, 16)
#define MPI_2INT CBSHIM_DT(16, 8)
#define MPI_FLOAT_INT CBSHIM_DT(17, 8)
#define MPI_DOUBLE_INT CBSHIM_DT(18, 16)
#define MPI_LONG_INT CBSHIM_DT(19, 16)
#define MPI_SHORT_INT CBSHIM_DT(20, 8)
#define MPI_LONG_DOUBLE_INT CBSHIM_DT(21, 32)
#define MPI_C_BOOL CBSHIM_DT(22, 1)
#define MPI_CXX_BOOL CBSHIM_DT(22, 1)
#define MPI_INT8_T CBSHIM_DT(23, 1)
#define MPI_INT16_T CBSHIM_DT(24, 2)
#define MPI_INT32_T CBSHIM_DT(25, 4)
#define MPI_INT64_T CBSHIM_DT(26, 8)
#define MPI_UINT8_T CBSHIM_DT(27, 1)
#define MPI_UINT16_T CBSHIM_DT(28, 2)
#define MPI_UINT32_T CBSHIM_DT(29, 4)
#define MPI_UINT64_T CBSHIM_DT(30, 8)
#define MPI_WCHAR CBSHIM_DT(31, 4)
#define CBSHIM_KIND_DERIVED 100

static inline void cbshim_unsupported(const char *what) {
  fprintf(stderr, "[mpi shim] %s is not available in the single-rank oracle build\n", what);
  abort();
}
static inline void cbshim_copy(const void *src, void *dst, long bytes) {
  if (src == MPI_IN_PLACE || src == dst || bytes <= 0 || src == NULL || dst == NULL) return;
  memmove(dst, src, (size_t)bytes);
}

/* ---- environment ---- */
static inline int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Init_thread(int *argc, char ***argv, int required, int *provided) {
  (void)argc; (void)argv; if (provided) *provided = required; return MPI_SUCCESS;
}
static inline int MPI_Query_thread(int *provided) { *provided = MPI_THREAD_MULTIPLE; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Finalized(int *flag) { *flag = 0; return MPI_SUCCESS; }
static inline int MPI_Initialized(int *flag) { *flag = 1; return MPI_SUCCESS; }
static inline int MPI_Abort(MPI_Comm c, int code) {
  (void)c; fprintf(stderr, "[mpi shim] MPI_Abort(code=%d)\n", code); exit(code ? code : 1); return 0;
}
static inline double MPI_Wtime(void) {
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline double MPI_Wtick(void) { return 1e-9; }
static inline int MPI_Get_processor_name(char *name, int *len) { strcpy(name, "localhost"); *len = 9; return MPI_SUCCESS; }
static inline int MPI_Error_string(int code, char *s, int *len) { *len = snprintf(s, MPI_MAX_ERROR_STRING, "mpi shim error %d", code); return MPI_SUCCESS; }
static inline int MPI_Pcontrol(const int level, ...) { (void)level; return MPI_SUCCESS; }
static inline int MPI_Alloc_mem(MPI_Aint size, MPI_Info info, void *baseptr) { (void)info; *(void **)baseptr = malloc((size_t)size); return MPI_SUCCESS; }
static inline int MPI_Free_mem(void *base) { free(base); return MPI_SUCCESS; }
static inline int MPI_Comm_set_errhandler(MPI_Comm c, MPI_Errhandler e) { (void)c; (void)e; return MPI_SUCCESS; }

/* ---- communicators / groups ---- */
static inline int MPI_Comm_rank(MPI_Comm c, int *r) { (void)c; *r = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int *s) { (void)c; *s = 1; return MPI_SUCCESS; }
static inline int MPI_Comm_dup(MPI_Comm c, MPI_Comm *n) { *n = c; return MPI_SUCCESS; }
static inline int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm *n) { (void)color; (void)key; *n = c; return MPI_SUCCESS; }
static inline int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm *n) { (void)g; *n = c; return MPI_SUCCESS; }
static inline int MPI_Comm_free(MPI_Comm *c) { *c = MPI_COMM_NULL; return MPI_SUCCESS; }
static inline int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int *result) { (void)a; (void)b; *result = MPI_IDENT; return MPI_SUCCESS; }
static inline int MPI_Comm_group(MPI_Comm c, MPI_Group *g) { (void)c; *g = 1; return MPI_SUCCESS; }
static inline int MPI_Group_incl(MPI_Group g, int n, const int *ranks, MPI_Group *ng) { (void)g; (void)n; (void)ranks; *ng = 1; return MPI_SUCCESS; }
static inline int MPI_Group_excl(MPI_Group g, int n, const int *ranks, MPI_Group *ng) { (void)g; (void)n; (void)ranks; *ng = 1; return MPI_SUCCESS; }
static inline int MPI_Group_free(MPI_Group *g) { *g = MPI_GROUP_NULL; return MPI_SUCCESS; }
static inline int MPI_Group_rank(MPI_Group g, int *r) { (void)g; *r = 0; return MPI_SUCCESS; }
static inline int MPI_Group_size(MPI_Group g, int *s) { (void)g; *s = 1; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }

/* ---- datatypes / ops ---- */
static inline int MPI_Type_size(MPI_Datatype dt, int *size) { *size = (int)CBSHIM_DT_SIZE(dt); return MPI_SUCCESS; }
static inline int MPI_Type_contiguous(int n, MPI_Datatype old, MPI_Datatype *nt) {
  long bytes = (long)n * CBSHIM_DT_SIZE(old);
  if (bytes >= (1L << 20)) cbshim_unsupported("MPI_Type_contiguous > 1 MiB element");
  *nt = CBSHIM_DT(CBSHIM_KIND_DERIVED, (int)bytes); return MPI_SUCCESS;
}
static inline int MPI_Type_create_struct(int count, const int *bl, const MPI_Aint *disp, const MPI_Datatype *types, MPI_Datatype *nt) {
  long extent = 0; int i;
  for (i = 0; i < count; ++i) { long e = disp[i] + (long)bl[i] * CBSHIM_DT_SIZE(types[i]); if (e > extent) extent = e; }
  extent = (extent + 7) & ~7L; /* struct users on this path are 8-byte aligned pairs */
  *nt = CBSHIM_DT(CBSHIM_KIND_DERIVED, (int)extent); return MPI_SUCCESS;
}
static inline int MPI_Type_commit(MPI_Datatype *dt) { (void)dt; return MPI_SUCCESS; }
static inline int MPI_Type_free(MPI_Datatype *dt) { *dt = MPI_DATATYPE_NULL; return MPI_SUCCESS; }
static inline int MPI_Get_address(const void *loc, MPI_Aint *a) { *a = (MPI_Aint)loc; return MPI_SUCCESS; }
static inline int MPI_Op_create(MPI_User_function *f, int commute, MPI_Op *op) { (void)f; (void)commute; *op = 1000; return MPI_SUCCESS; }
static inline int MPI_Op_free(MPI_Op *op) { *op = MPI_OP_NULL; return MPI_SUCCESS; }

/* ---- collectives: a single rank's contribution is the result ---- */
static inline int MPI_Bcast(void *b, int n, MPI_Datatype dt, int root, MPI_Comm c) { (void)b; (void)n; (void)dt; (void)root; (void)c; return MPI_SUCCESS; }
static inline int MPI_Ibcast(void *b, int n, MPI_Datatype dt, int root, MPI_Comm c, MPI_Request *r) { (void)b; (void)n; (void)dt; (void)root; (void)c; *r = 0; return MPI_SUCCESS; }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype dt, MPI_Op op, MPI_Comm c) { (void)op; (void)c; cbshim_copy(s, r, (long)n * CBSHIM_DT_SIZE(dt)); return MPI_SUCCESS; }
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype dt, MPI_Op op, int root, MPI_Comm c) { (void)op; (void)root; (void)c; cbshim_copy(s, r, (long)n * CBSHIM_DT_SIZE(dt)); return MPI_SUCCESS; }
static inline int MPI_Scan(const void *s, void *r, int n, MPI_Datatype dt, MPI_Op op, MPI_Comm c) { (void)op; (void)c; cbshim_copy(s, r, (long)n * CBSHIM_DT_SIZE(dt)); return MPI_SUCCESS; }
/* Exscan leaves rank 0's receive buffer untouched (MPI standard: undefined on rank 0). */
static inline int MPI_Exscan(const void *s, void *r, int n, MPI_Datatype dt, MPI_Op op, MPI_Comm c) { (void)s; (void)r; (void)n; (void)dt; (void)op; (void)c; return MPI_SUCCESS; }
static inline int MPI_Reduce_scatter(const void *s, void *r, const int *cnts, MPI_Datatype dt, MPI_Op op, MPI_Comm c) { (void)op; (void)c; cbshim_copy(s, r, (long)cnts[0] * CBSHIM_DT_SIZE(dt)); return MPI_SUCCESS; }
static inline int MPI_Allgather(const void *s, int sn, MPI_Datatype sdt, void *r, int rn, MPI_Datatype rdt, MPI_Comm c) { (void)rn; (void)rdt; (void)c; cbshim_copy(s, r, (long)sn * CBSHIM_DT_SIZE(sdt)); return MPI_SUCCESS; }
static inline int MPI_Allgatherv(const void *s, int sn, MPI_Datatype sdt, void *r, const int *rn, const int *displs, MPI_Datatype rdt, MPI_Comm c) {
  (void)rn; (void)c; if (s != MPI_IN_PLACE) cbshim_copy(s, (char *)r + (long)displs[0] * CBSHIM_DT_SIZE(rdt), (long)sn * CBSHIM_DT_SIZE(sdt)); return MPI_SUCCESS;
}
static inline int MPI_Gather(const void *s, int sn, MPI_Datatype sdt, void *r, int rn, MPI_Datatype rdt, int root, MPI_Comm c) { (void)rn; (void)rdt; (void)root; (void)c; cbshim_copy(s, r, (long)sn * CBSHIM_DT_SIZE(sdt)); return MPI_SUCCESS; }
static inline int MPI_Gatherv(const void *s, int sn, MPI_Datatype sdt, void *r, const int *rn, const int *displs, MPI_Datatype rdt, int root, MPI_Comm c) {
  (void)rn; (void)root; (void)c; if (s != MPI_IN_PLACE) cbshim_copy(s, (char *)r + (long)displs[0] * CBSHIM_DT_SIZE(rdt), (long)sn * CBSHIM_DT_SIZE(sdt)); return MPI_SUCCESS;
}
static inline int MPI_Scatter(const void *s, int sn, MPI_Datatype sdt, void *r, int rn, MPI_Datatype rdt, int root, MPI_Comm c) { (void)sn; (void)sdt; (void)root; (void)c; if (r != MPI_IN_PLACE) cbshim_copy(s, r, (long)rn * CBSHIM_DT_SIZE(rdt)); return MPI_SUCCESS; }
static inline int MPI_Scatterv(const void *s, const int *sn, const int *displs, MPI_Datatype sdt, void *r, int rn, MPI_Datatype rdt, int root, MPI_Comm c) {
  (void)sn; (void)root; (void)c; if (r != MPI_IN_PLACE) cbshim_copy((const char *)s + (long)displs[0] * CBSHIM_DT_SIZE(sdt), r, (long)rn * CBSHIM_DT_SIZE(rdt)); return MPI_SUCCESS;
}
static inline int MPI_Alltoall(const void *s, int sn, MPI_Datatype sdt, void *r, int rn, MPI_Datatype rdt, MPI_Comm c) { (void)rn; (void)rdt; (void)c; cbshim_copy(s, r, (long)sn * CBSHIM_DT_SIZE(sdt)); return MPI_SUCCESS; }
static inline int MPI_Alltoallv(const void *s, const int *sn, const int *sd, MPI_Datatype sdt, void *r, const int *rn, const int *rd, MPI_Datatype rdt, MPI_Comm c) {
  (void)rn; (void)c;
  if (s != MPI_IN_PLACE) cbshim_copy((const char *)s + (long)sd[0] * CBSHIM_DT_SIZE(sdt), (char *)r + (long)rd[0] * CBSHIM_DT_SIZE(rdt), (long)sn[0] * CBSHIM_DT_SIZE(sdt));
  return MPI_SUCCESS;
}
static inline int MPI_Sendrecv(const void *s, int sn, MPI_Datatype sdt, int dest, int stag, void *r, int rn, MPI_Datatype rdt, int src, int rtag, MPI_Comm c, MPI_Status *st) {
  (void)dest; (void)stag; (void)rn; (void)rdt; (void)src; (void)rtag; (void)c;
  cbshim_copy(s, r, (long)sn * CBSHIM_DT_SIZE(sdt));
  if (st) { st->MPI_SOURCE = 0; st->MPI_TAG = rtag; st->MPI_ERROR = 0; st->_count = sn; }
  return MPI_SUCCESS;
}

/* ---- point to point / one sided: unreachable at P=1 on the SpGEMM path ---- */
static inline int MPI_Send(const void *b, int n, MPI_Datatype dt, int dest, int tag, MPI_Comm c) { (void)b; (void)n; (void)dt; (void)dest; (void)tag; (void)c; cbshim_unsupported("MPI_Send"); return 1; }
static inline int MPI_Recv(void *b, int n, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Status *st) { (void)b; (void)n; (void)dt; (void)src; (void)tag; (void)c; (void)st; cbshim_unsupported("MPI_Recv"); return 1; }
static inline int MPI_Isend(const void *b, int n, MPI_Datatype dt, int dest, int tag, MPI_Comm c, MPI_Request *r) { (void)b; (void)n; (void)dt; (void)dest; (void)tag; (void)c; (void)r; cbshim_unsupported("MPI_Isend"); return 1; }
static inline int MPI_Issend(const void *b, int n, MPI_Datatype dt, int dest, int tag, MPI_Comm c, MPI_Request *r) { (void)b; (void)n; (void)dt; (void)dest; (void)tag; (void)c; (void)r; cbshim_unsupported("MPI_Issend"); return 1; }
static inline int MPI_Irecv(void *b, int n, MPI_Datatype dt, int src, int tag, MPI_Comm c, MPI_Request *r) { (void)b; (void)n; (void)dt; (void)src; (void)tag; (void)c; (void)r; cbshim_unsupported("MPI_Irecv"); return 1; }
static inline int MPI_Wait(MPI_Request *r, MPI_Status *st) { (void)r; (void)st; return MPI_SUCCESS; }
static inline int MPI_Waitall(int n, MPI_Request *r, MPI_Status *st) { (void)n; (void)r; (void)st; return MPI_SUCCESS; }
static inline int MPI_Test(MPI_Request *r, int *flag, MPI_Status *st) { (void)r; (void)st; *flag = 1; return MPI_SUCCESS; }
static inline int MPI_Get_count(const MPI_Status *st, MPI_Datatype dt, int *count) { (void)dt; *count = st ? st->_count : 0; return MPI_SUCCESS; }
static inline int MPI_Win_create(void *base, MPI_Aint size, int disp, MPI_Info info, MPI_Comm c, MPI_Win *w) { (void)base; (void)size; (void)disp; (void)info; (void)c; *w = 1; return MPI_SUCCESS; }
static inline int MPI_Win_free(MPI_Win *w) { *w = MPI_WIN_NULL; return MPI_SUCCESS; }
static inline int MPI_Win_fence(int a, MPI_Win w) { (void)a; (void)w; return MPI_SUCCESS; }
static inline int MPI_Win_lock(int t, int rank, int a, MPI_Win w) { (void)t; (void)rank; (void)a; (void)w; return MPI_SUCCESS; }
static inline int MPI_Win_unlock(int rank, MPI_Win w) { (void)rank; (void)w; return MPI_SUCCESS; }
static inline int MPI_Win_post(MPI_Group g, int a, MPI_Win w) { (void)g; (void)a; (void)w; return MPI_SUCCESS; }
static inline int MPI_Win_start(MPI_Group g, int a, MPI_Win w) { (void)g; (void)a; (void)w; return MPI_SUCCESS; }
static inline int MPI_Win_complete(MPI_Win w) { (void)w; return MPI_SUCCESS; }
static inline int MPI_Win_wait(MPI_Win w) { (void)w; return MPI_SUCCESS; }
static inline int MPI_Get(void *o, int on, MPI_Datatype odt, int rank, MPI_Aint disp, int tn, MPI_Datatype tdt, MPI_Win w) { (void)o; (void)on; (void)odt; (void)rank; (void)disp; (void)tn; (void)tdt; (void)w; cbshim_unsupported("MPI_Get"); return 1; }
static inline int MPI_Put(const void *o, int on, MPI_Datatype odt, int rank, MPI_Aint disp, int tn, MPI_Datatype tdt, MPI_Win w) { (void)o; (void)on; (void)odt; (void)rank; (void)disp; (void)tn; (void)tdt; (void)w; cbshim_unsupported("MPI_Put"); return 1; }
static inline int MPI_Info_create(MPI_Info *i) { *i = 1; return MPI_SUCCESS; }
static inline int MPI_Info_set(MPI_Info i, const char *k, const char *v) { (void)i; (void)k; (void)v; return MPI_SUCCESS; }
static inline int MPI_Info_free(MPI_Info *i) { *i = MPI_INFO_NULL; return MPI_SUCCESS; }

/* ---- file I/O: matrix readers/writers are outside the SpGEMM path ---- */
static inline int MPI_File_open(MPI_Comm c, const char *fn, int amode, MPI_Info info, MPI_File *fh) { (void)c; (void)fn; (void)amode; (void)info; (void)fh; cbshim_unsupported("MPI_File_open"); return 1; }
static inline int MPI_File_close(MPI_File *fh) { (void)fh; return MPI_SUCCESS; }
static inline int MPI_File_set_view(MPI_File fh, MPI_Offset disp, MPI_Datatype e, MPI_Datatype f, const char *rep, MPI_Info info) { (void)fh; (void)disp; (void)e; (void)f; (void)rep; (void)info; cbshim_unsupported("MPI_File_set_view"); return 1; }
static inline int MPI_File_write(MPI_File fh, const void *b, int n, MPI_Datatype dt, MPI_Status *st) { (void)fh; (void)b; (void)n; (void)dt; (void)st; cbshim_unsupported("MPI_File_write"); return 1; }
static inline int MPI_File_write_all(MPI_File fh, const void *b, int n, MPI_Datatype dt, MPI_Status *st) { (void)fh; (void)b; (void)n; (void)dt; (void)st; cbshim_unsupported("MPI_File_write_all"); return 1; }
static inline int MPI_File_read_at(MPI_File fh, MPI_Offset off, void *b, int n, MPI_Datatype dt, MPI_Status *st) { (void)fh; (void)off; (void)b; (void)n; (void)dt; (void)st; cbshim_unsupported("MPI_File_read_at"); return 1; }
static inline int MPI_File_get_size(MPI_File fh, MPI_Offset *sz) { (void)fh; (void)sz; cbshim_unsupported("MPI_File_get_size"); return 1; }

#ifdef __cplusplus
}
#endif
#endif /* CBGPU_ORACLE_MPI_SHIM_H */
