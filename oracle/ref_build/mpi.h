/* Minimal in-box MPI for building the UNMODIFIED HPDDM reference in a container
 * that has no MPI (SURVEY.md section 8c "partial route").  Ranks are forked
 * processes; messages travel through an anonymous shared mapping.  Only what the
 * reference's Schwarz / Krylov path touches is provided.  TEST INFRASTRUCTURE
 * ONLY: used to produce oracle/_ref/ and the golden vectors under tests/golden/.
 */
#ifndef HPDDM_SHIM_MPI_H
#define HPDDM_SHIM_MPI_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MPI_VERSION 3
#define MPI_SUBVERSION 1
#define MPI_SUCCESS 0

typedef int MPI_Comm;     /* index into a per-process table */
typedef int MPI_Group;
typedef int MPI_Datatype; /* enum below */
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct MPI_Status {
  int MPI_SOURCE, MPI_TAG, MPI_ERROR;
  size_t bytes_;
} MPI_Status;
typedef void(MPI_User_function)(void *, void *, int *, MPI_Datatype *);

#define MPI_COMM_NULL 0
#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_GROUP_NULL 0
#define MPI_REQUEST_NULL 0
#define MPI_DATATYPE_NULL 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_IN_PLACE ((void *)-1)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_UNDEFINED (-32766)

enum {
  MPI_CHAR = 1, MPI_UNSIGNED_CHAR, MPI_BYTE, MPI_SHORT, MPI_UNSIGNED_SHORT, MPI_INT, MPI_UNSIGNED, MPI_LONG, MPI_UNSIGNED_LONG,
  MPI_LONG_LONG, MPI_UNSIGNED_LONG_LONG, MPI_FLOAT, MPI_DOUBLE, MPI_C_COMPLEX, MPI_C_DOUBLE_COMPLEX, MPI_C_BOOL, MPI_LONG_DOUBLE
};
#define MPI_LONG_LONG_INT MPI_LONG_LONG
#define MPI_CXX_DOUBLE_COMPLEX MPI_C_DOUBLE_COMPLEX
#define MPI_CXX_FLOAT_COMPLEX MPI_C_COMPLEX
enum { MPI_SUM = 1, MPI_MAX, MPI_MIN, MPI_PROD, MPI_LOR, MPI_LAND, MPI_BOR, MPI_BAND, MPI_OP_USER_BASE = 100 };
enum { MPI_IDENT = 0, MPI_CONGRUENT = 1, MPI_SIMILAR = 2, MPI_UNEQUAL = 3 };

int MPI_Init(int *, char ***);
int MPI_Finalize(void);
int MPI_Finalized(int *);
int MPI_Initialized(int *);
int MPI_Abort(MPI_Comm, int);
double MPI_Wtime(void);
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Comm_dup(MPI_Comm, MPI_Comm *);
int MPI_Comm_free(MPI_Comm *);
int MPI_Comm_group(MPI_Comm, MPI_Group *);
int MPI_Comm_create(MPI_Comm, MPI_Group, MPI_Comm *);
int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm *);
int MPI_Comm_compare(MPI_Comm, MPI_Comm, int *);
int MPI_Group_incl(MPI_Group, int, const int *, MPI_Group *);
int MPI_Group_excl(MPI_Group, int, const int *, MPI_Group *);
int MPI_Group_free(MPI_Group *);
int MPI_Group_size(MPI_Group, int *);
int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *);
int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Wait(MPI_Request *, MPI_Status *);
int MPI_Waitall(int, MPI_Request *, MPI_Status *);
int MPI_Waitany(int, MPI_Request *, int *, MPI_Status *);
int MPI_Test(MPI_Request *, int *, MPI_Status *);
int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *);
int MPI_Barrier(MPI_Comm);
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Gather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Gatherv(const void *, int, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, int, MPI_Comm);
int MPI_Scatter(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Scatterv(const void *, const int *, const int *, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Allgatherv(const void *, int, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, MPI_Comm);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce(const void *, void *, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Op_create(MPI_User_function *, int, MPI_Op *);
int MPI_Op_free(MPI_Op *);
/* non-blocking collectives (HPDDM_ICOLLECTIVE is off; executed eagerly) */
int MPI_Igather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm, MPI_Request *);
int MPI_Igatherv(const void *, int, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, int, MPI_Comm, MPI_Request *);
int MPI_Iscatter(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm, MPI_Request *);
int MPI_Iscatterv(const void *, const int *, const int *, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm, MPI_Request *);
int MPI_Iallreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm, MPI_Request *);

#ifdef __cplusplus
}
#endif
#endif
