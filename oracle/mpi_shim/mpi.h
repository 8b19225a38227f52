/* oracle/mpi_shim/mpi.h -- TEST INFRASTRUCTURE, not product code.
 *
 * A thread-backed stand-in for the ten MPI entry points the HPCCG reference
 * uses (generate_matrix.cpp:207-208, make_local_matrix.cpp:75-76,185,305,353,
 * 389-411,485-534,546-583, exchange_externals.cpp:68-126, ddot.cpp:79-80,
 * compute_residual.cpp:73, mytimer.cpp:49-55).  No MPI exists in this image, so
 * the unmodified reference sources are compiled with -DUSING_MPI against this
 * header and every "rank" is a std::thread of one process (see mpi_shim.cpp).
 *
 * Semantics kept: MPI_Send is buffered (the reference sends before the peer
 * waits, make_local_matrix.cpp:395-397); messages match on (source, tag) in
 * FIFO order; MPI_ANY_SOURCE is honoured; MPI_Allreduce reduces in rank order
 * so the multi-rank oracle is itself deterministic.
 */
#ifndef HPCCG_ORACLE_MPI_SHIM_H
#define HPCCG_ORACLE_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

typedef struct MPI_Status {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
} MPI_Status;

typedef struct MPI_Request {
  void *buf;
  int count;
  MPI_Datatype datatype;
  int source;
  int tag;
  int active;
} MPI_Request;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_ANY_SOURCE (-1)

#define MPI_INT 1
#define MPI_DOUBLE 2

#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3

int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm comm);
int MPI_Wait(MPI_Request *req, MPI_Status *status);
double MPI_Wtime(void);
int MPI_Finalize(void);

/* Shim control (used by ref_driver.cpp only): run fn(rank, arg) on `size`
 * threads, each bound to its rank, and join them. */
void hpccg_shim_run(int size, void (*fn)(int rank, void *arg), void *arg);

#ifdef __cplusplus
}
#endif
#endif
