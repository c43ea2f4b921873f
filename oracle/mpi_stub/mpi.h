/* mpi.h -- TEST INFRASTRUCTURE ONLY: a single-rank loop-back stand-in for the handful of MPI calls the reference's
 * stand-alone transpose tool makes (cpp/exec/upsp_matrix_transpose.cpp: Init / Comm_rank / Comm_size / Barrier / Wtime /
 * Isend + Recv to itself / Get_count / Waitall / Finalize), so that the tool can be compiled from the reference tree
 * where no MPI exists (`make -C oracle ref`) and its own global_transpose can be run as a checker.  One rank only:
 * a send is queued (pointer + count), the next receive copies it out. */
#ifndef UPSP_ORACLE_MPI_STUB_H
#define UPSP_ORACLE_MPI_STUB_H
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR, count_; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_FLOAT 4
#define MPI_ANY_SOURCE (-1)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)
#define MPI_SUCCESS 0

static struct { const void* buf; int count; } mpi_stub_queue[64];
static int mpi_stub_head = 0, mpi_stub_tail = 0;

static inline int MPI_Init(int* argc, char*** argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int* rank) { (void)c; *rank = 0; return MPI_SUCCESS; }
static inline int MPI_Comm_size(MPI_Comm c, int* size) { (void)c; *size = 1; return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm c) { (void)c; return MPI_SUCCESS; }
static inline double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static inline int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request* req) {
  (void)t; (void)tag; (void)c;
  if (dest != 0 || mpi_stub_tail - mpi_stub_head >= 64) abort();
  mpi_stub_queue[mpi_stub_tail % 64].buf = buf;
  mpi_stub_queue[mpi_stub_tail % 64].count = count;
  ++mpi_stub_tail;
  if (req) *req = 0;
  return MPI_SUCCESS;
}
static inline int MPI_Recv(void* buf, int max_count, MPI_Datatype t, int source, int tag, MPI_Comm c, MPI_Status* st) {
  (void)source; (void)tag; (void)c;
  if (mpi_stub_head == mpi_stub_tail) abort();      /* nothing was sent: a real run would dead-lock */
  const int count = mpi_stub_queue[mpi_stub_head % 64].count;
  if (count > max_count) abort();
  memcpy(buf, mpi_stub_queue[mpi_stub_head % 64].buf, (size_t)count * (size_t)t);
  ++mpi_stub_head;
  if (st) { st->MPI_SOURCE = 0; st->MPI_TAG = 0; st->MPI_ERROR = 0; st->count_ = count; }
  return MPI_SUCCESS;
}
static inline int MPI_Get_count(const MPI_Status* st, MPI_Datatype t, int* count) { (void)t; *count = st->count_; return MPI_SUCCESS; }
static inline int MPI_Waitall(int n, MPI_Request* reqs, MPI_Status* sts) { (void)n; (void)reqs; (void)sts; return MPI_SUCCESS; }
#endif
