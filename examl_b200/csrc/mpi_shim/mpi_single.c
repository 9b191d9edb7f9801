/* mpi_single.c -- single-process implementation of mpi_shim/mpi.h (rank 0 of 1).
 * Collectives degenerate to copies.  Used for 1-GPU runs of the drop-in binary and for
 * building the CPU reference under oracle/_ref (no MPI exists in the image). */
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static size_t type_bytes(MPI_Datatype t)
{
  switch (t) {
  case MPI_INT:           return sizeof(int);
  case MPI_DOUBLE:        return sizeof(double);
  case MPI_UNSIGNED_LONG: return sizeof(unsigned long);
  case MPI_BYTE:          return 1;
  default:
    fprintf(stderr, "mpi_single: unsupported datatype %d\n", t);
    abort();
  }
}

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; return MPI_SUCCESS; }
int MPI_Finalize(void) { return MPI_SUCCESS; }
int MPI_Abort(MPI_Comm comm, int errorcode) { (void)comm; exit(errorcode ? errorcode : 1); }
int MPI_Barrier(MPI_Comm comm) { (void)comm; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; *rank = 0; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; *size = 1; return MPI_SUCCESS; }

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm)
{
  (void)op; (void)comm;
  if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf)
    memcpy(recvbuf, sendbuf, (size_t)count * type_bytes(type));
  return MPI_SUCCESS;
}

int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm)
{
  (void)root;
  return MPI_Allreduce(sendbuf, recvbuf, count, type, op, comm);
}

int MPI_Bcast(void *buffer, int count, MPI_Datatype type, int root, MPI_Comm comm)
{
  (void)buffer; (void)count; (void)type; (void)root; (void)comm;
  return MPI_SUCCESS;
}

int MPI_Scatterv(const void *sendbuf, const int *sendcounts, const int *displs, MPI_Datatype sendtype,
                 void *recvbuf, int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm)
{
  (void)sendcounts; (void)recvtype; (void)root; (void)comm;
  memcpy(recvbuf, (const char *)sendbuf + (size_t)displs[0] * type_bytes(sendtype),
         (size_t)recvcount * type_bytes(sendtype));
  return MPI_SUCCESS;
}

int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
                void *recvbuf, const int *recvcounts, const int *displs, MPI_Datatype recvtype,
                int root, MPI_Comm comm)
{
  (void)recvcounts; (void)recvtype; (void)root; (void)comm;
  memcpy((char *)recvbuf + (size_t)displs[0] * type_bytes(sendtype), sendbuf,
         (size_t)sendcount * type_bytes(sendtype));
  return MPI_SUCCESS;
}
