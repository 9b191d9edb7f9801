/* mpi.h -- the slice of the MPI C API that the unchanged ExaML host files use.
 *
 * ExaML's host code (axml.c, byteFile.c, communication.c, optimizeModel.c, searchAlgo.c)
 * includes <mpi.h> (axml.h:55) and calls exactly the entry points declared below
 * (call sites: axml.c:137,1309,1923,2137,2575-2579; byteFile.c:111-145;
 * communication.c:42-141; optimizeModel.c:1978-1979,2267,2274; searchAlgo.c:2196-2491).
 * No MPI installation exists in the build image, so this header plus ONE of
 *   mpi_single.c   (one process, one GPU; also used to build the CPU reference in oracle/_ref)
 *   mpi_nccl.cpp   (one process per GPU, rank/size from the launcher's environment,
 *                   collectives carried by NCCL over NVLink)
 * stands in for it.  Only what the reference calls is provided; this is not a general MPI.
 */
#ifndef EXAML_B200_MPI_SHIM_H
#define EXAML_B200_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;   /* must be comparable with == (communication.c:44-46) */
typedef int MPI_Op;

#define MPI_COMM_WORLD     0
#define MPI_SUCCESS        0

#define MPI_INT            1
#define MPI_DOUBLE         2
#define MPI_UNSIGNED_LONG  3
#define MPI_BYTE           4

#define MPI_SUM            1

#define MPI_IN_PLACE       ((void *)-1)

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int errorcode);
int MPI_Barrier(MPI_Comm comm);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
int MPI_Bcast(void *buffer, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Scatterv(const void *sendbuf, const int *sendcounts, const int *displs, MPI_Datatype sendtype,
                 void *recvbuf, int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm);
int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
                void *recvbuf, const int *recvcounts, const int *displs, MPI_Datatype recvtype,
                int root, MPI_Comm comm);

#ifdef __cplusplus
}
#endif
#endif
