/* mpi_tcp.c -- multi-process implementation of mpi_shim/mpi.h for images without an MPI.
 *
 * One process per GPU, started by examl_b200/launch.py (or torchrun --no-python, or any launcher that
 * exports RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT).  Only the control-plane collectives of the
 * unchanged ExaML host files travel here (a few ints/doubles per call: axml.c:2137,
 * optimizeModel.c:1978-1979, 2267, 2274, searchAlgo.c:2196-2491, the NCCL id bootstrap of
 * axml_backend.c); the two hot-path all-reduces run over NCCL inside the engine.  Topology: a star on
 * rank 0 over TCP; reductions are applied in rank order, so results are deterministic.
 */
#include "mpi.h"

#include <arpa/inet.h>
#include <errno.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>

static int g_rank = 0, g_size = 1;
static int *g_peer = NULL;       /* rank 0: socket of every other rank; others: g_peer[0] = socket to rank 0 */

static void die(const char *what)
{
  fprintf(stderr, "mpi_tcp rank %d: %s: %s\n", g_rank, what, strerror(errno));
  exit(70);
}

static size_t type_bytes(MPI_Datatype t)
{
  switch (t) {
  case MPI_INT:           return sizeof(int);
  case MPI_DOUBLE:        return sizeof(double);
  case MPI_UNSIGNED_LONG: return sizeof(unsigned long);
  case MPI_BYTE:          return 1;
  default: fprintf(stderr, "mpi_tcp: unsupported datatype %d\n", t); exit(70);
  }
}

static void send_all(int fd, const void *buf, size_t n)
{
  const char *p = (const char *)buf;
  while (n) {
    ssize_t k = send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) { if (k < 0 && errno == EINTR) continue; die("send"); }
    p += k; n -= (size_t)k;
  }
}

static void recv_all(int fd, void *buf, size_t n)
{
  char *p = (char *)buf;
  while (n) {
    ssize_t k = recv(fd, p, n, 0);
    if (k <= 0) { if (k < 0 && errno == EINTR) continue; die("recv (peer gone?)"); }
    p += k; n -= (size_t)k;
  }
}

static int env_int(const char *const *names, int dflt)
{
  int i;
  for (i = 0; names[i]; i++) {
    const char *v = getenv(names[i]);
    if (v && *v) return atoi(v);
  }
  return dflt;
}

int MPI_Init(int *argc, char ***argv)
{
  static const char *const rank_names[] = {"RANK", "PMI_RANK", "OMPI_COMM_WORLD_RANK", "SLURM_PROCID", NULL};
  static const char *const size_names[] = {"WORLD_SIZE", "PMI_SIZE", "OMPI_COMM_WORLD_SIZE", "SLURM_NTASKS", NULL};
  const char *addr = getenv("MASTER_ADDR"), *port_s = getenv("MASTER_PORT");
  int port = port_s ? atoi(port_s) : 29400, one = 1, i;
  const char *tok_s = getenv("MPI_TCP_TOKEN");
  const unsigned long long token = tok_s ? strtoull(tok_s, NULL, 16) : 0ULL;
  (void)argc; (void)argv;
  g_rank = env_int(rank_names, 0);
  g_size = env_int(size_names, 1);
  if (g_size <= 1) { g_size = 1; g_rank = 0; return MPI_SUCCESS; }
  if (!addr) addr = "127.0.0.1";
  g_peer = (int *)calloc((size_t)g_size, sizeof(int));

  if (g_rank == 0) {
    struct sockaddr_in sa;
    int ls = socket(AF_INET, SOCK_STREAM, 0);
    if (ls < 0) die("socket");
    setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    memset(&sa, 0, sizeof sa);
    /* listen on MASTER_ADDR only (loopback for examl_b200/launch.py), not on every interface; MASTER_ADDR given as a
       host name that is not a dotted quad: all interfaces as before.  MPI_TCP_TOKEN (set by the launcher) must be echoed
       by every rank that connects. */
    sa.sin_family = AF_INET; sa.sin_port = htons((unsigned short)port);
    if (inet_pton(AF_INET, addr, &sa.sin_addr) != 1) sa.sin_addr.s_addr = htonl(INADDR_ANY);
    if (bind(ls, (struct sockaddr *)&sa, sizeof sa) < 0) die("bind");
    if (listen(ls, g_size) < 0) die("listen");
    for (i = 1; i < g_size; i++) {
      int fd = accept(ls, NULL, NULL), r;
      if (fd < 0) die("accept");
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
      recv_all(fd, &r, sizeof r);
      { unsigned long long tok = 0; recv_all(fd, &tok, sizeof tok);
        if (tok != token) { fprintf(stderr, "mpi_tcp: connection with a wrong token refused\n"); close(fd); i--; continue; } }
      if (r < 1 || r >= g_size || g_peer[r]) { fprintf(stderr, "mpi_tcp: bad rank %d announced\n", r); exit(70); }
      g_peer[r] = fd;
    }
    close(ls);
  } else {
    struct addrinfo hints, *res = NULL;
    char portbuf[16];
    int fd = -1, tries;
    memset(&hints, 0, sizeof hints);
    hints.ai_family = AF_INET; hints.ai_socktype = SOCK_STREAM;
    snprintf(portbuf, sizeof portbuf, "%d", port);
    if (getaddrinfo(addr, portbuf, &hints, &res) != 0 || !res) die("getaddrinfo(MASTER_ADDR)");
    for (tries = 0; tries < 600; tries++) {                 /* rank 0 may not be listening yet */
      fd = socket(AF_INET, SOCK_STREAM, 0);
      if (fd < 0) die("socket");
      if (connect(fd, res->ai_addr, res->ai_addrlen) == 0) break;
      close(fd); fd = -1;
      usleep(100000);
    }
    freeaddrinfo(res);
    if (fd < 0) die("connect to rank 0");
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    send_all(fd, &g_rank, sizeof g_rank);
    send_all(fd, &token, sizeof token);
    g_peer[0] = fd;
  }
  return MPI_SUCCESS;
}

int MPI_Finalize(void)
{
  /* The sockets stay open until the process exits: exit handlers (the test probe's final dump) may still
     run collectives, and ExaML calls MPI_Finalize right before exit() everywhere. */
  if (g_size > 1) MPI_Barrier(MPI_COMM_WORLD);
  return MPI_SUCCESS;
}

/* Peers blocked in a collective notice the closed sockets (recv_all fails -> die) instead of waiting for ever. */
int MPI_Abort(MPI_Comm comm, int errorcode)
{
  int i;
  (void)comm;
  for (i = 0; g_peer && i < g_size; i++) if (g_peer[i] > 0) shutdown(g_peer[i], SHUT_RDWR);
  exit(errorcode > 0 && errorcode < 256 ? errorcode : 1);
}
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; *rank = g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; *size = g_size; return MPI_SUCCESS; }

static void add_into(void *acc, const void *x, int count, MPI_Datatype t)
{
  int i;
  switch (t) {
  case MPI_INT:           for (i = 0; i < count; i++) ((int *)acc)[i] += ((const int *)x)[i]; break;
  case MPI_DOUBLE:        for (i = 0; i < count; i++) ((double *)acc)[i] += ((const double *)x)[i]; break;
  case MPI_UNSIGNED_LONG: for (i = 0; i < count; i++) ((unsigned long *)acc)[i] += ((const unsigned long *)x)[i]; break;
  default: fprintf(stderr, "mpi_tcp: cannot reduce datatype %d\n", t); exit(70);
  }
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm)
{
  const size_t n = (size_t)count * type_bytes(type);
  int i;
  (void)op; (void)comm;
  if (sendbuf != MPI_IN_PLACE && sendbuf != recvbuf) memcpy(recvbuf, sendbuf, n);
  if (g_size == 1) return MPI_SUCCESS;
  if (g_rank == 0) {
    void *tmp = malloc(n ? n : 1);
    for (i = 1; i < g_size; i++) { recv_all(g_peer[i], tmp, n); add_into(recvbuf, tmp, count, type); }
    for (i = 1; i < g_size; i++) send_all(g_peer[i], recvbuf, n);
    free(tmp);
  } else {
    send_all(g_peer[0], recvbuf, n);
    recv_all(g_peer[0], recvbuf, n);
  }
  return MPI_SUCCESS;
}

int MPI_Reduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm)
{
  (void)root;     /* every rank receives the result; the reference follows each Reduce with a Bcast anyway */
  return MPI_Allreduce(sendbuf, recvbuf, count, type, op, comm);
}

int MPI_Barrier(MPI_Comm comm)
{
  int token = 0;
  return MPI_Allreduce(MPI_IN_PLACE, &token, 1, MPI_INT, MPI_SUM, comm);
}

int MPI_Bcast(void *buffer, int count, MPI_Datatype type, int root, MPI_Comm comm)
{
  const size_t n = (size_t)count * type_bytes(type);
  int i;
  (void)comm;
  if (g_size == 1) return MPI_SUCCESS;
  if (root != 0) {                       /* route through rank 0 */
    if (g_rank == root) send_all(g_peer[0], buffer, n);
    if (g_rank == 0) recv_all(g_peer[root], buffer, n);
  }
  if (g_rank == 0) { for (i = 1; i < g_size; i++) if (i != root) send_all(g_peer[i], buffer, n); }
  else if (g_rank != root) recv_all(g_peer[0], buffer, n);
  return MPI_SUCCESS;
}

int MPI_Scatterv(const void *sendbuf, const int *sendcounts, const int *displs, MPI_Datatype sendtype,
                 void *recvbuf, int recvcount, MPI_Datatype recvtype, int root, MPI_Comm comm)
{
  const size_t w = type_bytes(sendtype);
  int i;
  (void)recvtype; (void)comm;
  if (root != 0) { fprintf(stderr, "mpi_tcp: Scatterv root must be 0\n"); exit(70); }
  if (g_rank == 0) {
    for (i = 1; i < g_size; i++)
      send_all(g_peer[i], (const char *)sendbuf + (size_t)displs[i] * w, (size_t)sendcounts[i] * w);
    memcpy(recvbuf, (const char *)sendbuf + (size_t)displs[0] * w, (size_t)recvcount * w);
  } else
    recv_all(g_peer[0], recvbuf, (size_t)recvcount * w);
  return MPI_SUCCESS;
}

int MPI_Gatherv(const void *sendbuf, int sendcount, MPI_Datatype sendtype,
                void *recvbuf, const int *recvcounts, const int *displs, MPI_Datatype recvtype,
                int root, MPI_Comm comm)
{
  const size_t w = type_bytes(sendtype);
  int i;
  (void)recvtype; (void)comm;
  if (root != 0) { fprintf(stderr, "mpi_tcp: Gatherv root must be 0\n"); exit(70); }
  if (g_rank == 0) {
    memcpy((char *)recvbuf + (size_t)displs[0] * w, sendbuf, (size_t)sendcount * w);
    for (i = 1; i < g_size; i++)
      recv_all(g_peer[i], (char *)recvbuf + (size_t)displs[i] * w, (size_t)recvcounts[i] * w);
  } else
    send_all(g_peer[0], sendbuf, (size_t)sendcount * w);
  return MPI_SUCCESS;
}
