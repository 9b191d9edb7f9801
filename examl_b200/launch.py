"""Start one process per GPU for binaries linked against mpi_shim/mpi_tcp.c (no MPI in the image):

    python -m examl_b200.launch -np 8 -- bin/examl-b200 -s aln.binary -t start.tree -m GAMMA -n RUN

Each rank gets RANK, WORLD_SIZE, LOCAL_RANK, MASTER_ADDR=127.0.0.1 and a free MASTER_PORT -- the same
variables torchrun exports.  The first non-zero exit code is returned after the remaining ranks were
terminated (by PID, never by pattern).
"""
from __future__ import annotations

import argparse
import os
import socket
import subprocess
import sys
import time


def free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(np_: int, cmd, env=None, stdout=None, timeout=None) -> int:
    port = free_port()
    token = "%016x" % int.from_bytes(os.urandom(8), "little")      # mpi_tcp.c: rank 0 only accepts ranks that echo it
    procs = []
    for r in range(np_):
        e = dict(os.environ if env is None else env)
        e.update(RANK=str(r), WORLD_SIZE=str(np_), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                 MPI_TCP_TOKEN=token)
        out = stdout if (stdout is not None and r == 0) else (subprocess.DEVNULL if stdout is not None else None)
        procs.append(subprocess.Popen(cmd, env=e, stdout=out, stderr=subprocess.STDOUT if out is not None else None))
    t0 = time.time()
    rc = 0
    while procs:
        for p in list(procs):
            code = p.poll()
            if code is None:
                continue
            procs.remove(p)
            if code != 0 and rc == 0:
                rc = code
                for q in procs:
                    q.terminate()
        if timeout and time.time() - t0 > timeout:
            for q in procs:
                q.kill()
            return 124
        time.sleep(0.02)
    return rc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-np", type=int, required=True)
    ap.add_argument("cmd", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    cmd = a.cmd[1:] if a.cmd and a.cmd[0] == "--" else a.cmd
    sys.exit(launch(a.np, cmd))


if __name__ == "__main__":
    main()
