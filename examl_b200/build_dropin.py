"""Build the drop-in binary: the reference's UNCHANGED host code + the B200 back-end.

    bin/examl-b200         reference objects (axml.c, searchAlgo.c, optimizeModel.c, models.c, byteFile.c, ...)
                           + examl_b200/csrc/axml_backend.c + mpi shim, linked against libexaml_b200.so
    bin/examl-b200-probe   the same with oracle/ref_probe.c attached through ld --wrap (golden dumps and
                           the evaluateGeneric timing loop; used by tests and bench only)

The five objects newviewGenericSpecial.o, evaluateGenericSpecial.o, makenewzGenericSpecial.o,
evaluatePartialGenericSpecial.o and avxLikelihood.o of the reference are NOT linked: axml_backend.c provides their entry points.
Only possible where /root/reference exists (the reference objects are compiled there by
oracle/Makefile into oracle/_ref/obj/); elsewhere the prebuilt binaries are used as they are.
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OBJ = os.path.join(ROOT, "oracle", "_ref", "obj")
BIN = os.path.join(ROOT, "bin")
KEEP = ["axml", "optimizeModel", "trash", "searchAlgo", "topologies", "treeIO",
        "bipartitionList", "restartHashTable", "byteFile", "partitionAssignment", "communication", "quartets", "models"]
CFLAGS = ["-D__SIM_SSE3", "-D__AVX", "-D_OPTIMIZED_FUNCTIONS", "-msse3", "-mavx", "-D_GNU_SOURCE", "-D_USE_ALLREDUCE",
          "-O2", "-w"]


def build(force: bool = False) -> bool:
    exe = os.path.join(BIN, "examl-b200")
    if not os.path.exists(os.path.join(REF, "examl", "axml.c")):
        return os.path.exists(exe)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", "-j8"])
    os.makedirs(BIN, exist_ok=True)
    shim = os.path.join(HERE, "csrc", "mpi_shim")
    backend_o = os.path.join(OBJ, "axml_backend.o")
    subprocess.check_call(["gcc"] + CFLAGS + ["-I" + shim, "-I" + os.path.join(REF, "examl"), "-I" + os.path.join(ROOT, "include"),
                           "-c", "-o", backend_o, os.path.join(HERE, "csrc", "axml_backend.c")])
    # mpi_tcp.c: rank/size from the launcher's environment (examl_b200/launch.py, torchrun --no-python);
    # without WORLD_SIZE it is a single rank
    objs = [os.path.join(OBJ, "avx", k + ".o") for k in KEEP] + [backend_o, os.path.join(OBJ, "mpi_tcp.o")]
    link = ["-L" + HERE, "-lexaml_b200", "-Wl,-rpath,$ORIGIN/../examl_b200", "-lm"]
    subprocess.check_call(["gcc", "-o", exe] + objs + link)
    subprocess.check_call(["gcc", "-o", exe + "-probe"] + objs + [os.path.join(OBJ, "avx", "ref_probe.o")] +
                          ["-Wl,--wrap=evaluateGeneric", "-Wl,--wrap=makenewzGeneric", "-Wl,--wrap=evaluatePartialGeneric"] + link)
    return True


if __name__ == "__main__":
    print("built" if build(True) else "reference not available")
