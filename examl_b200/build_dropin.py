"""Build the drop-in binary: the reference's UNCHANGED host code + the B200 back-end.

    bin/examl-b200         reference objects (axml.c, searchAlgo.c, optimizeModel.c, models.c, byteFile.c, ...)
                           + examl_b200/csrc/axml_backend.c + mpi shim, linked against libexaml_b200.so
    bin/examl-b200-probe   the same with oracle/ref_probe.c attached through ld --wrap (golden dumps and
                           the evaluateGeneric timing loop; used by tests and bench only)
    bin/examl-b200-spr     examl-b200 with ONE reference file changed by the four-line patch of INTEGRATION.md:
    bin/examl-b200-spr-probe   rearrangeBIG (searchAlgo.c:800) brackets its addTraverseBIG calls with
                           batchedInsertionScanBegin / End, so that the lazy insertions of a pruned subtree are scored
                           by one exb_score_insertions call instead of two calls per candidate.  The patch
                           is applied to a build-time copy under oracle/_ref/obj/ (git-ignored); nothing of the
                           reference is stored in the repository.

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


PROTOTYPES = ("extern void batchedInsertionScanBegin(tree *tr, nodeptr p, nodeptr p1, nodeptr p2, int mintrav, int maxtrav);\n"
              "extern void batchedInsertionScanEnd(tree *tr);\n")
# (anchor inside rearrangeBIG, text inserted, before/after) -- the whole maintainer patch
SPR_PATCH = [
    ("if (! removeNodeBIG(tr, p,  tr->numBranches)) return badRear;", "batchedInsertionScanBegin(tr, p, p1, p2, mintrav, maxtrav);", "after"),
    ("hookup(p->next,       p1, p1z, tr->numBranches);", "batchedInsertionScanEnd(tr);", "before"),
    ("mintrav2 = mintrav > 2 ? mintrav : 2;", "batchedInsertionScanBegin(tr, q, q1, q2, mintrav2, maxtrav);", "after"),
    ("hookup(q->next,       q1, q1z, tr->numBranches);", "batchedInsertionScanEnd(tr);", "before"),
]


def patched_search_object() -> str:
    """searchAlgo.o with the batched-insertion hooks: the reference file is read where it lies, the four calls are
    inserted into rearrangeBIG in a build-time copy (oracle/_ref/obj/spr/, git-ignored) and compiled with the flags of
    the other reference objects."""
    src = open(os.path.join(REF, "examl", "searchAlgo.c")).read()
    a, b = src.index("int rearrangeBIG("), src.index("double treeOptimizeRapid(")
    body = src[a:b]
    for anchor, text, where in SPR_PATCH:
        assert body.count(anchor) == 1, f"patch anchor not found exactly once in rearrangeBIG: {anchor}"
        body = body.replace(anchor, anchor + "\n\t  " + text if where == "after" else text + "\n\t  " + anchor)
    out_dir = os.path.join(OBJ, "spr")
    os.makedirs(out_dir, exist_ok=True)
    c = os.path.join(out_dir, "searchAlgo.c")
    first_fn = src.index("boolean initrav")
    open(c, "w").write(src[:first_fn] + PROTOTYPES + src[first_fn:a] + body + src[b:])
    o = os.path.join(out_dir, "searchAlgo.o")
    subprocess.check_call(["gcc"] + CFLAGS + ["-I" + os.path.join(HERE, "csrc", "mpi_shim"), "-I" + os.path.join(REF, "examl"),
                           "-c", "-o", o, c])
    return o


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
    # the same two binaries with the batched-insertion hooks in rearrangeBIG
    spr_o = patched_search_object()
    objs_spr = [spr_o if o.endswith(os.sep + "searchAlgo.o") else o for o in objs]
    subprocess.check_call(["gcc", "-o", exe + "-spr"] + objs_spr + link)
    subprocess.check_call(["gcc", "-o", exe + "-spr-probe"] + objs_spr + [os.path.join(OBJ, "avx", "ref_probe.o")] +
                          ["-Wl,--wrap=evaluateGeneric", "-Wl,--wrap=makenewzGeneric", "-Wl,--wrap=evaluatePartialGeneric"] + link)
    return True


if __name__ == "__main__":
    print("built" if build(True) else "reference not available")
