"""TEST INFRASTRUCTURE.  Build tests/hostcheck/_build/examl-hostcheck: the reference's unchanged host objects
(oracle/_ref/obj, compiled from /root/reference by oracle/Makefile) + examl_b200/csrc/axml_backend.c + the
C ABI implemented on the CPU oracle (exb_on_oracle.c) + mpi_tcp.c + the probe.  Same link line as
examl_b200/build_dropin.py with libexaml_b200.so swapped for the oracle: it lets the CPU test suite drive the
drop-in's host logic.  Only possible where /root/reference exists; elsewhere a prebuilt binary is used if present."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from examl_b200.build_dropin import CFLAGS, KEEP, OBJ, REF   # noqa: E402  (the product's own object list and flags)

OUT = os.path.join(HERE, "_build")
EXE = os.path.join(OUT, "examl-hostcheck")
EXE_SPR = os.path.join(OUT, "examl-hostcheck-spr")     # searchAlgo.c with the batched-insertion hooks


def build(force: bool = False) -> str | None:
    if not os.path.exists(os.path.join(REF, "examl", "axml.c")):
        return EXE if os.path.exists(EXE) else None
    srcs = [os.path.join(ROOT, "examl_b200", "csrc", "axml_backend.c"), os.path.join(HERE, "exb_on_oracle.c"),
            os.path.join(ROOT, "oracle", "likelihood_oracle.c"), os.path.join(ROOT, "oracle", "ref_probe.c"), __file__]
    srcs.append(os.path.join(ROOT, "examl_b200", "build_dropin.py"))
    if not force and os.path.exists(EXE) and os.path.exists(EXE_SPR) and all(os.path.getmtime(EXE_SPR) >= os.path.getmtime(s) for s in srcs):
        return EXE
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", "-j8"])
    os.makedirs(OUT, exist_ok=True)
    shim = os.path.join(ROOT, "examl_b200", "csrc", "mpi_shim")
    inc = ["-I" + shim, "-I" + os.path.join(REF, "examl"), "-I" + os.path.join(ROOT, "include")]
    objs = []
    for src, flags in ((srcs[0], CFLAGS), (srcs[1], ["-O2"]), (srcs[2], ["-O2", "-ffp-contract=off"])):
        o = os.path.join(OUT, os.path.basename(src)[:-2] + ".o")
        subprocess.check_call(["gcc"] + flags + inc + ["-c", "-o", o, src])
        objs.append(o)
    objs += [os.path.join(OBJ, "avx", k + ".o") for k in KEEP] + [os.path.join(OBJ, "mpi_tcp.o"), os.path.join(OBJ, "avx", "ref_probe.o")]
    wrap = ["-Wl,--wrap=evaluateGeneric", "-Wl,--wrap=makenewzGeneric", "-Wl,--wrap=evaluatePartialGeneric", "-lm"]
    subprocess.check_call(["gcc", "-o", EXE] + objs + wrap)
    # the same harness with the batched-insertion hooks in rearrangeBIG (examl_b200/build_dropin.py: the maintainer patch)
    from examl_b200.build_dropin import patched_search_object
    spr_o = patched_search_object()
    subprocess.check_call(["gcc", "-o", EXE_SPR] + [spr_o if o.endswith(os.sep + "searchAlgo.o") else o for o in objs] + wrap)
    return EXE


if __name__ == "__main__":
    print(build(True))
