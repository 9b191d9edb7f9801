"""TEST INFRASTRUCTURE: ctypes loader for the CPU oracle (oracle/likelihood_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may import
this module.  It binds the oracle through examl_b200.capi.Engine with prefix `orc_`, so the
parity tests drive the CUDA engine and the oracle with byte-identical arguments.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "likelihood_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return LIB


def build_ref() -> bool:
    """Compile the unmodified reference into oracle/_ref (only where /root/reference exists)."""
    if not os.path.exists("/root/reference/examl/axml.c"):
        return os.path.exists(os.path.join(REF_DIR, "examl-AVX"))
    subprocess.check_call(["make", "-s", "-C", HERE, "ref", "-j8"])
    return True


def load() -> ctypes.CDLL:
    return ctypes.CDLL(build())


def OracleEngine(mxtips: int, num_partitions: int, num_branches: int = 1):
    from examl_b200.capi import Engine
    return Engine(mxtips, num_partitions, num_branches, lib=load(), prefix="orc_")
