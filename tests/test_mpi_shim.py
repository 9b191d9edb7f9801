"""mpi_shim/mpi_tcp.c against ExaML's own MPI usage: the UNMODIFIED reference (oracle/_ref/examl-AVX-mp) on
1, 2 and 3 CPU ranks -- sites assigned by the reference's partitionAssignment.c, alignment slices read by its
byteFile.c, collectives carried by the shim -- must end `-f e` on the same likelihood as the golden single-rank
run.  (CPU only; the same shim carries the control plane of the multi-GPU drop-in binary.)"""
import json
import os
import subprocess
import sys
import tempfile

import pytest

from dropin_util import PARSER, ROOT, write_inputs
from golden_util import load_json

EXE = os.path.join(ROOT, "oracle", "_ref", "examl-AVX-mp")


@pytest.mark.skipif(not (os.path.exists(EXE) and os.path.exists(PARSER)), reason="oracle/_ref not built")
@pytest.mark.parametrize("ranks,tag,lines", [(1, "dna49", None), (2, "dna49", None), (3, "dna49p", "from-fixture")])
def test_reference_on_n_ranks_over_tcp_shim(ranks, tag, lines):
    gold = load_json(f"{tag}.final.json")
    first = load_json(f"{tag}.first.json")
    if lines == "from-fixture":
        lo, lines = 1, []
        for i, part in enumerate(first["partitions"]):
            n = sum(part["wgt"])
            lines.append(f"DNA, p{i} = {lo}-{lo + n - 1}")
            lo += n
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree, _ = write_inputs(tag, tmp, lines)
        env = dict(os.environ, EXAML_PROBE_DUMP=os.path.join(tmp, "dump"), EXAML_PROBE_NOTIPS="1", PYTHONPATH=ROOT)
        r = subprocess.run([sys.executable, "-m", "examl_b200.launch", "-np", str(ranks), "--", EXE, "-s", binary, "-t", tree,
                            "-m", "GAMMA", "-n", "R", "-f", "e", "-w", tmp + "/"],
                           cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:]
        got = json.load(open(os.path.join(tmp, "dump.final.json")))["likelihood"]     # every rank writes the same value
        assert abs(got - gold["likelihood"]) <= 1e-9 * abs(gold["likelihood"]), (got, gold["likelihood"])
