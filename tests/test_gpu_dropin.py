"""The drop-in proof: the reference's UNCHANGED host code (axml.c, searchAlgo.c, optimizeModel.c, models.c,
byteFile.c, ... compiled from /root/reference) linked with examl_b200/csrc/axml_backend.c and the CUDA
engine runs ExaML's own workflows on the GPU and lands on the reference's likelihoods.

north_star bar: lnL within 1e-6 relative of the reference's AVX build on testData/49 and testData/140."""
import tempfile

import pytest

from dropin_util import have_binaries, run_dropin, write_inputs
from golden_util import load_json

pytestmark = pytest.mark.gpu
NORTH_STAR_RTOL = 1e-6


def _rel(a, b):
    return abs(a - b) / abs(b)


@pytest.mark.skipif(not have_binaries(), reason="bin/examl-b200-probe / oracle/_ref/parse-examl not built")
@pytest.mark.parametrize("tag,lines,extra", [
    ("dna49", None, []),
    ("dna49p", "from-fixture", []),
    ("dna49pM", "from-fixture", ["-M"]),
    ("aa60", ["LG, p0 = 1-300"], []),
])
def test_tree_evaluation_matches_reference(tag, lines, extra):
    """`-f e`: model + branch-length optimisation on a fixed tree (treeEvaluate, modOpt: thousands of
    evaluateGeneric / makenewzGeneric / newviewGeneric calls through the back-end)."""
    first_gold = load_json(f"{tag}.first.json")
    final_gold = load_json(f"{tag}.final.json")
    if lines == "from-fixture":
        lo, lines = 1, []
        import numpy as np
        for i, part in enumerate(first_gold["partitions"]):
            n = int(np.sum(part["wgt"]))
            lines.append(f"DNA, p{i} = {lo}-{lo + n - 1}")
            lo += n
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree, _ = write_inputs(tag, tmp, lines)
        r, dumps = run_dropin(binary, tree, tmp, ["-f", "e"] + extra)
        assert r.returncode == 0, r.stdout[-3000:]
        assert _rel(dumps["first"]["likelihood"], first_gold["likelihood"]) < 1e-11, (dumps["first"]["likelihood"], first_gold["likelihood"])
        got, want = dumps["final"]["likelihood"], final_gold["likelihood"]
        assert _rel(got, want) < NORTH_STAR_RTOL, (got, want)
        print(f"{tag}: drop-in final lnL {got!r} vs reference {want!r} (rel {_rel(got, want):.2e})")


@pytest.mark.skipif(not have_binaries(), reason="bin/examl-b200-probe / oracle/_ref/parse-examl not built")
def test_aa140_lg_tree_evaluation():
    """testData/140 (protein, LG): the other north_star data set."""
    first_gold = load_json("aa140.first.json")
    final_gold = load_json("aa140.final.json")
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree, _ = write_inputs("aa140", tmp, ["LG, p0 = 1-%d" % sum(first_gold["partitions"][0]["wgt"])])
        r, dumps = run_dropin(binary, tree, tmp, ["-f", "e"])
        assert r.returncode == 0, r.stdout[-3000:]
        assert _rel(dumps["first"]["likelihood"], first_gold["likelihood"]) < 1e-11
        assert _rel(dumps["final"]["likelihood"], final_gold["likelihood"]) < NORTH_STAR_RTOL, \
            (dumps["final"]["likelihood"], final_gold["likelihood"])


@pytest.mark.skipif(not have_binaries(), reason="bin/examl-b200-probe / oracle/_ref/parse-examl not built")
def test_spr_search_runs_and_improves():
    """Default mode (-f d): lazy-SPR tree search from testData/49.tree -- newviewGeneric / makenewzGeneric /
    evaluateGeneric(partial) in the access pattern of searchAlgo.c.  The search is chaotic in the last bits
    (a 1e-13 difference can flip an SPR acceptance), so the bar is: it terminates, and ends at least as
    good as where the reference's AVX build ended (-16221.368990) minus 1e-6 relative."""
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree, _ = write_inputs("dna49", tmp)
        r, dumps = run_dropin(binary, tree, tmp, ["-f", "d"])
        assert r.returncode == 0, r.stdout[-3000:]
        final = dumps["final"]["likelihood"]
        assert final >= -16221.368990 * (1 + NORTH_STAR_RTOL), final
        print("search final lnL", final)
