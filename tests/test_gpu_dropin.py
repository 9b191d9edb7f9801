"""The drop-in proof: the reference's UNCHANGED host code (axml.c, searchAlgo.c, optimizeModel.c, models.c,
byteFile.c, ... compiled from /root/reference) linked with examl_b200/csrc/axml_backend.c and the CUDA
engine runs ExaML's own workflows on the GPU and lands on the reference's likelihoods.

north_star bar: lnL within 1e-6 relative of the reference's AVX build on testData/49 and testData/140."""
import tempfile

import pytest

from dropin_util import DROPIN_PLAIN, REFERENCE, have_binaries, run_dropin, run_mode, small_dna_case, write_inputs
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
    ("dna49", None, ["-S"]),            # the reference's own -S run ends bit-identical to its dense run (dna49.final)
    ("aa60", ["LG, p0 = 1-300"], []),
    ("aa60lg4x", ["LG4X, p0 = 1-300"], []),
    ("aa60lg4m", ["LG4M, p0 = 1-300"], []),
    ("dna49psr", None, ["PSR"]),
    ("dna49ppsr", "from-fixture", ["PSR"]),
    ("aa60psr", ["LG, p0 = 1-300"], ["PSR"]),
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
        psr = "PSR" in extra
        r, dumps = run_dropin(binary, tree, tmp, ["-f", "e"] + [x for x in extra if x != "PSR"], model="PSR" if psr else "GAMMA")
        assert r.returncode == 0, r.stdout[-3000:]
        assert _rel(dumps["first"]["likelihood"], first_gold["likelihood"]) < 1e-11, (dumps["first"]["likelihood"], first_gold["likelihood"])
        got, want = dumps["final"]["likelihood"], final_gold["likelihood"]
        assert _rel(got, want) < NORTH_STAR_RTOL, (got, want)
        print(f"{tag}: drop-in final lnL {got!r} vs reference {want!r} (rel {_rel(got, want):.2e})")
        if psr:
            # the per-site rate scan: optimizeModel.c asked the back-end the very questions it asked the reference
            # (same sites, same rates, same stopping points) and got the same answers
            import numpy as np
            a, b = np.array(dumps["partial"]["partial"]), np.array(load_json(f"{tag}.partial.json")["partial"])
            assert a.shape == b.shape and np.array_equal(a[:, :3], b[:, :3]), "the rate scan took a different path"
            assert np.all(np.abs(a[:, 3] - b[:, 3]) <= 1e-9 * np.maximum(1.0, np.abs(b[:, 3])))


@pytest.mark.skipif(not have_binaries(), reason="bin/examl-b200-probe / oracle/_ref/parse-examl not built")
def test_host_newton_raphson_path():
    """The multi-rank form of makenewzGeneric (host state machine, exb_makenewz_prepare + one exb_core per iteration)
    on a single GPU: EXAML_B200_HOST_NR=1 selects it."""
    final_gold = load_json("dna49.final.json")
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree, _ = write_inputs("dna49", tmp)
        r, dumps = run_dropin(binary, tree, tmp, ["-f", "e"], env={"EXAML_B200_HOST_NR": "1"})
        assert r.returncode == 0, r.stdout[-3000:]
        assert _rel(dumps["final"]["likelihood"], final_gold["likelihood"]) < NORTH_STAR_RTOL


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


@pytest.mark.skipif(not have_binaries(), reason="bin/examl-b200-probe / oracle/_ref/parse-examl not built")
def test_two_ranks_two_gpus():
    """One process per GPU: ExaML's own partitionAssignment.c shards the sites, byteFile.c reads each rank's slice,
    the back-end uploads it to the rank's GPU; control-plane collectives over mpi_tcp.c, the two hot-path
    all-reduces over NCCL.  Same end state as the single-rank reference."""
    import os, subprocess, sys, json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from dropin_util import DROPIN, ROOT
    final_gold = load_json("dna49p.final.json")
    first_gold = load_json("dna49p.first.json")
    lo, lines = 1, []
    for i, part in enumerate(first_gold["partitions"]):
        n = sum(part["wgt"])
        lines.append(f"DNA, p{i} = {lo}-{lo + n - 1}")
        lo += n
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree, _ = write_inputs("dna49p", tmp, lines)
        env = dict(os.environ, EXAML_PROBE_DUMP=os.path.join(tmp, "dump"), EXAML_PROBE_NOTIPS="1", PYTHONPATH=ROOT)
        r = subprocess.run([sys.executable, "-m", "examl_b200.launch", "-np", "2", "--", DROPIN, "-s", binary, "-t", tree,
                            "-m", "GAMMA", "-n", "R", "-f", "e", "-w", tmp + "/"],
                           cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:]
        got = json.load(open(os.path.join(tmp, "dump.final.json")))["likelihood"]
        assert _rel(got, final_gold["likelihood"]) < NORTH_STAR_RTOL, (got, final_gold["likelihood"])
        print("2 ranks / 2 GPUs final lnL", got, "reference", final_gold["likelihood"])


@pytest.mark.skipif(not have_binaries(), reason="bin/examl-b200 / oracle/_ref not built")
@pytest.mark.parametrize("args", [["-m", "GAMMA", "-f", "d"], ["-m", "GAMMA", "-f", "o"], ["-m", "PSR", "-f", "d"],
                                  ["-m", "GAMMA", "-f", "q"], ["-m", "GAMMA", "-f", "d", "-M"]])
def test_search_and_quartet_workflows(args):
    """-f d / -f o searches, PSR search, -f q quartets, -M search on a small seeded alignment: the reference binary
    (CPU, run beside it) and the drop-in on the GPU must agree."""
    import numpy as np
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree = small_dna_case(tmp)
        r0, best0, q0 = run_mode(REFERENCE, binary, tree, tmp, "REF", args)
        r1, best1, q1 = run_mode(DROPIN_PLAIN, binary, tree, tmp, "GPU", args)
        assert r0.returncode == 0 and r1.returncode == 0, (r0.stdout[-1500:], r1.stdout[-1500:])
        if "q" in args:
            a, b = np.array(q0), np.array(q1)
            assert a.shape == b.shape and a.shape[0] == 630
            assert np.all(np.abs(a - b) <= 2e-6 * np.abs(a))
        else:
            assert best0 is not None and best1 is not None
            assert _rel(best1, best0) < NORTH_STAR_RTOL, (best1, best0)


@pytest.mark.skipif(not have_binaries(), reason="bin/examl-b200-spr / oracle/_ref/parse-examl not built")
@pytest.mark.parametrize("case", ["dna49", "small", "small-M", "aa"])
def test_batched_insertion_scoring_behind_the_search(case):
    """SURVEY 8 f-4 on the GPU: bin/examl-b200-spr = the drop-in with the maintainer patch in rearrangeBIG.
    Every pruned subtree's lazy insertions are scored by one exb_score_insertions call; testInsertBIG is answered from the
    table.  The search ends where the reference's does (testData/49: -16221.368990)."""
    import os, re, time
    from dropin_util import ROOT
    exe = os.path.join(ROOT, "bin", "examl-b200-spr")
    if not os.path.exists(exe):
        pytest.skip("bin/examl-b200-spr not built")
    env = {"EXAML_B200_SPR_STATS": "1"}
    with tempfile.TemporaryDirectory() as tmp:
        if case == "dna49":
            binary, tree, _ = write_inputs("dna49", tmp)
            args, best0 = ["-m", "GAMMA", "-f", "d"], -16221.368990
        elif case == "aa":
            from dropin_util import small_aa_case
            binary, tree = small_aa_case(tmp, ["LG, p0 = 1-120"], taxa=12)
            args = ["-m", "GAMMA", "-f", "d"]
            r0, best0, _ = run_mode(REFERENCE, binary, tree, tmp, "REF", args)
            assert r0.returncode == 0
        else:
            binary, tree = small_dna_case(tmp, taxa=20, sites=600, seed=13)
            args = ["-m", "GAMMA", "-f", "d"] + (["-M"] if case == "small-M" else [])
            r0, best0, _ = run_mode(REFERENCE, binary, tree, tmp, "REF", args)
            assert r0.returncode == 0
        t0 = time.time()
        r1, best1, _ = run_mode(exe, binary, tree, tmp, "SPR", args, env=env)
        dt = time.time() - t0
        assert r1.returncode == 0, r1.stdout[-3000:]
        assert best1 >= best0 * (1 + NORTH_STAR_RTOL), (best1, best0)
        m = re.search(r"batched insertion scoring: (\d+) scans, (\d+) engine calls, (\d+) candidates scored, (\d+) served from the table, (\d+) answered", r1.stdout)
        assert m, r1.stdout[-800:]
        scans, calls, scored, served, flushed = map(int, m.groups())
        assert scans > 10 and served > scans and flushed == 0, m.group(0)
        print(case, "search with batched insertion scoring: %.1f s, final lnL" % dt, best1, "reference", best0, "|", m.group(0))
