"""CPU coverage of the drop-in's HOST logic (examl_b200/csrc/axml_backend.c): the reference's unchanged host
objects + the back-end, linked against the C ABI implemented on the CPU oracle (tests/hostcheck/exb_on_oracle.c)
instead of libexaml_b200.so.  Runs ExaML's own workflows without a GPU and must land on the reference's results:
traversal descriptors and x flags, the host Newton-Raphson loop, -M masks, the anticipated PSR rate scan, and the
multi-rank plumbing over mpi_tcp.c.  (The same workflows run on the GPU in tests/test_gpu_dropin.py.)"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from dropin_util import PARSER, REFERENCE, ROOT, run_dropin, run_mode, small_aa_case, small_dna_case, write_inputs
from golden_util import load_json

sys.path.insert(0, os.path.join(ROOT, "tests", "hostcheck"))
import build as hostcheck_build  # noqa: E402

EXE = hostcheck_build.build()
pytestmark = pytest.mark.skipif(EXE is None or not os.path.exists(PARSER),
                                reason="tests/hostcheck/_build/examl-hostcheck or oracle/_ref/parse-examl not built")
# every workflow runs with the constants guard on: a model change that is not followed by a full traversal would abort
ENV = {"EXAML_B200_HOST_NR": "1", "EXAML_B200_CHECK_CONSTANTS": "1"}


def _rel(a, b):
    return abs(a - b) / abs(b)


def _partition_lines(first):
    lo, lines = 1, []
    for i, part in enumerate(first["partitions"]):
        n = int(np.sum(part["wgt"]))
        lines.append(f"DNA, p{i} = {lo}-{lo + n - 1}")
        lo += n
    return lines


@pytest.mark.parametrize("tag,partitioned,extra,model", [
    ("dna49", False, [], "GAMMA"),
    ("dna49pM", True, ["-M"], "GAMMA"),
    ("dna49psr", False, [], "PSR"),
    ("dna49ppsr", True, [], "PSR"),
])
def test_tree_evaluation_host_logic(tag, partitioned, extra, model):
    first_gold, final_gold = load_json(f"{tag}.first.json"), load_json(f"{tag}.final.json")
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree, _ = write_inputs(tag, tmp, _partition_lines(first_gold) if partitioned else None)
        r, dumps = run_dropin(binary, tree, tmp, ["-f", "e"] + extra, env=ENV, exe=EXE, model=model)
        assert r.returncode == 0, r.stdout[-3000:]
        assert _rel(dumps["first"]["likelihood"], first_gold["likelihood"]) < 1e-11
        assert _rel(dumps["final"]["likelihood"], final_gold["likelihood"]) < 1e-6, (dumps["final"]["likelihood"], final_gold["likelihood"])
        if model == "PSR":
            # optimizeRateCategories asked the back-end exactly what it asked the reference, and stopped where it stopped
            a, b = np.array(dumps["partial"]["partial"]), np.array(load_json(f"{tag}.partial.json")["partial"])
            assert a.shape == b.shape and np.array_equal(a[:, :3], b[:, :3])
            assert np.all(np.abs(a[:, 3] - b[:, 3]) <= 1e-9 * np.maximum(1.0, np.abs(b[:, 3])))


def test_two_ranks_host_logic():
    """ExaML's own site sharding over two ranks (mpi_tcp.c), PSR: per-rank rate scans, gathered and re-categorised on rank 0."""
    tag = "dna49ppsr"
    first_gold, final_gold = load_json(f"{tag}.first.json"), load_json(f"{tag}.final.json")
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree, _ = write_inputs(tag, tmp, _partition_lines(first_gold))
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
        env.update(ENV, EXAML_PROBE_DUMP=os.path.join(tmp, "dump"), EXAML_PROBE_NOTIPS="1", PYTHONPATH=ROOT)
        r = subprocess.run([sys.executable, "-m", "examl_b200.launch", "-np", "2", "--", EXE, "-s", binary, "-t", tree,
                            "-m", "PSR", "-n", "R", "-f", "e", "-w", tmp + "/"],
                           cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-3000:]
        got = json.load(open(os.path.join(tmp, "dump.final.json")))["likelihood"]
        assert _rel(got, final_gold["likelihood"]) < 1e-6, (got, final_gold["likelihood"])


@pytest.mark.parametrize("args", [["-m", "GAMMA", "-f", "d"], ["-m", "GAMMA", "-f", "o"], ["-m", "PSR", "-f", "d"],
                                  ["-m", "GAMMA", "-f", "q"], ["-m", "GAMMA", "-f", "d", "-M"]])
def test_search_and_quartet_workflows(args):
    """The other workflows of axml.c (-f d new rapid hill climbing, -f o the old search, -f q quartets; PSR search,
    -M) on a small seeded alignment: the reference binary and the back-end behind the same host code must agree."""
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree = small_dna_case(tmp)
        r0, best0, q0 = run_mode(REFERENCE, binary, tree, tmp, "REF", args)
        r1, best1, q1 = run_mode(EXE, binary, tree, tmp, "HC", args, env=ENV)
        assert r0.returncode == 0 and r1.returncode == 0, (r0.stdout[-1500:], r1.stdout[-1500:])
        if "q" in args:
            a, b = np.array(q0), np.array(q1)
            assert a.shape == b.shape and a.shape[0] == 630            # 210 quartets x 3 topologies
            assert np.all(np.abs(a - b) <= 2e-6 * np.abs(a))           # the file prints 6 decimals
        else:
            assert best0 is not None and best1 is not None
            assert _rel(best1, best0) < 1e-6, (best1, best0)


@pytest.mark.parametrize("args,taxa", [(["-m", "GAMMA", "-f", "d"], 24), (["-m", "GAMMA", "-f", "d", "-M"], 12), (["-m", "PSR", "-f", "d"], 12)])
def test_batched_insertion_scan_behind_the_search(args, taxa):
    """rearrangeBIG with the maintainer patch (batchedInsertionScanBegin / End around its addTraverseBIG
    calls, examl_b200/build_dropin.py SPR_PATCH).  The lazy insertions of every pruned subtree are scored by ONE
    exb_score_insertions call and testInsertBIG's newviewGeneric / evaluateGeneric pairs are answered from that table; the
    search must end where the reference's does.  Under PSR the engine has no batched scoring and the same binary takes
    the ordinary path."""
    import re
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree = small_dna_case(tmp, taxa=taxa, sites=400, seed=21)
        r0, best0, _ = run_mode(REFERENCE, binary, tree, tmp, "REF", args)
        r1, best1, _ = run_mode(hostcheck_build.EXE_SPR, binary, tree, tmp, "SPR", args, env=dict(ENV, EXAML_B200_SPR_STATS="1"))
        assert r0.returncode == 0 and r1.returncode == 0, (r0.stdout[-1500:], r1.stdout[-1500:])
        assert _rel(best1, best0) < 1e-6, (best1, best0)
        m = re.search(r"batched insertion scoring: (\d+) scans, (\d+) engine calls, (\d+) candidates scored, (\d+) served from the table, (\d+) answered", r1.stdout)
        assert m, r1.stdout[-800:]
        scans, calls, scored, served, flushed = map(int, m.groups())
        if "PSR" in args:
            assert scans == 0 and served == 0
        else:
            assert scans > 10 and calls == scans and served > scans and served <= scored and flushed == 0, m.group(0)


@pytest.mark.parametrize("case", ["dna49", "aa16"])
@pytest.mark.parametrize("which", ["plain", "batched-scan"])
def test_search_reproduces_the_references_evaluations(which, case):
    """tests/golden/<case>.search.trace.gz holds what the REFERENCE's own binary saw in tr->likelihood after each of its
    first evaluateGeneric calls of `-f d` (6000 on testData/49; 3000 on a seeded 16-taxon protein set under LG) -- model optimisation, then the lazy SPR insertions scored one by one by testInsertBIG
    (searchAlgo.c:682-735) and the evaluations in between.  The drop-in's host logic must issue the same evaluations in the
    same order and see the same numbers -- also when the insertions are answered from the table of the batched scan
    (exb_score_insertions), which is thereby pinned to the reference's own values."""
    import gzip
    import shutil
    import time
    gold = [l.split() for l in gzip.open(os.path.join(ROOT, "tests", "golden", case + ".search.trace.gz"), "rt")]
    exe = EXE if which == "plain" else hostcheck_build.EXE_SPR
    with tempfile.TemporaryDirectory() as tmp:
        extra = []
        if case == "aa16":
            for f in ("aa16.phy", "aa16.tree"):
                shutil.copy(os.path.join(ROOT, "tests", "golden", f), tmp)
            open(os.path.join(tmp, "lg.model"), "w").write("LG, p0 = 1-200\n")
            subprocess.run([PARSER, "-s", "aa16.phy", "-n", "aa16", "-m", "PROT", "-q", "lg.model"], cwd=tmp, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
            binary, tree = os.path.join(tmp, "aa16.binary"), os.path.join(tmp, "aa16.tree")
        else:
            binary, tree, _ = write_inputs(case, tmp)
        tf = os.path.join(tmp, "trace")
        # the probe traces the first evaluations; the run is cut short afterwards (whole searches are covered elsewhere)
        e = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
        e.update(ENV, EXAML_PROBE_TRACE=tf, EXAML_PROBE_TRACE_MAX=str(len(gold)))
        p = subprocess.Popen([exe, "-s", binary, "-t", tree, "-m", "GAMMA", "-n", "TR", "-f", "d", "-w", tmp + "/"] + extra, cwd=tmp, env=e,
                             stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
        t0 = time.time()
        while p.poll() is None and time.time() - t0 < 900:
            if os.path.exists(tf) and sum(1 for _ in open(tf)) >= len(gold):
                p.terminate()
                break
            time.sleep(0.5)
        p.wait(timeout=60)
        got = [l.split() for l in open(tf)][:len(gold)]
    assert len(got) == len(gold), (len(got), len(gold))
    nins = 0
    for i, (a, b) in enumerate(zip(gold, got)):
        assert a[:4] == b[:4], (i, a, b)                                        # same branch, same kind of traversal
        assert abs(float(a[4]) - float(b[4])) <= 1e-8 * abs(float(a[4])), (i, a, b)
        nins += a[3] == "0"
    assert nins > len(gold) // 2      # most of them partial evaluations: the insertions and what happens between them


def test_deferred_newviews_change_the_calls_not_the_results():
    """newviewGeneric only marks its nodes; the next call that needs them recomputes them inside its own descriptor
    (axml_backend.c).  Same search, same likelihood, no newview engine call left; EXAML_B200_EAGER_NEWVIEWS=1 is the
    call-by-call behaviour."""
    import re
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree = small_dna_case(tmp, taxa=14, sites=300, seed=8)
        out = {}
        for tag, extra in (("EAGER", {"EXAML_B200_EAGER_NEWVIEWS": "1"}), ("LAZY", {})):
            r, best, _ = run_mode(EXE, binary, tree, tmp, tag, ["-m", "GAMMA", "-f", "d"], env=dict(ENV, EXAML_B200_SPR_STATS="1", **extra))
            assert r.returncode == 0, r.stdout[-1500:]
            m = re.search(r"likelihood calls: (\d+) evaluateGeneric \(full\) \+ (\d+) \(partial\), (\d+) newviewGeneric \(\+ (\d+) deferred\), (\d+) makenewzGeneric; (\d+) descriptor entries", r.stdout)
            assert m, r.stdout[-600:]
            out[tag] = (best, [int(x) for x in m.groups()])
        (b0, c0), (b1, c1) = out["EAGER"], out["LAZY"]
        assert _rel(b1, b0) < 1e-9, (b0, b1)
        assert c0[2] > 100 and c0[3] == 0                  # eager: every newviewGeneric is an engine call
        assert c1[2] == 0 and c1[3] == c0[2]               # deferred: none is, the same calls were made by the host
        assert c1[0] == c0[0] and c1[1] == c0[1] and c1[4] == c0[4]
        assert c1[5] <= c0[5]                              # and no more node updates than before


def test_batched_insertion_scan_on_two_ranks():
    """The scan's per-candidate sums are all-reduced inside exb_score_insertions (here: the oracle harness over mpi_tcp.c);
    both ranks walk the same branches and must answer testInsertBIG identically -- the search ends where the reference's does."""
    import re
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree = small_dna_case(tmp, taxa=16, sites=500, seed=33)
        args = ["-m", "GAMMA", "-f", "d"]
        r0, best0, _ = run_mode(REFERENCE, binary, tree, tmp, "REF", args)
        assert r0.returncode == 0
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
        env.update(ENV, EXAML_B200_SPR_STATS="1", PYTHONPATH=ROOT)
        r = subprocess.run([sys.executable, "-m", "examl_b200.launch", "-np", "2", "--", hostcheck_build.EXE_SPR, "-s", binary, "-t", tree,
                            "-n", "SPR2", "-w", tmp + "/"] + args,
                           cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:]
        m = re.findall(r"Likelihood of best tree: (-?[0-9.]+)", r.stdout)
        assert m and _rel(float(m[-1]), best0) < 1e-6, (m, best0)
        st = re.search(r"batched insertion scoring: (\d+) scans, (\d+) engine calls, (\d+) candidates scored, (\d+) served", r.stdout)
        assert st and int(st.group(1)) > 10 and int(st.group(4)) > int(st.group(1)), r.stdout[-800:]


def test_restart_from_checkpoint():
    """-R: the search state comes back from ExaML_binaryCheckpoint (searchAlgo.c:1726); CLVs are not part of it, so the
    back-end must rebuild everything from the first full traversal after the restart."""
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree = small_dna_case(tmp, taxa=14, sites=400, seed=9)
        r0, best0, _ = run_mode(EXE, binary, tree, tmp, "A", ["-m", "GAMMA", "-f", "d"], env=ENV)
        assert r0.returncode == 0, r0.stdout[-1500:]
        ckpt = os.path.join(tmp, "ExaML_binaryCheckpoint.A_0")
        assert os.path.exists(ckpt)
        e = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
        e.update(ENV)
        r1 = subprocess.run([EXE, "-s", binary, "-R", ckpt, "-m", "GAMMA", "-n", "B", "-f", "d", "-w", tmp + "/"], cwd=tmp, env=e,
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r1.returncode == 0, r1.stdout[-1500:]
        import re
        m = re.findall(r"Likelihood of best tree: (-?[0-9.]+)", r1.stdout)
        assert m and _rel(float(m[-1]), best0) < 1e-6, (m, best0)


def test_auto_protein_model_next_to_lg4x():
    """--auto-prot=ml: optimizeModel.c tries every empirical AA matrix on the AUTO partition (the host rewrites EIGN / EV /
    EI / tipVector without notice, the back-end must pick every change up) while the neighbour partition runs LG4X
    (four eigen-systems per partition)."""
    import re
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree = small_aa_case(tmp, ["AUTO, p0 = 1-60", "LG4X, p1 = 61-120"])
        args = ["-m", "GAMMA", "-f", "e", "--auto-prot=ml"]
        r0, _, _ = run_mode(REFERENCE, binary, tree, tmp, "REF", args)
        r1, _, _ = run_mode(EXE, binary, tree, tmp, "HC", args, env=ENV)
        assert r0.returncode == 0 and r1.returncode == 0, (r0.stdout[-1500:], r1.stdout[-1500:])
        pick = lambda out: re.findall(r"best-scoring AA model: (\S+) likelihood (-?[0-9.]+)", out)
        final = lambda out: float(re.findall(r"Likelihood tree 0: (-?[0-9.]+)", out)[-1])
        assert pick(r0.stdout) and [m for m, _ in pick(r0.stdout)] == [m for m, _ in pick(r1.stdout)]
        assert _rel(final(r1.stdout), final(r0.stdout)) < 1e-6
