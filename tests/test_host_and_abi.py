"""CPU-side tests: the host logic mirrored from the reference, the model constants, and the shape of the
C-ABI library (it must load and export every symbol of include/examl_b200.h -- no compute without a GPU)."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from examl_b200 import data as exdata
from examl_b200 import model as exmodel
from examl_b200.likelihood import Likelihood
from examl_b200.phylo import INNER_INNER, TIP_INNER, TIP_TIP, Tree
from golden_util import alignment_of, constants_of, load_engine, load_json, traversal_of
from synth import build_engine, random_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    from examl_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "examl_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(exb_[a-z_0-9]+)\s*\(", hdr)))
    assert declared and set(declared) == set(capi.EXPORTS), (declared, capi.EXPORTS)
    lib = capi.load_library()
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    """Without a CUDA device the engine refuses to exist -- and says why."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from examl_b200.capi import Engine, EngineError
    with pytest.raises(EngineError, match="no CUDA device"):
        Engine(8, 1, 1)


def test_dropin_binary_fails_loudly_without_gpu():
    import torch
    exe = os.path.join(ROOT, "bin", "examl-b200")
    parser = os.path.join(ROOT, "oracle", "_ref", "parse-examl")
    if torch.cuda.is_available() or not (os.path.exists(exe) and os.path.exists(parser)):
        pytest.skip("needs the prebuilt drop-in binary and no GPU")
    from dropin_util import write_inputs
    with tempfile.TemporaryDirectory() as tmp:
        binary, tree, _ = write_inputs("dna49", tmp)
        r = subprocess.run([exe, "-s", binary, "-t", tree, "-m", "GAMMA", "-n", "X", "-f", "e", "-w", tmp + "/"],
                           cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode != 0
        assert "no CPU fallback" in r.stdout


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "examl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in src and "from oracle" not in src and "import oracle" not in src, f


# ---------------------------------------------------------------- traversal descriptors
@pytest.mark.parametrize("tag", ["dna49", "aa60", "dnaiid200"])
def test_descriptor_builder_reproduces_reference_lnl(oracle_factory, tag):
    """Our computeTraversalInfo on the parsed Newick tree gives a valid post-order descriptor: evaluated by the
    oracle it reproduces the reference's first-evaluation lnL (the node numbering differs from the reference's
    tree reader, the likelihood does not)."""
    first = load_json(f"{tag}.first.json")
    tree = Tree.from_newick(first["tree_newick"], first["names"])
    ti = tree.full_traversal()
    assert len(ti) == first["traversal"]["count"]
    ref_cases = sorted(e["tipCase"] for e in first["traversal"]["entries"][1:])
    assert sorted(e.tipCase for e in ti[1:]) == ref_cases
    done = set()
    for e in ti[1:]:                                   # children before parents, tip always in q
        for child, inner in ((e.qNumber, e.tipCase == INNER_INNER), (e.rNumber, e.tipCase != TIP_TIP)):
            assert (child > tree.mxtips) == inner
            assert child <= tree.mxtips or child in done
        done.add(e.pNumber)
    eng = load_engine(oracle_factory, first)
    lnl = float(np.sum(eng.evaluateIterative(ti)))
    assert abs(lnl - first["likelihood"]) <= 1e-11 * abs(lnl)


def test_partial_traversals_follow_orientation_flags(oracle_factory):
    tree, parts, models = random_case(14, [60], [4], seed=4)
    eng = build_engine(oracle_factory, tree, parts, models)
    lk = Likelihood(tree, eng, 1)
    base = lk.evaluateGeneric(tree.start, True)
    assert tree.partial_traversal(tree.start)[1:] == []            # everything already points at the root branch
    far = max(tree.branches(), key=lambda ab: ab[0].number + ab[1].number)[0]
    ti = tree.partial_traversal(far)                               # flips the x flags: the descriptor MUST be executed
    assert 0 < len(ti) - 1 < tree.mxtips - 2                       # only the path in between is re-oriented
    assert abs(float(eng.evaluateIterative(ti)[0]) - base) <= 1e-10 * abs(base)
    assert tree.partial_traversal(far)[1:] == []


def test_newton_raphson_improves_likelihood(oracle_factory):
    tree, parts, models = random_case(10, [300], [4], seed=8)
    eng = build_engine(oracle_factory, tree, parts, models)
    lk = Likelihood(tree, eng, 1)
    before = lk.evaluateGeneric(tree.start, True)
    for a, b in list(tree.branches()):
        z = lk.makenewzGeneric(a, b, a.z, 10)
        assert 1e-15 <= z[0] <= 1 - 1e-6
        a.z[:] = z
    assert lk.evaluateGeneric(tree.start, True) > before


# ---------------------------------------------------------------- model constants
def test_constants_reproduce_reference_likelihood(oracle_factory):
    """Constants regenerated by examl_b200.model from the reference's frequencies / rates (all 1, alpha 1)
    give the reference's first-evaluation lnL although the eigen-basis itself is not unique."""
    first = load_json("dna49.first.json")
    part = first["partitions"][0]
    mc = exmodel.make_constants(np.ones(6), np.array(part["frequencies"]), alpha=part["alpha"])
    np.testing.assert_allclose(mc.gammaRates, part["gammaRates"], rtol=1e-9)
    np.testing.assert_allclose(np.sort(mc.EIGN), np.sort(part["EIGN"]), rtol=1e-9, atol=1e-12)
    eng = load_engine(oracle_factory, first)
    eng.set_model(0, mc)
    lnl = eng.evaluateIterative(traversal_of(first))[0]
    assert abs(lnl - first["likelihood"]) <= 1e-10 * abs(lnl)


def test_constants_are_a_valid_markov_process():
    from scipy.linalg import expm
    rng = np.random.default_rng(1)
    for S in (4, 20):
        ex = rng.uniform(0.2, 3.0, S * (S - 1) // 2)
        f = rng.dirichlet(np.full(S, 8.0))
        mc = exmodel.make_constants(ex, f, alpha=0.7)
        r = np.zeros((S, S)); r[np.triu_indices(S, 1)] = ex; r = r + r.T
        Q = r * f[None, :]
        np.fill_diagonal(Q, 0.0)
        np.fill_diagonal(Q, -Q.sum(axis=1))
        Q /= -(f * np.diag(Q)).sum()                     # one expected substitution per unit time
        t = 0.37
        P = exmodel.transition_matrix(mc, np.exp(-t))
        np.testing.assert_allclose(P, expm(Q * t), atol=1e-12)
        np.testing.assert_allclose(mc.gammaRates.mean(), 1.0, rtol=1e-12)
        np.testing.assert_allclose(mc.tipVector.reshape(-1, S)[-1 if S == 20 else 15], np.minimum(mc.EV.reshape(S, S).sum(0), exmodel.MAX_TIP_EV))


# ---------------------------------------------------------------- data helpers
def test_pattern_compression_roundtrip_and_sharding():
    tree = Tree.random(9, seed=3)
    tips = exdata.synthetic_dna(tree, 500, seed=2, mode="signal", p_change=0.05)
    cols, wgt = exdata.compress_patterns(tips)
    assert wgt.sum() == 500 and cols.shape[1] < 500
    back = exdata.expand_patterns(cols, wgt)
    assert sorted(map(bytes, back.T)) == sorted(map(bytes, tips.T))
    for n in (1, 2, 3, 8):
        spans = [exdata.shard_sites(1001, r, n) for r in range(n)]
        assert spans[0][0] == 0 and spans[-1][1] == 1001
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
