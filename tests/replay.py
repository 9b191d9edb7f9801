"""Replay of the reference's golden states (tests/golden) through an engine.  The same code
checks the CPU oracle (pinning it to the reference) and the CUDA engine (parity proper)."""
from __future__ import annotations

import math

import numpy as np

from examl_b200.likelihood import Likelihood
from examl_b200.phylo import Entry, Tree
from golden_util import load_engine, load_json, set_models, traversal_of

# north_star tolerance: lnL within 1e-6 relative of the reference's AVX build.  The engines do
# far better (reduction-order effects only), so the tests pin a much tighter bound.
NORTH_STAR_RTOL = 1e-6
TIGHT_RTOL = 5e-12

CASES = {
    # tag: (has mkz, has final)
    "dna49": (True, True),
    "dna49p": (False, True),
    "dna49pM": (True, True),
    "aa140": (True, True),
    "dnaiid200": (False, False),
    "aa60": (True, True),
    # -m PSR: one rate category per site (CAT kernels); the end state carries the optimised per-site categories
    "dna49psr": (True, True),
    "dna49ppsr": (False, True),
    "aa60psr": (True, True),
    # LG4X / LG4M: per-category eigen-systems and category weights
    "aa60lg4x": (True, True),
    "aa60lg4m": (True, True),
}


def rel(a, b):
    return abs(a - b) / abs(b)


def replay_case(factory, tag, rtol=TIGHT_RTOL, **kw):
    has_mkz, has_final = CASES[tag]
    first = load_json(f"{tag}.first.json")
    eng = load_engine(factory, first, **kw)
    nparts = first["NumberOfModels"]

    # --- first full evaluation ---------------------------------------------------------------
    lh = eng.evaluateIterative(traversal_of(first))
    for m in range(nparts):
        assert rel(lh[m], first["perPartitionLH"][m]) < rtol, (tag, m, lh[m], first["perPartitionLH"][m])
    assert rel(float(np.sum(lh)), first["likelihood"]) < rtol
    ti0 = first["traversal"]["entries"][0]
    for m, part in enumerate(first["partitions"]):
        got = eng.read_scaler(m, ti0["p"]) + eng.read_scaler(m, ti0["q"])
        assert got == part["rootScaler"], (tag, m, got, part["rootScaler"])

    # --- first makenewz: descriptor + Newton-Raphson on the host, sum/core on the engine -------
    if has_mkz:
        mkz = load_json(f"{tag}.mkz.json")
        tree = Tree(first["mxtips"], first["numBranches"])      # topology not needed: descriptor is given
        lk = Likelihood(tree, eng, nparts)
        ti = traversal_of(mkz)
        z = lk._topLevelMakenewz(ti, list(mkz["makenewz"]["z0"]), mkz["makenewz"]["maxiter"])
        for got, want in zip(z, mkz["makenewz"]["result"]):
            assert rel(got, want) < 1e-9, (tag, z, mkz["makenewz"]["result"])

    # --- final state of `-f e`: optimised model and branch lengths ---------------------------------
    if has_final:
        final = load_json(f"{tag}.final.json")
        set_models(eng, final)
        lh = eng.evaluateIterative(traversal_of(final))
        for m in range(nparts):
            assert rel(lh[m], final["perPartitionLH"][m]) < rtol, (tag, m, lh[m], final["perPartitionLH"][m])
        assert rel(float(np.sum(lh)), final["likelihood"]) < rtol
    eng.close()


PARTIAL_CASES = ("dna49psr", "dna49ppsr", "aa60psr")


def replay_partial(factory, tag, rtol=1e-11, **kw):
    """evaluatePartialGeneric: the reference's recorded (partition, site, rate) -> value calls during its first
    optimizeRateCategories, replayed as one batch."""
    first = load_json(f"{tag}.first.json")
    part = load_json(f"{tag}.partial.json")
    eng = load_engine(factory, part, alignment_state=first, **kw)
    calls = np.array(part["partial"])
    got = eng.evaluatePartial(traversal_of(part), calls[:, 0].astype(np.int32), calls[:, 1].astype(np.int32), calls[:, 2])
    want = calls[:, 3]
    err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
    assert err.max() < rtol, (tag, int(err.argmax()), got[err.argmax()], want[err.argmax()])
    eng.close()
    return len(want)
