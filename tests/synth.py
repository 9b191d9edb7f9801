"""Seeded synthetic cases shared by the CPU and GPU tests."""
from __future__ import annotations

import numpy as np

from examl_b200 import data as exdata
from examl_b200 import model as exmodel
from examl_b200.phylo import Tree


def random_models(kinds, seed=0):
    """kinds: list of 4 / 20.  GTR-like constants with non-degenerate spectra."""
    rng = np.random.default_rng(seed)
    out = []
    for i, s in enumerate(kinds):
        if s == 4:
            rates = list(rng.uniform(0.5, 3.0, 5)) + [1.0]
            f = rng.dirichlet(np.full(4, 10.0))
            out.append(exmodel.make_constants(np.array(rates), f, alpha=float(rng.uniform(0.3, 2.0))))
        else:
            out.append(exmodel.poisson_aa(alpha=float(rng.uniform(0.3, 2.0)), seed=seed * 100 + i))
    return out


def random_case(mxtips, widths, kinds, seed=0, mode="signal", num_branches=1, compress=False):
    """Returns (tree, [(states, tips, wgt)], models)."""
    tree = Tree.random(mxtips, seed=seed + 1, num_branches=num_branches)
    rng = np.random.default_rng(seed + 2)
    for a, _b in tree.branches():
        z = list(rng.uniform(0.35, 0.97, num_branches))
        a.z[:] = z
    parts = []
    for i, (w, s) in enumerate(zip(widths, kinds)):
        if s == 4:
            tips = exdata.synthetic_dna(tree, w, seed=seed * 10 + i, mode=mode, p_change=0.1, gap_frac=0.03)
        else:
            tips = exdata.synthetic_aa(tree, w, seed=seed * 10 + i, mode=mode, p_change=0.1, gap_frac=0.03)
        if compress and w > 0:
            tips, wgt = exdata.compress_patterns(tips)
        else:
            wgt = rng.integers(1, 4, w).astype(np.int32)
        parts.append((s, tips, wgt))
    return tree, parts, random_models(kinds, seed)


def random_categories(parts, max_cats=25, seed=0):
    """PSR-style per-site rate categories: (perSiteRates[ncat], rateCategory[width]) per partition."""
    rng = np.random.default_rng(seed + 77)
    out = []
    for i, (s, tips, wgt) in enumerate(parts):
        ncat = int(rng.integers(1, max_cats + 1))
        rates = np.sort(rng.gamma(2.0, 0.5, ncat)) + 0.01
        out.append((rates, rng.integers(0, ncat, wgt.shape[0]).astype(np.int32)))
    return out


def build_engine(factory, tree, parts, models, cats=None, max_cats=25, **kw):
    """cats: None for GAMMA, else the list from random_categories (CAT / PSR)."""
    eng = factory(tree.mxtips, len(parts), tree.numBranches, **kw)
    if cats is not None:
        eng.set_rate_model(0, max_cats)
    for m, (s, tips, wgt) in enumerate(parts):
        eng.set_partition(m, s, wgt, tips)
    eng.commit()
    for m, mc in enumerate(models):
        eng.set_model(m, mc)
        if cats is not None:
            eng.set_categories(m, cats[m][0], cats[m][1])
    return eng
