"""Parity of the CUDA engine (through the C ABI) against the reference's goldens and the CPU oracle.

Tolerances: north_star asks for lnL within 1e-6 relative of the reference's AVX build.  The CUDA
path differs from the reference only in summation order and in exp/log ulps, so these tests hold it
to 5e-12 relative on lnL, 1e-11 on CLV entries and exact equality on scaler counts."""
import numpy as np
import pytest

from examl_b200.likelihood import Likelihood
from replay import CASES, PARTIAL_CASES, replay_case, replay_partial
from synth import build_engine, random_case, random_categories

pytestmark = pytest.mark.gpu

CLV_RTOL = 1e-11


@pytest.mark.parametrize("tag", list(CASES))
def test_reference_goldens(cuda_factory, tag):
    replay_case(cuda_factory, tag)


@pytest.mark.parametrize("tile", [32, 64, 256, 1024])
def test_goldens_any_tile_size(cuda_factory, tile):
    replay_case(cuda_factory, "dna49", tile_sites=tile)
    replay_case(cuda_factory, "aa60", tile_sites=tile)


def _compare_engines(gpu, cpu, tree, parts, mask=None, clv_rtol=CLV_RTOL):
    ti = tree.full_traversal()
    lg = gpu.evaluateIterative(ti, mask)
    lc = cpu.evaluateIterative(ti, mask)
    for m, (s, tips, wgt) in enumerate(parts):
        if mask is not None and not mask[m]:
            continue
        if wgt.shape[0] == 0:
            assert lg[m] == 0.0
            continue
        assert abs(lg[m] - lc[m]) <= 5e-12 * abs(lc[m]), (m, lg[m], lc[m])
        for e in ti[1:]:
            a, b = gpu.read_clv(m, e.pNumber), cpu.read_clv(m, e.pNumber)
            scale = np.abs(b).max(axis=(1, 2), keepdims=True)
            assert np.all(np.abs(a - b) <= clv_rtol * scale), (m, e.pNumber)
            assert gpu.read_scaler(m, e.pNumber) == cpu.read_scaler(m, e.pNumber)
    return ti


@pytest.mark.parametrize("widths,kinds", [
    ([1], [4]), ([31], [4]), ([33], [4]), ([257], [4]), ([1000], [4]),
    ([1], [20]), ([65], [20]), ([300], [20]),
    ([100, 0, 37, 64], [4, 4, 20, 4]),
])
def test_clv_scaler_sum_core_vs_oracle(cuda_factory, oracle_factory, widths, kinds):
    tree, parts, models = random_case(24, widths, kinds, seed=len(widths) * 7 + widths[0])
    gpu = build_engine(cuda_factory, tree, parts, models, tile_sites=64)
    cpu = build_engine(oracle_factory, tree, parts, models)
    _compare_engines(gpu, cpu, tree, parts)
    # makenewz at an inner branch: sum table + derivatives for a few trial values
    p = tree.start.back.next.back
    if tree.isTip(p.number):
        p = tree.start.back.next.next.back
    ti = tree.partial_traversal(p)
    gpu.makenewzIterative(ti)
    cpu.makenewzIterative(ti)
    for m, (s, tips, wgt) in enumerate(parts):
        if wgt.shape[0]:
            a, b = gpu.read_sumtable(m), cpu.read_sumtable(m)
            assert np.all(np.abs(a - b) <= CLV_RTOL * np.abs(b).max(axis=(1, 2), keepdims=True))
    for z in (0.9, 0.5, 0.05, 1e-15):
        g1, g2 = gpu.execCore([np.log(z)])
        c1, c2 = cpu.execCore([np.log(z)])
        assert abs(g1[0] - c1[0]) <= 1e-10 * max(1.0, abs(c1[0])), (z, g1, c1)
        assert abs(g2[0] - c2[0]) <= 1e-10 * max(1.0, abs(c2[0])), (z, g2, c2)


def test_scaling_path_iid_dna(cuda_factory, oracle_factory):
    """i.i.d. columns on 150 taxa: every site is rescaled at least once (SURVEY.md 4 iv)."""
    tree, parts, models = random_case(150, [500], [4], seed=3, mode="iid")
    gpu = build_engine(cuda_factory, tree, parts, models, tile_sites=128)
    cpu = build_engine(oracle_factory, tree, parts, models)
    ti = _compare_engines(gpu, cpu, tree, parts)
    root = gpu.read_scaler(0, ti[0].pNumber) + gpu.read_scaler(0, ti[0].qNumber)
    assert root >= int(parts[0][2].sum()), root


def test_masks_and_per_partition_branches(cuda_factory, oracle_factory):
    tree, parts, models = random_case(16, [90, 40, 130], [4, 20, 4], seed=11, num_branches=3)
    gpu = build_engine(cuda_factory, tree, parts, models, tile_sites=32)
    cpu = build_engine(oracle_factory, tree, parts, models)
    _compare_engines(gpu, cpu, tree, parts)
    mask = [True, False, True]
    before = gpu.evaluateIterative(tree.full_traversal())
    out = before.copy()
    out[1] = 123.0                                 # a masked partition's slot must not be touched
    gpu.evaluateIterative(tree.full_traversal(), mask, out=out)
    assert out[1] == 123.0 and out[0] == before[0] and out[2] == before[2]
    # derivative slots follow the partitions under -M
    p = tree.start.back
    ti = tree.partial_traversal(p)
    gpu.makenewzIterative(ti)
    cpu.makenewzIterative(ti)
    lz = [np.log(0.8), np.log(0.6), np.log(0.4)]
    g1, g2 = gpu.execCore(lz, [True, True, False])
    c1, c2 = cpu.execCore(lz, [True, True, False])
    assert g1[2] == 0.0 and g2[2] == 0.0
    np.testing.assert_allclose(g1, c1, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(g2, c2, rtol=1e-10, atol=1e-10)


def test_root_invariance_and_newton_raphson(cuda_factory, oracle_factory):
    """Pulley principle: lnL is the same at every branch (partial traversals re-orient CLVs);
    then one smoothing sweep of makenewzGeneric must not lower lnL and must track the oracle."""
    tree, parts, models = random_case(20, [400], [4], seed=5)
    gpu = build_engine(cuda_factory, tree, parts, models, tile_sites=64)
    lk = Likelihood(tree, gpu, 1)
    base = lk.evaluateGeneric(tree.start, True)
    for a, b in list(tree.branches())[:12]:
        assert abs(lk.evaluateGeneric(a, False) - base) <= 1e-10 * abs(base)
    tree2, parts2, models2 = random_case(20, [400], [4], seed=5)
    cpu = build_engine(oracle_factory, tree2, parts2, models2)
    lk2 = Likelihood(tree2, cpu, 1)
    lk2.evaluateGeneric(tree2.start, True)
    for (a, b), (a2, b2) in list(zip(tree.branches(), tree2.branches()))[:10]:
        z = lk.makenewzGeneric(a, b, a.z, 10)
        z2 = lk2.makenewzGeneric(a2, b2, a2.z, 10)
        assert abs(z[0] - z2[0]) <= 1e-9 * z2[0]
        a.z[:] = z
        a2.z[:] = z2
    after = lk.evaluateGeneric(tree.start, True)
    assert after >= base - 1e-9 * abs(base)
    assert abs(after - lk2.evaluateGeneric(tree2.start, True)) <= 1e-10 * abs(after)


def test_bitwise_repeatability(cuda_factory):
    """optimizeModel.c:3030 asserts that two consecutive evaluations agree bit for bit."""
    tree, parts, models = random_case(30, [5000, 777], [4, 20], seed=9)
    gpu = build_engine(cuda_factory, tree, parts, models)
    ti = tree.full_traversal()
    a = gpu.evaluateIterative(ti).copy()
    for _ in range(3):
        b = gpu.evaluateIterative(ti)
        assert a.tobytes() == b.tobytes()


def test_full_size_linearity_dna(cuda_factory, oracle_factory):
    """1 M sites x 40 taxa: the alignment is a 1 000-site block repeated 1 000 times, so
    lnL(1 M) = 1 000 x lnL(block) -- a size-independent check at BASELINE.json's site count."""
    tree, parts, models = random_case(40, [1000], [4], seed=21)
    s, tips, wgt = parts[0]
    wgt = np.ones_like(wgt)
    cpu = build_engine(oracle_factory, tree, [(s, tips, wgt)], models)
    ti = tree.full_traversal()
    block = cpu.evaluateIterative(ti)[0]
    big = [(s, np.tile(tips, (1, 1000)), np.ones(1000 * 1000, dtype=np.int32))]
    gpu = build_engine(cuda_factory, tree, big, models)
    total = gpu.evaluateIterative(ti)[0]
    assert abs(total - 1000.0 * block) <= 1e-9 * abs(total), (total, block)


def test_errors_are_loud(cuda_factory):
    from examl_b200.capi import EngineError
    tree, parts, models = random_case(8, [10], [4], seed=1)
    gpu = cuda_factory(8, 1, 1)
    with pytest.raises(EngineError):
        gpu.set_partition(0, 2, parts[0][2], parts[0][1])       # binary data: not implemented
    gpu.set_partition(0, 4, parts[0][2], parts[0][1])
    gpu.commit()
    with pytest.raises(EngineError):
        gpu.evaluateIterative(tree.full_traversal())             # no model constants yet
    gpu.set_model(0, models[0])
    bad = tree.full_traversal()
    bad[1].pNumber = 3                                           # a tip cannot be recomputed
    with pytest.raises(EngineError):
        gpu.evaluateIterative(bad)


@pytest.mark.parametrize("widths,kinds,nb", [([400], [4], 1), ([150, 90], [4, 20], 1), ([120, 60, 200], [4, 20, 4], 3)])
def test_device_newton_raphson_matches_host_loop(cuda_factory, oracle_factory, widths, kinds, nb):
    """exb_makenewz (the whole NR loop in one cooperative kernel) against the host loop on the oracle:
    same z for every branch of a smoothing sweep, same curvature flags, same end likelihood."""
    tree, parts, models = random_case(18, widths, kinds, seed=23, num_branches=nb)
    tree2, parts2, models2 = random_case(18, widths, kinds, seed=23, num_branches=nb)
    gpu = build_engine(cuda_factory, tree, parts, models, tile_sites=64)
    cpu = build_engine(oracle_factory, tree2, parts2, models2)
    lk = Likelihood(tree, gpu, len(widths), device_newton_raphson=True)
    lk2 = Likelihood(tree2, cpu, len(widths))
    lk.evaluateGeneric(tree.start, True)
    lk2.evaluateGeneric(tree2.start, True)
    for (a, b), (a2, b2) in zip(list(tree.branches()), list(tree2.branches())):
        z = lk.makenewzGeneric(a, b, a.z, 10)
        z2 = lk2.makenewzGeneric(a2, b2, a2.z, 10)
        assert lk.nr_iterations == lk2.nr_iterations
        assert lk.curvatOK == lk2.curvatOK
        for u, v in zip(z, z2):
            assert abs(u - v) <= 1e-9 * v, (z, z2)
        a.z[:] = z
        a2.z[:] = z2
    l1, l2 = lk.evaluateGeneric(tree.start, True), lk2.evaluateGeneric(tree2.start, True)
    assert abs(l1 - l2) <= 1e-10 * abs(l2)


@pytest.mark.parametrize("widths,kinds,nb", [([1], [4], 1), ([700], [4], 1), ([50], [20], 1), ([500], [20], 1),
                                             ([300, 0, 420, 33], [4, 4, 20, 20], 1), ([130, 400], [4, 20], 2)])
def test_psr_cat_model_vs_oracle(cuda_factory, oracle_factory, widths, kinds, nb):
    """-m PSR: one rate category per site (CAT kernels): CLVs, scalers, lnL, sum table, derivatives and the
    device-resident Newton-Raphson against the oracle's CAT restatement."""
    tree, parts, models = random_case(22, widths, kinds, seed=41 + len(widths), num_branches=nb)
    tree2, parts2, models2 = random_case(22, widths, kinds, seed=41 + len(widths), num_branches=nb)
    cats = random_categories(parts, max_cats=25, seed=5)
    gpu = build_engine(cuda_factory, tree, parts, models, cats=cats, tile_sites=64)
    cpu = build_engine(oracle_factory, tree2, parts2, models2, cats=cats)
    _compare_engines(gpu, cpu, tree, parts)
    lk = Likelihood(tree, gpu, len(widths), device_newton_raphson=True)
    lk2 = Likelihood(tree2, cpu, len(widths))
    lk.evaluateGeneric(tree.start, True)
    lk2.evaluateGeneric(tree2.start, True)
    for (a, b), (a2, b2) in list(zip(tree.branches(), tree2.branches()))[:8]:
        ti = tree.partial_traversal(a)
        ti2 = tree2.partial_traversal(a2)
        gpu.makenewzIterative(ti); cpu.makenewzIterative(ti2)
        for m, (s, tips, wgt) in enumerate(parts):
            if wgt.shape[0]:
                x, y = gpu.read_sumtable(m), cpu.read_sumtable(m)
                assert np.all(np.abs(x - y) <= CLV_RTOL * np.abs(y).max(axis=(1, 2), keepdims=True))
        lz = [np.log(0.7)] * nb
        g1, g2 = gpu.execCore(lz); c1, c2 = cpu.execCore(lz)
        np.testing.assert_allclose(g1, c1, rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(g2, c2, rtol=1e-10, atol=1e-10)
        z = lk.makenewzGeneric(a, b, a.z, 10)
        z2 = lk2.makenewzGeneric(a2, b2, a2.z, 10)
        for u, v in zip(z, z2):
            assert abs(u - v) <= 1e-9 * v
        a.z[:] = z; a2.z[:] = z2
    l1, l2 = lk.evaluateGeneric(tree.start, True), lk2.evaluateGeneric(tree2.start, True)
    assert abs(l1 - l2) <= 1e-10 * abs(l2)


def test_psr_scaling_path(cuda_factory, oracle_factory):
    tree, parts, models = random_case(150, [400, 300], [4, 20], seed=3, mode="iid")
    cats = random_categories(parts, max_cats=12, seed=2)
    gpu = build_engine(cuda_factory, tree, parts, models, cats=cats, max_cats=12)
    cpu = build_engine(oracle_factory, tree, parts, models, cats=cats, max_cats=12)
    ti = _compare_engines(gpu, cpu, tree, parts)
    assert gpu.read_scaler(0, ti[0].pNumber) + gpu.read_scaler(0, ti[0].qNumber) > 0


@pytest.mark.parametrize("cat", [False, True])
def test_large_tree_entry_windows(cuda_factory, oracle_factory, cat):
    """700 taxa: the descriptor (698 entries) spans eight staging windows of 96 entries; the tip rows and the
    descriptor are staged per window, the carry registers live across windows.  GAMMA and PSR, DNA and AA."""
    tree, parts, models = random_case(700, [70, 40], [4, 20], seed=77)
    cats = random_categories(parts, max_cats=9, seed=1) if cat else None
    gpu = build_engine(cuda_factory, tree, parts, models, cats=cats, max_cats=9, tile_sites=64)
    cpu = build_engine(oracle_factory, tree, parts, models, cats=cats, max_cats=9)
    ti = tree.full_traversal()
    lg, lc = gpu.evaluateIterative(ti), cpu.evaluateIterative(ti)
    assert np.all(np.abs(lg - lc) <= 5e-12 * np.abs(lc)), (lg, lc)
    for e in ti[1::37]:
        for m in range(2):
            a, b = gpu.read_clv(m, e.pNumber), cpu.read_clv(m, e.pNumber)
            assert np.all(np.abs(a - b) <= CLV_RTOL * np.abs(b).max(axis=(1, 2), keepdims=True))
            assert gpu.read_scaler(m, e.pNumber) == cpu.read_scaler(m, e.pNumber)
    # evaluate with no newview at all, and a partial traversal whose length is not a multiple of the window
    lg0, lc0 = gpu.evaluateIterative(ti[:1]), cpu.evaluateIterative(ti[:1])
    assert np.all(np.abs(lg0 - lc0) <= 5e-12 * np.abs(lc0))
    sub = ti[:1] + ti[-200:]
    lg1, lc1 = gpu.evaluateIterative(sub), cpu.evaluateIterative(sub)
    assert np.all(np.abs(lg1 - lc1) <= 5e-12 * np.abs(lc1))
    gpu.makenewzIterative(sub); cpu.makenewzIterative(sub)
    g1, g2 = gpu.execCore([np.log(0.5)]); c1, c2 = cpu.execCore([np.log(0.5)])
    np.testing.assert_allclose(g1, c1, rtol=1e-10); np.testing.assert_allclose(g2, c2, rtol=1e-10)


def test_lg4_mixed_with_ordinary_partitions(cuda_factory, oracle_factory):
    """An LG4X partition (per-category eigen-systems + weights, taken from the reference's dump) next to an ordinary
    AA and a DNA partition: CLVs, lnL, sum table, derivatives and the device NR against the oracle."""
    from golden_util import load_json
    lg = load_json("aa60lg4x.final.json")["partitions"][0]
    tree, parts, models = random_case(25, [130, 90, 77], [20, 20, 4], seed=61)
    tree2, parts2, models2 = random_case(25, [130, 90, 77], [20, 20, 4], seed=61)

    def build(factory, tr, **kw):
        eng = factory(tr.mxtips, 3, 1, **kw)
        for m, (s, tips, wgt) in enumerate(parts):
            eng.set_partition(m, s, wgt, tips)
        eng.commit()
        eng.set_model_lg4(0, lg["EIGN_LG4"], lg["EV_LG4"], lg["EI_LG4"], lg["tipVector_LG4"], lg["gammaRates"], lg["weights"])
        eng.set_model(1, models[1]); eng.set_model(2, models[2])
        return eng
    gpu, cpu = build(cuda_factory, tree, tile_sites=64), build(oracle_factory, tree2)
    # The optimised LG4X model has very unequal category rates: in the fast categories the state-space vector is
    # nearly stationary, its eigen-coordinates are differences of almost equal terms, and two correct FP64
    # evaluations with different summation order differ by ~1e-10 of the site maximum there (lnL still agrees to
    # 5e-12, checked inside).  Hence the looser bound on individual CLV entries for this case only.
    _compare_engines(gpu, cpu, tree, parts, clv_rtol=1e-8)
    lk = Likelihood(tree, gpu, 3, device_newton_raphson=True)
    lk2 = Likelihood(tree2, cpu, 3)
    lk.evaluateGeneric(tree.start, True); lk2.evaluateGeneric(tree2.start, True)
    for (a, b), (a2, b2) in list(zip(tree.branches(), tree2.branches()))[:6]:
        z, z2 = lk.makenewzGeneric(a, b, a.z, 10), lk2.makenewzGeneric(a2, b2, a2.z, 10)
        assert abs(z[0] - z2[0]) <= 1e-9 * z2[0]
        a.z[:] = z; a2.z[:] = z2
    l1, l2 = lk.evaluateGeneric(tree.start, True), lk2.evaluateGeneric(tree2.start, True)
    assert abs(l1 - l2) <= 1e-10 * abs(l2)


# ---- evaluatePartialGeneric (the per-site rate scan of -m PSR) --------------------------------------------
@pytest.mark.parametrize("tag", PARTIAL_CASES)
def test_evaluate_partial_reference_goldens(cuda_factory, tag):
    """The reference's own evaluatePartialGeneric calls (partition, site, rate) -> value, recorded during its first
    optimizeRateCategories, answered in one batch."""
    assert replay_partial(cuda_factory, tag) >= 160


@pytest.mark.parametrize("mxtips,widths,kinds,nb,mode", [
    (22, [300, 0, 420, 33], [4, 4, 20, 20], 1, "signal"),
    (30, [130, 400], [4, 20], 2, "signal"),
    (150, [500, 200], [4, 20], 1, "iid"),        # every vector is rescaled several times
    (700, [70, 40], [4, 20], 1, "signal"),       # deep descriptor, many live slots
])
def test_evaluate_partial_vs_oracle(cuda_factory, oracle_factory, mxtips, widths, kinds, nb, mode):
    tree, parts, models = random_case(mxtips, widths, kinds, seed=91, num_branches=nb, mode=mode)
    cats = random_categories(parts, max_cats=8, seed=4)
    gpu = build_engine(cuda_factory, tree, parts, models, cats=cats, max_cats=8)
    cpu = build_engine(oracle_factory, tree, parts, models, cats=cats, max_cats=8)
    ti = tree.full_traversal()
    rng = np.random.default_rng(12)
    mo, si, ra = [], [], []
    for m, w in enumerate(widths):
        for rep in range(3):                       # every site, several rates, shuffled across partitions
            mo += [m] * w; si += list(range(w)); ra += list(rng.gamma(2.0, 0.6, w) + 1e-3)
    perm = rng.permutation(len(mo))
    mo, si, ra = np.array(mo)[perm], np.array(si)[perm], np.array(ra)[perm]
    g = gpu.evaluatePartial(ti, mo, si, ra)
    c = cpu.evaluatePartial(ti, mo, si, ra)
    assert np.all(np.isfinite(g))
    assert np.all(np.abs(g - c) <= 1e-11 * np.maximum(1.0, np.abs(c))), np.abs(g - c).max()
    # consistency with the ordinary evaluate: the sum over the sites at their categories' rates is the partition's lnL
    lh = gpu.evaluateIterative(ti)
    for m, w in enumerate(widths):
        if w:
            own = gpu.evaluatePartial(ti, [m] * w, np.arange(w), np.asarray(cats[m][0])[cats[m][1]])
            assert abs(own.sum() - lh[m]) <= 1e-10 * abs(lh[m]), (m, own.sum(), lh[m])
    # an empty batch is legal, the CLVs of the engine were not disturbed
    assert gpu.evaluatePartial(ti, [], [], []).shape == (0,)
    lh2 = gpu.evaluateIterative(ti[:1])
    assert np.array_equal(lh, lh2)


def test_evaluate_partial_errors(cuda_factory):
    from examl_b200.capi import EngineError
    tree, parts, models = random_case(12, [40], [4], seed=5)
    gam = build_engine(cuda_factory, tree, parts, models)
    with pytest.raises(EngineError, match="CAT"):
        gam.evaluatePartial(tree.full_traversal(), [0], [0], [1.0])
    cats = random_categories(parts, max_cats=4, seed=1)
    psr = build_engine(cuda_factory, tree, parts, models, cats=cats, max_cats=4)
    ti = tree.full_traversal()
    with pytest.raises(EngineError, match="FULL traversal"):
        psr.evaluatePartial(ti[:1] + ti[3:], [0], [0], [1.0])
    with pytest.raises(EngineError, match="site"):
        psr.evaluatePartial(ti, [0], [40], [1.0])
    with pytest.raises(EngineError, match="partition"):
        psr.evaluatePartial(ti, [1], [0], [1.0])


def test_many_partitions_shared_memory_constants(cuda_factory, oracle_factory):
    """More than 16 DNA partitions in one class: one launch for all of them, EI / EV from shared memory instead of
    kernel parameters (the C5 shape: thousands of small partitions per GPU), with -M style per-partition branches
    exercised through the mask."""
    widths = [40 + 7 * (i % 5) for i in range(24)] + [0, 64]
    kinds = [4] * len(widths)
    tree, parts, models = random_case(20, widths, kinds, seed=123)
    gpu = build_engine(cuda_factory, tree, parts, models)
    cpu = build_engine(oracle_factory, tree, parts, models)
    _compare_engines(gpu, cpu, tree, parts)
    mask = np.array([i % 3 != 0 for i in range(len(widths))], dtype=np.uint8)
    _compare_engines(gpu, cpu, tree, parts, mask=mask)


# ---- the shapes of BASELINE configs 4 and 5 -----------------------------------------------------------------
def test_c4_shape_aa_500_taxa_vs_oracle(cuda_factory, oracle_factory):
    """C4's shape: LG+GAMMA on the 500-taxon bench tree (6 entry windows, DMMA path, scaling fires for AA at this
    depth), 51 200 sites on the GPU; the oracle checks a site sample -- sites are independent, so the lnL of the
    first 1 536 sites of the big engine (read through a weight mask) must equal a 1 536-site oracle run."""
    import bench
    w = bench.make_workload("aa", 1, sites=51_200)
    tips = bench.make_sites(w.tree, 0, w.sites, "aa")
    mc = bench.lg_model(0.5)
    n = 1536
    wgt = np.zeros(w.sites, dtype=np.int32)
    wgt[:n] = 1
    gpu = cuda_factory(w.taxa, 2, 1)
    gpu.set_partition(0, 20, np.ones(w.sites, dtype=np.int32), tips)      # all sites, weight 1
    gpu.set_partition(1, 20, wgt[:4096].copy(), tips[:, :4096].copy())    # the sample again as its own partition (ragged tile)
    gpu.commit()
    gpu.set_model(0, mc); gpu.set_model(1, mc)
    cpu = oracle_factory(w.taxa, 1, 1)
    cpu.set_partition(0, 20, np.ones(n, dtype=np.int32), tips[:, :n].copy())
    cpu.set_model(0, mc)
    got = gpu.evaluateIterative(w.ti)
    want = cpu.evaluateIterative(w.ti)[0]
    assert abs(got[1] - want) <= 1e-10 * abs(want), (got[1], want)
    # CLV of the root's inner node for the sample sites: big partition == small partition == oracle, scalers included
    root = w.ti[0].qNumber
    a, b = gpu.read_clv(0, root)[:n], gpu.read_clv(1, root)[:n]
    c = cpu.read_clv(0, root)
    assert np.array_equal(a, b)
    scale = np.abs(c).max(axis=(1, 2), keepdims=True)
    assert np.all(np.abs(a - c) <= 1e-10 * scale)
    assert cpu.read_scaler(0, root) > 0, "the AA scaling path must fire at this depth"
    assert gpu.read_scaler(1, root) == cpu.read_scaler(0, root)
    # linearity over the site axis: lnL(all) = sum over 4 quarter engines
    q = w.sites // 4
    tot = 0.0
    for k in range(4):
        e = cuda_factory(w.taxa, 1, 1)
        e.set_partition(0, 20, np.ones(q, dtype=np.int32), tips[:, k * q:(k + 1) * q].copy())
        e.commit(); e.set_model(0, mc)
        tot += e.evaluateIterative(w.ti)[0]
        e.close()
    assert abs(tot - got[0]) <= 1e-10 * abs(tot), (tot, got[0])


def test_c5_shape_1250_partitions_vs_oracle(cuda_factory, oracle_factory):
    """C5's per-GPU shape: 1 250 DNA partitions (here x 96 sites to stay small) with per-partition models on the
    354-taxon tree, masks on; every partition's lnL against the oracle (8 distinct (data, model) combinations)."""
    import bench
    w = bench.make_workload("c5", 1)
    W, P = 96, 1250
    tips = bench.make_sites(w.tree, 0, 8 * W, "dna")
    models = [bench.gtr_model(m) for m in range(8)]
    gpu = cuda_factory(w.taxa, P, 1)
    for m in range(P):
        gpu.set_partition(m, 4, np.ones(W, dtype=np.int32), tips[:, (m % 8) * W:(m % 8 + 1) * W].copy())
    gpu.commit()
    for m in range(P):
        gpu.set_model(m, models[(m // 8) % 8])            # data block m % 8, model (m / 8) % 8: 64 combinations
    cpu = oracle_factory(w.taxa, 64, 1)
    for j in range(64):
        cpu.set_partition(j, 4, np.ones(W, dtype=np.int32), tips[:, (j % 8) * W:(j % 8 + 1) * W].copy())
        cpu.set_model(j, models[j // 8])
    want64 = cpu.evaluateIterative(w.ti)
    want = np.array([want64[(m % 8) + 8 * ((m // 8) % 8)] for m in range(P)])
    got = gpu.evaluateIterative(w.ti).copy()
    assert np.all(np.abs(got - want) <= 1e-10 * np.abs(want)), np.max(np.abs(got - want) / np.abs(want))
    # masked: executed partitions recomputed, masked ones untouched (eval:856-859)
    mask = (np.arange(P) % 3 != 0).astype(np.uint8)
    out = np.full(P, 7.0)
    gpu.evaluateIterative(w.ti, mask, out=out)
    assert np.array_equal(out[mask == 1], got[mask == 1]) and np.all(out[mask == 0] == 7.0)


# ---- mixed precision: the AA newview on tcgen05 (3xTF32, FP32 accumulators in tensor memory) -----------------------
def _tc_factory(cuda_factory):
    def make(mxtips, nparts, nbranches=1, **kw):
        e = cuda_factory(mxtips, nparts, nbranches, **kw)
        e.set_aa_precision(1)
        return e
    return make


def test_tf32x3_reference_goldens_within_north_star_bar(cuda_factory):
    """exb_set_aa_precision(EXB_AA_TF32X3): the reference's own testData/140 state and the aa60 synthetic state.  Bar =
    north_star's 1e-6 relative on lnL (measured ~1e-8, tools/mixed_precision_study.py); the FP64 path is held to 5e-12."""
    from golden_util import load_engine, load_json, traversal_of
    for tag in ("aa140", "aa60"):
        st = load_json(f"{tag}.first.json")
        eng = load_engine(_tc_factory(cuda_factory), st)
        lh = eng.evaluateIterative(traversal_of(st))
        assert abs(lh.sum() - st["likelihood"]) <= 1e-6 * abs(st["likelihood"]), (tag, lh.sum(), st["likelihood"])
        ref = load_engine(cuda_factory, st)
        lr = ref.evaluateIterative(traversal_of(st))
        print(tag, "tf32x3 lnL", lh.sum(), "fp64 lnL", lr.sum(), "rel", abs(lh.sum() - lr.sum()) / abs(lr.sum()))
        assert abs(lh.sum() - lr.sum()) <= 1e-6 * abs(lr.sum())


@pytest.mark.parametrize("mxtips,widths,mode", [(24, [300], "signal"), (150, [517, 40], "iid"), (400, [1000], "signal")])
def test_tf32x3_vs_oracle(cuda_factory, oracle_factory, mxtips, widths, mode):
    """CLVs (relative to the site's largest entry), scaler counts up to borderline sites, lnL, repeatability, ragged
    tiles, several partitions, and the makenewz path on top of tcgen05-written CLVs."""
    tree, parts, models = random_case(mxtips, widths, [20] * len(widths), seed=31, mode=mode)
    gpu = build_engine(_tc_factory(cuda_factory), tree, parts, models)
    cpu = build_engine(oracle_factory, tree, parts, models)
    ti = tree.full_traversal()
    lg, lc = gpu.evaluateIterative(ti).copy(), cpu.evaluateIterative(ti)
    assert np.all(np.abs(lg - lc) <= 1e-6 * np.abs(lc)), (lg, lc)
    assert np.array_equal(gpu.evaluateIterative(ti), lg)                   # bitwise repeatable
    worst = 0.0
    for m, (s, tips, wgt) in enumerate(parts):
        tot_g = tot_c = 0
        for e in ti[1:]:
            a, b = gpu.read_clv(m, e.pNumber), cpu.read_clv(m, e.pNumber)
            sg, sc = gpu.read_scaler(m, e.pNumber), cpu.read_scaler(m, e.pNumber)
            tot_g, tot_c = sg, sc
            if sg == sc:                          # same scaling decisions: the vectors are directly comparable
                scale = np.abs(b).max(axis=(1, 2), keepdims=True)
                worst = max(worst, float((np.abs(a - b) / scale).max()))
        assert abs(tot_g - tot_c) <= 0.02 * max(tot_c, 50), (tot_g, tot_c)
    assert worst <= 5e-2, worst
    # partial traversal + Newton-Raphson on the mixed-precision CLVs agrees with the oracle's branch length to 1e-4
    from examl_b200.likelihood import Likelihood
    lk, lk2 = Likelihood(tree, gpu, len(parts)), Likelihood(tree, cpu, len(parts))
    p = tree.start.back
    z1 = lk.makenewzGeneric(p, p.back, p.z, 10)
    z2 = lk2.makenewzGeneric(p, p.back, p.z, 10)
    assert abs(z1[0] - z2[0]) <= 1e-4 * z2[0], (z1, z2)


# ---- -S: no CLV blocks for (node, tile) pairs that are all-gap below the node -------------------------------------------
def _gappy_case(mxtips, widths, kinds, seed, mode="signal", missing=0.5):
    """Phylogenomic missing data: in every partition a random subset of the taxa is entirely absent (gap code)."""
    tree, parts, models = random_case(mxtips, widths, kinds, seed=seed, mode=mode)
    rng = np.random.default_rng(seed + 5)
    out = []
    for s, tips, wgt in parts:
        tips = tips.copy()
        absent = rng.random(mxtips) < missing
        absent[rng.integers(0, mxtips)] = False           # somebody is always there
        tips[absent, :] = 15 if s == 4 else 22
        out.append((s, tips, wgt))
    return tree, out, models


@pytest.mark.parametrize("mxtips,widths,kinds,tile,mode", [
    (30, [600, 500, 300, 200], [4, 4, 4, 20], 256, "signal"),
    (64, [130, 257, 96, 400], [4, 20, 4, 20], 64, "iid"),          # small tiles, scaling fires
    (120, [40 + 9 * (i % 7) for i in range(20)], [4] * 20, 256, "signal"),   # > 16 DNA partitions: constants from shared memory
])
def test_sev_matches_dense_oracle_and_saves_blocks(cuda_factory, oracle_factory, mxtips, widths, kinds, tile, mode):
    tree, parts, models = _gappy_case(mxtips, widths, kinds, seed=77, mode=mode)

    def sev_factory(mx, npart, nb=1, **kw):
        e = cuda_factory(mx, npart, nb, tile_sites=tile, **kw)
        e.set_sev(True)
        return e
    gpu = build_engine(sev_factory, tree, parts, models)
    cpu = build_engine(oracle_factory, tree, parts, models)
    _compare_engines(gpu, cpu, tree, parts)                       # lnL, every CLV (gap tiles expand to the gap column), scalers
    used, dense, nbytes = gpu.sev_stats()
    assert 0 < used < dense, (used, dense)          # some (node, tile) pairs own no block (how many: test_sev_footprint_follows_the_data)
    # masks, then a different root (partial traversals re-orient nodes: blocks are handed back and taken)
    mask = np.array([i % 2 == 0 for i in range(len(widths))], dtype=np.uint8)
    _compare_engines(gpu, cpu, tree, parts, mask=mask)
    from examl_b200.likelihood import Likelihood
    tree2 = _gappy_case(mxtips, widths, kinds, seed=77, mode=mode)[0]          # same tree, own orientation flags
    tree.full_traversal(); tree2.full_traversal()
    lk, lk2 = Likelihood(tree, gpu, len(parts)), Likelihood(tree2, cpu, len(parts))
    for num in (2, mxtips // 2, mxtips, mxtips + 3, 2 * mxtips - 3):
        node, node2 = tree.nodep[num], tree2.nodep[num]
        a, b = lk.evaluateGeneric(node, False), lk2.evaluateGeneric(node2, False)
        assert abs(a - b) <= 1e-10 * abs(b), (num, a, b)
        z1 = lk.makenewzGeneric(node, node.back, node.z, 10)
        z2 = lk2.makenewzGeneric(node2, node2.back, node2.z, 10)
        assert abs(z1[0] - z2[0]) <= 1e-8 * z2[0]
    a, b = lk.evaluateGeneric(tree.start, True), lk2.evaluateGeneric(tree2.start, True)
    assert abs(a - b) <= 1e-10 * abs(b)
    used2, _, _ = gpu.sev_stats()
    assert used2 == used, (used, used2)                            # back at the original orientation: same blocks in use
    # the dense engine computes the same numbers
    dense_eng = build_engine(lambda mx, npart, nb=1, **kw: cuda_factory(mx, npart, nb, tile_sites=tile, **kw), tree, parts, models)
    assert np.allclose(dense_eng.evaluateIterative(tree.full_traversal()), gpu.evaluateIterative(tree.full_traversal()), rtol=1e-12, atol=0)


def test_sev_footprint_follows_the_data(cuda_factory):
    """A supermatrix that is 3/4 missing data: the slabs the engine allocated hold far less than the dense arena."""
    tree, parts, models = _gappy_case(200, [20000] * 6, [4] * 6, seed=3, missing=0.75)
    e = cuda_factory(200, 6, 1)
    e.set_sev(True)
    for m, (s, tips, wgt) in enumerate(parts):
        e.set_partition(m, s, wgt, tips)
    e.commit()
    for m, mc in enumerate(models):
        e.set_model(m, mc)
    e.evaluateIterative(tree.full_traversal())
    used, dense, nbytes = e.sev_stats()
    dense_bytes = 198 * 120000 * 128
    print("SEV blocks", used, "of", dense, "slab bytes", nbytes, "dense arena", dense_bytes)
    assert used < 0.75 * dense and nbytes < 0.8 * dense_bytes       # measured on B200: 64 543 of 93 852 blocks, 2.1 of 3.0 GB


# ---- batched scoring of lazy SPR insertions (exb_score_insertions) ----------------------------------------
@pytest.mark.parametrize("mxtips,widths,kinds,nb,tile,mode,maxtrav", [
    (16, [700], [4], 1, None, "signal", 5),                    # one DNA partition: constants as kernel parameters
    (40, [300, 200], [4, 20], 1, 1024, "iid", 8),              # several chunks per tile, scaling fires, windows of 48 / 96 entries
    (24, [150, 90, 60], [4, 20, 4], 3, 64, "signal", 6),       # -M: one branch slot per partition; small tiles
    (30, [40 + 7 * (i % 5) for i in range(20)], [4] * 20, 1, None, "signal", 4),   # > 16 DNA partitions: constants from shared memory
    (150, [96], [20], 1, None, "iid", 12),                     # AA, hundreds of candidates: the scan is cut into several calls
])
def test_score_insertions_vs_oracle(cuda_factory, oracle_factory, mxtips, widths, kinds, nb, tile, mode, maxtrav):
    """The reference scores one insertion per newviewGeneric + evaluateGeneric pair (testInsertBIG searchAlgo.c:682); the
    engine takes the whole walk of addTraverseBIG as one mixed descriptor.  Same descriptor on the CUDA engine and on the
    oracle (whose orc_score_insertions is pinned to the candidate-by-candidate procedure by tests/test_spr_batch.py)."""
    from examl_b200.likelihood import Likelihood
    from examl_b200.spr import build_scan, remove_node, score_scan
    from spr_util import prunable_nodes
    tree, parts, models = random_case(mxtips, widths, kinds, seed=31, mode=mode, num_branches=nb)
    tree2 = random_case(mxtips, widths, kinds, seed=31, mode=mode, num_branches=nb)[0]
    kw = {} if tile is None else {"tile_sites": tile}
    gpu = build_engine(lambda mx, npart, nbr=1: cuda_factory(mx, npart, nbr, **kw), tree, parts, models)
    cpu = build_engine(oracle_factory, tree2, parts, models)
    assert gpu.can_score_insertions()
    lk, lk2 = Likelihood(tree, gpu, len(parts)), Likelihood(tree2, cpu, len(parts))
    lk.evaluateGeneric(tree.start, True); lk2.evaluateGeneric(tree2.start, True)
    rng = np.random.default_rng(8)
    nodes, nodes2 = prunable_nodes(tree), prunable_nodes(tree2)
    mask = np.array([1] * len(parts), dtype=np.uint8)
    for rnd, pick in enumerate(rng.choice(len(nodes), size=3, replace=False)):
        p, p_2 = nodes[pick], nodes2[pick]
        zqr = [a * b for a, b in zip(p.next.z, p.next.next.z)]
        saved = [(p.next.back, list(p.next.z), p.next.next.back, list(p.next.next.z)) for p in (p, p_2)]
        p1, p2 = remove_node(tree, p, zqr)
        q1, q2 = remove_node(tree2, p_2, zqr)
        for l, a, b in ((lk, p1, p2), (lk2, q1, q2)):
            l.newviewGeneric(a); l.newviewGeneric(b)
        ti, cands = build_scan(tree, p, p1, p2, 1, maxtrav)
        ti2, cands2 = build_scan(tree2, p_2, q1, q2, 1, maxtrav)
        assert [c.number for c in cands] == [c.number for c in cands2] and len(cands) > 0
        if rnd == 2 and len(parts) > 1:
            mask[1] = 0                                            # a masked partition reads 0
        got = score_scan(gpu, ti, len(cands), max_entries=mxtips + 1, execute_model=mask)
        want = score_scan(cpu, ti2, len(cands2), max_entries=mxtips + 1, execute_model=mask)
        assert np.all(np.abs(got - want) <= 5e-12 * np.abs(want)), np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))
        assert np.all(got[:, mask == 0] == 0.0)
        mask[:] = 1
        # the same call again gives the same bits
        assert np.array_equal(got, score_scan(gpu, ti, len(cands), max_entries=mxtips + 1)) or rnd == 2
        from examl_b200.phylo import hookup
        for pp, sv in ((p, saved[0]), (p_2, saved[1])):
            hookup(pp.next, sv[0], sv[1]); hookup(pp.next.next, sv[2], sv[3])
        lk.newviewGeneric(p); lk2.newviewGeneric(p_2)
        a, b = lk.evaluateGeneric(p, False), lk2.evaluateGeneric(p_2, False)
        assert abs(a - b) <= 5e-12 * abs(b)
    # the tree's own CLVs and scalers are what the ordinary path leaves behind
    a, b = lk.evaluateGeneric(tree.start, True), lk2.evaluateGeneric(tree2.start, True)
    assert abs(a - b) <= 5e-12 * abs(b)


def test_score_insertions_refused_where_not_built(cuda_factory):
    """PSR and -S engines say so (the drop-in then scores candidate by candidate)."""
    tree, parts, models = random_case(8, [64], [4], seed=1)
    e = cuda_factory(8, 1, 1)
    e.set_sev(True)
    e.set_partition(0, 4, parts[0][2], parts[0][1]); e.commit(); e.set_model(0, models[0])
    assert not e.can_score_insertions()
    from examl_b200.capi import EngineError
    with pytest.raises(EngineError):
        e.scoreInsertions(tree.full_traversal()[:1], 0)
