#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: million site-CLV updates per second
(newview) of a full-traversal evaluateGeneric, on 1/2/4/8 B200 (sites / partitions sharded like
partitionAssignment.c, one all-reduce per evaluation), next to the reference's own CPU build timed on
the host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--workload dna|aa|c5] [--no-secondary] [--sites S] [--taxa T]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs[2..4]):
  dna  C3  synthetic DNA GTR+GAMMA4, 1 M sites x 200 taxa, 1 partition           <- the headline line
  aa   C4  synthetic AA LG+GAMMA4, 200 k sites x 500 taxa, 1 partition (FP64 tensor-pipe kernel)
  c5   C5  10 k DNA partitions x 1 k sites, 354-taxon tree (the topology of testData/354.tree), per-partition
           GTR models; 1 250 whole partitions per GPU (= one GPU's share of C5; N = 8 is the full config)
The default invocation prints ONE JSON line for `dna` and carries the other two in its "secondary" object (plus C4 in the
tcgen05 mode and, on one GPU, `spr_scan`: the batched insertion scan of SURVEY 8 f-4 against the candidate-by-candidate path).

A "step" is one evaluateGeneric(tr, tr->start, fullTraversal=TRUE): (taxa-2) newviews over all
sites + evaluate + all-reduce (the reference's own timing loop).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from examl_b200 import data as exdata  # noqa: E402
from examl_b200 import model as exmodel  # noqa: E402
from examl_b200.phylo import Tree  # noqa: E402

TREE_SEED = 20260922
CHUNK = 1 << 16                      # sites are generated in seeded chunks so any rank can make its slice
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")
FALLBACK_HBM_GBS = 6650.0            # /opt/skills/guides/B200_PROFILING.md fallback
DMMA_PEAK_TFLOPS = 37.1              # measured on B200 with tools/microbench/fp64_rates.cu (DMMA.8x8x4, 0.258 / clk / SM)
METRIC = "million site-lnL updates/sec (newview)"
C5_PARTITIONS, C5_WIDTH, C5_PER_GPU = 10_000, 1_000, 1_250


# ------------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------------
def lg_model(alpha: float = 0.5) -> exmodel.ModelConstants:
    """LG as the reference's initGeneric builds it (eigen-system dumped from the reference, tests/golden/
    make_bench_fixtures.py) with the Gamma rates of `alpha`."""
    g = json.load(open(os.path.join(GOLDEN, "lg_eigensystem.json")))
    return exmodel.ModelConstants(20, np.array(g["EIGN"]), np.array(g["EV"]), np.array(g["EI"]), np.array(g["tipVector"]),
                                  exmodel.gamma_rates(alpha), np.array(g["frequencies"]))


def gtr_model(k: int, alpha_shift: float = 0.0) -> exmodel.ModelConstants:
    """One of 8 fixed GTR+GAMMA parameter sets (C5: every partition has its own model; partition m uses set m % 8)."""
    rng = np.random.default_rng(900 + k % 8)
    rates = list(rng.uniform(0.5, 3.0, 5)) + [1.0]
    f = rng.dirichlet(np.full(4, 10.0))
    return exmodel.make_constants(np.array(rates), f, alpha=float(rng.uniform(0.3, 2.0)) + alpha_shift)


def model_for(kind: str, alpha: float = 0.5):
    if kind == "dna":
        return exmodel.make_constants(np.array([1.0, 2.0, 1.0, 1.0, 2.0, 1.0]), np.array([0.3, 0.2, 0.2, 0.3]), alpha)
    return lg_model(alpha)


def make_sites(tree: Tree, lower: int, upper: int, kind: str) -> np.ndarray:
    """Tip codes of global sites [lower, upper): chunk k always comes from seed 1000+k."""
    out = []
    k0, k1 = lower // CHUNK, (upper - 1) // CHUNK
    for k in range(k0, k1 + 1):
        if kind == "dna":
            blk = exdata.synthetic_dna(tree, CHUNK, seed=1000 + k, mode="signal", p_change=0.08, gap_frac=0.01)
        else:
            blk = exdata.synthetic_aa(tree, CHUNK, seed=1000 + k, mode="signal", p_change=0.08, gap_frac=0.01)
        a = max(lower, k * CHUNK) - k * CHUNK
        b = min(upper, (k + 1) * CHUNK) - k * CHUNK
        out.append(blk[:, a:b])
    return np.ascontiguousarray(np.concatenate(out, axis=1))


def tree354() -> Tree:
    return Tree.from_newick(open(os.path.join(GOLDEN, "tree354.newick")).read(), [f"T{i}" for i in range(1, 355)])


def make_workload(name: str, world: int, sites=None, taxa=None, model="gamma") -> SimpleNamespace:
    """Global description of a workload: tree, partitions (global index -> [lower, upper) on the global site axis),
    owner rank of every partition / slice."""
    w = SimpleNamespace(name=name, model=model)
    if name in ("dna", "aa"):
        w.kind = name
        w.states = 4 if name == "dna" else 20
        w.sites = sites or (1_000_000 if name == "dna" else 200_000)
        w.taxa = taxa or (200 if name == "dna" else 500)
        w.tree = Tree.random(w.taxa, seed=TREE_SEED)
        w.nparts = 1
        w.part_width = w.sites
        w.config = "C3" if name == "dna" else "C4"
        w.desc = (f"synthetic {'DNA GTR' if w.states == 4 else 'AA LG'}+{'GAMMA4' if model == 'gamma' else 'PSR (25 per-site rate categories)'}, "
                  f"{w.sites} sites x {w.taxa} taxa, 1 partition, full-traversal evaluateGeneric")
        w.scaling = "strong"
    else:
        w.kind = "dna"
        w.states = 4
        w.taxa = 354
        w.tree = tree354()
        w.part_width = C5_WIDTH
        w.nparts = min(C5_PARTITIONS, C5_PER_GPU * world)
        w.sites = w.nparts * C5_WIDTH
        w.config = "C5"
        w.desc = (f"synthetic partitioned DNA, {w.nparts} partitions x {C5_WIDTH} sites (per-partition GTR+GAMMA4 models), 354-taxon tree "
                  f"(topology of testData/354.tree), {C5_PER_GPU} whole partitions per GPU"
                  + (" = the full C5" if w.nparts == C5_PARTITIONS else f" = one GPU's share of the 10 000-partition C5, x{world}"))
        w.scaling = "weak"
    w.ti = w.tree.full_traversal()
    w.tip_cases = w.tree.tip_case_counts(w.ti)
    return w


def local_slices(w: SimpleNamespace, rank: int, world: int):
    """[(partition, lower, upper)] owned by `rank` on the global site axis.  One partition: a contiguous slice of its
    sites; many equal partitions: whole partitions, contiguous blocks (what partitionAssignment.c:398 arrives at)."""
    if w.nparts == 1:
        lo, up = exdata.shard_sites(w.sites, rank, world)
        return [(0, lo, up)]
    p0, p1 = exdata.shard_sites(w.nparts, rank, world)
    return [(m, m * w.part_width, (m + 1) * w.part_width) for m in range(p0, p1)]


def psr_categories(n_sites: int, lower: int = 0):
    """25 per-site rate categories, assigned pseudo-randomly but reproducibly per global site index."""
    rates = np.linspace(0.05, 3.0, 25)
    idx = ((np.arange(lower, lower + n_sites, dtype=np.int64) * 2654435761) >> 7) % 25
    return rates, idx.astype(np.int32)


def algorithmic_bytes(ti, n_sites: int, states: int, cats_per_site: int = 4) -> float:
    """bytes a perfect per-node implementation must move for this traversal."""
    span = cats_per_site * states * 8
    per = {0: 2 + span, 1: span + 1 + span, 2: 3 * span}
    b = sum(per[e.tipCase] for e in ti[1:]) * n_sites
    b += (span + 5) * n_sites            # evaluate at a tip branch (start is a tip): span + 1 B tip + 4 B weight
    return float(b)


def algorithmic_flops(ti, n_sites: int, states: int, cats_per_site: int = 4) -> float:
    """FP64 flops a tip-table implementation (the reference's, and ours) needs: per site update and category two
    SxS matrix-vector products for inner children (a table look-up for tips), the element-wise product and the SxS
    back-transformation (~9 440 flop per INNER_INNER AA site update, ~352 for DNA)."""
    mv = 2 * states * states * cats_per_site           # one child, all categories of a site
    per = {0: mv + cats_per_site * states, 1: 2 * mv + cats_per_site * states, 2: 3 * mv + cats_per_site * states}
    return float(sum(per[e.tipCase] for e in ti[1:]) * n_sites)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def start(self):
        """Start sampling (call before the warm-up: nvidia-smi needs ~0.2 s to deliver its first line)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = [r for t, r in self.rows if self.t0 is None or (self.t0 - 0.03 <= t <= (self.t1 or 1e30) + 0.03)]
        if not rows:                              # timed region shorter than the sampling interval: nearest samples
            rows = [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], 0.0, set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own OMP-AVX build on the host cores (oracle/_ref),
# bounded sample of the same workload; falls back to the single-threaded C port of the oracle.
# ------------------------------------------------------------------------------------------------------
def usable_cores() -> int:
    """Cores this process may actually run on: affinity mask, capped by the cgroup CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except (OSError, ValueError):
        pass
    return n


def run_reference(w: SimpleNamespace, sample_sites: int, steps: int, warmup: int):
    """The reference's CPU implementation on the first `sample_sites` sites of the workload (C5: the first
    sample_sites / 1000 partitions, with a partition file)."""
    cores = usable_cores()
    tree = w.tree
    if w.nparts > 1:
        sample_sites = max(w.part_width, sample_sites // w.part_width * w.part_width)
    tips = make_sites(tree, 0, sample_sites, w.kind)
    exe = os.path.join(REF_DIR, "examl-OMP-AVX")
    parser = os.path.join(REF_DIR, "parse-examl")
    what = ("all" if sample_sites == w.sites else "first") + f" {sample_sites} sites"
    if w.nparts > 1:
        what += f" = {sample_sites // w.part_width} partitions"
    sample = f"{what} of the workload, {tree.mxtips} taxa, {steps} full-traversal evaluateGeneric calls"
    if os.path.exists(exe) and os.path.exists(parser):
        with tempfile.TemporaryDirectory() as tmp:
            exdata.write_phylip(os.path.join(tmp, "a.phy"), tree.names, tips, w.states)
            del tips
            open(os.path.join(tmp, "a.tree"), "w").write(tree.to_newick() + "\n")
            pargs = ["-m", "DNA"] if w.kind == "dna" else ["-m", "PROT", "-q", os.path.join(tmp, "a.model")]
            if w.kind != "dna":
                open(os.path.join(tmp, "a.model"), "w").write(f"LG, p0 = 1-{sample_sites}\n")
            elif w.nparts > 1:
                with open(os.path.join(tmp, "a.model"), "w") as f:
                    for m in range(sample_sites // w.part_width):
                        f.write(f"DNA, g{m} = {m * w.part_width + 1}-{(m + 1) * w.part_width}\n")
                pargs += ["-q", os.path.join(tmp, "a.model")]
            subprocess.run([parser, "-s", "a.phy", "-n", "a", "-c"] + pargs, cwd=tmp, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
            env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}   # a single process
            env.update(OMP_NUM_THREADS=str(cores), EXAML_PROBE_BENCH=str(steps), EXAML_PROBE_WARMUP=str(warmup))
            r = subprocess.run([exe, "-s", "a.binary", "-t", "a.tree", "-m", "PSR" if w.model == "psr" else "GAMMA", "-n", "B", "-f", "e", "-w", tmp + "/"],
                               cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            for line in r.stdout.splitlines():
                if line.startswith("PROBE_BENCH "):
                    d = json.loads(line[len("PROBE_BENCH "):])
                    return {"value": d["site_updates_per_s"] / 1e6, "unit": "M site-updates/s", "cores": cores,
                            "kind": "reference", "sample": sample + " (examl-OMP-AVX, unmodified reference sources, "
                            f"OMP_NUM_THREADS={cores}; no MPI in the image)", "seconds": d["seconds"],
                            "patterns": d["sites"], "lnl": d["likelihood"]}
            raise RuntimeError("reference run produced no PROBE_BENCH line:\n" + r.stdout[-2000:])
    # port: oracle/likelihood_oracle.c, one core
    from oracle.oracle import OracleEngine
    np_ = sample_sites // w.part_width if w.nparts > 1 else 1
    eng = OracleEngine(tree.mxtips, np_, 1)
    if w.model == "psr":
        eng.set_rate_model(0, 25)
    for m in range(np_):
        blk = tips[:, m * w.part_width:(m + 1) * w.part_width] if w.nparts > 1 else tips
        eng.set_partition(m, w.states, np.ones(blk.shape[1], dtype=np.int32), np.ascontiguousarray(blk))
        eng.set_model(m, gtr_model(m) if w.nparts > 1 else model_for(w.kind))
    if w.model == "psr":
        eng.set_categories(0, *psr_categories(sample_sites))
    for _ in range(warmup):
        eng.evaluateIterative(w.ti)
    t0 = time.perf_counter()
    for _ in range(steps):
        lnl = eng.evaluateIterative(w.ti).sum()
    dt = time.perf_counter() - t0
    return {"value": steps * (tree.mxtips - 2) * sample_sites / dt / 1e6, "unit": "M site-updates/s", "cores": 1,
            "kind": "port", "sample": sample + " (oracle/likelihood_oracle.c, single thread)", "seconds": dt,
            "patterns": sample_sites, "lnl": float(lnl)}


def default_cpu_sample(w: SimpleNamespace, full: bool) -> int:
    """Sites of the CPU run.  The reference arm of the headline config runs the WHOLE workload (1 M sites: 25 GB of
    CLVs fit the host); AA / C5 and the in-line cpu_baseline use a bounded sample."""
    if w.name == "dna":
        return w.sites if full else min(w.sites, 200_000)
    if w.name == "aa":
        return min(w.sites, 12_500)            # 1/16 of C4
    return min(w.sites, 78 * w.part_width)     # ~1/16 of one GPU's share of C5


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def parity_check(w: SimpleNamespace, local_rank: int, lnl_parts: np.ndarray):
    """Proof of work: the CUDA path against the C port of the oracle on a bounded sample of THIS workload's shape
    (same tree, same taxa, same models).  dna / aa: a fresh single-GPU engine on the first n sites; c5: the
    per-partition lnL of the timed engine itself for the first 8 partitions (8 distinct models)."""
    from examl_b200.capi import Engine
    from oracle.oracle import OracleEngine
    if w.nparts > 1:
        np_ = 8
        tips = make_sites(w.tree, 0, np_ * w.part_width, w.kind)
        o = OracleEngine(w.taxa, np_, 1)
        for m in range(np_):
            o.set_partition(m, 4, np.ones(w.part_width, dtype=np.int32), np.ascontiguousarray(tips[:, m * w.part_width:(m + 1) * w.part_width]))
            o.set_model(m, gtr_model(m))
        want = o.evaluateIterative(w.ti)
        got = lnl_parts[:np_]
        return {"checked": f"per-partition lnL of partitions 0..7 of the timed engine ({w.part_width} sites each, 8 distinct models) vs oracle",
                "gpu_lnl": float(got.sum()), "oracle_lnl": float(want.sum()),
                "rel_err": float(np.max(np.abs(got - want) / np.abs(want)))}
    n = min(w.sites, 200_000 if w.states == 4 else 3_072)
    tips = make_sites(w.tree, 0, n, w.kind)
    mc = model_for(w.kind)
    e2 = Engine(w.taxa, 1, 1, device=local_rank)
    if getattr(w, "aa_precision", 0):
        e2.set_aa_precision(w.aa_precision)
    o2 = OracleEngine(w.taxa, 1, 1)
    for e in (e2, o2):
        e.set_partition(0, w.states, np.ones(n, dtype=np.int32), tips)
    e2.commit()
    e2.set_model(0, mc)
    o2.set_model(0, mc)
    g, c = float(e2.evaluateIterative(w.ti)[0]), float(o2.evaluateIterative(w.ti)[0])
    e2.close()
    return {"checked": f"first {n} sites x {w.taxa} taxa, full traversal, fresh engine vs oracle", "sample_sites": n,
            "gpu_lnl": g, "oracle_lnl": c, "rel_err": abs(g - c) / abs(c)}


def run_b200(w: SimpleNamespace, args, rank: int, world: int, local_rank: int, steps: int, warmup: int,
             sampler=None, cpu_baseline=True, parity=True) -> dict | None:
    import torch
    import torch.distributed as dist
    from examl_b200.capi import Engine

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    mine = local_slices(w, rank, world)
    t_gen = time.perf_counter()
    lo, up = mine[0][1], mine[-1][2]
    tips = make_sites(w.tree, lo, up, w.kind)
    t_gen = time.perf_counter() - t_gen
    n_local = up - lo

    t_setup = time.perf_counter()
    eng = Engine(w.taxa, w.nparts, 1, device=local_rank, tile_sites=args.tile)
    if getattr(w, "aa_precision", 0):
        eng.set_aa_precision(w.aa_precision)
    if w.model == "psr":
        eng.set_rate_model(0, 25)
    owned = {m for m, _, _ in mine}
    empty_w, empty_t = np.zeros(0, dtype=np.int32), np.zeros((w.taxa, 0), dtype=np.uint8)
    for m in range(w.nparts):
        if m not in owned:
            eng.set_partition(m, w.states, empty_w, empty_t)       # another rank's partition: width 0 here
    for m, a, b in mine:
        eng.set_partition(m, w.states, np.ones(b - a, dtype=np.int32), np.ascontiguousarray(tips[:, a - lo:b - lo]))
    del tips
    eng.commit()
    models = [gtr_model(m) for m in range(8)] if w.nparts > 1 else [model_for(w.kind)]
    for m in range(w.nparts):
        eng.set_model(m, models[m % len(models)])
    if w.model == "psr":
        eng.set_categories(0, *psr_categories(n_local, lo))
    if world > 1:
        uid = [eng.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(rank, world, uid[0])
    t_setup = time.perf_counter() - t_setup

    out = np.zeros(w.nparts)
    ti = w.ti
    tip = eng.pack(ti)                      # the descriptor as the C ABI takes it (host arrays)
    for _ in range(max(warmup, 3)):
        eng.evaluateIterative(tip, out=out)
    lnl_parts = out.copy()

    # ---- timed region 1: inputs resident (descriptor identical every step) -> `value`
    #      K steps between two barriers.  The engine's CUDA events (call start | traverse start | traverse end | call
    #      end, on its own stream) are recorded in the SAME steps, so that `value` and the kernel times the roofline
    #      is computed from see the same clocks (a box under sw_power_cap slows down within ~100 ms of load); reading
    #      them costs one cudaEventSynchronize on an already finished call per step
    eng.set_timing(True)
    barrier()
    if sampler:
        sampler.mark_begin()
    launches0 = eng.launch_count()
    dev_ms, phases = 0.0, np.zeros(3)
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.evaluateIterative(tip, out=out)
        dev_ms += eng.last_kernel_ms()
        phases += np.array(eng.last_phase_ms())
    barrier()
    wall = time.perf_counter() - t0
    launches = eng.launch_count() - launches0
    eng.set_timing(False)

    # ---- timed region 2: end to end through the C ABI with host buffers: every step the host hands over a
    #      new descriptor (a changed branch length) and new model constants (a changed alpha; C5: of 8 partitions),
    #      and reads the per-partition lnL back
    if w.nparts > 1:
        step_models = [[gtr_model(m, 0.01 * (i + 1)) for m in range(8)] for i in range(steps)]
    else:
        step_models = [[model_for(w.kind, 0.5 + 0.01 * (i + 1))] for i in range(steps)]
    ti_var = w.tree.full_traversal()
    packed = []
    for i in range(steps):                  # host-side descriptors of the steps (what ExaML's C host code builds per call)
        ti_var[1].qz[0] = 0.9 - 0.001 * i
        packed.append(eng.pack(ti_var))
    reset = sorted(owned)[:8] if w.nparts > 1 else [0]
    barrier()
    t1 = time.perf_counter()
    for i in range(steps):
        for j, m in enumerate(reset):
            eng.set_model(m, step_models[i][j % len(step_models[i])])
        eng.evaluateIterative(packed[i], out=out)
    barrier()
    wall_e2e = time.perf_counter() - t1
    if sampler:
        sampler.mark_end()
    is_p2p = eng.comm_is_p2p() if world > 1 else False

    times = torch.tensor([wall, wall_e2e, dev_ms / 1e3] + list(phases / 1e3), dtype=torch.float64, device="cuda")
    per_rank = None
    if world > 1:
        gathered = [torch.empty_like(times) for _ in range(world)]
        dist.all_gather(gathered, times)
        per_rank = [[round(1e3 * float(v) / steps, 4) for v in g.tolist()] for g in gathered]   # ms per step, every rank
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    wall, wall_e2e, dev_s, ph_tab, ph_trav, ph_fin = [float(x) for x in times.tolist()]

    res = None
    if rank == 0:
        updates = float(steps) * (w.taxa - 2) * w.sites
        model_doubles = (4 + 4 + 16 + 16 + 64) if w.states == 4 else (20 + 4 + 400 + 400 + 460)
        h2d = (w.taxa + 1) * 16 + 2 * (w.taxa + 1) * 8 + 8 + 16 + w.nparts + len(reset) * model_doubles * 8   # staged descriptor + mask + constants
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
        cats = 4 if w.model == "gamma" else 1
        trav_s = ph_trav / steps                           # the dominant kernel(s): the traverse launch(es) of one step
        abytes = algorithmic_bytes(ti, n_local, w.states, cats)
        achieved = abytes / trav_s / 1e9
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{w.name}:{w.sites}x{w.taxa}")
            if tj and world == 1 and w.model == "gamma" and not getattr(w, "aa_precision", 0):
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        dram_frac = (traffic / trav_s / 1e9 / peak) if traffic else None
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
        clv_write_bytes = float((w.taxa - 2) * n_local * w.states * cats * 8)
        if w.states == 4:
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "dram_frac": dram_frac,
                        "kernel": "traverse_dna_kernel", "kernel_ms": 1e3 * trav_s, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": abytes,
                        "write_floor_frac": clv_write_bytes / trav_s / 1e9 / peak,
                        "note": "child CLVs are consumed from registers / L2, so DRAM traffic is ~half of the algorithmic bytes "
                                "and frac can exceed 1; dram_frac = measured DRAM bytes per launch (ncu, profiles/traffic.json) / "
                                "kernel time / peak is the physical utilisation; write_floor_frac = mandatory CLV write stream / "
                                "kernel time / peak (the design's own ceiling, DESIGN.md 5)"}
        else:
            aflops = algorithmic_flops(ti, n_local, w.states, cats)
            tf = aflops / trav_s / 1e12
            roofline = {"bound": "tensor", "achieved": tf, "peak": DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tf / DMMA_PEAK_TFLOPS,
                        "traffic": traffic, "dram_frac": dram_frac,
                        "kernel": "traverse_aa_tc_kernel (tcgen05 kind::tf32; frac is against the FP64 DMMA peak for comparison)"
                                  if getattr(w, "aa_precision", 0) else "traverse_aa_kernel", "kernel_ms": 1e3 * trav_s,
                        "peak_source": "FP64 tensor pipe (DMMA.8x8x4) measured on this pool with tools/microbench/fp64_rates.cu; "
                                       "MEASURED_PEAKS.json has no FP64 figure",
                        "algorithmic_flops_per_launch": aflops, "hbm_gbs_algorithmic": achieved,
                        "hbm_frac_algorithmic": achieved / peak,
                        "note": "issued flops are 1.2x the algorithmic ones (N padded 20 -> 24 in the 8x8x4 tiles)"}
        allreduce = ("none (1 GPU)" if world == 1 else
                     "inside finish_kernel over CUDA-IPC peer buffers (NVLink)" if is_p2p else "ncclAllReduce after the kernel")
        res = {
            "metric": METRIC, "value": updates / wall / 1e6, "unit": "M site-updates/s",
            "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
            "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": w.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w.desc, "baseline_config": w.config, "tip_cases_TT_TI_II": w.tip_cases,
                       "partitions": w.nparts, "sites_per_gpu": n_local,
                       "parallelism": (f"site-sharded dp{world}" if w.nparts == 1 else f"whole partitions per rank, dp{world}"),
                       "allreduce": allreduce, "allreduce_bytes": 8 * w.nparts,
                       "l2": "inputs_larger_than_L2 (CLV arena %.1f GB per GPU)" % ((w.taxa - 2) * n_local * w.states * cats * 8 / 1e9),
                       "value_timing": "wall clock of the K steps between two barriers (+ cuda synchronize), max over ranks; "
                                       "kernel_ms_per_step / kernel_phases_ms: CUDA events around the kernels of the same K steps "
                                       "(engine stream), max over ranks",
                       "lnl": float(lnl_parts.sum())},
            "e2e": {"value": updates / wall_e2e / 1e6, "unit": "M site-updates/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * w.nparts,
                    "note": "exb_set_model + exb_evaluate with host buffers each step (new descriptor, new model "
                            "constants in, per-partition lnL out); alignment resident since start-up like the reference's readMyData"},
            "gpu_launches": launches,
            "kernel_ms_per_step": 1e3 * dev_s / steps,
            "kernel_phases_ms": {"tables": 1e3 * ph_tab / steps, "traverse": 1e3 * ph_trav / steps, "finish": 1e3 * ph_fin / steps},
            "kernel_value": updates / dev_s / 1e6,
            "roofline": roofline,
            "setup": {"generate_s": t_gen, "engine_s": t_setup},
        }
        if per_rank:
            res["per_rank_ms_per_step"] = {"columns": ["wall", "wall_e2e", "kernels", "tables", "traverse", "finish"], "ranks": per_rank}
    # free the arena before anything else allocates (parity engine, next workload)
    eng.close()
    if rank == 0:
        if parity:
            try:
                res["parity"] = parity_check(w, local_rank, lnl_parts)
            except Exception as ex:
                res["parity"] = {"error": repr(ex)[:300]}
        if cpu_baseline:
            try:
                ref = run_reference(w, args.cpu_sample_sites or default_cpu_sample(w, full=False), 3, 1)
                res["cpu_baseline"] = {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as ex:       # the baseline is a reported number, never a reason to lose the bench line
                res["cpu_baseline"] = {"error": repr(ex)[:300]}
    return res


def secondary_summary(r: dict) -> dict:
    """What the headline line carries about another workload."""
    keep = ("value", "unit", "ms_per_step", "steps", "kernel_ms_per_step", "kernel_phases_ms", "kernel_value", "scaling",
            "roofline", "parity", "cpu_baseline", "gpu_launches", "per_rank_ms_per_step")
    s = {k: r[k] for k in keep if k in r}
    s["config"] = r["config"]
    s["e2e"] = {k: r["e2e"][k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")}
    s["kernel_ms"] = r["roofline"]["kernel_ms"]
    s["finish_ms"] = r["kernel_phases_ms"]["finish"]
    s["allreduce_bytes"] = r["config"]["allreduce_bytes"]
    if r["config"]["baseline_config"] == "C5":
        ph = r["kernel_phases_ms"]
        host = r["ms_per_step"] - r["kernel_ms_per_step"]
        parts = {"traverse kernel": ph["traverse"], "finish kernel (scaler chains, reduction, all-reduce)": ph["finish"],
                 "d-table kernel": ph["tables"], "host + launch + copies": host}
        s["limiter"] = max(parts, key=parts.get)
        s["step_breakdown_ms"] = parts
    return s


def run_insertion_scan(local_rank: int) -> dict:
    """SURVEY 8 f-4 at scale, single GPU: all lazy SPR insertions of a few pruned subtrees within a rearrangement radius,
    scored (a) by ONE exb_score_insertions call per pruned subtree and (b) the reference's way, one newviewGeneric +
    evaluateGeneric pair per candidate (testInsertBIG searchAlgo.c:682; tests/spr_util.reference_scan) through the same
    engine -- so the two tables must agree to rounding (the pulley principle), and the first one against the oracle is
    tests/test_gpu_parity.py::test_score_insertions_vs_oracle."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from examl_b200.capi import Engine
    from examl_b200.likelihood import Likelihood
    from examl_b200.phylo import hookup
    from examl_b200.spr import build_scan, remove_node, score_scan
    from spr_util import prunable_nodes, reference_scan
    taxa, sites, radius, nprune = 200, 100_000, 5, 6
    tree = Tree.random(taxa, seed=20260922)
    rng = np.random.default_rng(11)
    for a, _b in tree.branches():
        a.z[:] = [float(rng.uniform(0.35, 0.97))]
    tips = make_sites(tree, 0, sites, "dna")
    eng = Engine(taxa, 1, 1, device=local_rank)
    eng.set_partition(0, 4, np.ones(sites, dtype=np.int32), tips)
    eng.commit()
    eng.set_model(0, model_for("dna"))
    eng.set_timing(False)
    lk = Likelihood(tree, eng, 1)
    lk.evaluateGeneric(tree.start, True)
    nodes = prunable_nodes(tree)
    tb = tr = worst = 0.0
    ncand = 0
    launches0 = eng.launch_count()
    for pick in rng.choice(len(nodes), size=nprune, replace=False):
        p = nodes[pick]
        saved = (p.next.back, list(p.next.z), p.next.next.back, list(p.next.next.z))
        p1, p2 = remove_node(tree, p, [a * b for a, b in zip(p.next.z, p.next.next.z)])
        lk.newviewGeneric(p1); lk.newviewGeneric(p2)           # removeNodeBIG's makenewzGeneric leaves both looking at each other
        ti, cands = build_scan(tree, p, p1, p2, 1, radius)
        score_scan(eng, ti, len(cands), eng.max_entries())      # untimed: scratch CLVs are allocated on first use
        t0 = time.perf_counter()
        got = score_scan(eng, ti, len(cands), eng.max_entries())
        tb += time.perf_counter() - t0
        t0 = time.perf_counter()
        want = reference_scan(lk, p, p1, p2, 1, radius)
        tr += time.perf_counter() - t0
        ncand += len(cands)
        for k, (_q, lh) in enumerate(want):
            worst = max(worst, abs(got[k][0] - lh[0]) / abs(lh[0]))
        hookup(p.next, saved[0], saved[1]); hookup(p.next.next, saved[2], saved[3])
        lk.newviewGeneric(p)
    out = {"metric": "lazy SPR insertions scored per second (exb_score_insertions, one call per pruned subtree)",
           "value": ncand / tb, "unit": "insertions/s", "candidates": ncand, "pruned_subtrees": nprune,
           "ms_batched": 1e3 * tb, "ms_candidate_by_candidate": 1e3 * tr, "speedup": tr / tb,
           "parity": {"checked": "batched table vs newviewGeneric + evaluateGeneric per candidate on the same engine",
                      "rel_err": float(worst)},
           "gpu_launches": eng.launch_count() - launches0,
           "config": {"workload": f"synthetic DNA GTR+GAMMA4, {sites} sites x {taxa} taxa, rearrangement radius {radius}",
                      "reference_procedure": "testInsertBIG searchAlgo.c:682 (insertBIG non-thorough + newviewGeneric + evaluateGeneric)"}}
    eng.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dna", choices=["dna", "aa", "c5"])
    ap.add_argument("--model", default="gamma", choices=["gamma", "psr"],
                    help="rate heterogeneity: GAMMA (the headline config) or PSR/CAT (one rate category per site, 25 categories)")
    ap.add_argument("--sites", type=int, default=None)
    ap.add_argument("--taxa", type=int, default=None)
    ap.add_argument("--tile", type=int, default=None)
    ap.add_argument("--cpu-sample-sites", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="only the selected workload (default: the dna line also carries aa (C4) and c5 in `secondary`)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    w = make_workload(args.workload, world, args.sites, args.taxa, args.model)

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = args.cpu_sample_sites or default_cpu_sample(w, full=True)
        ref = run_reference(w, sample, max(args.steps, 1), args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": ref["value"], "unit": "M site-updates/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * ref["seconds"] / max(args.steps, 1), "higher_is_better": True,
                "scaling": w.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w.desc, "baseline_config": w.config, "tip_cases_TT_TI_II": w.tip_cases,
                           "partitions": w.nparts, "cpu_sites": ref["patterns"], "cores": ref["cores"]},
                "cpu_baseline": {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": ref["value"], "unit": "M site-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; examl_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    line = run_b200(w, args, rank, world, local_rank, args.steps, args.warmup, sampler,
                    cpu_baseline=not args.no_cpu_baseline and world == 1)
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        line["clocks"] = clocks

    if args.workload == "dna" and args.model == "gamma" and not args.no_secondary and args.sites is None and args.taxa is None:
        sec = {}
        for name, key in (("aa", "aa_c4"), ("c5", "c5"), ("aa", "aa_c4_tcgen05_tf32x3")):
            try:
                w2 = make_workload(name, world)
                if key.endswith("tf32x3"):
                    # the same C4 workload with the newview on the tcgen05 tensor cores (mixed precision, opt-in)
                    w2.aa_precision = 1
                    w2.desc += " -- newview on tcgen05 in 3xTF32 (exb_set_aa_precision), CLVs stored as FP64"
                r2 = run_b200(w2, args, rank, world, local_rank, min(args.steps, 10), args.warmup, None,
                              cpu_baseline=not args.no_cpu_baseline and world == 1)
                if rank == 0:
                    sec[key] = secondary_summary(r2)
            except Exception as ex:                      # a secondary must never cost the headline line
                if world > 1:
                    raise                                 # ... but ranks must not diverge
                sec[key] = {"error": repr(ex)[:400]}
        if world == 1:
            try:                                          # the batched insertion scan against the per-candidate path
                sec["spr_scan"] = run_insertion_scan(local_rank)
            except Exception as ex:
                sec["spr_scan"] = {"error": repr(ex)[:400]}
        if rank == 0:
            line["secondary"] = sec
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
