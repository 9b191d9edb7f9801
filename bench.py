#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: million site-CLV updates per second
(newview) of a full-traversal evaluateGeneric on synthetic DNA, 1 M sites x 200 taxa, 4 Gamma
rates, 1 partition, GTR+Gamma -- on 1/2/4/8 B200 (sites sharded like partitionAssignment.c, one
NCCL all-reduce per evaluation), next to the reference's own CPU build timed on the host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--sites S] [--taxa T] [--workload dna|aa]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from examl_b200 import data as exdata  # noqa: E402
from examl_b200 import model as exmodel  # noqa: E402
from examl_b200.phylo import Tree  # noqa: E402

TREE_SEED = 20260922
CHUNK = 1 << 16                      # sites are generated in seeded chunks so any rank can make its slice
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
FALLBACK_HBM_GBS = 6650.0            # /opt/skills/guides/B200_PROFILING.md fallback


def make_sites(tree: Tree, lower: int, upper: int, workload: str) -> np.ndarray:
    """Tip codes of global sites [lower, upper): chunk k always comes from seed 1000+k."""
    out = []
    k0, k1 = lower // CHUNK, (upper - 1) // CHUNK
    for k in range(k0, k1 + 1):
        if workload == "dna":
            blk = exdata.synthetic_dna(tree, CHUNK, seed=1000 + k, mode="signal", p_change=0.08, gap_frac=0.01)
        else:
            blk = exdata.synthetic_aa(tree, CHUNK, seed=1000 + k, mode="signal", p_change=0.08, gap_frac=0.01)
        a = max(lower, k * CHUNK) - k * CHUNK
        b = min(upper, (k + 1) * CHUNK) - k * CHUNK
        out.append(blk[:, a:b])
    return np.ascontiguousarray(np.concatenate(out, axis=1))


def model_for(workload: str, alpha: float = 0.5):
    if workload == "dna":
        return exmodel.make_constants(np.array([1.0, 2.0, 1.0, 1.0, 2.0, 1.0]), np.array([0.3, 0.2, 0.2, 0.3]), alpha)
    return exmodel.poisson_aa(alpha=alpha, seed=4)     # generic 20-state GTR-like matrix (LG-shaped work)


def psr_categories(n_sites: int, lower: int = 0):
    """25 per-site rate categories, assigned pseudo-randomly but reproducibly per global site index."""
    rates = np.linspace(0.05, 3.0, 25)
    idx = ((np.arange(lower, lower + n_sites, dtype=np.int64) * 2654435761) >> 7) % 25
    return rates, idx.astype(np.int32)


def algorithmic_bytes(ti, n_sites: int, states: int, cats_per_site: int = 4) -> float:
    """bytes a perfect implementation must move for this traversal."""
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


DMMA_PEAK_TFLOPS = 37.1      # measured on B200 with tools/microbench/fp64_rates.cu (DMMA.8x8x4, 0.258 / clk / SM)


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


def run_reference(tree: Tree, workload: str, sample_sites: int, steps: int, warmup: int, model: str = "gamma"):
    cores = usable_cores()
    tips = make_sites(tree, 0, sample_sites, workload)
    exe = os.path.join(REF_DIR, "examl-OMP-AVX")
    parser = os.path.join(REF_DIR, "parse-examl")
    sample = f"first {sample_sites} sites of the workload, {tree.mxtips} taxa, {steps} full-traversal evaluateGeneric calls"
    if os.path.exists(exe) and os.path.exists(parser):
        with tempfile.TemporaryDirectory() as tmp:
            exdata.write_phylip(os.path.join(tmp, "a.phy"), tree.names, tips, 4 if workload == "dna" else 20)
            open(os.path.join(tmp, "a.tree"), "w").write(tree.to_newick() + "\n")
            pargs = ["-m", "DNA"] if workload == "dna" else ["-m", "PROT", "-q", os.path.join(tmp, "lg.model")]
            if workload != "dna":
                open(os.path.join(tmp, "lg.model"), "w").write(f"LG, p0 = 1-{sample_sites}\n")
            subprocess.run([parser, "-s", "a.phy", "-n", "a", "-c"] + pargs, cwd=tmp, check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
            env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}   # a single process
            env.update(OMP_NUM_THREADS=str(cores), EXAML_PROBE_BENCH=str(steps), EXAML_PROBE_WARMUP=str(warmup))
            r = subprocess.run([exe, "-s", "a.binary", "-t", "a.tree", "-m", "PSR" if model == "psr" else "GAMMA", "-n", "B", "-f", "e", "-w", tmp + "/"],
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
    mc = model_for(workload)
    eng = OracleEngine(tree.mxtips, 1, 1)
    if model == "psr":
        eng.set_rate_model(0, 25)
    eng.set_partition(0, mc.states, np.ones(sample_sites, dtype=np.int32), tips)
    eng.set_model(0, mc)
    if model == "psr":
        eng.set_categories(0, *psr_categories(sample_sites))
    ti = tree.full_traversal()
    for _ in range(warmup):
        eng.evaluateIterative(ti)
    t0 = time.perf_counter()
    for _ in range(steps):
        lnl = eng.evaluateIterative(ti)[0]
    dt = time.perf_counter() - t0
    return {"value": steps * (tree.mxtips - 2) * sample_sites / dt / 1e6, "unit": "M site-updates/s", "cores": 1,
            "kind": "port", "sample": sample + " (oracle/likelihood_oracle.c, single thread)", "seconds": dt,
            "patterns": sample_sites, "lnl": float(lnl)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dna", choices=["dna", "aa"])
    ap.add_argument("--model", default="gamma", choices=["gamma", "psr"],
                    help="rate heterogeneity: GAMMA (the headline config) or PSR/CAT (one rate category per site, 25 categories)")
    ap.add_argument("--sites", type=int, default=None)
    ap.add_argument("--taxa", type=int, default=None)
    ap.add_argument("--tile", type=int, default=None)
    ap.add_argument("--cpu-sample-sites", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    sites = args.sites or (1_000_000 if args.workload == "dna" else 200_000)
    taxa = args.taxa or (200 if args.workload == "dna" else 500)
    states = 4 if args.workload == "dna" else 20
    metric = "million site-lnL updates/sec (newview)"
    workload_name = (f"synthetic {'DNA GTR' if states == 4 else 'AA (20-state GTR-like, LG-shaped)'}+{'GAMMA4' if args.model == 'gamma' else 'PSR (25 per-site rate categories)'}, "
                     f"{sites} sites x {taxa} taxa, 1 partition, full-traversal evaluateGeneric")
    tree = Tree.random(taxa, seed=TREE_SEED)
    ti = tree.full_traversal()
    ntt, nti, nii = tree.tip_case_counts(ti)

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = args.cpu_sample_sites or min(sites, 200_000 if states == 4 else 10_000)
        ref = run_reference(tree, args.workload, sample, max(args.steps, 1), args.warmup, args.model)
        line = {"impl": "reference", "metric": metric, "value": ref["value"], "unit": "M site-updates/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * ref["seconds"] / max(args.steps, 1), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name, "tip_cases_TT_TI_II": [ntt, nti, nii]},
                "cpu_baseline": {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": ref["value"], "unit": "M site-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from examl_b200.capi import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; examl_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    lower, upper = exdata.shard_sites(sites, rank, world)
    t_gen = time.perf_counter()
    tips = make_sites(tree, lower, upper, args.workload)
    t_gen = time.perf_counter() - t_gen
    wgt = np.ones(upper - lower, dtype=np.int32)
    mc = model_for(args.workload)

    eng = Engine(taxa, 1, 1, device=local_rank, tile_sites=args.tile)
    if args.model == "psr":
        eng.set_rate_model(0, 25)
    eng.set_partition(0, states, wgt, tips)
    eng.commit()
    eng.set_model(0, mc)
    if args.model == "psr":
        eng.set_categories(0, *psr_categories(upper - lower, lower))
    if world > 1:
        uid = [eng.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(rank, world, uid[0])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    out = np.zeros(1)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    tip = eng.pack(ti)                      # the descriptor as the C ABI takes it (host arrays)
    for _ in range(max(args.warmup, 3)):
        eng.evaluateIterative(tip, out=out)
    lnl_resident = float(out[0])

    # ---- timed region 1: inputs resident (descriptor identical every step) -> `value`
    #      K steps between two barriers, engine-internal event timing off (it costs an event synchronise per call)
    eng.set_timing(False)
    barrier()
    if sampler:
        sampler.mark_begin()
    launches0 = eng.launch_count()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.evaluateIterative(tip, out=out)
    barrier()
    wall = time.perf_counter() - t0
    launches = eng.launch_count() - launches0

    # ---- timed region 1b: the same K steps with CUDA events around the kernels of every step (engine stream):
    #      the kernel time the roofline is computed from
    eng.set_timing(True)
    barrier()
    dev_ms = 0.0
    for _ in range(args.steps):
        eng.evaluateIterative(tip, out=out)
        dev_ms += eng.last_kernel_ms()
    barrier()
    eng.set_timing(False)

    # ---- timed region 2: end to end through the C ABI with host buffers: every step the host hands over a
    #      new descriptor (a changed branch length) and new model constants (a changed alpha), and reads lnL back
    alphas = [0.5 + 0.01 * (i + 1) for i in range(args.steps)]
    models = [model_for(args.workload, a) for a in alphas]
    ti_var = tree.full_traversal()
    packed = []
    for i in range(args.steps):             # host-side descriptors of the steps (what ExaML's C host code builds per call)
        ti_var[1].qz[0] = 0.9 - 0.001 * i
        packed.append(eng.pack(ti_var))
    barrier()
    t1 = time.perf_counter()
    for i in range(args.steps):
        eng.set_model(0, models[i])
        eng.evaluateIterative(packed[i], out=out)
    barrier()
    wall_e2e = time.perf_counter() - t1
    if sampler:
        sampler.mark_end()
    clocks = sampler.stop() if sampler else None

    times = torch.tensor([wall, wall_e2e, dev_ms / 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    wall, wall_e2e, dev_s = [float(x) for x in times.tolist()]

    if rank == 0:
        updates = float(args.steps) * (taxa - 2) * sites
        model_doubles = (4 + 4 + 16 + 16 + 64) if states == 4 else (20 + 4 + 400 + 400 + 460)
        h2d = (taxa + 1) * 16 + 2 * (taxa + 1) * 8 + 8 + 16 + model_doubles * 8   # staged descriptor + constants
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
        n_local = upper - lower
        abytes = algorithmic_bytes(ti, n_local, states, 4 if args.model == "gamma" else 1)
        achieved = abytes / (dev_s / args.steps) / 1e9
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.workload}:{sites}x{taxa}")
            if tj and world == 1 and args.model == "gamma":
                traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        dram_frac = (traffic / (dev_s / args.steps) / 1e9 / peak) if traffic else None
        if states == 4:
            roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "dram_frac": dram_frac, "kernel": "traverse_dna_kernel",
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)",
                        "algorithmic_bytes_per_launch": abytes,
                        "note": "child CLVs are consumed from registers / L2, so DRAM traffic is ~half of the algorithmic bytes "
                                "and frac can exceed 1; dram_frac = measured DRAM bytes per launch (ncu, profiles/traffic.json) / "
                                "kernel time / peak is the physical utilisation; the design's own ceiling is the CLV write "
                                "stream (see DESIGN.md 5)"}
        else:
            aflops = algorithmic_flops(ti, n_local, states, 4 if args.model == "gamma" else 1)
            tf = aflops / (dev_s / args.steps) / 1e12
            roofline = {"bound": "tensor", "achieved": tf, "peak": DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tf / DMMA_PEAK_TFLOPS,
                        "traffic": traffic, "dram_frac": dram_frac, "kernel": "traverse_aa_kernel",
                        "peak_source": "FP64 tensor pipe (DMMA.8x8x4) measured on this pool with tools/microbench/fp64_rates.cu; "
                                       "MEASURED_PEAKS.json has no FP64 figure",
                        "algorithmic_flops_per_launch": aflops, "hbm_gbs_algorithmic": achieved,
                        "note": "issued flops are 1.2x the algorithmic ones (N padded 20 -> 24 in the 8x8x4 tiles); ncu: DMMA pipe 65 % active"}
        line = {
            "metric": metric, "value": updates / wall / 1e6, "unit": "M site-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name, "tip_cases_TT_TI_II": [ntt, nti, nii],
                       "sites_per_gpu": n_local, "parallelism": f"site-sharded dp{world}",
                       "allreduce": ("none (1 GPU)" if world == 1 else
                                     "inside finish_kernel over CUDA-IPC peer buffers (NVLink)" if eng.comm_is_p2p() else "ncclAllReduce after the kernel"),
                       "l2": "inputs_larger_than_L2 (CLV arena %.1f GB per GPU)" % ((taxa - 2) * n_local * states * 32 / 1e9),
                       "value_timing": "wall clock of the K steps between two barriers (+ cuda synchronize), max over ranks; "
                                       "kernel_ms_per_step: CUDA events around the kernels of each step (engine stream), max over ranks",
                       "lnl": lnl_resident},
            "e2e": {"value": updates / wall_e2e / 1e6, "unit": "M site-updates/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                    "note": "exb_set_model + exb_evaluate with host buffers each step (new descriptor, new model "
                            "constants in, lnL out); alignment resident since start-up like the reference's readMyData"},
            "gpu_launches": launches,
            "kernel_ms_per_step": 1e3 * dev_s / args.steps,
            "kernel_value": updates / dev_s / 1e6,
            "roofline": roofline,
            "clocks": clocks,
            "setup": {"generate_s": t_gen},
        }
        if not args.no_cpu_baseline and world == 1:
            sample = args.cpu_sample_sites or min(sites, 200_000 if states == 4 else 10_000)
            try:
                ref = run_reference(tree, args.workload, sample, 3, 1, args.model)
                line["cpu_baseline"] = {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample")}
                if states == 4 and args.model == "gamma":
                    # proof of work: the same sample through the CUDA engine must give the reference's lnL
                    e2 = Engine(taxa, 1, 1, device=local_rank)
                    e2.set_partition(0, states, np.ones(sample, dtype=np.int32), make_sites(tree, 0, sample, args.workload))
                    e2.commit()
                    # the reference starts from alpha = 1, GTR rates 1, empirical frequencies -- not our constants;
                    # so compare like with like through the C port of the oracle instead
                    from oracle.oracle import OracleEngine
                    o2 = OracleEngine(taxa, 1, 1)
                    o2.set_partition(0, states, np.ones(sample, dtype=np.int32), make_sites(tree, 0, sample, args.workload))
                    o2.set_model(0, mc)
                    e2.set_model(0, mc)
                    g, c = float(e2.evaluateIterative(ti)[0]), float(o2.evaluateIterative(ti)[0])
                    line["parity"] = {"sample_sites": sample, "gpu_lnl": g, "oracle_lnl": c, "rel_err": abs(g - c) / abs(c)}
                    e2.close()
            except Exception as ex:       # the baseline is a reported number, never a reason to lose the bench line
                line["cpu_baseline"] = {"error": repr(ex)[:300]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
