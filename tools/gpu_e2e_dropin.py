"""Whole-program comparison on a mid-size data set: ExaML's own `-f e` (model + branch-length optimisation on a fixed
tree: thousands of evaluateGeneric / makenewzGeneric / newviewGeneric calls) through the drop-in binary on one B200 vs
the unmodified reference (OMP-AVX, all usable cores).   usage: tools/gpu_e2e_dropin.py [taxa] [sites] [mode] [GAMMA|PSR]"""
import os, sys, time, subprocess, tempfile, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from examl_b200 import data as exdata
from examl_b200.phylo import Tree
sys.path.insert(0, ROOT); import bench

taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 200
sites = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
mode = sys.argv[3] if len(sys.argv) > 3 else "e"
rate_model = sys.argv[4] if len(sys.argv) > 4 else "GAMMA"
tree = Tree.random(taxa, seed=3)
tips = exdata.synthetic_dna(tree, sites, seed=9, mode="signal", p_change=0.08)
with tempfile.TemporaryDirectory() as tmp:
    exdata.write_phylip(os.path.join(tmp, "a.phy"), tree.names, tips, 4)
    open(os.path.join(tmp, "a.tree"), "w").write(tree.to_newick() + "\n")
    subprocess.run([os.path.join(ROOT, "oracle/_ref/parse-examl"), "-s", "a.phy", "-n", "a", "-m", "DNA"], cwd=tmp, check=True,
                   stdout=subprocess.DEVNULL)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    res = {}
    for name, exe, extra_env in [x for x in (("b200", os.path.join(ROOT, "bin/examl-b200-probe"), {}),
                                 ("reference", os.path.join(ROOT, "oracle/_ref/examl-OMP-AVX"), {"OMP_NUM_THREADS": str(bench.usable_cores())})) if not (x[0] == "reference" and os.environ.get("E2E_SKIP_REF"))]:
        e = dict(env); e.update(extra_env); e["EXAML_PROBE_DUMP"] = os.path.join(tmp, name); e["EXAML_PROBE_NOTIPS"] = "1"
        t0 = time.time()
        r = subprocess.run([exe, "-s", "a.binary", "-t", "a.tree", "-m", rate_model, "-n", name, "-f", mode, "-w", tmp + "/"], cwd=tmp, env=e,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        dt = time.time() - t0
        lnl = json.load(open(os.path.join(tmp, name + ".final.json")))["likelihood"] if r.returncode == 0 else None
        res[name] = (dt, lnl)
        print(f"{name:10s} -m {rate_model} -f {mode}: wall {dt:8.1f} s  final lnL {lnl!r}  rc={r.returncode}", flush=True)
        if r.returncode: print(r.stdout[-1500:])
    if all(v[1] for v in res.values()):
        print("speed-up %.1fx (reference on %d cores), lnL rel diff %.2e" % (res["reference"][0] / res["b200"][0], bench.usable_cores(),
              abs(res["b200"][1] - res["reference"][1]) / abs(res["reference"][1])))
