#!/usr/bin/env python
"""Regenerate tests/golden/*.json.gz from the REFERENCE ITSELF.

Runs only in the build container (needs /root/reference and oracle/_ref, see oracle/Makefile):
the reference's own AVX binary -- compiled from the unmodified sources, with oracle/ref_probe.c
attached through ld --wrap -- is run on the reference's testData and on seeded synthetic
alignments, and dumps, with %.17g:
  *.first  state after the first full evaluateGeneric (axml.c:2317): model constants, the
           traversal descriptor, lnL, per-partition lnL, root scaler counts, the alignment
           as the reference holds it (pattern-compressed tip codes + weights)
  *.mkz    state after the first makenewzGeneric (descriptor, z0, resulting z)
  *.final  state at exit of `-f e` (optimised model + branch lengths) and its lnL
  *.partial (PSR) state during the first optimizeRateCategories and the first 160 evaluatePartialGeneric
           calls of every partition: (partition, site, rate, value)
  dna49.search.trace.gz   the first 6000 evaluateGeneric calls of the default search (-f d) on testData/49, one line each:
           "E <p> <p->back> <fullTraversal> <tr->likelihood>" -- 561 model-optimisation evaluations, then the lazy SPR
           insertions testInsertBIG scores one by one (searchAlgo.c:735) and the evaluations in between.  The drop-in
           must produce the same sequence, also when the insertions are answered from the batched scan's table.
The parity tests replay those states through the oracle and the CUDA engine.

Usage: python tests/golden/make_goldens.py            (about a minute)
"""
import gzip
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF = "/root/reference"
BIN = os.path.join(ROOT, "oracle", "_ref")

from examl_b200 import data as exdata          # noqa: E402
from examl_b200.phylo import Tree              # noqa: E402


def run(cmd, cwd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, cwd=cwd, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout)
        raise SystemExit(f"command failed: {' '.join(cmd)}")
    return r.stdout


def store(name, obj):
    path = os.path.join(HERE, name + ".json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(obj, separators=(",", ":")).encode())
    print(f"  wrote {os.path.relpath(path, ROOT)} ({os.path.getsize(path)} bytes)  lnL={obj['likelihood']!r}")


def case(tag, alignment, tree_file, parser_args, examl_args, dumps=("first", "mkz", "final"), exit_first=False, model="GAMMA"):
    with tempfile.TemporaryDirectory() as tmp:
        run([os.path.join(BIN, "parse-examl"), "-s", alignment, "-n", "aln"] + parser_args, tmp)
        env = {"EXAML_PROBE_DUMP": os.path.join(tmp, "dump")}
        if exit_first:
            env["EXAML_PROBE_EXIT_FIRST"] = "1"
        out = run([os.path.join(BIN, "examl-AVX"), "-s", "aln.binary", "-t", tree_file, "-m", model,
                   "-n", "G", "-w", tmp + "/"] + examl_args, tmp, env)
        newick = open(tree_file).read().strip()
        for d in dumps:
            obj = json.load(open(os.path.join(tmp, f"dump.{d}.json")))
            if d == "first":
                obj["tree_newick"] = newick
            store(f"{tag}.{d}", obj)
        return out


def search_trace(tag, alignment, tree_file, parser_args, lines=6000, extra=()):
    with tempfile.TemporaryDirectory() as tmp:
        run([os.path.join(BIN, "parse-examl"), "-s", alignment, "-n", "aln"] + parser_args, tmp)
        tf = os.path.join(tmp, "trace")
        run([os.path.join(BIN, "examl-AVX"), "-s", "aln.binary", "-t", tree_file, "-m", "GAMMA", "-n", "T", "-w", tmp + "/", "-f", "d"] + list(extra),
            tmp, {"EXAML_PROBE_TRACE": tf, "EXAML_PROBE_TRACE_MAX": str(lines)})
        path = os.path.join(HERE, tag + ".search.trace.gz")
        open(path, "wb").write(gzip.compress(open(tf, "rb").read(), 9, mtime=0))
        print(f"  wrote {os.path.relpath(path, ROOT)} ({os.path.getsize(path)} bytes)")


def main():
    if not os.path.exists(os.path.join(BIN, "examl-AVX")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", "-j8"])
    td = os.path.join(REF, "testData")
    print("testData/49, default search: the sequence of evaluateGeneric results")
    search_trace("dna49", f"{td}/49", f"{td}/49.tree", ["-m", "DNA"])
    # (no such trace for testData/49 + 49.model under -M: its fourth partition carries no signal for some branches, the
    #  derivatives there are rounding noise of 1e-20 -- the reference's own execCore and the oracle's agree on every other
    #  slot to 12 digits and differ by 40 % on that one --, and a Newton-Raphson step taken from noise sends any two
    #  implementations down different paths that only meet again at the end of the search)
    with tempfile.TemporaryDirectory() as tmp2:
        print("synthetic AA 'signal', 16 taxa x 200 columns, LG: the same for protein data")
        tree = Tree.random(16, seed=19)
        tips = exdata.synthetic_aa(tree, 200, seed=23, mode="signal", p_change=0.15, gap_frac=0.02)
        phy = os.path.join(tmp2, "aa16.phy")
        exdata.write_phylip(phy, tree.names, tips, 20)
        tf = os.path.join(tmp2, "aa16.tree")
        open(tf, "w").write(tree.to_newick() + "\n")
        q = os.path.join(tmp2, "lg.model")
        open(q, "w").write("LG, p0 = 1-200\n")
        search_trace("aa16", phy, tf, ["-m", "PROT", "-q", q], lines=3000)
        # the inputs of this case travel with the trace (the test rebuilds the binary alignment from them)
        import shutil
        shutil.copy(phy, os.path.join(HERE, "aa16.phy")); shutil.copy(tf, os.path.join(HERE, "aa16.tree"))
    with tempfile.TemporaryDirectory() as tmp:
        print("testData/49, DNA, unpartitioned, -f e")
        case("dna49", f"{td}/49", f"{td}/49.tree", ["-m", "DNA"], ["-f", "e"])
        print("testData/49 + 49.model (4 DNA partitions), -f e")
        case("dna49p", f"{td}/49", f"{td}/49.tree", ["-m", "DNA", "-q", f"{td}/49.model"], ["-f", "e"], ("first", "final"))
        print("testData/49 + 49.model, -M (per-partition branch lengths), -f e")
        case("dna49pM", f"{td}/49", f"{td}/49.tree", ["-m", "DNA", "-q", f"{td}/49.model"], ["-f", "e", "-M"])
        print("testData/140, AA, single LG partition, -f e")
        lg = os.path.join(tmp, "lg.model")
        open(lg, "w").write("LG, p0 = 1-1104\n")
        case("aa140", f"{td}/140", f"{td}/140.tree", ["-m", "PROT", "-q", lg], ["-f", "e"])
        print("synthetic iid DNA, 200 taxa x 600 columns (every site is rescaled), first evaluation only")
        tree = Tree.random(200, seed=20260922)
        tips = exdata.synthetic_dna(tree, 600, seed=7, mode="iid", gap_frac=0.01)
        phy = os.path.join(tmp, "iid200.phy")
        exdata.write_phylip(phy, tree.names, tips, 4)
        tf = os.path.join(tmp, "iid200.tree")
        open(tf, "w").write(tree.to_newick() + "\n")
        case("dnaiid200", phy, tf, ["-m", "DNA"], ["-f", "e"], ("first",), exit_first=True)
        print("synthetic AA 'signal', 60 taxa x 300 columns, LG, first evaluation + first makenewz")
        tree = Tree.random(60, seed=11)
        tips = exdata.synthetic_aa(tree, 300, seed=5, mode="signal", p_change=0.15)
        phy = os.path.join(tmp, "aa60.phy")
        exdata.write_phylip(phy, tree.names, tips, 20)
        tf = os.path.join(tmp, "aa60.tree")
        open(tf, "w").write(tree.to_newick() + "\n")
        lg2 = os.path.join(tmp, "lg2.model")
        open(lg2, "w").write("LG, p0 = 1-300\n")
        case("aa60", phy, tf, ["-m", "PROT", "-q", lg2], ["-f", "e"], ("first", "mkz", "final"))
        print("LG4X / LG4M (one eigen-system per GAMMA category + category weights) on the same AA set")
        for mname in ("LG4X", "LG4M"):
            q = os.path.join(tmp, mname + ".model")
            open(q, "w").write(f"{mname}, p0 = 1-300\n")
            case("aa60" + mname.lower(), phy, tf, ["-m", "PROT", "-q", q], ["-f", "e"], ("first", "mkz", "final"))
        print("PSR (per-site rate categories, -m PSR): the same AA set, and testData/49")
        case("aa60psr", phy, tf, ["-m", "PROT", "-q", lg2], ["-f", "e"], ("first", "mkz", "final", "partial"), model="PSR")
        case("dna49psr", f"{td}/49", f"{td}/49.tree", ["-m", "DNA"], ["-f", "e"], ("first", "mkz", "final", "partial"), model="PSR")
        case("dna49ppsr", f"{td}/49", f"{td}/49.tree", ["-m", "DNA", "-q", f"{td}/49.model"], ["-f", "e"], ("first", "final", "partial"), model="PSR")


if __name__ == "__main__":
    main()
