"""Helpers to run the reference-side binaries (oracle/_ref/parse-examl, bin/examl-b200-probe) on the
alignments stored in the golden fixtures."""
from __future__ import annotations

import json
import os
import subprocess

import numpy as np

from examl_b200 import data as exdata
from golden_util import alignment_of, load_json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARSER = os.path.join(ROOT, "oracle", "_ref", "parse-examl")
DROPIN = os.path.join(ROOT, "bin", "examl-b200-probe")
DROPIN_PLAIN = os.path.join(ROOT, "bin", "examl-b200")


def have_binaries() -> bool:
    return os.path.exists(PARSER) and os.path.exists(DROPIN)


def write_inputs(tag: str, tmp: str, partition_lines=None):
    """Rebuild a PHYLIP alignment + tree file from <tag>.first.json and parse it to ExaML's binary format.
    The fixture holds the pattern-compressed alignment; expanding each pattern by its weight gives the
    same multiset of columns as the original file."""
    first = load_json(f"{tag}.first.json")
    mxtips = first["mxtips"]
    blocks, states = [], first["partitions"][0]["states"]
    for part in first["partitions"]:
        tips, wgt = alignment_of(part, mxtips)
        blocks.append(exdata.expand_patterns(tips, wgt))
    aln = np.concatenate(blocks, axis=1)
    phy = os.path.join(tmp, f"{tag}.phy")
    exdata.write_phylip(phy, first["names"], aln, states)
    tree = os.path.join(tmp, f"{tag}.tree")
    open(tree, "w").write(first["tree_newick"] + "\n")
    args = [PARSER, "-s", phy, "-n", tag, "-m", "DNA" if states == 4 else "PROT"]
    if partition_lines:
        q = os.path.join(tmp, f"{tag}.model")
        open(q, "w").write("\n".join(partition_lines) + "\n")
        args += ["-q", q]
    subprocess.run(args, cwd=tmp, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
    return os.path.join(tmp, f"{tag}.binary"), tree, first


def run_dropin(binary: str, tree: str, tmp: str, extra_args, env=None, exe=DROPIN, model="GAMMA"):
    e = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    e["EXAML_PROBE_DUMP"] = os.path.join(tmp, "dump")
    e.update(env or {})
    r = subprocess.run([exe, "-s", binary, "-t", tree, "-m", model, "-n", "RUN", "-w", tmp + "/"] + extra_args,
                       cwd=tmp, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    dumps = {}
    for d in ("first", "mkz", "final", "partial"):
        p = os.path.join(tmp, f"dump.{d}.json")
        if os.path.exists(p):
            dumps[d] = json.load(open(p))
    return r, dumps


REFERENCE = os.path.join(ROOT, "oracle", "_ref", "examl-AVX")


def small_dna_case(tmp: str, taxa: int = 10, sites: int = 300, seed: int = 5):
    """A seeded synthetic alignment small enough for the reference's search / quartet modes to run in well under a
    second; returns (binary alignment, start tree)."""
    from examl_b200.phylo import Tree
    tree = Tree.random(taxa, seed=seed)
    tips = exdata.synthetic_dna(tree, sites, seed=seed + 1, mode="signal", p_change=0.15, gap_frac=0.02)
    exdata.write_phylip(os.path.join(tmp, "s.phy"), tree.names, tips, 4)
    open(os.path.join(tmp, "s.tree"), "w").write(tree.to_newick() + "\n")
    subprocess.run([PARSER, "-s", "s.phy", "-n", "s", "-m", "DNA"], cwd=tmp, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
    return os.path.join(tmp, "s.binary"), os.path.join(tmp, "s.tree")


def run_mode(exe: str, binary: str, tree: str, tmp: str, run_id: str, args, env=None, timeout=900):
    """Run one ExaML workflow; returns (CompletedProcess, 'Likelihood of best tree' or None, quartet table or None)."""
    import re
    e = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    e.update(env or {})
    r = subprocess.run([exe, "-s", binary, "-t", tree, "-n", run_id, "-w", tmp + "/"] + args, cwd=tmp, env=e,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    m = re.findall(r"Likelihood of best tree: (-?[0-9.]+)", r.stdout)
    quartets = None
    qf = os.path.join(tmp, f"ExaML_quartets.{run_id}")
    if os.path.exists(qf):
        quartets = [[float(x) for x in re.findall(r"-[0-9]+\.[0-9]+", line)] for line in open(qf) if "|" in line]
        quartets = [q for q in quartets if q]
    return r, (float(m[-1]) if m else None), quartets


def small_aa_case(tmp: str, partition_lines, taxa: int = 8, sites: int = 120, seed: int = 4):
    """Small seeded protein alignment with a partition file (e.g. ["AUTO, p0 = 1-60", "LG4X, p1 = 61-120"])."""
    from examl_b200.phylo import Tree
    tree = Tree.random(taxa, seed=seed)
    tips = exdata.synthetic_aa(tree, sites, seed=seed - 1, mode="signal", p_change=0.2, gap_frac=0.02)
    exdata.write_phylip(os.path.join(tmp, "s.phy"), tree.names, tips, 20)
    open(os.path.join(tmp, "s.tree"), "w").write(tree.to_newick() + "\n")
    open(os.path.join(tmp, "q.model"), "w").write("\n".join(partition_lines) + "\n")
    subprocess.run([PARSER, "-s", "s.phy", "-n", "s", "-m", "PROT", "-q", "q.model"], cwd=tmp, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
    return os.path.join(tmp, "s.binary"), os.path.join(tmp, "s.tree")
