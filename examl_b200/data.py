"""Alignment-side data formats for the likelihood path: tip-code tables, synthetic alignments
(the SURVEY.md 8(d) generators), pattern compression and PHYLIP output (the input format of
the reference's parser, which produces the binary alignment ExaML reads).

Tip codes (globalVariables.h:60-98): DNA is a 4-bit mask A=1 C=2 G=4 T=8 with 15 = gap/N;
AA is 0..19 in the order ARNDCQEGHILKMFPSTWYV, 20 = B, 21 = Z, 22 = gap/X.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from .phylo import Tree

DNA_LETTERS = "_ACMGRSVTWYHKDB-"               # inverseMeaningDNA, globalVariables.h:66
AA_LETTERS = "ARNDCQEGHILKMFPSTWYVBZ-"          # inverseMeaningPROT, globalVariables.h:67
DNA_GAP, AA_GAP = 15, 22


def _evolve(tree: Tree, n_sites: int, rng: np.random.Generator, n_states: int, p_change: float) -> np.ndarray:
    """States 0..n_states-1 at every tip, evolved down the tree from a uniform root: along each
    branch a site keeps its state with probability 1-p_change, else draws a uniform state."""
    out = np.zeros((tree.mxtips, n_sites), dtype=np.uint8)
    root = tree.start.back
    start_states = rng.integers(0, n_states, n_sites, dtype=np.uint8)
    out[tree.start.number - 1] = start_states
    stack = [(root, start_states)]
    while stack:
        node, parent = stack.pop()
        hit = rng.random(n_sites) < p_change
        mine = np.where(hit, rng.integers(0, n_states, n_sites, dtype=np.uint8), parent)
        if tree.isTip(node.number):
            out[node.number - 1] = mine
        else:
            stack.append((node.next.back, mine))
            stack.append((node.next.next.back, mine))
    return out


def synthetic_dna(tree: Tree, n_sites: int, seed: int = 1, mode: str = "signal",
                  p_change: float = 0.08, gap_frac: float = 0.01) -> np.ndarray:
    """uint8 [mxtips, n_sites] DNA tip codes.  mode "signal": evolved down `tree` (scaling never
    fires at 200 taxa); mode "iid": uniform random columns (every site is rescaled)."""
    rng = np.random.default_rng(seed)
    if mode == "signal":
        s = _evolve(tree, n_sites, rng, 4, p_change)
    elif mode == "iid":
        s = rng.integers(0, 4, (tree.mxtips, n_sites), dtype=np.uint8)
    else:
        raise ValueError(mode)
    codes = (np.uint8(1) << s).astype(np.uint8)
    if gap_frac > 0:
        codes[rng.random(codes.shape) < gap_frac] = DNA_GAP
    return codes


def synthetic_aa(tree: Tree, n_sites: int, seed: int = 1, mode: str = "signal",
                 p_change: float = 0.08, gap_frac: float = 0.01) -> np.ndarray:
    """uint8 [mxtips, n_sites] AA tip codes (0..19, 22 = gap)."""
    rng = np.random.default_rng(seed)
    if mode == "signal":
        s = _evolve(tree, n_sites, rng, 20, p_change)
    else:
        s = rng.integers(0, 20, (tree.mxtips, n_sites), dtype=np.uint8)
    if gap_frac > 0:
        s[rng.random(s.shape) < gap_frac] = AA_GAP
    return s


def compress_patterns(tips: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Collapse identical alignment columns: returns (unique columns [mxtips, n_patterns], weights)."""
    cols, inverse = np.unique(tips.T, axis=0, return_inverse=True)
    wgt = np.bincount(inverse.reshape(-1), minlength=cols.shape[0]).astype(np.int32)
    return np.ascontiguousarray(cols.T), wgt


def expand_patterns(tips: np.ndarray, wgt: np.ndarray) -> np.ndarray:
    return np.repeat(tips, np.asarray(wgt, dtype=np.int64), axis=1)


def write_phylip(path: str, names: Sequence[str], tips: np.ndarray, states: int) -> None:
    letters = np.frombuffer((DNA_LETTERS if states == 4 else AA_LETTERS).encode(), dtype=np.uint8)
    with open(path, "wb") as f:
        f.write(f"{tips.shape[0]} {tips.shape[1]}\n".encode())
        for name, row in zip(names, tips):
            f.write(name.encode() + b" " + letters[row].tobytes() + b"\n")


def shard_sites(width: int, rank: int, nranks: int) -> Tuple[int, int]:
    """Contiguous slice [lower, upper) of a single partition owned by `rank`: every rank gets
    floor or ceil of width/nranks sites, as partitionAssignment.c:398 does for one partition."""
    base, rem = divmod(width, nranks)
    lower = rank * base + min(rank, rem)
    return lower, lower + base + (1 if rank < rem else 0)
