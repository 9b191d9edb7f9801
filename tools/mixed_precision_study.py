#!/usr/bin/env python
"""Can the 20-state contraction run on tcgen05 (tf32 / bf16 operands, FP32 accumulators in TMEM) and still meet the
north_star bar (lnL within 1e-6 relative of the FP64 reference)?  A numerical experiment instead of an argument.

The full-traversal likelihood of an AA alignment is recomputed with every matrix product of the path
    Y = (d (.) X) * EI^T          X3 = (Y1 (.) Y2) * EV
done the way a tensor-core kernel with reduced-precision operands would do it, everything else exactly as the CUDA
engine does it, and compared with the FP64 result node by node:

  f64        the reference arithmetic (what traverse_aa_kernel computes with DMMA)
  fp32       operands and accumulators IEEE single (what an ideal 3xTF32 scheme approaches)
  tf32x3     operands split hi + lo into two TF32 numbers (11-bit significands), three tensor-core products
             hi*hi + hi*lo + lo*hi, FP32 accumulation  (the "3xTF32" scheme of CUTLASS)
  bf16x6     operands split into three BF16 pieces, the six products that matter, FP32 accumulation
  tf32x1     operands rounded to TF32 once (what kind::tf32 gives without splitting)

Range: CLV entries span 2^-256 .. 1 between the reference's rescalings -- outside FP32 -- so the reduced modes
rescale EVERY site at EVERY node by an exact power of two ("per-site exponent pre-scaling": the site's largest
entry is brought into [0.5, 1)) and carry the exponent as an integer next to the vector.

Workloads: the reference's own testData/140 (LG, 1041 patterns; golden state dumped from the reference, lnL
-129315.0588145853) and a C4-shaped synthetic case (500 taxa, LG).  Output: profiles/r2_mixed_precision_study.json
and a table on stdout.   usage: python tools/mixed_precision_study.py [--sites 2048]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_util import alignment_of, constants_of, load_json, traversal_of  # noqa: E402

F32 = np.float32


def round_tf32(x: np.ndarray) -> np.ndarray:
    """Round FP32 to TF32 (10 explicit mantissa bits), round-to-nearest-even on the dropped 13 bits."""
    u = x.astype(F32).view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & ~np.uint64(0x1FFF)
    return u.astype(np.uint32).view(F32)


def round_bf16(x: np.ndarray) -> np.ndarray:
    u = x.astype(F32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & ~np.uint64(0xFFFF)
    return u.astype(np.uint32).view(F32)


def mm32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """[n, 20] x [20, 20] with FP32 products and FP32 accumulation in k order (what an MMA accumulator does)."""
    acc = np.zeros((a.shape[0], b.shape[1]), dtype=F32)
    for k in range(a.shape[1]):
        acc = (acc + a[:, k:k + 1] * b[k:k + 1, :]).astype(F32)
    return acc


def product(mode: str, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """a [n, 20] (activations) times b [20, 20] (constants)."""
    if mode == "f64":
        return a @ b
    a32, b32 = a.astype(F32), b.astype(F32)
    if mode == "fp32":
        return mm32(a32, b32)
    if mode == "tf32x1":
        return mm32(round_tf32(a32), round_tf32(b32))
    if mode == "tf32x3":
        ah, bh = round_tf32(a32), round_tf32(b32)
        al, bl = round_tf32((a32 - ah).astype(F32)), round_tf32((b32 - bh).astype(F32))
        return (mm32(al, bh) + mm32(ah, bl) + mm32(ah, bh)).astype(F32)      # small terms first
    if mode == "bf16x6":
        def split3(x):
            h = round_bf16(x); r = (x - h).astype(F32)
            m = round_bf16(r); lo = round_bf16((r - m).astype(F32))
            return h, m, lo
        a0, a1, a2 = split3(a32)
        b0, b1, b2 = split3(b32)
        s = mm32(a2, b0) + mm32(a0, b2) + mm32(a1, b1)
        s = (s + mm32(a1, b0) + mm32(a0, b1)).astype(F32)
        return (s + mm32(a0, b0)).astype(F32)
    raise ValueError(mode)


def traverse(mode: str, state: dict, tips: np.ndarray, wgt: np.ndarray, mc, ti, collect=None):
    """Full-traversal lnL.  CLV x[node]: [n, 4, 20] plus an integer exponent per site (value = x * 2^exp); under f64
    the reference's own scaling rule (all 80 entries < 2^-256 -> * 2^256) is used instead, exp counts in units of 256."""
    S, n = 20, tips.shape[1]
    low = mode != "f64"
    dt = F32 if low else np.float64
    EI, EV = mc.EI.reshape(S, S), mc.EV.reshape(S, S)
    tipv = mc.tipVector.reshape(23, S)
    mxtips = state["mxtips"]
    clv, expo = {}, {}

    def child(num, z):
        """Y[c] = (d_c (.) x_c) EI^T for the child at branch value z: [n, 4, 20]."""
        lz = np.log(max(z, 1e-15))
        out = np.empty((n, 4, S), dtype=dt)
        for c in range(4):
            d = np.exp(mc.EIGN * mc.gammaRates[c] * lz); d[0] = 1.0
            x = tipv[tips[num - 1]] if num <= mxtips else clv[num][:, c, :]
            out[:, c, :] = product(mode, (x.astype(dt) * d.astype(dt)).astype(dt), EI.T)
        return out, (np.zeros(n, dtype=np.int64) if num <= mxtips else expo[num])

    for e in ti[1:]:
        yq, eq = child(e.qNumber, e.qz[0])
        yr, er = child(e.rNumber, e.rz[0])
        z = (yq * yr).astype(dt)
        x3 = np.empty_like(z)
        for c in range(4):
            x3[:, c, :] = product(mode, z[:, c, :], EV)
        ex = eq + er
        if low:                                  # exact power-of-two normalisation of every site
            m = np.abs(x3).max(axis=(1, 2))
            k = np.where(m > 0, np.floor(np.log2(np.maximum(m, 1e-300))).astype(np.int64) + 1, 0)
            x3 = np.ldexp(x3, -k[:, None, None].astype(np.int32)).astype(dt)
            ex = ex + k
        elif e.tipCase != 0:
            small = np.abs(x3).max(axis=(1, 2)) < 2.0 ** -256
            x3[small] *= 2.0 ** 256
            ex = ex - 256 * small.astype(np.int64)
        clv[e.pNumber], expo[e.pNumber] = x3, ex
        if collect is not None:
            collect[e.pNumber] = (x3.astype(np.float64), ex.copy())
    # evaluate at the root branch (p is a tip)
    b = ti[0]
    lz = np.log(max(b.qz[0], 1e-15))
    xa = tipv[tips[b.pNumber - 1]]
    term = np.zeros(n)
    for c in range(4):
        d = np.exp(mc.EIGN * mc.gammaRates[c] * lz); d[0] = 1.0
        if low:
            term += ((xa.astype(dt) * clv[b.qNumber][:, c, :]).astype(dt) * d.astype(dt)).astype(dt).sum(axis=1, dtype=dt).astype(np.float64)
        else:
            term += (xa * clv[b.qNumber][:, c, :] * d).sum(axis=1)
    site = np.log(0.25 * np.abs(term)) + expo[b.qNumber] * np.log(2.0)
    return float((wgt * site).sum()), site


def depth_of(ti, mxtips):
    depth = {}
    for e in ti[1:]:
        depth[e.pNumber] = 1 + max(depth.get(e.qNumber, 0), depth.get(e.rNumber, 0))
    return depth


def study(name, state, tips, wgt, mc, ti, modes):
    ref_nodes = {}
    ref, ref_site = traverse("f64", state, tips, wgt, mc, ti, ref_nodes)
    depth = depth_of(ti, state["mxtips"])
    out = {"case": name, "taxa": state["mxtips"], "patterns": int(tips.shape[1]), "max_depth": max(depth.values()),
           "lnl_f64": ref, "modes": {}}
    for mode in modes:
        nodes = {}
        lnl, site = traverse(mode, state, tips, wgt, mc, ti, nodes)
        # CLV error per node relative to the site's largest entry, after bringing both to the same exponent
        by_depth = {}
        for p, (x, ex) in nodes.items():
            xr, exr = ref_nodes[p]
            a = np.ldexp(x, (ex - ex.max()).astype(np.int32)[:, None, None])          # exponents differ per convention:
            br = np.ldexp(xr, (exr - ex.max()).astype(np.int32)[:, None, None])        # compare true values
            scale = np.abs(br).max(axis=(1, 2))
            ok = scale > 0
            err = (np.abs(a - br).max(axis=(1, 2))[ok] / scale[ok]).max() if ok.any() else 0.0
            by_depth[depth[p]] = max(by_depth.get(depth[p], 0.0), float(err))
        out["modes"][mode] = {"lnl": lnl, "rel_err_lnl": abs(lnl - ref) / abs(ref),
                              "max_abs_err_site_lnl": float(np.abs(site - ref_site).max()),
                              "clv_rel_err_by_depth": {str(k): by_depth[k] for k in sorted(by_depth)}}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=1024, help="sites of the C4-shaped synthetic case")
    args = ap.parse_args()
    modes = ["fp32", "tf32x3", "bf16x6", "tf32x1"]
    results = []

    st = load_json("aa140.first.json")
    part = st["partitions"][0]
    tips, wgt = alignment_of(part, st["mxtips"])
    results.append(study("testData/140 (LG), reference golden state", st, tips, wgt, constants_of(part), traversal_of(st), modes))
    assert abs(results[0]["lnl_f64"] - st["likelihood"]) <= 1e-9 * abs(st["likelihood"]), (results[0]["lnl_f64"], st["likelihood"])

    import bench
    w = bench.make_workload("aa", 1, sites=args.sites)
    t2 = bench.make_sites(w.tree, 0, args.sites, "aa")
    results.append(study(f"C4 shape: synthetic AA, LG+GAMMA, {args.sites} sites x 500 taxa", {"mxtips": w.taxa}, t2,
                         np.ones(args.sites), bench.lg_model(0.5), w.ti, modes))

    for r in results:
        print(f"\n{r['case']}: {r['taxa']} taxa, {r['patterns']} patterns, tree depth {r['max_depth']}, lnL(f64) = {r['lnl_f64']:.9f}")
        print(f"  {'mode':8s} {'rel err lnL':>12s} {'max |d lnL_site|':>17s}   CLV error (rel. to site max) at depth 1 / mid / max")
        for m, v in r["modes"].items():
            ks = list(v["clv_rel_err_by_depth"])
            pick = [ks[0], ks[len(ks) // 2], ks[-1]]
            print(f"  {m:8s} {v['rel_err_lnl']:12.3e} {v['max_abs_err_site_lnl']:17.3e}   " +
                  " / ".join(f"{v['clv_rel_err_by_depth'][k]:.2e}" for k in pick))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "profiles", "r2_mixed_precision_study.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
