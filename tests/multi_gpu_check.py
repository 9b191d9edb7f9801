"""Launched under torchrun with N ranks (one per GPU): the alignment is sharded by sites exactly like
partitionAssignment.c does for a single partition; every rank must obtain the same, globally reduced lnL and
derivatives (NCCL all-reduce inside exb_evaluate / exb_core), equal to the unsharded values.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

from examl_b200 import data as exdata
from examl_b200.capi import Engine
from examl_b200.likelihood import Likelihood
from synth import build_engine, random_case


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tree, parts, models = random_case(40, [100003, 0, 5000], [4, 4, 20], seed=17)

    def factory(mxtips, nparts, nb=1, **kw):
        return Engine(mxtips, nparts, nb, device=local, **kw)

    # reference values: the whole alignment on this rank's GPU, no communicator
    full = build_engine(factory, tree, parts, models)
    ti = tree.full_traversal()
    want = full.evaluateIterative(ti).copy()
    full.makenewzIterative(tree.partial_traversal(tree.start))
    want_d = full.execCore([np.log(0.7)])
    # the whole Newton-Raphson loop on one GPU (device kernel) and through the host loop
    inner = tree.start.back
    ti_nr = tree.partial_traversal(inner)          # built once: building it re-orients the tree's x flags
    z0 = list(inner.z)
    want_z, want_ok, want_it = full.makenewz(ti_nr, z0, 10)
    host_z = Likelihood(tree, full, len(parts)).makenewzGeneric(inner, inner.back, z0, 10)
    # an insertion scan (exb_score_insertions) on the whole alignment: the table the sharded engines must reproduce
    # (last on this engine: it recomputes CLVs for the pruned topology)
    from examl_b200.spr import build_scan, remove_node, score_scan
    from spr_util import prunable_nodes
    scan_tree = random_case(40, [100003, 0, 5000], [4, 4, 20], seed=17)[0]       # a copy: pruning re-wires it
    def holds_start(rec):                      # is tip 1 (the traversal root) inside the subtree behind rec?
        if scan_tree.isTip(rec.number):
            return rec.number == 1
        return holds_start(rec.next.back) or holds_start(rec.next.next.back)
    sp = [x for x in prunable_nodes(scan_tree) if not holds_start(x.back)][11]
    sp1, sp2 = remove_node(scan_tree, sp, [a * b for a, b in zip(sp.next.z, sp.next.next.z)])
    Likelihood(scan_tree, full, len(parts)).evaluateGeneric(scan_tree.start, True)
    scan_ti, scan_cands = build_scan(scan_tree, sp, sp1, sp2, 1, 4)
    want_scan = score_scan(full, scan_ti, len(scan_cands), full.max_entries())
    full.close()

    # sharded: rank r owns a contiguous slice of every partition (partition 1 is empty everywhere)
    shard = []
    for s, tips, wgt in parts:
        lo, hi = exdata.shard_sites(wgt.shape[0], rank, world)
        shard.append((s, np.ascontiguousarray(tips[:, lo:hi]), wgt[lo:hi]))
    eng = build_engine(factory, tree, shard, models)
    uid = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(rank, world, uid[0])
    got = eng.evaluateIterative(ti).copy()
    eng.makenewzIterative(tree.partial_traversal(tree.start))
    got_d = eng.execCore([np.log(0.7)])
    # back-to-back exchanges (the receive buffers are double-buffered by exchange number): same bits every time
    stable = True
    for it in range(200):
        again = eng.evaluateIterative(ti[:1] if it % 2 else ti)
        stable = stable and bool(np.array_equal(again, got))
        if it % 50 == 0:
            d = eng.execCore([np.log(0.7)])
            stable = stable and d[0][0] == got_d[0][0] and d[1][0] == got_d[1][0]
    mode = "p2p-in-kernel" if eng.comm_is_p2p() else "nccl"
    # device-resident Newton-Raphson across the ranks (one in-kernel exchange per iteration): same decisions as on one
    # GPU, same bits on every rank, and the ordinary calls keep working afterwards (the exchange counter moved on)
    nr_ok, got_z, got_it = True, np.zeros(1), -1
    if eng.comm_is_p2p():
        got_z, got_ok, got_it = eng.makenewz(ti_nr, z0, 10)
        nr_ok = got_it == want_it and abs(got_z[0] - want_z[0]) <= 1e-9 * want_z[0] and abs(got_z[0] - host_z[0]) <= 1e-9 * host_z[0]
        for _ in range(3):
            z2, _, it2 = eng.makenewz(ti_nr, z0, 10)
            nr_ok = nr_ok and z2[0] == got_z[0] and it2 == got_it
        nr_ok = nr_ok and bool(np.array_equal(eng.evaluateIterative(ti), got))
    else:
        try:
            eng.makenewz(ti_nr, z0, 10)
            nr_ok = False                       # must refuse without peer access
        except Exception:
            pass

    # the insertion scan across the ranks: per-candidate sums are all-reduced inside the call
    eng.evaluateIterative(scan_tree.full_traversal())       # the pruned topology's CLVs on the sharded engine
    got_scan = score_scan(eng, scan_ti, len(scan_cands), eng.max_entries())
    scan_ok = len(scan_cands) > 3 and bool(np.all(np.abs(got_scan - want_scan) <= 1e-11 * np.maximum(np.abs(want_scan), 1.0)))
    scan_ok = scan_ok and bool(np.array_equal(eng.evaluateIterative(ti), got))

    ok = stable and nr_ok and scan_ok and bool(np.all(np.abs(got - want) <= 1e-11 * np.maximum(np.abs(want), 1.0)))
    ok = ok and abs(got_d[0][0] - want_d[0][0]) <= 1e-9 * max(1.0, abs(want_d[0][0]))
    ok = ok and abs(got_d[1][0] - want_d[1][0]) <= 1e-9 * max(1.0, abs(want_d[1][0]))
    # every rank holds bit-identical results (the all-reduce is part of the call)
    t = torch.tensor(np.concatenate([got, got_d[0], got_d[1], got_z]), device="cuda")
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    ok = ok and all(torch.equal(gathered[0], g) for g in gathered)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(flag.item()) else "FAIL", "allreduce", mode, "world", world, "lnL", got.tolist(), "want", want.tolist(),
              "d", [float(got_d[0][0]), float(got_d[1][0])], "want_d", [float(want_d[0][0]), float(want_d[1][0])],
              "device NR z", float(got_z[0]), "iterations", got_it, "1-GPU", float(want_z[0]), want_it, "host loop", float(host_z[0]),
              "insertion scan", len(scan_cands), "candidates", "ok" if scan_ok else "MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
