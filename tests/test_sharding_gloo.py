"""N > 1 host logic on CPU: two processes (gloo) each own a site slice assigned the way
partitionAssignment.c does for one partition; local per-partition lnL / derivatives from the CPU oracle are
summed with an all-reduce (the role NCCL plays inside exb_evaluate / exb_core) and must equal the unsharded
values -- including a partition that is empty on one rank."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    from examl_b200 import data as exdata
    from oracle.oracle import OracleEngine
    from synth import build_engine, random_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tree, parts, models = random_case(12, [101, 1, 40], [4, 4, 20], seed=31)
    ti = tree.full_traversal()
    shard = []
    for s, tips, wgt in parts:
        lo, hi = exdata.shard_sites(wgt.shape[0], rank, world)
        shard.append((s, np.ascontiguousarray(tips[:, lo:hi]), wgt[lo:hi]))
    eng = build_engine(OracleEngine, tree, shard, models)
    lh = torch.tensor(eng.evaluateIterative(ti))
    dist.all_reduce(lh)                      # evaluateGenericSpecial.c:968-975
    eng.makenewzIterative(tree.partial_traversal(tree.start))
    d1, d2 = eng.execCore([np.log(0.6)])
    d = torch.tensor([d1[0], d2[0]])
    dist.all_reduce(d)                       # makenewzGenericSpecial.c:1241-1248
    # an insertion scan (exb_score_insertions): candidate-major per-partition sums, all-reduced like evaluate's output
    from examl_b200.spr import build_scan, remove_node, score_scan
    from spr_util import prunable_nodes
    p = prunable_nodes(tree)[5]
    p1, p2 = remove_node(tree, p, [a * b for a, b in zip(p.next.z, p.next.next.z)])
    sti, cands = build_scan(tree, p, p1, p2, 1, 4)
    scan = torch.tensor(score_scan(eng, sti, len(cands), 10_000))
    dist.all_reduce(scan)
    if rank == 0:
        q.put((lh.tolist(), d.tolist(), [p[2].shape[0] for p in shard], scan.tolist()))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(oracle_factory):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from synth import build_engine, random_case
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    lh, d, widths0, scan = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert widths0[1] == 1 or widths0[1] == 0           # the 1-site partition lives on exactly one rank
    tree, parts, models = random_case(12, [101, 1, 40], [4, 4, 20], seed=31)
    eng = build_engine(oracle_factory, tree, parts, models)
    want = eng.evaluateIterative(tree.full_traversal())
    np.testing.assert_allclose(lh, want, rtol=1e-12)
    eng.makenewzIterative(tree.partial_traversal(tree.start))
    d1, d2 = eng.execCore([np.log(0.6)])
    np.testing.assert_allclose(d, [d1[0], d2[0]], rtol=1e-10)
    from examl_b200.spr import build_scan, remove_node, score_scan
    from spr_util import prunable_nodes
    p = prunable_nodes(tree)[5]
    p1, p2 = remove_node(tree, p, [a * b for a, b in zip(p.next.z, p.next.next.z)])
    sti, cands = build_scan(tree, p, p1, p2, 1, 4)
    np.testing.assert_allclose(scan, score_scan(eng, sti, len(cands), 10_000), rtol=1e-11)
    assert len(cands) > 3
