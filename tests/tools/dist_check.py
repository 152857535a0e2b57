"""torchrun --nproc-per-node N scripts/dist_check.py -- sharded overlap join on N GPUs vs the CPU oracle.
Every rank holds a random slice of both tables (mixed contigs); after the contig all-to-all each rank joins its
contigs; the union of the per-rank (global_read, global_variant) pairs must equal the oracle's pair set."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402  (checker only)
from polars_bio_b200 import dist as pbd, engine  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_contigs = 11
    rng = np.random.default_rng(7)  # same global tables on every rank; each takes its slice
    N, M = 200_000, 60_000
    pc = rng.integers(-1, n_contigs, N).astype(np.int32); ps = rng.integers(0, 2_000_000, N).astype(np.int32)
    pe = (ps + rng.integers(1, 300, N)).astype(np.int32)
    bc = rng.integers(0, n_contigs - 1, M).astype(np.int32); bs = rng.integers(0, 2_000_000, M).astype(np.int32)
    be = (bs + rng.integers(1, 3000, M)).astype(np.int32)
    sl = lambda a, n: a[n * rank // world: n * (rank + 1) // world]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    lp = [t(sl(x, N)) for x in (pc, ps, pe)]
    lb = [t(sl(x, M)) for x in (bc, bs, be)]
    # path 1: table-at-a-time primitives
    hist = pbd.contig_histogram(lp[0], n_contigs) + pbd.contig_histogram(lb[0], n_contigs)
    owner = pbd.owner_table(hist, world)
    pbase, ptotal = pbd.row_id_base(lp[0].numel(), dev)
    bbase, btotal = pbd.row_id_base(lb[0].numel(), dev)
    assert ptotal == N and btotal == M
    q1 = pbd.shard_table(*lp, n_contigs, owner, pbase)
    x1 = pbd.shard_table(*lb, n_contigs, owner, bbase)
    # path 2: the fused exchange (one all_reduce, one count all-to-all) must deliver the same rows
    os.environ["PBGPU_EXCHANGE"] = "nccl"
    (q2, x2), owner2 = pbd.shard_tables([tuple(lp), tuple(lb)], n_contigs)
    assert torch.equal(owner, owner2.cpu())
    for u, v in ((q1, q2), (x1, x2)):
        for col_u, col_v in zip(u, v):
            assert torch.equal(col_u, col_v)
    # path 3: the exchange over NVLink peer memory (CUDA IPC arenas, plan + scatter kernels): same rows, same order,
    # over several steps (both arena parities) and after the arenas had to grow
    os.environ["PBGPU_EXCHANGE"] = "peer"
    for it in range(4):
        (q3, x3), owner3 = pbd.shard_tables([tuple(lp), tuple(lb)], n_contigs)
        assert torch.equal(owner, owner3.cpu())
        for u, v in ((q2, q3), (x2, x3)):
            for col_u, col_v in zip(u, v):
                assert torch.equal(col_u, col_v), ("peer exchange differs from the NCCL exchange", it)
    for it in range(3):  # overlapped mode: every table on its own stream, an event per table
        ready = []
        (x4, q4), _ = pbd.shard_tables([tuple(lb), tuple(lp)], n_contigs, ready=ready)
        assert len(ready) == 2
        for u, v, ev in ((x2, x4, ready[0]), (q2, q4, ready[1])):
            torch.cuda.current_stream().wait_event(ev)
            for col_u, col_v in zip(u, v):
                assert torch.equal(col_u, col_v), ("overlapped peer exchange differs from the NCCL exchange", it)
    kind = pbd.exchange_kind()
    if kind == "peer":  # a second pair of tables, 3x larger on this group: forces a collective arena growth
        big_p = [torch.cat([x, x, x]) for x in lp]
        big_b = [torch.cat([x, x, x]) for x in lb]
        os.environ["PBGPU_EXCHANGE"] = "nccl"
        (qn, xn), _ = pbd.shard_tables([tuple(big_p), tuple(big_b)], n_contigs)
        os.environ["PBGPU_EXCHANGE"] = "peer"
        (qp, xp), _ = pbd.shard_tables([tuple(big_p), tuple(big_b)], n_contigs)
        for u, v in ((qn, qp), (xn, xp)):
            for col_u, col_v in zip(u, v):
                assert torch.equal(col_u, col_v), "peer exchange differs after growing"
        (q2, x2), _ = pbd.shard_tables([tuple(lp), tuple(lb)], n_contigs)
    qc, qs, qe, qrow = q2
    xc, xs, xe, xrow = x2
    ix = engine.DeviceIndex(xc, xs, xe, n_contigs)
    a, b = ix.overlap_pairs(qc, qs, qe, engine.FILTER_STRICT)
    cnt = ix.count_overlaps(qc, qs, qe, engine.FILTER_STRICT)
    assert int(cnt.sum()) == a.numel()
    ga = pbd.translate(a, qrow).cpu().numpy().view(np.uint32).astype(np.int64)
    gb = pbd.translate(b, xrow).cpu().numpy().view(np.uint32).astype(np.int64)
    mine = ga * M + gb
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    if rank == 0:
        got = np.sort(np.concatenate(parts))
        oa, ob = oracle.Index(bc, bs, be, n_contigs).overlap_pairs(pc, ps, pe, True)
        want = np.sort(oa.astype(np.int64) * M + ob.astype(np.int64))
        assert len(got) == len(want) and np.array_equal(got, want), (len(got), len(want))
        print(f"DIST_CHECK_OK world={world} pairs={len(got)} exchange={kind}")
    # the other strategy (single-contig case): replicate the indexed table, probes stay where they are
    assert pbd.join_strategy(torch.tensor([1e7]), torch.tensor([1e6]), world) == "replicate"
    (ac, as_, ae), bbase2, bsizes = pbd.replicate_table(*lb)
    assert bbase2 == bbase and sum(bsizes) == M and ac.numel() == M
    assert np.array_equal(ac.cpu().numpy(), bc) and np.array_equal(as_.cpu().numpy(), bs) and np.array_equal(ae.cpu().numpy(), be)
    ixr = engine.DeviceIndex(ac, as_, ae, n_contigs)
    ra, rb = ixr.overlap_pairs(*lp, engine.FILTER_STRICT)
    mine_r = (ra.cpu().numpy().view(np.uint32).astype(np.int64) + pbase) * M + rb.cpu().numpy().view(np.uint32).astype(np.int64)
    parts = [None] * world
    dist.all_gather_object(parts, mine_r)
    if rank == 0:
        got_r = np.sort(np.concatenate(parts))
        assert np.array_equal(got_r, want), "replicated-build join differs from the oracle"
        print(f"DIST_CHECK_REPLICATE_OK world={world} pairs={len(got_r)}")
    # the one-call form, both strategies and the automatic choice
    for strat in ("shard", "replicate", None):
        oa_, ob_, used = pbd.overlap(tuple(lp), tuple(lb), n_contigs, engine.FILTER_STRICT, strategy=strat)
        mine_o = oa_.cpu().numpy().view(np.uint32).astype(np.int64) * M + ob_.cpu().numpy().view(np.uint32).astype(np.int64)
        parts = [None] * world
        dist.all_gather_object(parts, mine_o)
        if rank == 0:
            assert np.array_equal(np.sort(np.concatenate(parts)), want), ("dist.overlap differs from the oracle", strat, used)
            print(f"DIST_CHECK_ONE_CALL_OK strategy={strat} used={used}")
    dist.barrier()
    pbd.close_peer_exchanges()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
