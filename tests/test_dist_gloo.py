"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in polars_bio_b200/dist.py: owner table, global
histogram, row-id bases and the record all-to-all.  The CUDA pack / unpack kernels around them are covered by
the gpu-marked test below (needs 2 GPUs) and by `bench.py --gpus N`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from polars_bio_b200 import dist as pbd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        n, n_contigs = 5000 + 137 * rank, 7
        c = rng.integers(-1, n_contigs, n).astype(np.int32)  # includes null keys (-1)
        s = rng.integers(0, 10_000, n).astype(np.int32)
        e = (s + rng.integers(1, 100, n)).astype(np.int32)
        hist = pbd.contig_histogram(torch.from_numpy(c), n_contigs)
        owner = pbd.owner_table(hist, world)
        base, total = pbd.row_id_base(n, torch.device("cpu"))
        # what the pack kernel produces: records stably grouped by destination rank, null keys dropped
        dest = np.where(c >= 0, owner.numpy()[np.clip(c, 0, n_contigs - 1)], world)
        order = np.argsort(dest, kind="stable")
        kept = int((dest < world).sum())
        rec = np.stack([c, s, e, (base + np.arange(n)).astype(np.int32)], axis=1)[order].astype(np.int32)
        counts = torch.from_numpy(np.bincount(dest[dest < world], minlength=world).astype(np.int64))
        got = pbd.exchange_records(torch.from_numpy(np.ascontiguousarray(rec)), counts).numpy()
        assert (owner.numpy()[got[:, 0]] == rank).all()          # every received row belongs here
        gathered = [None] * world
        dist.all_gather_object(gathered, (got.tolist(), rec[:kept].tolist(), hist.tolist(), owner.tolist(), base, total))
        q.put((rank, gathered))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_exchange():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = results[0]
    received = sorted(tuple(r) for part in g for r in part[0])
    sent = sorted(tuple(r) for part in g for r in part[1])
    assert received == sent and len(sent) > 8000                  # nothing lost, nothing duplicated
    assert g[0][2] == g[1][2] and g[0][3] == g[1][3]              # same histogram and owner table on both ranks
    assert g[0][4] == 0 and g[1][4] == 5000 and g[0][5] == g[1][5] == 10137
    rows = [r[3] for r in sent]
    assert len(set(rows)) == len(rows)                             # global row ids are unique


def test_owner_table_lpt_balance_and_determinism():
    w = torch.tensor([248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50, 156, 57],
                     dtype=torch.float64)  # GRCh38 chr1..22,X,Y (Mb)
    for world in (1, 2, 4, 8):
        o = pbd.owner_table(w, world)
        assert torch.equal(o, pbd.owner_table(w.clone(), world))
        load = torch.zeros(world, dtype=torch.float64).index_add_(0, o.long(), w)
        assert load.max() / load.mean() < 1.10                      # chr1 alone is 8 % of the genome
        assert set(o.tolist()) == set(range(world))


@pytest.mark.gpu
def test_sharded_join_two_gpus_matches_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(root, "tests", "tools", "dist_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_CHECK_OK" in r.stdout


def test_peer_layout_matches_bruteforce():
    """Host restatement of peer_plan_kernel (the GPU tests compare the kernel with it): regions in source order."""
    import numpy as np

    rng = np.random.default_rng(3)
    for world, T, nc in ((1, 1, 3), (2, 2, 11), (4, 2, 25), (8, 3, 5)):
        g = rng.integers(0, 1000, (world, T, nc + 1)).astype(np.int64)
        g[:, :, nc] = g[:, :, :nc].sum(axis=2) + rng.integers(0, 50, (world, T))  # slice size incl. null-keyed rows
        owner = pbd.owner_table(torch.from_numpy(g[:, :, :nc].sum(axis=(0, 1))), world).numpy()
        for rank in range(world):
            lay = pbd.peer_layout(g, rank, [10 ** 9] * T)
            assert np.array_equal(lay["owner"].numpy(), owner)
            for t in range(T):
                assert int(lay["base"][t]) == int(g[:rank, t, nc].sum())
                for d in range(world):
                    cols = owner == d
                    assert int(lay["rows"][t, d]) == int(g[rank, t, :nc][cols].sum())
                    assert int(lay["offset"][t, d]) == int(g[:rank, t, :nc][:, cols].sum())
                assert int(lay["received"][t]) == int(g[:, t, :nc][:, owner == rank].sum())
                assert int(lay["need"][t]) == max(int(g[:, t, :nc][:, owner == d].sum()) for d in range(world))
            assert not lay["overflow"]
        assert pbd.peer_layout(g, 0, [1] * T)["overflow"]


def _replicate_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = []
        for even in (False, True):
            n = 1000 if even else 700 + 300 * rank
            rng = np.random.default_rng(50 + rank)
            c = torch.from_numpy(rng.integers(0, 5, n).astype(np.int32))
            s = torch.from_numpy(rng.integers(0, 10_000, n).astype(np.int32))
            e = s + 7
            (ac, as_, ae), base, sizes = pbd.replicate_table(c, s, e)
            out.append((ac.tolist(), as_.tolist(), ae.tolist(), base, sizes, c.tolist(), s.tolist()))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_replicate_table():
    """Every rank gets the whole table in global row order (the replicate-the-indexed-side strategy of SURVEY 8e)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_replicate_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in range(2):
        r0, r1 = results[0][k], results[1][k]
        assert r0[:3] == r1[:3]                                   # identical whole table on both ranks
        assert r0[0] == r0[5] + r1[5] and r0[1] == r0[6] + r1[6]  # = rank 0's slice followed by rank 1's
        assert r0[3] == 0 and r1[3] == len(r0[5]) and r0[4] == r1[4] == [len(r0[5]), len(r1[5])]
        assert r0[2] == [x + 7 for x in r0[1]]


def test_join_strategy():
    one = torch.tensor([10_000_000.0]); onev = torch.tensor([1_000_000.0])
    assert pbd.join_strategy(one, onev, 1) == "shard"
    assert pbd.join_strategy(one, onev, 8) == "replicate"                      # a single contig cannot be sharded by contig
    even_p, even_b = torch.full((24,), 4e6), torch.full((24,), 3.75e6)
    assert pbd.join_strategy(even_p, even_b, 8) == "shard"                     # config 3: 100M x 90M over 24 contigs
    assert pbd.join_strategy(torch.full((24,), 2e6), torch.full((24,), 1e5), 2) == "replicate"  # tiny indexed side
    assert pbd.join_strategy(torch.full((24,), 2e6), torch.full((24,), 2e5), 2) == "shard"      # ... but not beyond L2
    skew = torch.tensor([9e6] + [1e5] * 10)
    assert pbd.join_strategy(skew, skew / 10, 4) == "replicate"                 # one contig dominates
