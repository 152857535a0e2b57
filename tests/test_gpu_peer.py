"""GPU tests of the peer-memory contig exchange (csrc/peer.cuh, polars_bio_b200.dist.PeerExchange).

Several ranks are simulated on ONE device: every simulated rank gets its own receive arenas (plain torch buffers) and
runs the plan + scatter kernels with the addresses of all arenas -- exactly what the ranks of a node do with CUDA-IPC
mappings (tests/tools/dist_check.py covers the IPC + NCCL part on N GPUs).  Expected result: destination d receives,
for every table, the rows whose contig it owns, ordered by global row id (= the stable NCCL exchange).  Bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _slices(rng, world, n_total, nc, null_frac=0.05):
    c = rng.integers(0, nc, n_total).astype(np.int32)
    c[rng.random(n_total) < null_frac] = -1
    c[rng.random(n_total) < 0.01] = nc + 3  # out-of-range code = null key
    s = rng.integers(0, 1_000_000, n_total).astype(np.int32)
    e = (s + rng.integers(1, 500, n_total)).astype(np.int32)
    cuts = np.sort(rng.integers(0, n_total + 1, world - 1)) if world > 1 else np.array([], dtype=np.int64)
    bounds = np.concatenate([[0], cuts, [n_total]]).astype(np.int64)
    return (c, s, e), bounds


def _gathered(tables, bounds_list, world, nc):
    g = np.zeros((world, len(tables), nc + 1), dtype=np.int64)
    for t, ((c, _, _), b) in enumerate(zip(tables, bounds_list)):
        for r in range(world):
            cc = c[b[r]: b[r + 1]]
            ok = (cc >= 0) & (cc < nc)
            g[r, t, :nc] = np.bincount(cc[ok], minlength=nc)
            g[r, t, nc] = len(cc)
    return g


def _run_simulated(world, nc, sizes, seed, cap_scale=1.3, flags=True):
    """flags=True: histograms and completion travel through the (simulated) control blocks, as on a node; the phases
    are enqueued rank by rank on ONE stream, so every wait finds its flags already raised (a missing or misplaced flag
    would show as a timeout).  flags=False: plan from explicit all-gathered histograms."""
    from polars_bio_b200 import _native, dist as pbd

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(seed)
    tables, bounds = [], []
    for n in sizes:
        t, b = _slices(rng, world, n, nc)
        tables.append(t); bounds.append(b)
    T = len(sizes)
    g = _gathered(tables, bounds, world, nc)
    lay = [pbd.peer_layout(g, r, [1 << 40] * T) for r in range(world)]
    need = lay[0]["need"].tolist()
    caps = [pbd._round_cap(int(x * cap_scale) + 1) for x in need]
    arena_bytes = 16 * sum(caps)
    arenas = [[torch.full((arena_bytes,), 0xAB, dtype=torch.uint8, device=dev) for _ in range(world)] for _ in range(2)]
    ptrs = [[a.data_ptr() for a in arenas[p]] for p in range(2)]
    ctl_bytes = int(_native.lib().pbgpu_peer_ctl_bytes(world, T, nc))
    ctl = [torch.zeros(ctl_bytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    exs = [pbd.PeerExchange(nc, caps, dev, arenas=ptrs, ctls=[c.data_ptr() for c in ctl] if flags else None, world=world, rank=r)
           for r in range(world)]
    gd = torch.from_numpy(g).to(dev)
    owner = lay[0]["owner"].numpy()
    for step in range(3):  # both parities, and the first one again
        local = [[tuple(torch.from_numpy(np.ascontiguousarray(col[b[r]: b[r + 1]])).to(dev) for col in t) for t, b in zip(tables, bounds)]
                 for r in range(world)]
        pars = [ex.begin_step(local[r]) for r, ex in enumerate(exs)]
        assert pars == [step & 1] * world
        E = pbd.PeerExchange
        if flags:
            for r in range(world):
                exs[r].phase(E.PH_HIST)  # histograms + publication
            for r in range(world):
                exs[r].phase(E.PH_PLAN)
        else:
            for r in range(world):
                exs[r].phase_plan(pars[r], gd)
        for r in range(world):
            for t in range(T):
                exs[r].phase(E.PH_SCATTER | E.PH_SIGNAL if flags else E.PH_SCATTER, t)
        if flags:
            for r in range(world):
                for t in range(T):
                    exs[r].phase(E.PH_WAIT, t)
        torch.cuda.synchronize()
        for d in range(world):
            assert int(exs[d].status.item()) == 0, "a wait kernel timed out"
            out, need_d = exs[d].collect(step & 1)
            assert out is not None
            assert [int(x) for x in need_d] == [int(x) for x in need]
            assert np.array_equal(exs[d].owner[:nc].cpu().numpy(), owner)
            for t, ((c, s, e), b) in enumerate(zip(tables, bounds)):
                ok = (c >= 0) & (c < nc)
                sel = np.zeros(len(c), dtype=bool)
                sel[ok] = owner[c[ok]] == d
                rows = np.nonzero(sel)[0]
                gc, gs, ge, grow = (x.cpu().numpy() for x in out[t])
                assert len(gc) == len(rows) == int(lay[d]["received"][t])
                assert np.array_equal(gc, c[rows]) and np.array_equal(gs, s[rows]) and np.array_equal(ge, e[rows])
                assert np.array_equal(grow.view(np.uint32), rows.astype(np.uint32))
        if flags:  # the control blocks hold exactly this step's flags and everybody's histograms
            for d in range(world):
                words = ctl[d].view(torch.int64).cpu().numpy()
                assert np.array_equal(words[:world], np.full(world, step + 1))
                for t in range(T):
                    assert np.array_equal(words[16 + 16 * t: 16 + 16 * t + world], np.full(world, step + 1))
                assert np.array_equal(words[128: 128 + g.size].reshape(g.shape), g)
    del local
    return exs, caps, need


@pytest.mark.parametrize("flags", [True, False])
@pytest.mark.parametrize("world,nc,sizes", [(1, 3, [5000]), (2, 11, [100_003, 20_001]), (4, 25, [300_000, 70_000]),
                                            (8, 24, [250_000, 33_333, 1000]), (16, 97, [120_000, 0, 5, 64_000]),
                                            (3, 1, [10_000, 10_000]), (4, 5000, [200_000, 50_000])])
def test_simulated_ranks_match_stable_exchange(world, nc, sizes, flags):
    _run_simulated(world, nc, sizes, seed=world * 1000 + nc, flags=flags)


def test_missing_peer_times_out_instead_of_hanging(monkeypatch):
    """A rank whose peer never publishes: the plan kernel gives up after the timeout and reports status 2."""
    import subprocess, sys, os, textwrap

    code = textwrap.dedent("""
        import torch, numpy as np
        from polars_bio_b200 import _native, dist as pbd
        dev = torch.device("cuda:0")
        world, nc, T = 2, 3, 1
        arenas = [[torch.zeros(16 * 64, dtype=torch.uint8, device=dev) for _ in range(world)] for _ in range(2)]
        ctl = [torch.zeros(int(_native.lib().pbgpu_peer_ctl_bytes(world, T, nc)), dtype=torch.uint8, device=dev) for _ in range(world)]
        ex = pbd.PeerExchange(nc, [64], dev, arenas=[[a.data_ptr() for a in arenas[p]] for p in range(2)],
                              ctls=[c.data_ptr() for c in ctl], world=world, rank=0)
        tab = [tuple(torch.zeros(10, dtype=torch.int32, device=dev) for _ in range(3))]
        par = ex.begin_step(tab); ex.phase(0)  # histograms, publication, plan: rank 1 never publishes
        ex.phase(0, 0)                          # scatter (skipped: the plan failed), signal, wait
        torch.cuda.synchronize()
        try:
            ex.collect(par)
            print("NO_ERROR")
        except RuntimeError as e:
            print("TIMEOUT_REPORTED", int(ex.status.item()))
    """)
    env = dict(os.environ, PBGPU_PEER_TIMEOUT_MS="300")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=root, timeout=240)
    assert "TIMEOUT_REPORTED 2" in r.stdout, (r.stdout, r.stderr[-2000:])


def test_overflow_is_detected_and_nothing_is_written():
    from polars_bio_b200 import dist as pbd

    dev = torch.device("cuda:0")
    world, nc = 2, 4
    rng = np.random.default_rng(5)
    t, b = _slices(rng, world, 50_000, nc)
    g = _gathered([t], [b], world, nc)
    need = int(pbd.peer_layout(g, 0, [1 << 40])["need"][0])
    caps = [pbd._round_cap(need // 2)]
    arenas = [[torch.full((16 * caps[0],), 0xAB, dtype=torch.uint8, device=dev) for _ in range(world)] for _ in range(2)]
    ptrs = [[a.data_ptr() for a in arenas[p]] for p in range(2)]
    gd = torch.from_numpy(g).to(dev)
    for r in range(world):
        ex = pbd.PeerExchange(nc, caps, dev, arenas=ptrs, world=world, rank=r)
        tl = [tuple(torch.from_numpy(np.ascontiguousarray(col[b[r]: b[r + 1]])).to(dev) for col in t)]
        par = ex.begin_step(tl)
        ex.phase_plan(par, gd)
        ex.phase(pbd.PeerExchange.PH_SCATTER, 0)
        out, need_r = ex.collect(par)
        assert out is None and int(need_r[0]) == need
    torch.cuda.synchronize()
    for a in arenas[0] + arenas[1]:
        assert bool((a == 0xAB).all())


def test_single_process_shard_tables_peer_equals_nccl_path(monkeypatch):
    """world = 1 (no process group): shard_tables through the peer arena (own memory) and through pack/unpack must
    return identical columns, and the arena grows when a later call brings more rows."""
    from polars_bio_b200 import dist as pbd

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    nc = 7
    pbd.close_peer_exchanges()
    try:
        for n, m in ((40_000, 9_000), (400_000, 90_000)):  # the second call overflows the first call's arenas
            (pc, ps, pe), _ = _slices(rng, 1, n, nc)
            (bc, bs, be), _ = _slices(rng, 1, m, nc)
            tabs = [tuple(torch.from_numpy(x).to(dev) for x in (pc, ps, pe)), tuple(torch.from_numpy(x).to(dev) for x in (bc, bs, be))]
            monkeypatch.setenv("PBGPU_EXCHANGE", "nccl")
            ref, owner_ref = pbd.shard_tables(tabs, nc)
            monkeypatch.setenv("PBGPU_EXCHANGE", "peer")
            got, owner = pbd.shard_tables(tabs, nc)
            assert pbd.exchange_kind() == "peer"
            assert torch.equal(owner.cpu(), owner_ref.cpu())
            for u, v in zip(ref, got):
                for cu, cv in zip(u, v):
                    assert torch.equal(cu, cv)
    finally:
        pbd.close_peer_exchanges()
