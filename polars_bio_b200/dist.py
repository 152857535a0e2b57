"""Multi-GPU sharding of the interval join: one process per GPU, contigs are the unit of distribution.

The join key is contig equality, so contigs are independent.  Rows arrive on arbitrary
ranks; one exchange step moves every row to the rank that owns its contig, after which each rank joins
its contigs with no further communication:

    per-contig row histogram (all_reduce)  ->  owner table by LPT bin packing (chr1 is ~10x chrY)
    -> K8 pack: stable bucket by destination rank into 16-byte records (contig,start,end,global_row)
    -> NCCL all-to-all of the records over NVLink (counts first, then payload)
    -> unpack to columns -> local index build / count / emit -> pair row ids translated back to global ids.

``torch.distributed`` is the plumbing (process group, all_to_all_single); packing, unpacking and id
translation are kernels of libpbgpu.so.  The collectives run on the backend of the tensors' device, so
the exchange logic is testable on CPU with gloo (tests/test_dist_gloo.py) given pre-packed records.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def owner_table(weights: torch.Tensor, world: int) -> torch.Tensor:
    """contig -> rank by longest-processing-time-first bin packing over per-contig weights.
    Deterministic (ties broken by contig id), so every rank computes the same table."""
    w = weights.detach().to("cpu", torch.float64).tolist()
    order = sorted(range(len(w)), key=lambda c: (-w[c], c))
    load = [0.0] * world
    owner = [0] * len(w)
    for c in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[c] = r
        load[r] += w[c]
    return torch.tensor(owner, dtype=torch.int32)


def contig_histogram(contig: torch.Tensor, n_contigs: int, group=None) -> torch.Tensor:
    """Global number of rows per contig (null-keyed rows ignored)."""
    ok = (contig >= 0) & (contig < n_contigs)
    h = torch.bincount(contig[ok].long(), minlength=n_contigs).to(torch.int64)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(h, group=group)
    return h


def row_id_base(n_local: int, device, group=None) -> Tuple[int, int]:
    """(first global row id of this rank's slice, total rows): exclusive prefix of the slice sizes."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0, n_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    sizes[rank] = n_local
    dist.all_reduce(sizes, group=group)
    return int(sizes[:rank].sum().item()), int(sizes.sum().item())


def exchange_records(packed: torch.Tensor, send_counts: torch.Tensor, group=None) -> torch.Tensor:
    """All-to-all of 16-byte records.  ``packed``: int32 [rows, 4] grouped by destination rank;
    ``send_counts``: int64 [world] rows per destination.  Returns the int32 [received, 4] records."""
    world = dist.get_world_size(group)
    send_counts = send_counts.to(packed.device, torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    send_list = [int(x) for x in send_counts.tolist()]
    recv_list = [int(x) for x in recv_counts.tolist()]
    kept = sum(send_list)
    out = torch.empty((sum(recv_list), 4), dtype=torch.int32, device=packed.device)
    dist.all_to_all_single(out, packed[:kept].contiguous(), output_split_sizes=recv_list, input_split_sizes=send_list, group=group)
    assert len(send_list) == world
    return out


def shard_table(contig: torch.Tensor, start: torch.Tensor, end: torch.Tensor, n_contigs: int, owner: torch.Tensor,
                base_row: int, group=None):
    """Move a table's rows to their contig owners.  Inputs: this rank's slice (int32 CUDA columns).
    Returns (contig, start, end, global_row) columns of the rows this rank now owns."""
    import ctypes

    from . import _native
    from .engine import _stream_ptr

    L = _native.lib()
    dev = contig.device
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = contig.numel()
    packed = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
    counts = torch.empty(world, dtype=torch.int64, device=dev)
    owner_d = owner.to(dev, torch.int32).contiguous()
    with torch.cuda.device(dev):
        _native.check(L.pbgpu_pack_by_owner(contig.data_ptr(), start.data_ptr(), end.data_ptr(), n, owner_d.data_ptr(), n_contigs,
                                            world, ctypes.c_uint32(base_row), packed.data_ptr(), counts.data_ptr(), _stream_ptr(dev)))
        recv = exchange_records(packed, counts, group) if world > 1 else packed[: int(counts.sum().item())]
        r = recv.shape[0]
        c2 = torch.empty(r, dtype=torch.int32, device=dev); s2 = torch.empty_like(c2); e2 = torch.empty_like(c2)
        row2 = torch.empty_like(c2)
        _native.check(L.pbgpu_unpack_records(recv.data_ptr(), r, c2.data_ptr(), s2.data_ptr(), e2.data_ptr(), row2.data_ptr(), _stream_ptr(dev)))
    return c2, s2, e2, row2


def shard_tables(tables, n_contigs: int, group=None, trace: Optional[list] = None):
    """Exchange several tables at once (typically the probe and the build side) with as few host round trips as
    possible: ONE all_reduce (per-contig histogram of all tables + slice sizes), ONE count all-to-all, one payload
    all-to-all per table.  ``tables``: list of (contig, start, end) int32 CUDA columns (this rank's slices).
    Returns (list of (contig, start, end, global_row) owned columns, owner table)."""
    import ctypes
    import time

    from . import _native
    from .engine import _stream_ptr

    L = _native.lib()
    dev = tables[0][0].device
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    T = len(tables)
    t0 = time.perf_counter()
    with torch.cuda.device(dev):
        sp = _stream_ptr(dev)
        # [hist (n_contigs) | slice sizes (T x world)] -> one all_reduce
        meta = torch.zeros(n_contigs + T * world, dtype=torch.int64, device=dev)
        for t, (c, _, _) in enumerate(tables):
            _native.check(L.pbgpu_contig_histogram(c.data_ptr(), c.numel(), n_contigs, meta.data_ptr(), sp))
            meta[n_contigs + t * world + rank] = c.numel()
        if world > 1:
            dist.all_reduce(meta, group=group)
        meta_h = meta.cpu()
        owner = owner_table(meta_h[:n_contigs], world)
        sizes = meta_h[n_contigs:].view(T, world)
        bases = [int(sizes[t, :rank].sum()) for t in range(T)]
        owner_d = owner.to(dev, non_blocking=True)
        if trace is not None: trace.append(("histogram+owner", time.perf_counter() - t0)); t0 = time.perf_counter()
        packed, counts = [], torch.empty((T, world), dtype=torch.int64, device=dev)
        for t, (c, s, e) in enumerate(tables):
            n = c.numel()
            pk = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
            _native.check(L.pbgpu_pack_by_owner(c.data_ptr(), s.data_ptr(), e.data_ptr(), n, owner_d.data_ptr(), n_contigs, world,
                                                ctypes.c_uint32(bases[t]), pk.data_ptr(), counts[t].data_ptr(), sp))
            packed.append(pk)
        if world > 1:
            recv_counts = torch.empty_like(counts)
            # counts[t, r] -> rank r ; as one all-to-all over the transposed [world, T] layout
            send_t = counts.t().contiguous()
            recv_t = torch.empty_like(send_t)
            dist.all_to_all_single(recv_t, send_t, group=group)
            both = torch.stack([send_t, recv_t]).cpu()  # one D2H
            send_l, recv_l = both[0].t().tolist(), both[1].t().tolist()  # [T][world]
        else:
            send_l = counts.cpu().tolist()
            recv_l = send_l
        if trace is not None: trace.append(("pack+counts", time.perf_counter() - t0)); t0 = time.perf_counter()
        out = []
        for t in range(T):
            kept, r = int(sum(send_l[t])), int(sum(recv_l[t]))
            if world > 1:
                recv = torch.empty((max(r, 1), 4), dtype=torch.int32, device=dev)
                dist.all_to_all_single(recv[:r], packed[t][:kept], output_split_sizes=[int(x) for x in recv_l[t]],
                                       input_split_sizes=[int(x) for x in send_l[t]], group=group)
            else:
                recv = packed[t]
            c2 = torch.empty(r, dtype=torch.int32, device=dev); s2 = torch.empty_like(c2); e2 = torch.empty_like(c2)
            row2 = torch.empty_like(c2)
            _native.check(L.pbgpu_unpack_records(recv.data_ptr(), r, c2.data_ptr(), s2.data_ptr(), e2.data_ptr(), row2.data_ptr(), sp))
            out.append((c2, s2, e2, row2))
        if trace is not None:
            torch.cuda.synchronize(dev)
            trace.append(("all-to-all+unpack", time.perf_counter() - t0))
    return out, owner


def translate(local_rows: torch.Tensor, global_of_local: torch.Tensor) -> torch.Tensor:
    """Pair buffer positions -> global row ids (in place)."""
    from . import _native
    from .engine import _stream_ptr

    dev = local_rows.device
    with torch.cuda.device(dev):
        _native.check(_native.lib().pbgpu_translate_rows(local_rows.data_ptr(), local_rows.numel(), global_of_local.data_ptr(),
                                                         local_rows.data_ptr(), _stream_ptr(dev)))
    return local_rows
