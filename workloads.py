"""Synthetic workloads of BASELINE.json `configs`, shared by bench.py, the scale tests and the tools.

Every table is generated chunk by chunk (1 Mi rows per chunk, numpy Generator seeded with [seed, chunk]) so that any
row range can be produced without the rest: a rank of the N-GPU job generates only its slice, the CPU arm only the
contigs it samples, and the union is the same table whatever the split.  Rows of all contigs are mixed (arbitrary
order, like row groups read from unsorted files).  Contig codes are 0..23 = chr1..chr22, chrX, chrY.
"""
from __future__ import annotations

import numpy as np

GRCH38 = np.array([248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
                   133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
                   58617616, 64444167, 46709983, 50818468, 156040895, 57227415], dtype=np.int64)
CONTIG_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]
CHUNK = 1 << 20
_P = GRCH38 / GRCH38.sum()


def _chunks(lo: int, hi: int):
    k = lo // CHUNK
    while k * CHUNK < hi:
        a, b = max(lo, k * CHUNK), min(hi, (k + 1) * CHUNK)
        yield k, a - k * CHUNK, b - k * CHUNK
        k += 1


def _rows(fn, seed: int, total: int, lo: int, hi: int):
    hi = min(hi, total)
    parts = []
    for k, a, b in _chunks(lo, hi):
        rows = min(CHUNK, total - k * CHUNK)
        c, s, e = fn(np.random.default_rng([seed, k]), rows)
        parts.append((c[a:b], s[a:b], e[a:b]))
    if not parts:
        z = np.zeros(0, np.int32)
        return z, z.copy(), z.copy()
    return tuple(np.ascontiguousarray(np.concatenate([p[j] for p in parts])) for j in range(3))


def _contigs(rng, n):
    return rng.choice(24, size=n, p=_P).astype(np.int32)


def _uniform_on(rng, c, width):
    return (rng.random(len(c)) * (GRCH38[c] - width)).astype(np.int64).astype(np.int32)


# ---- config 3: 100 M WGS reads x 90 M gnomAD-like variants, 24 contigs ------------------------------------------
def _c3_reads(rng, n):
    c = _contigs(rng, n)
    s = _uniform_on(rng, c, 150)
    return c, s, (s + 150).astype(np.int32)


def _c3_variants(rng, m):
    c = _contigs(rng, m)
    s = _uniform_on(rng, c, 200)
    ln = np.where(rng.random(m) < 0.9, 1, rng.geometric(0.2, m) + 1).astype(np.int32)  # 90 % SNV, 10 % indels
    return c, s, (s + ln).astype(np.int32)


C3_READS, C3_VARIANTS = 100_000_000, 90_000_000


def config3_reads(lo: int = 0, hi: int = C3_READS, total: int = C3_READS):
    """Rows [lo, hi) of the reads table: 150 bp, contig ~ GRCh38 length, start uniform (seed 3)."""
    return _rows(_c3_reads, 3, total, lo, hi)


def config3_variants(lo: int = 0, hi: int = C3_VARIANTS, total: int = C3_VARIANTS):
    """Rows [lo, hi) of the variants table: 90 % 1-bp SNVs, 10 % indels of length Geom(0.2)+1 (seed 4)."""
    return _rows(_c3_variants, 4, total, lo, hi)


def rank_slice(total: int, rank: int, world: int):
    """Contiguous block of rows rank `rank` starts with (rows are in arbitrary contig order, so a block holds rows of
    every contig); the global row id of a row is its position in the whole table."""
    return total * rank // world, total * (rank + 1) // world


# ---- config 2: one contig, 10 M reads x 1 M SNVs ------------------------------------------------------------------
CHR1_LEN = int(GRCH38[0])


def config2(n_reads: int = 10_000_000, n_variants: int = 1_000_000, seed_shift: int = 0):
    r1 = np.random.default_rng(1 + seed_shift)
    r2 = np.random.default_rng(2 + seed_shift)
    ps = r1.integers(0, CHR1_LEN - 150, n_reads, dtype=np.int64).astype(np.int32)
    pe = (ps + 150).astype(np.int32)
    bs = r2.integers(0, CHR1_LEN - 1, n_variants, dtype=np.int64).astype(np.int32)
    be = (bs + 1).astype(np.int32)
    return (np.zeros(n_reads, np.int32), ps, pe), (np.zeros(n_variants, np.int32), bs, be), 1


# ---- config 4: nearest k=1, 50 M queries x 5 M targets ------------------------------------------------------------
def _c4_queries(rng, n):
    return _c3_reads(rng, n)


def _c4_targets(rng, m):
    c = _contigs(rng, m)
    ln = np.clip(rng.lognormal(5.5, 1.0, m), 1, 100_000).astype(np.int32)
    s = _uniform_on(rng, c, 100_001)
    return c, s, (s + ln).astype(np.int32)


def config4(n: int = 50_000_000, m: int = 5_000_000):
    return _rows(_c4_queries, 5, n, 0, n), _rows(_c4_targets, 6, m, 0, m), 24


# ---- config 5: skewed output, 20 M reads x 200 k exons (~1e9 pairs at full size) -----------------------------------
def config5(n: int = 20_000_000, m: int = 200_000):
    """Exons stacked in gene-like loci (about 50 isoform exons per locus, starts jittered by < 120 bp, lengths
    log-normal with median ~134 and a tail clipped at 91,671 like tests/data/exons) so that a read on a locus overlaps
    most of them; 80 % of the reads start within +-75 bp of a random exon start, 20 % are uniform."""
    rng_r, rng_e = np.random.default_rng(7), np.random.default_rng(8)
    n_loci = max(1, m // 50)
    lc = _contigs(rng_e, n_loci)
    lpos = _uniform_on(rng_e, lc, 200_000)
    which = np.arange(m) % n_loci
    bc = lc[which]
    ln = np.clip(rng_e.lognormal(4.9, 1.1, m), 10, 91_671).astype(np.int32)
    bs = (lpos[which] + rng_e.integers(0, 120, m)).astype(np.int32)
    be = (bs + ln).astype(np.int32)
    on = rng_r.random(n) < 0.8
    ex = rng_r.integers(0, m, n)
    pc = np.where(on, bc[ex], _contigs(rng_r, n)).astype(np.int32)
    ps = np.where(on, bs[ex] + rng_r.integers(-75, 76, n), _uniform_on(rng_r, pc, 150)).astype(np.int32)
    ps = np.maximum(ps, 0)
    pe = (ps + 150).astype(np.int32)
    return (pc, ps, pe), (bc, bs, be), 24
