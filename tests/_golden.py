"""Loaders for the committed golden fixtures (tests/golden/, made by make_golden.py)."""
import json
import os

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def fixtures():
    with open(os.path.join(GOLDEN, "ref_fixtures.json")) as f:
        return json.load(f)


def exons_fbrain():
    z = np.load(os.path.join(GOLDEN, "exons_fbrain.npz"))
    return {k: z[k] for k in z.files}


def exons_fbrain_frames():
    """The two parquet fixtures as pandas frames with the reference's column names/dtypes."""
    z = exons_fbrain()
    names = z["contigs"]
    ex = pd.DataFrame({"contig": names[z["exons_chrom"]], "pos_start": z["exons_start"], "pos_end": z["exons_end"]})
    fb = pd.DataFrame({"contig": names[z["fbrain_chrom"]], "pos_start": z["fbrain_start"], "pos_end": z["fbrain_end"]})
    return ex, fb


def sort_all(df: pd.DataFrame) -> pd.DataFrame:
    """The reference's order normalisation (tests/_expected.py:204-215): sort by every column."""
    return df.sort_values(by=list(df.columns)).reset_index(drop=True)


def synth(n, n_contigs, span, max_len, seed, zero_len_frac=0.0):
    """Seeded random intervals: (contig code, start, end) int32 arrays."""
    rng = np.random.default_rng(seed)
    c = rng.integers(0, n_contigs, n).astype(np.int32)
    s = rng.integers(0, span, n).astype(np.int32)
    ln = rng.integers(1, max_len + 1, n).astype(np.int32)
    if zero_len_frac:
        ln[rng.random(n) < zero_len_frac] = 0
    return c, s, (s + ln).astype(np.int32)
