"""Option types of the range-operation boundary -- Python mirrors of the PyO3 classes in
/root/reference/src/option.rs:6-112 (same names, fields, defaults and enum values)."""
from __future__ import annotations

import enum
from dataclasses import dataclass
from typing import List, Optional, Tuple


class FilterOp(enum.IntEnum):          # option.rs:96-99
    Weak = 0     # 1-based closed intervals:   a.start <= b.end and a.end >= b.start
    Strict = 1   # 0-based half-open intervals: a.start <  b.end and a.end >  b.start


class RangeOp(enum.IntEnum):           # option.rs:103-112
    Overlap = 0
    Complement = 1
    Cluster = 2
    Nearest = 3
    Coverage = 4
    Subtract = 5
    CountOverlapsNaive = 6
    Merge = 7


class OverlapOutputMode(enum.IntEnum):  # option.rs:89-92
    Join = 0
    Left = 1


@dataclass
class RangeOptions:                     # option.rs:8-85
    range_op: RangeOp
    filter_op: Optional[FilterOp] = None
    suffixes: Optional[Tuple[str, str]] = None
    columns_1: Optional[List[str]] = None
    columns_2: Optional[List[str]] = None
    on_cols: Optional[List[str]] = None
    overlap_alg: Optional[str] = None
    overlap_low_memory: Optional[bool] = None
    nearest_k: Optional[int] = None
    include_overlaps: Optional[bool] = None
    compute_distance: Optional[bool] = None
    min_dist: Optional[int] = None
    view_table: Optional[str] = None
    view_columns: Optional[List[str]] = None
    overlap_output: Optional[OverlapOutputMode] = None
    distinct_output: Optional[bool] = None


GPU_RANGE_OPS = (RangeOp.Overlap, RangeOp.Nearest, RangeOp.Coverage, RangeOp.CountOverlapsNaive)
