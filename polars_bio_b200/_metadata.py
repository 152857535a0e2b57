"""Coordinate-system metadata: decides Strict (0-based, half-open) vs Weak (1-based, closed).

Same contract as /root/reference/polars_bio/_metadata.py:80-362:
  * pandas:  ``df.attrs["coordinate_system_zero_based"]``
  * polars:  ``df.config_meta`` (polars-config-meta) when that plugin is installed
  * pyarrow: schema metadata key ``coordinate_system_zero_based`` (extension of this package)
  * paths:   no metadata
Missing metadata -> MissingCoordinateSystemError in strict mode, else the global option with a
UserWarning; differing inputs -> CoordinateSystemMismatchError.
"""
from __future__ import annotations

import warnings
from typing import Any, Optional

from .constants import POLARS_BIO_COORDINATE_SYSTEM_CHECK, POLARS_BIO_COORDINATE_SYSTEM_ZERO_BASED
from .context import _flag
from .exceptions import CoordinateSystemMismatchError, MissingCoordinateSystemError

COORDINATE_SYSTEM_KEY = "coordinate_system_zero_based"


def _is_pandas(obj: Any) -> bool:
    try:
        import pandas as pd
    except ImportError:  # pragma: no cover
        return False
    return isinstance(obj, pd.DataFrame)


def _is_arrow(obj: Any) -> bool:
    try:
        import pyarrow as pa
    except ImportError:  # pragma: no cover
        return False
    return isinstance(obj, (pa.Table, pa.RecordBatch, pa.RecordBatchReader))


def set_coordinate_system(df, zero_based: bool):
    """Tag a frame; returns the tagged object (pyarrow tables are immutable, so a new one)."""
    if hasattr(df, "config_meta"):
        df.config_meta.set(**{COORDINATE_SYSTEM_KEY: zero_based})
        return df
    if _is_pandas(df):
        df.attrs[COORDINATE_SYSTEM_KEY] = zero_based
        return df
    if _is_arrow(df) and hasattr(df, "replace_schema_metadata"):
        md = dict(df.schema.metadata or {})
        md[COORDINATE_SYSTEM_KEY.encode()] = b"true" if zero_based else b"false"
        return df.replace_schema_metadata(md)
    try:  # plain polars frames without the config_meta plugin: keep a side attribute
        object.__setattr__(df, "_pb_coordinate_system_zero_based", zero_based)
        return df
    except Exception as exc:
        raise TypeError(f"Cannot set coordinate system on {type(df).__name__}") from exc


def get_coordinate_system(df) -> Optional[bool]:
    if hasattr(df, "config_meta"):
        return df.config_meta.get_metadata().get(COORDINATE_SYSTEM_KEY)
    if _is_pandas(df):
        return df.attrs.get(COORDINATE_SYSTEM_KEY)
    if _is_arrow(df):
        md = df.schema.metadata or {}
        v = md.get(COORDINATE_SYSTEM_KEY.encode())
        return None if v is None else v.decode().lower() == "true"
    if isinstance(df, str):
        return None
    return getattr(df, "_pb_coordinate_system_zero_based", None)


def _describe(df) -> str:
    if isinstance(df, str):
        return f"file path '{df}'"
    if _is_pandas(df):
        return "Pandas DataFrame"
    return type(df).__module__.split(".")[0].capitalize() + " " + type(df).__name__


def _hint(df) -> str:
    if _is_pandas(df):
        return 'Set df.attrs["coordinate_system_zero_based"] = True (0-based) or False (1-based).'
    if isinstance(df, str):
        return ('Paths carry no metadata; disable strict checking with '
                'pb.set_option("datafusion.bio.coordinate_system_check", False).')
    return "Set it with df.config_meta.set(coordinate_system_zero_based=True/False) or pb.set_coordinate_system(df, ...)."


def _resolve_missing(missing) -> bool:
    if _flag(POLARS_BIO_COORDINATE_SYSTEM_CHECK):
        first = missing[0]
        raise MissingCoordinateSystemError(f"{_describe(first)} is missing coordinate system metadata.\n\n{_hint(first)}")
    zero_based = _flag(POLARS_BIO_COORDINATE_SYSTEM_ZERO_BASED)
    warnings.warn(
        f"Coordinate system metadata is missing for: {', '.join(_describe(m) for m in missing)}. "
        f"Using global POLARS_BIO_COORDINATE_SYSTEM_ZERO_BASED setting ({'0-based' if zero_based else '1-based'}).",
        UserWarning, stacklevel=5)
    return zero_based


def validate_coordinate_systems(df1, df2, ctx=None) -> bool:
    """True -> 0-based (FilterOp.Strict); False -> 1-based (FilterOp.Weak)."""
    cs1, cs2 = get_coordinate_system(df1), get_coordinate_system(df2)
    missing = [d for d, c in ((df1, cs1), (df2, cs2)) if c is None]
    if missing:
        fallback = _resolve_missing(missing)
        cs1 = fallback if cs1 is None else cs1
        cs2 = fallback if cs2 is None else cs2
    if bool(cs1) != bool(cs2):
        name = lambda c: "0-based" if c else "1-based"
        raise CoordinateSystemMismatchError(
            f"Coordinate system mismatch: first input uses {name(cs1)} coordinates, second input uses {name(cs2)} coordinates.")
    return bool(cs1)


def validate_coordinate_system_single(df, ctx=None) -> bool:
    cs = get_coordinate_system(df)
    if cs is None:
        cs = _resolve_missing([df])
    return bool(cs)
