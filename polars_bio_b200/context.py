"""Session options -- the flag surface of /root/reference/polars_bio/context.py:27-140 and
/root/reference/src/context.rs:35-98, without DataFusion: a process-wide string key/value store.

Keys the GPU engine reads: ``datafusion.bio.coordinate_system_zero_based``,
``datafusion.bio.coordinate_system_check``, ``datafusion.execution.batch_size`` (output batch cap),
``bio.interval_join_low_memory``.  ``bio.interval_join_algorithm`` and
``datafusion.execution.target_partitions`` are accepted and recorded (the GPU engine has one
algorithm and no CPU partitions); unknown keys are stored, never rejected, like context.rs:91-98.
"""
from __future__ import annotations

import numbers
import threading
from typing import Optional

from .constants import (BATCH_SIZE, INTERVAL_JOIN_ALGORITHM, POLARS_BIO_COORDINATE_SYSTEM_CHECK,
                        POLARS_BIO_COORDINATE_SYSTEM_ZERO_BASED, TARGET_PARTITIONS)


class BioSessionContext:
    """Stand-in for the PyO3 ``BioSessionContext`` (src/context.rs:11-18)."""

    def __init__(self):
        self._lock = threading.Lock()
        self._opts = {
            TARGET_PARTITIONS: "1",                           # context.py:36
            POLARS_BIO_COORDINATE_SYSTEM_ZERO_BASED: "false",  # context.py:45 (1-based by default)
            POLARS_BIO_COORDINATE_SYSTEM_CHECK: "false",       # context.py:48 (lenient by default)
            INTERVAL_JOIN_ALGORITHM: "gpu",                    # the reference default is "coitrees"
            BATCH_SIZE: "8192",
        }

    def set_option(self, key: str, value, temporary: bool = False) -> None:
        if isinstance(value, bool):
            value = "true" if value else "false"
        elif isinstance(value, numbers.Number):
            value = str(value)
        with self._lock:
            self._opts[str(key)] = str(value)

    def get_option(self, key: str) -> Optional[str]:
        with self._lock:
            return self._opts.get(key)

    def sync_options(self) -> None:  # the reference pushes options into DataFusion here (context.rs:49-54)
        return None


_CTX = BioSessionContext()
ctx = _CTX


def set_option(key, value) -> None:
    """``pb.set_option`` (context.py:84-98)."""
    _CTX.set_option(key, value)


def get_option(key) -> Optional[str]:
    """``pb.get_option`` (context.py:101-118)."""
    return _CTX.get_option(key)


def _flag(key: str) -> bool:
    v = _CTX.get_option(key)
    return v is not None and v.lower() == "true"
