"""polars_bio_b200 -- B200-native interval-join engine behind polars-bio's range-operation API.

Only the hot path of the reference is here: ``overlap`` / ``nearest`` /
``count_overlaps`` / ``coverage`` and the unary sweeps ``merge`` / ``cluster`` / ``complement`` / ``subtract``
(the aliases of /root/reference/polars_bio/__init__.py:133-140),
computed by hand-written sm_100a kernels in libpbgpu.so (C ABI: include/pbgpu.h).
There is no CPU fallback: without the CUDA library or a GPU the calls raise.
"""
__version__ = "0.1.0"

from . import _native  # noqa: F401  (does not dlopen until first use)
from ._metadata import get_coordinate_system, set_coordinate_system  # noqa: F401
from .context import ctx, get_option, set_option  # noqa: F401
from .exceptions import CoordinateSystemMismatchError, MissingCoordinateSystemError  # noqa: F401
from .logging import set_loglevel  # noqa: F401
from .options import FilterOp, OverlapOutputMode, RangeOp, RangeOptions  # noqa: F401
from .range_op import (IntervalOperations, cluster, complement, count_overlaps, coverage, merge, nearest,  # noqa: F401
                       overlap, subtract)
from .range_op_io import (LazyRangeSource, RangeResult, RangeSession, range_lazy_scan, range_operation_frame,  # noqa: F401
                          range_operation_lazy, range_operation_scan)
from . import polars_ext  # noqa: F401,E402  (registers LazyFrame.pb when polars is present)

POLARS_BIO_MAX_THREADS = "datafusion.execution.target_partitions"  # /root/reference/polars_bio/__init__.py:142
