"""polars_bio_b200 -- B200-native interval-join engine behind polars-bio's range-operation API.

Only the hot path of the reference is here: ``overlap`` / ``nearest`` /
``count_overlaps`` / ``coverage`` computed by hand-written sm_100a kernels in libpbgpu.so
(C ABI: include/pbgpu.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _native  # noqa: F401  (does not load the library until first use)
