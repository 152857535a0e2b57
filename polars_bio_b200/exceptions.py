"""Error types of the range-operation API (same names as /root/reference/polars_bio/exceptions.py)."""


class CoordinateSystemMismatchError(Exception):
    """The two inputs of a range operation disagree on 0-based vs 1-based coordinates."""


class MissingCoordinateSystemError(Exception):
    """An input carries no coordinate-system metadata and strict checking is on
    (``datafusion.bio.coordinate_system_check = true``)."""
