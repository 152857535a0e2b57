"""Names shared with the reference (/root/reference/polars_bio/constants.py)."""
DEFAULT_INTERVAL_COLUMNS = ["chrom", "start", "end"]
DEFAULT_BATCH_SIZE = 8192

# session option keys (same strings as the reference, so user code that sets them keeps working)
POLARS_BIO_COORDINATE_SYSTEM_ZERO_BASED = "datafusion.bio.coordinate_system_zero_based"
POLARS_BIO_COORDINATE_SYSTEM_CHECK = "datafusion.bio.coordinate_system_check"
INTERVAL_JOIN_ALGORITHM = "bio.interval_join_algorithm"
INTERVAL_JOIN_LOW_MEMORY = "bio.interval_join_low_memory"
TARGET_PARTITIONS = "datafusion.execution.target_partitions"
BATCH_SIZE = "datafusion.execution.batch_size"
