"""`polars.io.plugins.register_io_source` of the stand-in (TEST INFRASTRUCTURE): the source is called lazily as
io_source(with_columns, predicate, n_rows, batch_size) and must yield DataFrames."""
from .. import LazyFrame, Schema


def register_io_source(io_source, *, schema, validate_schema: bool = False):
    sch = schema if isinstance(schema, Schema) else Schema(dict(schema))
    return LazyFrame(io_source, sch)
