"""STAND-IN for the polars API surface polars_bio_b200 touches -- TEST INFRASTRUCTURE, not polars.

polars is not installable in the build image, so the branches of the package that only run when `import polars` works
(`LazyFrame.pb` / `DataFrame.pb` namespaces, polars inputs of `_df_to_reader` / `_prepare_lazy_stream_input`, polars
outputs of `convert_result`, the IO-plugin source of `range_lazy_scan`) would otherwise never execute.  This module
implements just enough of the documented polars behaviour over pyarrow for tests/tools/polars_standin_check.py to drive
those branches: eager `DataFrame`, deferred `LazyFrame` with projection / predicate / row-limit pushdown into an
IO-plugin generator called with (with_columns, predicate, n_rows, batch_size) exactly like
`polars.io.plugins.register_io_source` calls it, `pl.from_arrow`, `pl.col(...)` comparisons, namespace registration.
It proves that our code follows the protocol as documented; it cannot prove compatibility with a given polars release.
"""
from __future__ import annotations

import pyarrow as pa
import pyarrow.compute as pc

from . import api  # noqa: F401
from .api import _NAMESPACES

__version__ = "0.0-standin"


class Expr:
    """`pl.col(name)` and comparisons of it with a scalar: enough for a predicate and a projection."""

    def __init__(self, name=None, fn=None):
        self._name, self._fn = name, fn

    class _Meta:
        def __init__(self, e):
            self._e = e

        def root_names(self):
            return [self._e._name] if self._e._name is not None else []

    @property
    def meta(self):
        return Expr._Meta(self)

    def _cmp(self, op, other):
        name = self._name
        return Expr(name, lambda t: getattr(pc, op)(t.column(name), pa.scalar(other)))

    def __gt__(self, o): return self._cmp("greater", o)
    def __ge__(self, o): return self._cmp("greater_equal", o)
    def __lt__(self, o): return self._cmp("less", o)
    def __le__(self, o): return self._cmp("less_equal", o)
    def __eq__(self, o): return self._cmp("equal", o)  # noqa: E704

    def evaluate(self, table: pa.Table):
        return self._fn(table) if self._fn is not None else table.column(self._name)


def col(name: str) -> Expr:
    return Expr(name)


def _names(cols):
    if len(cols) == 1 and isinstance(cols[0], (list, tuple)):
        cols = cols[0]
    return [c if isinstance(c, str) else c.meta.root_names()[0] for c in cols]


class _WithNamespaces:
    def __getattr__(self, item):
        ns = _NAMESPACES.get((type(self).__name__, item))
        if ns is None:
            raise AttributeError(item)
        return ns(self)


class Schema(dict):
    def to_arrow(self) -> pa.Schema:
        return pa.schema([pa.field(k, v) for k, v in self.items()])

    def names(self):
        return list(self.keys())


class DataFrame(_WithNamespaces):
    def __init__(self, data=None):
        if isinstance(data, pa.RecordBatch):
            data = pa.Table.from_batches([data])
        elif isinstance(data, dict):
            data = pa.table(data)
        elif data is None:
            data = pa.table({})
        self._t: pa.Table = data

    # -- what polars_bio_b200 calls -------------------------------------------------------------------
    def to_arrow(self) -> pa.Table:
        return self._t

    def lazy(self) -> "LazyFrame":
        t = self._t

        def source(with_columns, predicate, n_rows, batch_size):
            yield DataFrame(t)

        return LazyFrame(source, Schema({f.name: f.type for f in t.schema}))

    def filter(self, predicate) -> "DataFrame":
        return DataFrame(self._t.filter(predicate.evaluate(self._t)))

    def select(self, *cols) -> "DataFrame":
        return DataFrame(self._t.select(_names(cols)))

    def head(self, n: int = 5) -> "DataFrame":
        return DataFrame(self._t.slice(0, n))

    def __arrow_c_stream__(self, requested_schema=None):
        return self._t.__arrow_c_stream__(requested_schema)

    # -- conveniences for the tests -------------------------------------------------------------------
    @property
    def schema(self) -> Schema:
        return Schema({f.name: f.type for f in self._t.schema})

    @property
    def columns(self):
        return self._t.column_names

    @property
    def height(self) -> int:
        return self._t.num_rows

    def __len__(self):
        return self._t.num_rows


class LazyFrame(_WithNamespaces):
    """A deferred frame over an IO-plugin style generator: nothing runs before collect() / collect_batches()."""

    def __init__(self, source, schema: Schema, with_columns=None, predicate=None, n_rows=None):
        self._source, self._schema = source, schema
        self._with_columns, self._predicate, self._n_rows = with_columns, predicate, n_rows

    def _derive(self, **kw):
        args = dict(with_columns=self._with_columns, predicate=self._predicate, n_rows=self._n_rows)
        args.update(kw)
        return LazyFrame(self._source, self._schema, **args)

    def select(self, *cols) -> "LazyFrame":
        return self._derive(with_columns=_names(cols))

    def filter(self, predicate) -> "LazyFrame":
        return self._derive(predicate=predicate)

    def head(self, n: int = 5) -> "LazyFrame":
        return self._derive(n_rows=int(n))

    limit = head

    def lazy(self) -> "LazyFrame":
        return self

    def collect_schema(self) -> Schema:
        if self._with_columns is None:
            return self._schema
        return Schema({k: self._schema[k] for k in self._with_columns})

    def _frames(self, batch_size=None):
        left = self._n_rows
        for df in self._source(self._with_columns, self._predicate, self._n_rows, batch_size):
            if not isinstance(df, DataFrame):
                raise TypeError("an IO source must yield polars DataFrames")
            # polars applies what the source did not: the projection and the row limit (the predicate was handed over)
            if self._with_columns is not None and df.columns != list(self._with_columns):
                df = df.select(self._with_columns)
            if left is not None:
                if left <= 0:
                    return
                df = df.head(left)
                left -= df.height
            yield df

    def collect(self) -> DataFrame:
        tabs = [df.to_arrow() for df in self._frames()]
        sch = self.collect_schema().to_arrow()
        return DataFrame(pa.concat_tables(tabs) if tabs else sch.empty_table())

    def collect_batches(self, lazy=True, engine="streaming", chunk_size=None):
        sch = self.collect_schema().to_arrow()

        def gen():
            for df in self._frames(chunk_size):
                for b in df.to_arrow().cast(sch).to_batches():
                    yield b

        return pa.RecordBatchReader.from_batches(sch, gen())


def from_arrow(data) -> DataFrame:
    if isinstance(data, pa.RecordBatch):
        data = pa.Table.from_batches([data])
    return DataFrame(data)
