"""`Table` -- the user-facing entry point, same surface as `vinum.Table`
(vinum/api/table.py:17-470): `from_pydict / from_arrow / from_pandas`, `sql`, `sql_pd`,
`explain`, `head`, `schema`, `to_arrow`, `to_pandas`, `to_string`; plus the pyarrow IO
wrappers `read_csv / read_json / read_parquet` (vinum/io/arrow.py:64-190).

`Table.sql()` parses the SELECT with this package's parser (the reference's needs the pglast C
extension) and runs it on the device operators (vinum_b200.sql.engine); the result is again a
host `Table`, like the reference's MaterializeTableOperator produces (algebra.py:290-295).
"""
from __future__ import annotations

from typing import Dict

import pyarrow as pa

from .sql.engine import execute_sql, execute_sql_stream
from .sql.parser import parse_sql


class Table:
    def __init__(self, arrow_table: pa.Table):
        self._table = arrow_table
        self.last_stats: dict = {}

    # ----------------------------------------------------------- constructors
    @classmethod
    def from_pydict(cls, pydict: Dict) -> "Table":
        return cls(pa.Table.from_pydict(pydict))

    @classmethod
    def from_arrow(cls, arrow_table: pa.Table) -> "Table":
        return cls(arrow_table)

    @classmethod
    def from_pandas(cls, data_frame) -> "Table":
        return cls(pa.Table.from_pandas(data_frame))

    def shard(self, group=None):
        """This table as ONE rank's row range of a logical table spread over the GPUs of a `torchrun`
        job: `shard().sql(query)` is collective and returns the whole answer on rank 0
        (vinum_b200.sharded)."""
        from .sharded import ShardedTable
        return ShardedTable(self._table, group)

    # ------------------------------------------------------------------ query
    def sql(self, query: str) -> "Table":
        """Run a SELECT over this table (vinum/api/table.py:191-274)."""
        stats: dict = {}
        out = Table(execute_sql(query, self._table, stats=stats))
        self.last_stats = stats
        return out

    def sql_pd(self, query: str):
        """`sql(query).to_pandas()` (table.py:276-356)."""
        return self.sql(query).to_pandas()

    def explain(self, query: str, print_query_tree: bool = False) -> str:
        """Textual plan: the syntax tree and the operator chain the engine will run
        (table.py:358-410)."""
        q = parse_sql(query, self._table.schema.names)
        lines = []
        if print_query_tree:
            lines.append(repr(q))
        chain = ["scan(used columns -> HBM)"]
        agg = q.distinct or q.has_group_clause or any(_has_agg(e) for e in q.select)
        if q.where is not None:
            chain.append("where: fused predicate / mask into aggregate" if agg else "filter (compaction kernel)")
        if agg:
            chain.append("hash aggregate (device)")
        if q.having is not None:
            chain.append("having filter")
        if q.order_by:
            chain.append("radix sort + gather")
        chain.append("project")
        if q.limit is not None:
            chain.append(f"slice(limit={q.limit}, offset={q.offset})")
        chain.append("materialize -> host table")
        lines.append(" -> ".join(chain))
        text = "\n".join(lines)
        print(text)
        return text

    # ------------------------------------------------------------- accessors
    def head(self, n: int):
        return self._table.slice(0, n).to_pandas()

    @property
    def schema(self) -> pa.Schema:
        return self._table.schema

    def to_arrow(self) -> pa.Table:
        return self._table

    def to_pandas(self):
        return self._table.to_pandas()

    def to_pydict(self) -> Dict:
        return self._table.to_pydict()

    def to_string(self) -> str:
        return self._table.to_string() if hasattr(self._table, "to_string") else str(self._table)

    def __str__(self) -> str:
        return self.to_string()


def _has_agg(e) -> bool:
    from .sql.ast import contains_aggregate
    return contains_aggregate(e)


class StreamReader:
    """A stream of record batches as query input (vinum/api/stream_reader.py:12-94): the file
    need not fit in memory.  Batches are coalesced to device-sized chunks; an aggregate query
    folds each chunk into the device aggregate, any other query keeps only the rows that pass
    its WHERE."""

    def __init__(self, reader):
        self._reader = reader
        self.last_stats: dict = {}

    @property
    def schema(self) -> pa.Schema:
        return self._reader.schema

    def sql(self, query: str) -> Table:
        stats: dict = {}
        out = Table(execute_sql_stream(query, self._reader, stats=stats))
        self.last_stats = stats
        return out

    def sql_pd(self, query: str):
        return self.sql(query).to_pandas()


# ------------------------------------------------------------------------- IO
def stream_csv(input_file, read_options=None, parse_options=None, convert_options=None) -> StreamReader:
    """vn.stream_csv (vinum/io/arrow.py:9-61): pyarrow's streaming CSV reader as a query source."""
    import pyarrow.csv
    return StreamReader(pyarrow.csv.open_csv(input_file, read_options=read_options, parse_options=parse_options,
                                             convert_options=convert_options))


def read_csv(input_file, read_options=None, parse_options=None, convert_options=None) -> Table:
    import pyarrow.csv
    return Table(pyarrow.csv.read_csv(input_file, read_options=read_options, parse_options=parse_options,
                                      convert_options=convert_options))


def read_json(input_file, read_options=None, parse_options=None) -> Table:
    import pyarrow.json
    return Table(pyarrow.json.read_json(input_file, read_options=read_options, parse_options=parse_options))


def read_parquet(source, columns=None, **kwargs) -> Table:
    import pyarrow.parquet
    return Table(pyarrow.parquet.read_table(source, columns=columns, **kwargs))
