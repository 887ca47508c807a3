"""SQL front end: parser (stand-in for the reference's pglast parser) + device query engine."""
from .ast import Column, Expression, Literal, Op, Query, SortOrder  # noqa: F401
from .parser import ParserError, parse_sql  # noqa: F401
