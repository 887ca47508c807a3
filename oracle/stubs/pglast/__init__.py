"""Stub of pglast (pinned ==1.17 by the reference's setup.py:35; a C extension that is
not installable here).  It exists only so that `import vinum` works when the golden
generator imports the reference's Python operators; the SQL parser is never called.
TEST INFRASTRUCTURE ONLY."""


class Node:  # pragma: no cover - placeholder type
    pass


def parse_sql(sql):  # pragma: no cover
    raise RuntimeError("pglast stub: the SQL parser is not available in this environment")
