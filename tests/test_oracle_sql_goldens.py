"""oracle/join.cpp against the SQL known answers of tests/sql_goldens.py (OtherCondition, several equal conditions)."""
import pytest

import sql_goldens
from nested_loop import nested_loop_join
from test_oracle_join import run_oracle


@pytest.mark.parametrize("case", sql_goldens.cases(), ids=lambda c: c[0])
def test_oracle_sql_join_goldens(case):
    name, plan, left, right, count, rows = case
    got = run_oracle(plan, left, right)
    assert len(got) == count
    if rows is not None:
        assert sorted(got) == sorted(rows)
    assert sorted(nested_loop_join(plan, left, right), key=str) == sorted(got, key=str)   # the independent restatement agrees too
