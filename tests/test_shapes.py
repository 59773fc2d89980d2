"""auto_shape / tt_shape must reproduce the reference's choices exactly."""
import json
import os

import pytest
from hypothesis import given, settings, strategies as st

from helpers import GOLDEN, oracle
from tensorized_rnn_b200.shapes import auto_shape, tt_shape


def _table():
    with open(os.path.join(GOLDEN, "shapes.json")) as f:
        return json.load(f)


def test_auto_shape_matches_reference_table():
    t = _table()["auto_shape"]
    assert len(t) > 2000
    for key, want in t.items():
        n, d = map(int, key.split(","))
        assert auto_shape(n, d=d) == want, (n, d)


def test_tt_shape_matches_reference_table():
    for e in _table()["tt_shape"]:
        got = tt_shape(e["in"], e["hidden"], e["n_cores"], e["n_gates"], new_core=e["new_core"])
        assert got == e["shape"], e


def test_survey_golden_rows():
    # the rows SURVEY.md section 8a-1 lists explicitly
    assert tt_shape(1, 256, 2, 4) == [[1, 1], [32, 32]]
    assert tt_shape(256, 256, 2, 3) == [[16, 16], [24, 32]]
    assert tt_shape(40, 256, 3, 4) == [[2, 4, 5], [8, 8, 16]]
    assert tt_shape(40, 256, 4, 4) == [[2, 2, 2, 5], [4, 4, 8, 8]]
    assert tt_shape(1024, 1024, 4, 4) == [[4, 4, 8, 8], [8, 8, 8, 8]]
    assert tt_shape(768, 768, 4, 4) == [[4, 4, 6, 8], [6, 8, 8, 8]]


@settings(max_examples=150, deadline=None)
@given(n=st.integers(min_value=1, max_value=5000), d=st.integers(min_value=1, max_value=4))
def test_auto_shape_matches_oracle_restatement(n, d):
    # the oracle follows the reference's sympy/scipy procedure step by step
    assert auto_shape(n, d=d) == oracle.auto_shape(n, d)


def test_bad_arguments():
    with pytest.raises(AssertionError):
        tt_shape(4, 4, 2, 4, new_core="middle")
    with pytest.raises(NotImplementedError):
        auto_shape(12, d=2, criterion="var")
    with pytest.raises(ValueError):
        auto_shape(0, d=2)
