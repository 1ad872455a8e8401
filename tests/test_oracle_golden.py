"""CPU: the C oracle must reproduce the reference's golden vectors bit-exactly (pins the oracle)."""
import numpy as np
import pytest

import golden_util as gu


@pytest.mark.parametrize("name", gu.fixture_names())
def test_oracle_matches_reference_golden(oracle_lib, name):
    fx = gu.load(name)
    gu.replay(gu.OracleBackend(oracle_lib, fx), fx)


def test_golden_fixtures_present():
    names = gu.fixture_names()
    assert len(names) >= 8
    assert any("cramped" in n for n in names)
