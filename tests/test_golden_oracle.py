"""Standalone oracle against the committed fixtures recorded from the reference core (tests/golden/make_golden.py)."""
import pytest

from checkers import oracle_lp
from common import GOLDEN_CASES, build_from_golden, check_against_golden, load_golden


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("native", [True, False])
def test_oracle_reproduces_reference_fixture(name, native):
    g = load_golden(name)
    check_against_golden(build_from_golden(oracle_lp("f64"), g, native=native), g, exact=True)


@pytest.mark.parametrize("name", ["grid_potts", "grid_trunclin"])
def test_oracle_grid_api_reproduces_reference_fixture(name):
    g = load_golden(name)
    check_against_golden(build_from_golden(oracle_lp("f64"), g, use_grid_api=True), g, exact=True)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_f32_tracks_reference_fixture(name):
    g = load_golden(name)
    check_against_golden(build_from_golden(oracle_lp("f32"), g), g, exact=False, rtol=1e-5)
