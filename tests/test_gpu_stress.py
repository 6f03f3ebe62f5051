"""A short pass of the randomised kernel-path cross-check (tests/stress_paths.py): seeded random networks x every
formulation x with / without a status mask -- the default, warp-tile, CTA-tile and staged paths agree bitwise and the
default path matches the oracle."""
import pytest

import stress_paths

pytestmark = pytest.mark.gpu


def test_random_networks_all_paths_agree():
    assert stress_paths.main(10) == 0
