"""The assertions the parity tests rely on must themselves fail when they should."""
import numpy as np
import pytest

from helpers import max_abs


def test_max_abs_rejects_one_sided_nan_and_inf():
    assert max_abs([1.0, np.nan], [1.5, np.nan]) == 0.5          # NaN on both sides: same value
    assert max_abs([np.inf, -np.inf], [np.inf, -np.inf]) == 0.0
    for a, b in (([np.nan], [1.0]), ([1.0], [np.nan]), ([np.inf], [1.0]), ([np.inf], [-np.inf]), ([0.0], [-np.inf])):
        with pytest.raises(AssertionError):
            max_abs(a, b)
    assert max_abs(np.zeros((0, 3)), np.zeros((0, 3))) == 0.0
    with pytest.raises(AssertionError):
        max_abs(np.zeros(3), np.zeros(4))
