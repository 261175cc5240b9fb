"""Host-side helpers that stay on the CPU (post-processing of root statistics in `act()`)."""
import numpy as np


def stable_normalizer(x: np.ndarray, temp: float) -> np.ndarray:
    """x[i]**temp / sum_i x[i]**temp, normalised by the max first (reference alphazero/helpers.py:9-27)."""
    x = (x / np.max(x)) ** temp
    return np.abs(x / np.sum(x))
