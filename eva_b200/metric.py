"""Accuracy metric of the reference's tests and examples (the `eva.metric` module)."""
import numpy as np


def valuation_mse(a, b):
    """mean over the named vectors of their mean squared difference; both valuations (name -> sequence of
    numbers) must name the same vectors with the same lengths"""
    names = sorted(a)
    if names != sorted(b):
        raise ValueError("Valuations must have the same keys")
    per_vector = []
    for name in names:
        u, v = np.asarray(a[name], dtype=np.float64), np.asarray(b[name], dtype=np.float64)
        if u.shape != v.shape:
            raise ValueError("Values must have the same length")
        per_vector.append(float(np.mean((u - v) ** 2)))
    return float(np.mean(per_vector))
