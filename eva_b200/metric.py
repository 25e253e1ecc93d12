"""Accuracy metric used by the reference's tests and examples (python/eva/metric.py:6-19)."""


def valuation_mse(a, b):
    """Mean squared error between two valuations (dicts name -> list of numbers)."""
    if set(a.keys()) != set(b.keys()):
        raise ValueError("Valuations must have the same keys")
    total = 0.0
    for k in a.keys():
        if len(a[k]) != len(b[k]):
            raise ValueError("Values must have the same length")
        total += sum((x - y) ** 2 for x, y in zip(a[k], b[k])) / len(a[k])
    return total / len(a)
