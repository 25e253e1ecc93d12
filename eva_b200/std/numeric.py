"""Library helpers (reference python/eva/std/numeric.py:5-21)."""
from .. import py_to_eva


def horizontal_sum(x):
    """Sum of all vector elements, replicated into every slot (log2(n) rotations)."""
    x = py_to_eva(x)
    i = 1
    while i < x.program.vec_size:
        x = x + (x << i)
        i <<= 1
    return x
