"""Library helpers on top of the DSL (the reference ships eva.std.numeric.horizontal_sum)."""
from .. import py_to_eva


def horizontal_sum(x):
    """Every slot of the result holds the sum of all slots of x: rotate-and-add doubling, log2(vec_size) steps."""
    acc = py_to_eva(x)
    shift, width = 1, acc.program.vec_size
    while shift < width:
        acc = acc + (acc << shift)
        shift *= 2
    return acc
