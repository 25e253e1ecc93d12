"""eva_b200 -- B200-native (sm_100a) CKKS execution backend behind the EVA API.

Same user surface as the reference's ``eva`` package (python/eva/__init__.py):
``EvaProgram``, ``Input``, ``Output``, ``Expr`` operator overloads, ``evaluate``,
``py_to_eva``; submodules ``ckks`` (compiler), ``b200`` (key generation /
encrypt / execute / decrypt; drop-in for ``eva.seal``), ``metric`` and ``std``.
The hot path -- homomorphic evaluation of the compiled term DAG -- runs in
hand-written CUDA kernels reached through the C-ABI of include/evab200.h.  There
is no CPU fallback: without the CUDA library or without a GPU the backend raises.
"""
import numbers

from ._eva_b200 import (Op, Program, Term, Type, evaluate, set_num_threads)  # noqa: F401

__version__ = "0.1.0"

_current_program = None


def _curr():
    if _current_program is None:
        raise RuntimeError("No Program in context")
    return _current_program


def _py_to_term(x, program):
    if isinstance(x, Expr):
        return x.term
    if isinstance(x, list):
        return program._make_dense_constant(x)
    if isinstance(x, numbers.Number):
        return program._make_uniform_constant(x)
    if isinstance(x, Term):
        return x
    raise TypeError("No conversion to Term available for " + str(x))


def py_to_eva(x, program=None):
    """Maps Expr instances, Terms, lists and numbers to Expr (constants are created in `program`)."""
    if isinstance(x, Expr):
        return x
    if program is None:
        program = _curr()
    return Expr(_py_to_term(x, program), program)


class Expr:
    """Operator-overloading wrapper around a native Term of an EvaProgram."""

    def __init__(self, term, program):
        self.term = term
        self.program = program

    def _bin(self, op, lhs, rhs):
        return Expr(self.program._make_term(op, [lhs, rhs]), self.program)

    def __add__(self, other): return self._bin(Op.Add, self.term, _py_to_term(other, self.program))
    def __radd__(self, other): return self._bin(Op.Add, _py_to_term(other, self.program), self.term)
    def __sub__(self, other): return self._bin(Op.Sub, self.term, _py_to_term(other, self.program))
    def __rsub__(self, other): return self._bin(Op.Sub, _py_to_term(other, self.program), self.term)
    def __mul__(self, other): return self._bin(Op.Mul, self.term, _py_to_term(other, self.program))
    def __rmul__(self, other): return self._bin(Op.Mul, _py_to_term(other, self.program), self.term)

    def __pow__(self, exponent):
        if exponent < 1:
            raise ValueError("exponent must be greater than zero, got " + str(exponent))
        result = self.term
        for _ in range(exponent - 1):
            result = self.program._make_term(Op.Mul, [result, self.term])
        return Expr(result, self.program)

    def __lshift__(self, rotation): return Expr(self.program._make_left_rotation(self.term, rotation), self.program)
    def __rshift__(self, rotation): return Expr(self.program._make_right_rotation(self.term, rotation), self.program)
    def __neg__(self): return Expr(self.program._make_term(Op.Negate, [self.term]), self.program)


class EvaProgram(Program):
    """A Program that is also a context manager for the Input/Output free functions."""

    def __init__(self, name, vec_size):
        super().__init__(name, vec_size)

    def __enter__(self):
        global _current_program
        if _current_program is not None:
            raise RuntimeError("There is already an EVA Program in context")
        _current_program = self

    def __exit__(self, exc_type, exc_value, exc_traceback):
        global _current_program
        if _current_program is not self:
            raise RuntimeError("This program is not currently in context")
        _current_program = None


def Input(name, is_encrypted=True):
    program = _curr()
    return Expr(program._make_input(name, Type.Cipher if is_encrypted else Type.Raw), program)


def Output(name, expr):
    program = _curr()
    program._make_output(name, _py_to_term(expr, program))



def save(obj, path):
    """reference python/eva/__init__.py `save`: protobuf files of eva/serialization (see serialization.py)"""
    from . import serialization
    serialization.save(obj, path)


def load(path):
    """reference python/eva/__init__.py `load`"""
    from . import serialization
    return serialization.load(path)
