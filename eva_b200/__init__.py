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
import threading

from ._eva_b200 import (Op, Program, Term, Type, evaluate, set_num_threads)  # noqa: F401

__version__ = "0.1.0"

# the program whose ``with`` block is open (the DSL's free functions Input / Output and the
# constant conversions build into it); one per thread
_scope = threading.local()


def _open_program():
    prog = getattr(_scope, "program", None)
    if prog is None:
        raise RuntimeError("No Program in context")
    return prog


def _as_term(value, program):
    """Term for a DSL operand: an Expr / Term as is, a number or a list as a new Constant of `program`"""
    if isinstance(value, Expr):
        return value.term
    if isinstance(value, Term):
        return value
    if isinstance(value, numbers.Number):
        return program._make_uniform_constant(value)
    if isinstance(value, list):
        return program._make_dense_constant(value)
    raise TypeError("No conversion to Term available for " + str(value))


def py_to_eva(x, program=None):
    """Expr for x: Expr instances pass through; Terms, lists and numbers are wrapped (constants are
    created in `program`, default: the program in context)."""
    if isinstance(x, Expr):
        return x
    program = program if program is not None else _open_program()
    return Expr(_as_term(x, program), program)


class Expr:
    """A Term of an EvaProgram with Python operators: + - * (with Exprs, numbers or lists on either side),
    ** (positive integer powers), << >> (slot rotations) and unary minus."""

    def __init__(self, term, program):
        self.term, self.program = term, program

    def _emit(self, op, *operands):
        return Expr(self.program._make_term(op, [_as_term(o, self.program) for o in operands]), self.program)

    def __neg__(self):
        return self._emit(Op.Negate, self)

    def __pow__(self, exponent):
        if exponent < 1:
            raise ValueError("exponent must be greater than zero, got " + str(exponent))
        acc = self
        for _ in range(1, exponent):     # x * x * ... (a chain: the compiler balances / relinearizes it)
            acc = acc._emit(Op.Mul, acc, self)
        return acc

    def __lshift__(self, rotation):
        return Expr(self.program._make_left_rotation(self.term, rotation), self.program)

    def __rshift__(self, rotation):
        return Expr(self.program._make_right_rotation(self.term, rotation), self.program)


def _binary(op, reflected):
    def method(self, other):
        return self._emit(op, other, self) if reflected else self._emit(op, self, other)
    return method


for _name, _op in (("add", Op.Add), ("sub", Op.Sub), ("mul", Op.Mul)):
    setattr(Expr, "__%s__" % _name, _binary(_op, False))
    setattr(Expr, "__r%s__" % _name, _binary(_op, True))


class EvaProgram(Program):
    """A Program usable as a ``with`` block: inside it Input(...) / Output(...) add terms to it."""

    def __enter__(self):
        if getattr(_scope, "program", None) is not None:
            raise RuntimeError("There is already an EVA Program in context")
        _scope.program = self

    def __exit__(self, exc_type, exc_value, exc_traceback):
        if getattr(_scope, "program", None) is not self:
            raise RuntimeError("This program is not currently in context")
        _scope.program = None


def Input(name, is_encrypted=True):
    """a named program input: a ciphertext, or (is_encrypted=False) a raw vector"""
    prog = _open_program()
    return Expr(prog._make_input(name, Type.Cipher if is_encrypted else Type.Raw), prog)


def Output(name, expr):
    """name `expr` as a program output"""
    prog = _open_program()
    prog._make_output(name, _as_term(expr, prog))


def save(obj, path):
    """reference python/eva/__init__.py `save`: protobuf files of eva/serialization (see serialization.py)"""
    from . import serialization
    serialization.save(obj, path)


def load(path):
    """reference python/eva/__init__.py `load`"""
    from . import serialization
    return serialization.load(path)
