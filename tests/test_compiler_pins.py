"""The exact parameter pins of the reference's own test-suite (the only bit-level facts the
reference tests assert): tests/features.py:129,133 and tests/bug_fixes.py:68.  No GPU."""
from eva import EvaProgram, Input, Output
from eva.ckks import CKKSCompiler


def test_reduction_balancer_pins():
    prog = EvaProgram('ReductionTree', vec_size=16384)
    with prog:
        x1, x2, x3, x4 = Input('x1'), Input('x2'), Input('x3'), Input('x4')
        Output('y', (x1 * (x2 * (x3 * x4))) + (x1 + (x2 + (x3 + x4))))
    prog.set_output_ranges(20)
    prog.set_input_scales(60)
    _, params, _ = CKKSCompiler({'rescaler': 'always', 'balance_reductions': 'false', 'warn_vec_size': 'false'}).compile(prog)
    assert list(params.prime_bits) == [60, 20, 60, 60, 60, 60]
    _, params, _ = CKKSCompiler({'rescaler': 'always', 'balance_reductions': 'true', 'warn_vec_size': 'false'}).compile(prog)
    assert list(params.prime_bits) == [60, 20, 60, 60, 60]
    assert params.poly_modulus_degree == 32768


def test_output_rescaled_pin():
    prog = EvaProgram('OutputRescaled', vec_size=4)
    with prog:
        x = Input('x')
        Output('y', x * x)
    prog.set_output_ranges(20)
    prog.set_input_scales(60)
    _, params, _ = CKKSCompiler({'rescaler': 'lazy_waterline', 'warn_vec_size': 'false'}).compile(prog)
    assert list(params.prime_bits) == [60, 20, 60, 60]


def test_benchmark_program_parameters():
    """SURVEY.md Appendix B facts for the BASELINE configs"""
    from tests_programs import harris, polynomial, sobel, wide
    for prog, bits, n, rots in ((sobel(), [60] * 5, 16384, [0, 1, 2, 64, 65, 66, 128, 129, 130]),
                                (harris(), [60] * 5, 16384, [0, 1, 2, 64, 65, 66, 128, 129, 130]),
                                (polynomial(), [60, 60, 60], 8192, []),
                                (wide(64), [60, 50, 60], 16384, list(range(64)))):
        _, params, _ = CKKSCompiler({'warn_vec_size': 'false'}).compile(prog)
        assert list(params.prime_bits) == bits and params.poly_modulus_degree == n and sorted(params.rotations) == rots
