"""save / load in the reference's file format (eva/serialization/*.proto): wire-level structure and
round trips of programs, parameters and signatures (CPU); keys and ciphertexts on the GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from eva_b200 import evaluate, program_io, serialization as ser  # noqa: E402

FIXTURES = sorted(f[:-5] for f in os.listdir(os.path.join(os.path.dirname(__file__), "golden", "programs")) if f.endswith(".json"))


def _varint(b, i):
    v = s = 0
    while True:
        c = b[i]; i += 1
        v |= (c & 0x7F) << s; s += 7
        if not c & 0x80:
            return v, i


def _fields(b):
    """generic protobuf wire decoder: [(field number, wire type, value)]"""
    out, i = [], 0
    while i < len(b):
        key, i = _varint(b, i)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(b, i)
        elif wt == 1:
            v, i = b[i:i + 8], i + 8
        elif wt == 2:
            n, i = _varint(b, i)
            v, i = b[i:i + n], i + n
        elif wt == 5:
            v, i = b[i:i + 4], i + 4
        else:
            raise AssertionError("unexpected wire type %d" % wt)
        out.append((fn, wt, v))
    return out


def test_wire_format_matches_the_reference_schema():
    """known_type.proto: KnownType{1: Any{1: type_url, 2: value}, 2: creator}; eva.proto: Program{1 ir_version = 2,
    2 name, 3 vec_size, 4 terms{1 op, 2 packed operands, 3 attributes{1 key, 2|3|4|5 value}}, 5 inputs, 6 outputs}"""
    d = program_io.load_json("sobel")
    prog = program_io.build_program(d)[0]
    top = _fields(ser.dumps(prog))
    assert [f[0] for f in top] == [1, 2] and top[1][2].startswith(b"EVA ")
    any_ = dict((f[0], f[2]) for f in _fields(top[0][2]))
    assert any_[1] == b"type.googleapis.com/eva.msg.Program"
    pf = _fields(any_[2])
    scalars = {f[0]: f[2] for f in pf if f[0] in (1, 2, 3)}
    assert scalars[1] == 2 and scalars[2] == d["name"].encode() and scalars[3] == d["vec_size"]
    terms = [f[2] for f in pf if f[0] == 4]
    assert len(terms) == len(d["terms"]) and len([f for f in pf if f[0] == 5]) == 1 and len([f for f in pf if f[0] == 6]) == 1
    ops = {"Input": 1, "Output": 2, "Constant": 3, "Negate": 10, "Add": 11, "Sub": 12, "Mul": 13, "RotateLeftConst": 14,
           "RotateRightConst": 15, "Relinearize": 20, "ModSwitch": 21, "Rescale": 22, "Encode": 23}
    seen_ops = []
    for i, t in enumerate(terms):
        tf = _fields(t)
        op = [f[2] for f in tf if f[0] == 1]
        seen_ops.append(op[0] if op else 0)
        for f in tf:
            if f[0] == 2:   # packed operand indices: all smaller than this term's index (topological order)
                j = 0
                while j < len(f[2]):
                    v, j = _varint(f[2], j)
                    assert v < i
            if f[0] == 3:
                af = dict((g[0], g[2]) for g in _fields(f[2]))
                assert 1 <= af[1] <= 7 and len(af) <= 2
    import collections
    assert collections.Counter(seen_ops) == collections.Counter(ops[t["op"]] for t in d["terms"])
    # rotation attribute: key 2, sint32 (zigzag) in field 3
    rot = [dict((g[0], g[2]) for g in _fields(f[2])) for t in terms for f in _fields(t) if f[0] == 3]
    rots = sorted(((a[3] >> 1) ^ -(a[3] & 1)) for a in rot if a[1] == 2)
    assert rots == sorted(t["rotation"] for t in d["terms"] if "rotation" in t)


@pytest.mark.parametrize("name", FIXTURES)
def test_program_roundtrip(name):
    d = program_io.load_json(name)
    if "error" in d:
        pytest.skip("fixture records a compile error")
    prog, params, sig, _ = program_io.build_program(d)
    p2 = ser.loads(ser.dumps(prog))
    assert ser.program_to_msg(p2).SerializeToString(deterministic=True) == ser.program_to_msg(prog).SerializeToString(deterministic=True)
    assert p2.name == prog.name and p2.vec_size == prog.vec_size and set(p2.inputs) == set(prog.inputs) and set(p2.outputs) == set(prog.outputs)
    ps, sg = ser.loads(ser.dumps(params)), ser.loads(ser.dumps(sig))
    assert (list(ps.prime_bits), set(ps.rotations), ps.poly_modulus_degree) == (list(params.prime_bits), set(params.rotations), params.poly_modulus_degree)
    assert sg.vec_size == sig.vec_size
    assert {k: (int(v.input_type), v.scale, v.level) for k, v in sg.inputs.items()} == {k: (int(v.input_type), v.scale, v.level) for k, v in sig.inputs.items()}
    rng = np.random.default_rng(0)
    x = {k: list(rng.uniform(-1, 1, prog.vec_size)) for k in prog.inputs}
    a, b = evaluate(prog, x), evaluate(p2, x)
    assert all(np.array_equal(a[k], b[k]) for k in a)


def test_errors():
    with pytest.raises(TypeError):
        ser.dumps(42)
    with pytest.raises(RuntimeError):
        ser.loads(b"\x0a\x05hello")
    known = ser._cls("KnownType")()
    known.contents.type_url = "type.googleapis.com/eva.msg.Nope"
    with pytest.raises(RuntimeError, match="Unknown inner message type"):
        ser.loads(known.SerializeToString())
    d = program_io.load_json("polynomial")
    msg = ser.program_to_msg(program_io.build_program(d)[0])
    msg.ir_version = 1
    with pytest.raises(RuntimeError, match="version mismatch"):
        ser.program_from_msg(msg)


@pytest.mark.gpu
def test_client_server_flow_through_files(tmp_path):
    """reference examples/serialization.py and tests/features.py:154 (test_serialization): every object
    crosses a file between compile, keygen, encrypt, execute and decrypt"""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import serialization as example
    assert example.main(str(tmp_path)) < 0.01
    # a public-context file holds no secret key material
    from eva_b200 import load
    blob = open(tmp_path / "server.sealpublic", "rb").read()
    sk = load(str(tmp_path / "client.sealsecret"))._export()["secret_key"]
    assert sk.tobytes()[:4096] not in blob


@pytest.mark.parametrize("seed", range(8))
def test_compiled_random_programs_roundtrip(seed):
    """DSL -> CKKSCompiler -> save/load: the compiled program (Encode / Rescale / Relinearize / ModSwitch terms with
    their attributes), the parameters and the signature survive the file format"""
    from eva_b200 import EvaProgram, Input, Output
    from eva_b200.ckks import CKKSCompiler
    rng = np.random.default_rng(500 + seed)
    vec = 64
    prog = EvaProgram("ser%d" % seed, vec_size=vec)
    with prog:
        x, y = Input("x"), Input("y", is_encrypted=bool(seed % 3))
        acc = x * float(rng.uniform(0.5, 1.5)) + y
        for _ in range(int(rng.integers(1, 4))):
            k = int(rng.integers(4))
            acc = acc * x if k == 0 else (acc << int(rng.integers(1, vec))) + acc if k == 1 else acc * [float(v) for v in rng.uniform(-1, 1, vec)] if k == 2 else acc - x
        Output("z", acc)
    prog.set_input_scales(30)
    prog.set_output_ranges(20)
    compiled, params, sig = CKKSCompiler(config={"warn_vec_size": "false"}).compile(prog)
    c2, p2, s2 = ser.loads(ser.dumps(compiled)), ser.loads(ser.dumps(params)), ser.loads(ser.dumps(sig))
    assert ser.program_to_msg(c2).SerializeToString(deterministic=True) == ser.program_to_msg(compiled).SerializeToString(deterministic=True)
    assert (list(p2.prime_bits), set(p2.rotations), p2.poly_modulus_degree) == (list(params.prime_bits), set(params.rotations), params.poly_modulus_degree)
    assert {k: (int(v.input_type), v.scale, v.level) for k, v in s2.inputs.items()} == {k: (int(v.input_type), v.scale, v.level) for k, v in sig.inputs.items()}
    vals = {"x": list(rng.uniform(-1, 1, vec)), "y": list(rng.uniform(-1, 1, vec))}
    a, b = evaluate(compiled, vals), evaluate(c2, vals)
    assert all(np.array_equal(a[k], b[k]) for k in a)
