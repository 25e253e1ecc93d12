"""The repo's own CKKS compiler (eva_b200/csrc/host/compiler.hpp) versus the
REFERENCE compiler: every fixture in tests/golden/programs was produced by the
reference's own code (oracle/_ref/eva_ref_compile); here the same source
programs are compiled by this repo and compared:
  * encryption parameters (prime_bits, rotations, poly_modulus_degree) and signature: equal;
  * the compiled DAG: equal as a multiset of canonical term signatures (op,
    attributes, operand signatures) -- i.e. isomorphic up to term numbering.
    Commutative reductions may associate differently (the reference's traversal
    order depends on pointer hashing), so Add/Mul operand signatures are sorted;
  * plaintext semantics of the compiled program equal the source program's.
No GPU needed."""
import collections
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from eva_b200 import Op, Program, Type, evaluate
from eva_b200.ckks import CKKSCompiler

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "programs", "*.json")))


def build_source(lines):
    prog, T, scale, rng, cfg = None, {}, 0, 0, {}
    for line in lines:
        w = line.split()
        if not w:
            continue
        if w[0] == "program":
            prog = Program(w[1], int(w[2]))
        elif w[0] == "input":
            T[int(w[1])] = prog._make_input(w[2], {"cipher": Type.Cipher, "raw": Type.Raw, "plain": Type.Plain}[w[3]])
        elif w[0] == "uconst":
            T[int(w[1])] = prog._make_uniform_constant(float(w[2]))
        elif w[0] == "dconst":
            T[int(w[1])] = prog._make_dense_constant([float(x) for x in w[3:3 + int(w[2])]])
        elif w[0] == "term":
            T[int(w[1])] = prog._make_term(getattr(Op, w[2]), [T[int(x)] for x in w[3:]])
        elif w[0] == "rotl":
            T[int(w[1])] = prog._make_left_rotation(T[int(w[2])], int(w[3]))
        elif w[0] == "rotr":
            T[int(w[1])] = prog._make_right_rotation(T[int(w[2])], int(w[3]))
        elif w[0] == "output":
            prog._make_output(w[1], T[int(w[2])])
        elif w[0] == "scales":
            scale = int(w[1])
        elif w[0] == "ranges":
            rng = int(w[1])
    prog.set_output_ranges(rng)
    prog.set_input_scales(scale)
    return prog


def sig_of_terms(terms):
    """terms: list of dicts {id, op, args, attrs...} in topological order -> Counter of canonical hashes"""
    h = {}
    for t in terms:
        a = [h[x] for x in t["args"]]
        if t["op"] in ("Add", "Mul"):
            a = sorted(a)
        const = [float(x) for x in t["const"]] if "const" in t else None
        key = json.dumps([t["op"], t.get("rotation"), t.get("divisor"), t.get("scale"), t.get("level"), t.get("type"),
                          const, t.get("name"), a], sort_keys=True)
        h[t["id"]] = hashlib.sha1(key.encode()).hexdigest()
    return collections.Counter(h.values())


def terms_of_program(prog):
    names = {t.index: n for n, t in list(prog.inputs.items()) + list(prog.outputs.items())}
    out = []
    for t in prog.terms():
        a = t.attributes
        d = {"id": t.index, "op": str(t.op).split(".")[1], "args": [o.index for o in t.operands]}
        if "RotationAttribute" in a: d["rotation"] = a["RotationAttribute"]
        if "RescaleDivisorAttribute" in a: d["divisor"] = a["RescaleDivisorAttribute"]
        if "EncodeAtScaleAttribute" in a: d["scale"] = a["EncodeAtScaleAttribute"]
        if "EncodeAtLevelAttribute" in a: d["level"] = a["EncodeAtLevelAttribute"]
        if "TypeAttribute" in a: d["type"] = str(a["TypeAttribute"]).split(".")[1]
        if "ConstantValueAttribute" in a:
            c = a["ConstantValueAttribute"]
            d["const"] = c if len(set(c)) > 1 else c[:1]
        if t.index in names: d["name"] = names[t.index]
        out.append(d)
    return out


def golden_terms(d):
    names = {v: k for k, v in list(d["inputs"].items()) + list(d["outputs"].items())}
    out = []
    for t in d["terms"]:
        t = dict(t)
        t.pop("range", None)
        if t["id"] in names: t["name"] = names[t["id"]]
        if "const" in t and len(set(t["const"])) == 1: t["const"] = t["const"][:1]
        out.append(t)
    return out


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-5] for p in FIXTURES])
def test_compiler_matches_reference(path):
    d = json.load(open(path))
    src = build_source(d["source"])
    if "error" in d:   # the reference compiler rejects this program/config: same exception text expected
        with pytest.raises(RuntimeError) as ei:
            CKKSCompiler({k: str(v) for k, v in d["config"].items()} | {"warn_vec_size": "false"}).compile(src)
        assert str(ei.value) == d["error"]
        return
    compiled, params, sig = CKKSCompiler({k: str(v) for k, v in d["config"].items()} | {"warn_vec_size": "false"}).compile(src)
    assert list(params.prime_bits) == d["prime_bits"]
    assert sorted(params.rotations) == d["rotations"]
    assert params.poly_modulus_degree == d["poly_modulus_degree"]
    assert sig.vec_size == d["vec_size"]
    assert {k: (str(v.input_type).split(".")[1], v.scale, v.level) for k, v in sig.inputs.items()} == \
           {k: (v["type"], v["scale"], v["level"]) for k, v in d["signature"].items()}
    mine, ref = terms_of_program(compiled), golden_terms(d)
    hist = lambda ts: collections.Counter((t["op"], t.get("scale"), t.get("level"), t.get("divisor"), t.get("rotation")) for t in ts)
    assert hist(mine) == hist(ref)
    assert sig_of_terms(mine) == sig_of_terms(ref)
    # plaintext semantics: compiled == source (reference tests/common.py:25, MSE < 1e-10)
    rng = np.random.default_rng(0)
    inputs = {name: list(rng.uniform(-2, 2, d["vec_size"])) for name in d["signature"]}
    a, b = evaluate(src, inputs), evaluate(compiled, inputs)
    for k in a:
        assert np.mean((np.array(a[k]) - np.array(b[k])) ** 2) < 1e-10


def test_unset_scale_and_unknown_input_errors():
    p = Program("e", 8)
    x = p._make_input("x", Type.Cipher)
    p._make_output("y", x)
    with pytest.raises(RuntimeError, match="scale for input x"):
        CKKSCompiler().compile(p)
    p.set_input_scales(30); p.set_output_ranges(10)
    with pytest.raises((IndexError, KeyError, RuntimeError)):
        evaluate(p, {"nope": [0.0] * 8})
    with pytest.raises((ValueError, RuntimeError)):
        Program("bad", 6)
