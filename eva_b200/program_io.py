"""Loads a compiled EVA program from its JSON dump (term list + parameters +
signature; the format written by tests/golden/gen_golden.py from the reference
compiler's output) into an eva_b200 Program that the executor can run."""
import json
import os

from . import Op, Program, Type
from .ckks import CKKSEncodingInfo, CKKSParameters, CKKSSignature

HERE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")
_ATTR = {"rotation": "RotationAttribute", "divisor": "RescaleDivisorAttribute", "scale": "EncodeAtScaleAttribute",
         "level": "EncodeAtLevelAttribute", "range": "RangeAttribute"}


def load_json(name):
    """a reference-compiled program by name, or any path to a .json / .json.gz dump"""
    path = name if os.path.sep in name or name.endswith((".json", ".gz")) else os.path.join(HERE, "golden", "programs", name + ".json")
    if not os.path.exists(path) and os.path.exists(path + ".gz"):
        path += ".gz"
    if path.endswith(".gz"):
        import gzip
        with gzip.open(path, "rt") as f:
            return json.load(f)
    with open(path) as f:
        return json.load(f)


def build_program(d):
    """returns (program, params, signature, {fixture term id -> Term})"""
    p = Program(d["name"], d["vec_size"])
    in_names = {v: k for k, v in d["inputs"].items()}
    out_names = {v: k for k, v in d["outputs"].items()}
    terms = {}
    for t in d["terms"]:
        op = getattr(Op, t["op"])
        args = [terms[a] for a in t["args"]]
        if op == Op.Input:
            term = p._make_input(in_names[t["id"]], getattr(Type, t["type"]))
        elif op == Op.Constant:
            c = t["const"]
            term = p._make_uniform_constant(c[0]) if len(c) == 1 else p._make_dense_constant(c)
        elif op == Op.Output:
            term = p._make_output(out_names[t["id"]], args[0])
        elif op == Op.RotateLeftConst:
            term = p._make_left_rotation(args[0], t["rotation"])
        elif op == Op.RotateRightConst:
            term = p._make_right_rotation(args[0], t["rotation"])
        else:
            term = p._make_term(op, args)
        attrs = {_ATTR[k]: t[k] for k in _ATTR if k in t and not (k == "rotation")}
        if op not in (Op.Input,) and "type" in t:
            attrs["TypeAttribute"] = getattr(Type, t["type"])
        if attrs:
            term._set_attributes(attrs)
        terms[t["id"]] = term
    params = CKKSParameters(d["prime_bits"], set(d["rotations"]), d["poly_modulus_degree"])
    sig = CKKSSignature(d["vec_size"], {k: CKKSEncodingInfo(getattr(Type, v["type"]), v["scale"], v["level"])
                                        for k, v in d["signature"].items()})
    return p, params, sig, terms
