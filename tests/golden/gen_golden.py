"""Generates tests/golden/programs/*.json by running the REFERENCE's own
compiler (oracle/_ref/eva_ref_compile, built by oracle/ref_compiler/build.sh
from /root/reference) on the benchmark and test programs.  Only runs in the
build container (the GPU box has no /root/reference); the JSON fixtures are
committed.  The mini-DSL below mirrors python/eva/__init__.py:57-163 (operand
order of __radd__/__rmul__, ** as repeated Mul, << / >> rotations).

Usage: python tests/golden/gen_golden.py
"""
import json
import numbers
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PROBE = os.path.join(ROOT, "oracle", "_ref", "eva_ref_compile")


class Prog:
    def __init__(self, name, vec_size):
        self.lines = ["program %s %d" % (name, vec_size)]
        self.n = 0
        self.name, self.vec_size = name, vec_size

    def _id(self):
        self.n += 1
        return self.n

    def term(self, op, args):
        i = self._id(); self.lines.append("term %d %s %s" % (i, op, " ".join(map(str, args)))); return i

    def const(self, x):
        i = self._id()
        if isinstance(x, list):
            self.lines.append("dconst %d %d %s" % (i, len(x), " ".join(repr(float(v)) for v in x)))
        else:
            self.lines.append("uconst %d %r" % (i, float(x)))
        return i

    def to_term(self, x):
        if isinstance(x, Expr):
            return x.t
        if isinstance(x, (list, numbers.Number)):
            return self.const(x)
        raise TypeError(x)

    def input(self, name, is_encrypted=True):
        i = self._id(); self.lines.append("input %d %s %s" % (i, name, "cipher" if is_encrypted else "raw")); return Expr(i, self)

    def output(self, name, e):
        self.lines.append("output %s %d" % (name, self.to_term(e)))

    def text(self, scale, rng, config=None, eval_inputs=None):
        ls = list(self.lines) + ["scales %d" % scale, "ranges %d" % rng]
        for k, v in (config or {}).items():
            ls.append("config %s %s" % (k, v))
        for k, v in (eval_inputs or {}).items():
            ls.append("evalinput %s %d %s" % (k, len(v), " ".join(repr(float(x)) for x in v)))
        return "\n".join(ls) + "\n"


class Expr:
    def __init__(self, t, p):
        self.t, self.p = t, p

    def __add__(self, o): return Expr(self.p.term("Add", [self.t, self.p.to_term(o)]), self.p)
    def __radd__(self, o): return Expr(self.p.term("Add", [self.p.to_term(o), self.t]), self.p)
    def __sub__(self, o): return Expr(self.p.term("Sub", [self.t, self.p.to_term(o)]), self.p)
    def __rsub__(self, o): return Expr(self.p.term("Sub", [self.p.to_term(o), self.t]), self.p)
    def __mul__(self, o): return Expr(self.p.term("Mul", [self.t, self.p.to_term(o)]), self.p)
    def __rmul__(self, o): return Expr(self.p.term("Mul", [self.p.to_term(o), self.t]), self.p)

    def __pow__(self, e):
        r = self.t
        for _ in range(e - 1):
            r = self.p.term("Mul", [r, self.t])
        return Expr(r, self.p)

    def __lshift__(self, r):
        i = self.p._id(); self.p.lines.append("rotl %d %d %d" % (i, self.t, r)); return Expr(i, self.p)

    def __rshift__(self, r):
        i = self.p._id(); self.p.lines.append("rotr %d %d %d" % (i, self.t, r)); return Expr(i, self.p)

    def __neg__(self): return Expr(self.p.term("Negate", [self.t]), self.p)


# ---- programs (reference examples/image_processing.py:12-100, README.md:93-105) ----
def convolution(image, width, filt):
    for i in range(len(filt)):
        for j in range(len(filt[0])):
            rotated = image << i * width + j
            partial = rotated * filt[i][j]
            convolved = partial if (i == 0 and j == 0) else convolved + partial
    return convolved


def convolutionXY(image, width, filt):
    for i in range(len(filt)):
        for j in range(len(filt[0])):
            rotated = image << (i * width + j)
            horizontal = rotated * filt[i][j]
            vertical = rotated * filt[j][i]
            if i == 0 and j == 0:
                Ix, Iy = horizontal, vertical
            else:
                Ix = Ix + horizontal
                Iy = Iy + vertical
    return Ix, Iy


SOBEL_FILTER = [[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]]


def sobel(h=64, w=64):
    p = Prog("sobel", h * w)
    image = p.input("image")
    a1, a2, a3 = 2.2137874823876622, -1.0984324107372518, 0.17254603006834726
    conv_hor, conv_ver = convolutionXY(image, w, SOBEL_FILTER)
    dsq = conv_hor ** 2 + conv_ver ** 2
    dsq2 = dsq * dsq
    dsq3 = dsq2 * dsq
    p.output("image", dsq * a1 + dsq2 * a2 + dsq3 * a3)
    return p, 25, 10


def harris(h=64, w=64):
    p = Prog("harris", h * w)
    image = p.input("image")
    pool = [[1, 1, 1]] * 3
    c = 0.04
    Ix, Iy = convolutionXY(image, w, SOBEL_FILTER)
    Ixx, Iyy, Ixy = Ix ** 2, Iy ** 2, Ix * Iy
    Sxx, Syy, Sxy = convolution(Ixx, w, pool), convolution(Iyy, w, pool), convolution(Ixy, w, pool)
    det = Sxx * Syy - Sxy * Sxy
    trace = Sxx + Syy
    p.output("image", det - trace ** 2 * c)
    return p, 30, 20


def polynomial():
    p = Prog("poly", 1024)
    x = p.input("x")
    p.output("y", 3 * x ** 2 + 5 * x - 2)
    return p, 30, 30


def wide(nprod):
    """BASELINE config 5: nprod independent rot(x,i%64)*rot(y,(i//64)%64) products, tree-summed."""
    p = Prog("wide%d" % nprod, 8192)
    x, y = p.input("x"), p.input("y")
    terms = [(x << (i % 64)) * (y << ((i // 64) % 64)) for i in range(nprod)]
    while len(terms) > 1:
        terms = [terms[i] + terms[i + 1] for i in range(0, len(terms), 2)]
    p.output("z", terms[0])
    return p, 40, 30


def sobel_large():
    """reference tests/large_programs.py:10-53 shape: 90x90 padded into vec 8192, scale 45, range 20."""
    p = Prog("sobel8192", 8192)
    image = p.input("image")
    a1, a2, a3 = 2.2137874823876622, -1.0984324107372518, 0.17254603006834726
    ch, cv = convolutionXY(image, 90, SOBEL_FILTER)
    dsq = ch ** 2 + cv ** 2
    dsq2 = dsq * dsq
    dsq3 = dsq2 * dsq
    p.output("image", dsq * a1 + dsq2 * a2 + dsq3 * a3)
    return p, 45, 20


def features_programs():
    """reference tests/features.py style small programs"""
    out = []
    for opname in ("add", "sub", "mul"):
        for enc1 in (True, False):
            for enc2 in (True, False):
                p = Prog("bin_%s_%d%d" % (opname, enc1, enc2), 64)
                a, b = p.input("a", enc1), p.input("b", enc2)
                p.output("y", a + b if opname == "add" else a - b if opname == "sub" else a * b)
                out.append((p, 30, 30))
    p = Prog("unary", 64)
    x = p.input("x")
    p.output("neg", -x); p.output("cube", x ** 3); p.output("const", p_expr_const(p, 7.5))
    out.append((p, 30, 30))
    for rot in (-2, -1, 0, 1):
        p = Prog("rot_%s" % str(rot).replace("-", "m"), 8)
        x = p.input("x")
        p.output("l", x << rot); p.output("r", x >> rot)
        out.append((p, 30, 30))
    p = Prog("mixed", 64)
    a, b, c = p.input("a"), p.input("b", False), p.input("c")
    p.output("y", (a * b + c) * [float(i % 5) for i in range(64)] - b)
    out.append((p, 30, 30))
    p = Prog("transparent", 64)     # tests/features.py:135  x - x + x*0
    x = p.input("x")
    p.output("y", x - x + x * 0)
    out.append((p, 30, 30))
    p = Prog("hsum", 2048)           # tests/std.py horizontal_sum
    x = p.input("x")
    i = 1
    while i < 2048:
        x = x + (x << i)
        i <<= 1
    p.output("y", x)
    out.append((p, 30, 30))
    p = Prog("deep", 1024)           # tests/bug_fixes.py:10 high inner scale style
    x = p.input("x")
    p.output("y", (x * x * x * x) * 0.5 + x)
    out.append((p, 60, 30))
    return out


def p_expr_const(p, v):
    return Expr(p.const(v), p)


def run(p, scale, rng, config=None, eval_inputs=None):
    txt = p.text(scale, rng, config, eval_inputs)
    res = subprocess.run([PROBE], input=txt.encode(), stdout=subprocess.PIPE, check=True).stdout.decode()
    d = json.loads(res)
    d["source"] = txt.splitlines()
    d["config"] = config or {}
    d["input_scale"], d["output_range"] = scale, rng
    return d


def main():
    out = os.path.join(HERE, "programs")
    os.makedirs(out, exist_ok=True)
    jobs = [("sobel", sobel(), None), ("harris", harris(), None), ("polynomial", polynomial(), None),
            ("wide64", wide(64), None), ("sobel8192", sobel_large(), None),
            ("wide4096", wide(4096), None)]     # BASELINE config 5: >= 4096 parallel ciphertext multiplications (stored gzipped)
    for resc in ("lazy_waterline", "eager_waterline", "always", "minimum"):
        jobs.append(("sobel_%s" % resc, sobel(), {"rescaler": resc}))
    jobs.append(("sobel_nobalance", sobel(), {"balance_reductions": "false"}))
    jobs.append(("sobel_eager_relin", sobel(), {"lazy_relinearize": "false"}))
    jobs.append(("polynomial_192q", polynomial(), {"security_level": "192", "quantum_safe": "true"}))
    for (p, s, r) in features_programs():
        jobs.append(("feat_" + p.name, (p, s, r), None))
    for name, (p, s, r), cfg in jobs:
        d = run(p, s, r, cfg)
        # the BASELINE workloads ship with the product (workloads/), everything else is a test fixture
        dest = os.path.join(HERE, "..", "..", "workloads") if name in ("sobel", "harris", "polynomial", "wide4096") else out
        os.makedirs(dest, exist_ok=True)
        if name == "wide4096":
            import gzip
            d.pop("source", None)    # the 16k-line source dump is reproducible from wide(4096)
            with gzip.open(os.path.join(dest, name + ".json.gz"), "wt") as f:
                json.dump(d, f, separators=(",", ":"))
        else:
            with open(os.path.join(dest, name + ".json"), "w") as f:
                json.dump(d, f, separators=(",", ":"))
        print(name, d.get("error") or (d["poly_modulus_degree"], d["prime_bits"], d["rotations"], len(d["terms"])))


if __name__ == "__main__":
    main()
