"""Executes a compiled program (golden JSON) on the CPU ORACLE, term by term,
following the dispatch of the reference's SEALExecutor::operator()
(eva/seal/seal_executor.h:279-404).  Test infrastructure / CPU baseline only.
Supports serial execution and a dependency-counting thread pool equivalent to
MulticoreProgramTraversal (eva/common/multicore_program_traversal.h:55-79);
the oracle's C calls release the GIL, so threads scale across cores."""
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from oracle import oracle as o

CIPHER_OPS = ("Add", "Sub", "Mul", "Negate", "RotateLeftConst", "RotateRightConst", "Relinearize", "ModSwitch", "Rescale")


class OracleProgram:
    def __init__(self, d, orc):
        self.d, self.orc = d, orc
        self.terms = {t["id"]: t for t in d["terms"]}
        self.order = [t["id"] for t in d["terms"]]
        self.vec = d["vec_size"]
        self.in_names = {v: k for k, v in d["inputs"].items()}
        self.out_names = {v: k for k, v in d["outputs"].items()}
        self.rk = None
        self.gks = {}

    def prepare_keys(self):
        self.rk = self.orc.relin_key()
        for t in self.d["terms"]:
            if t["op"] in ("RotateLeftConst", "RotateRightConst") and t["rotation"] != 0:
                steps = t["rotation"] if t["op"] == "RotateLeftConst" else -t["rotation"]
                elt = o.galois_elt_from_step(self.orc.N, steps)
                if elt not in self.gks:
                    self.gks[elt] = self.orc.galois_key(elt)

    def cipher_op_count(self, kinds=None):
        kinds = kinds or self.kinds()
        return sum(1 for t in self.d["terms"] if t["op"] in CIPHER_OPS and kinds[t["id"]] == "cipher")

    def kinds(self):
        k = {}
        for t in self.d["terms"]:
            a = [k[x] for x in t["args"]]
            if t["op"] == "Input":
                k[t["id"]] = {"Cipher": "cipher", "Raw": "raw", "Plain": "plain"}[t["type"]]
            elif t["op"] == "Constant":
                k[t["id"]] = "raw"
            elif t["op"] == "Encode":
                k[t["id"]] = "plain"
            elif t["op"] == "Output":
                k[t["id"]] = a[0]
            elif all(x == "raw" for x in a):
                k[t["id"]] = "raw"
            else:
                k[t["id"]] = "cipher"
        return k

    def exec_term(self, t, V):
        """V: id -> ("cipher", array, scale) | ("plain", array, scale) | ("raw", list)"""
        orc, op = self.orc, t["op"]
        A = [V[x] for x in t["args"]]
        if op == "Input":
            return V[t["id"]]
        if op == "Constant":
            c = t["const"]
            return ("raw", np.tile(np.array(c, dtype=np.float64), self.vec // len(c)))
        if op == "Encode":
            ell = orc.k - 1 - t["level"]
            return ("plain", orc.encode(A[0][1], 2.0 ** t["scale"], ell), 2.0 ** t["scale"])
        if op == "Output":
            return A[0]
        if all(a[0] == "raw" for a in A):
            x = A[0][1]
            if op == "Add": return ("raw", x + A[1][1])
            if op == "Sub": return ("raw", x - A[1][1])
            if op == "Mul": return ("raw", x * A[1][1])
            if op == "Negate": return ("raw", -x)
            if op == "RotateLeftConst": return ("raw", np.roll(x, -t["rotation"]))
            if op == "RotateRightConst": return ("raw", np.roll(x, t["rotation"]))
        if op in ("Add", "Sub", "Mul"):
            x, y = A
            if x[0] != "cipher":               # seal_executor.h:115-119,153-157: swap so the cipher is first
                assert op != "Sub"
                x, y = y, x
            if y[0] == "cipher":
                if op == "Add": return ("cipher", orc.add(x[1], y[1]), x[2])
                if op == "Sub": return ("cipher", orc.sub(x[1], y[1]), x[2])
                if t["args"][0] == t["args"][1]: return ("cipher", orc.square(x[1]), x[2] * x[2])
                return ("cipher", orc.mul(x[1], y[1]), x[2] * y[2])
            if op == "Add": return ("cipher", orc.add_plain(x[1], y[1]), x[2])
            if op == "Sub": return ("cipher", orc.sub_plain(x[1], y[1]), x[2])
            return ("cipher", orc.mul_plain(x[1], y[1]), x[2] * y[2])
        x = A[0]
        if op == "Negate": return ("cipher", orc.negate(x[1]), x[2])
        if op in ("RotateLeftConst", "RotateRightConst"):
            steps = t["rotation"] if op == "RotateLeftConst" else -t["rotation"]
            if steps == 0: return ("cipher", x[1].copy(), x[2])
            elt = o.galois_elt_from_step(orc.N, steps)
            return ("cipher", orc.rotate(x[1], steps, self.gks[elt]), x[2])
        if op == "Relinearize": return ("cipher", orc.relinearize(x[1], self.rk), x[2])
        if op == "ModSwitch": return ("cipher", orc.mod_switch(x[1]), x[2])
        if op == "Rescale": return ("cipher", orc.rescale(x[1]), x[2] / 2.0 ** t["divisor"])   # seal_executor.h:214
        raise RuntimeError("Unhandled op " + op)

    def run(self, inputs, threads=1, keep=None):
        """inputs: name -> value tuple.  returns id -> value for every term; with keep=<ids> a value is dropped
        after its last use unless listed (the 16k-term wide DAG would otherwise hold ~10 GB of intermediates)."""
        V = {}
        for tid, name in self.in_names.items():
            V[tid] = inputs[name]
        left = None
        if keep is not None:
            keep = set(keep)
            left = {}
            for tid in self.order:
                for a in self.terms[tid]["args"]:
                    left[a] = left.get(a, 0) + 1

        def release(tid):
            for a in self.terms[tid]["args"]:
                left[a] -= 1
                if left[a] == 0 and a not in keep:
                    V.pop(a, None)
        if threads <= 1:
            for tid in self.order:
                V[tid] = self.exec_term(self.terms[tid], V)
                if left is not None:
                    release(tid)
            return V
        uses = {tid: [] for tid in self.order}
        pending = {}
        for tid in self.order:
            args = self.terms[tid]["args"]
            pending[tid] = len(args)
            for a in args:
                uses[a].append(tid)
        lock = threading.Lock()
        done = threading.Event()
        remaining = [len(self.order)]
        pool = ThreadPoolExecutor(max_workers=threads)

        errors = []

        def work(tid):
            try:
                V[tid] = self.exec_term(self.terms[tid], V)
            except BaseException as e:  # surface worker failures instead of hanging
                errors.append(e)
                done.set()
                return
            ready = []
            with lock:
                if left is not None:
                    release(tid)
                for u in uses[tid]:
                    pending[u] -= 1
                    if pending[u] == 0:
                        ready.append(u)
                remaining[0] -= 1
                if remaining[0] == 0:
                    done.set()
            for u in ready:
                pool.submit(work, u)
        initial = [tid for tid in self.order if pending[tid] == 0]   # snapshot before any worker runs
        for tid in initial:
            pool.submit(work, tid)
        done.wait()
        pool.shutdown(wait=True)
        if errors:
            raise errors[0]
        return V


def run_many(op, inputs_list, threads):
    """Runs several instances of one compiled program concurrently on ONE thread pool
    (instances x DAG parallelism), the CPU analogue of a batched GPU step."""
    uses = {tid: [] for tid in op.order}
    base_pending = {}
    for tid in op.order:
        args = op.terms[tid]["args"]
        base_pending[tid] = len(args)
        for a in args:
            uses[a].append(tid)
    lock = threading.Lock()
    done = threading.Event()
    errors = []
    Vs = []
    pend = []
    for inputs in inputs_list:
        V = {}
        for tid, name in op.in_names.items():
            V[tid] = inputs[name]
        Vs.append(V)
        pend.append(dict(base_pending))
    remaining = [len(op.order) * len(inputs_list)]
    pool = ThreadPoolExecutor(max_workers=max(1, threads))

    def work(i, tid):
        try:
            Vs[i][tid] = op.exec_term(op.terms[tid], Vs[i])
        except BaseException as e:
            errors.append(e)
            done.set()
            return
        ready = []
        with lock:
            for u in uses[tid]:
                pend[i][u] -= 1
                if pend[i][u] == 0:
                    ready.append(u)
            remaining[0] -= 1
            if remaining[0] == 0:
                done.set()
        for u in ready:
            pool.submit(work, i, u)
    initial = [(i, tid) for i in range(len(inputs_list)) for tid in op.order if base_pending[tid] == 0]
    for i, tid in initial:
        pool.submit(work, i, tid)
    done.wait()
    pool.shutdown(wait=True)
    if errors:
        raise errors[0]
    return Vs
