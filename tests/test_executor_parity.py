"""Executor parity on whole programs: the product executor (C++ plan + CUDA
streams/graph, C-ABI kernels) versus the oracle runner, on programs compiled by
the REFERENCE compiler (tests/golden/programs).  Bit-exact on every intermediate
ciphertext; decrypted outputs also checked against the reference's own
acceptance criterion (tests/common.py:34, MSE < 0.01) using reference_outputs."""
import numpy as np
import pytest

from eva_b200 import program_io as gl
from oracle import oracle as o
from oracle_exec import OracleProgram

pytestmark = pytest.mark.gpu


def run_both(name, seed=1, use_graph=True, streams=8, lo=-1.0, hi=1.0, cache=True, fuse_sums=False):
    """fuse_sums=False materialises every term (all intermediates are compared); with fuse_sums=True the
    inner nodes of fused multiply_plain / add trees do not exist and are skipped (debug_value -> None)"""
    from eva_b200 import b200
    d = gl.load_json(name)
    prog, params, sig, terms = gl.build_program(d)
    N = d["poly_modulus_degree"]
    orc = o.Oracle(N, d["prime_bits"]).keygen(seed)
    op = OracleProgram(d, orc)
    op.prepare_keys()
    assert b200.create_coeff_modulus(N, d["prime_bits"]) == orc.primes
    pub = b200.context_from_raw_keys(N, orc.primes, op.rk, {int(e): k for e, k in op.gks.items()})
    pub.set_options(num_streams=streams, use_graph=use_graph, cache_constants=cache, fuse_sums=fuse_sums)
    rng = np.random.default_rng(seed)
    inputs_o, val, plain_inputs = {}, b200.B200Valuation(), {}
    for name_, info in d["signature"].items():
        x = rng.uniform(lo, hi, d["vec_size"])
        plain_inputs[name_] = x
        ell = orc.k - 1 - info["level"]
        if info["type"] == "Cipher":
            ct = orc.encrypt(orc.encode(x, 2.0 ** info["scale"], ell), seed=11)
            inputs_o[name_] = ("cipher", ct, 2.0 ** info["scale"])
            val.set_cipher(name_, ct, 2.0 ** info["scale"])
        elif info["type"] == "Raw":
            inputs_o[name_] = ("raw", x)
            val.set_raw(name_, list(x))
        else:
            pt = orc.encode(x, 2.0 ** info["scale"], ell)
            inputs_o[name_] = ("plain", pt, 2.0 ** info["scale"])
            val.set_plain(name_, pt, 2.0 ** info["scale"])
    V = op.run(inputs_o)
    out = pub.execute(prog, val)
    out2 = pub.execute(prog, val)   # second run replays the captured graph
    n_checked = n_fused = 0
    for t in d["terms"]:
        want = V[t["id"]]
        got = pub.debug_value(prog, terms[t["id"]].index)
        if got is None and fuse_sums and want[0] == "cipher":
            n_fused += 1
            continue
        if want[0] == "raw":
            assert np.allclose(np.asarray(got), want[1])
            continue
        arr, scale = got
        w = want[1] if want[0] == "cipher" else want[1][None]
        assert arr.shape == w.shape, (t, arr.shape, w.shape)
        assert scale == want[2], (t, scale, want[2])
        assert np.array_equal(arr, w), "term %d (%s) differs" % (t["id"], t["op"])
        n_checked += 1
    for oname, oid in d["outputs"].items():
        for res in (out, out2):
            kind, arr, scale = res.get(oname)
            want = V[oid]
            if kind == "cipher":
                assert np.array_equal(arr, want[1]) and scale == want[2]
    assert pub.cipher_op_count(prog) == op.cipher_op_count()
    if fuse_sums:
        return d, orc, V, plain_inputs, n_checked, n_fused
    return d, orc, V, plain_inputs, n_checked


@pytest.mark.parametrize("name", ["polynomial", "feat_bin_mul_11", "feat_bin_sub_01", "feat_unary", "feat_rot_m1", "feat_rot_0",
                                  "feat_mixed", "feat_transparent", "feat_hsum", "feat_deep"])
def test_small_programs_bit_exact(name):
    run_both(name)


def test_sobel_bit_exact_and_accurate():
    d, orc, V, x, n = run_both("sobel", lo=0.0, hi=0.2)   # smooth-image range: stays inside set_output_ranges(10)
    assert n >= 61
    assert sum(1 for t in d["terms"] if t["op"] in ("Add", "Sub", "Mul", "Negate", "RotateLeftConst", "RotateRightConst",
                                                    "Relinearize", "ModSwitch", "Rescale")) == 61
    oid = d["outputs"]["image"]
    dec = orc.decode(orc.decrypt(V[oid][1]), V[oid][2])[:d["vec_size"]]
    # plaintext semantics of the Sobel program computed directly (reference examples/image_processing.py:39-63)
    img = x["image"]
    F = [[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]]
    ix = sum(np.roll(img, -(i * 64 + j)) * F[i][j] for i in range(3) for j in range(3))
    iy = sum(np.roll(img, -(i * 64 + j)) * F[j][i] for i in range(3) for j in range(3))
    dsq = ix ** 2 + iy ** 2
    ref = dsq * 2.2137874823876622 + dsq ** 2 * -1.0984324107372518 + dsq ** 3 * 0.17254603006834726
    assert np.mean((dec - ref) ** 2) < 0.01


def test_harris_bit_exact():
    d, orc, V, x, n = run_both("harris")
    assert n >= 143


@pytest.mark.parametrize("graph,streams", [(False, 1), (False, 8), (True, 4)])
def test_scheduler_modes_agree(graph, streams):
    run_both("sobel", use_graph=graph, streams=streams)


@pytest.mark.parametrize("name", ["sobel", "feat_mixed", "polynomial"])
def test_constants_encoded_every_run(name):
    """cache_constants=False: Encode terms run on the GPU inside every execute (as the reference does)"""
    run_both(name, cache=False)


def test_wide_dag():
    run_both("wide64")


def test_without_hoisted_rotations():
    """hoist_rotations=False: every rotation runs the plain evab_rotate path (default shares the inverse NTT);
    uniform_encode=False: scalar constants go through the full FFT encoder (default: one-pass encoder)"""
    from eva_b200 import b200
    orig = b200.B200Public.set_options
    try:
        b200.B200Public.set_options = lambda self, **kw: orig(self, **{**kw, "hoist_rotations": False, "uniform_encode": False})
        run_both("sobel", lo=0.0, hi=0.2, cache=False)
    finally:
        b200.B200Public.set_options = orig


@pytest.mark.parametrize("chunk", [1, 3])
def test_rotation_chunk_sizes(chunk):
    """rotation_chunk=1: one evab_rotate_modup_prepared call per rotation; 3: the 8 rotations of Sobel's image in chunks of 3 + 3 + 2
    (default 16: all in one evab_rotate_modup_many call) -- every intermediate identical to the oracle either way"""
    from eva_b200 import b200
    orig = b200.B200Public.set_options
    try:
        b200.B200Public.set_options = lambda self, **kw: orig(self, **{**kw, "rotation_chunk": chunk})
        run_both("sobel", lo=0.0, hi=0.2)
        run_both("harris")
    finally:
        b200.B200Public.set_options = orig


@pytest.mark.parametrize("name,min_fused", [("sobel", 30), ("harris", 40), ("polynomial", 0), ("feat_hsum", 0), ("feat_mixed", 0), ("wide64", 1)])
def test_fused_sums_bit_exact(name, min_fused):
    """default executor mode: trees of multiply_plain / add evaluated by one kernel (evab_sum_terms);
    every materialised term and every output still matches the oracle bit for bit"""
    r = run_both(name, lo=0.0, hi=0.2, fuse_sums=True)
    assert r[5] >= min_fused, r[5]


def test_batched_execute_matches_single():
    """execute_batch (one plan, kernels batched over instances) == separate execute() calls, bit for bit,
    with raw inputs differing per instance (feat_mixed) and with rotations / key switching (sobel)."""
    from eva_b200 import b200
    for name in ("feat_mixed", "sobel"):
        d = gl.load_json(name)
        prog, params, sig, terms = gl.build_program(d)
        N = d["poly_modulus_degree"]
        orc = o.Oracle(N, d["prime_bits"]).keygen(3)
        op = OracleProgram(d, orc)
        op.prepare_keys()
        pub = b200.context_from_raw_keys(N, orc.primes, op.rk, {int(e): k for e, k in op.gks.items()})
        pub.set_options(num_streams=4, use_graph=True, cache_constants=False, fuse=3, fuse_sums=False)   # one plan, 3 instances per launch; every term materialised
        rng = np.random.default_rng(5)
        vals, inputs_o = [], []
        for b in range(3):
            val, io_ = b200.B200Valuation(), {}
            for name_, info in d["signature"].items():
                x = rng.uniform(0, 0.2, d["vec_size"])
                ell = orc.k - 1 - info["level"]
                if info["type"] == "Cipher":
                    ct = orc.encrypt(orc.encode(x, 2.0 ** info["scale"], ell), seed=100 + b)
                    val.set_cipher(name_, ct, 2.0 ** info["scale"]); io_[name_] = ("cipher", ct, 2.0 ** info["scale"])
                else:
                    val.set_raw(name_, list(x)); io_[name_] = ("raw", x)
            vals.append(val); inputs_o.append(io_)
        outs = pub.execute_batch(prog, vals)
        outs2 = pub.execute_batch(prog, vals)
        for b in range(3):
            V = op.run(inputs_o[b])
            for oname, oid in d["outputs"].items():
                for res in (outs[b], outs2[b], pub.execute(prog, vals[b])):
                    kind, arr, scale = res.get(oname)
                    assert np.array_equal(arr, V[oid][1]) and scale == V[oid][2]
            for t in d["terms"]:   # every intermediate of every instance
                want = V[t["id"]]
                if want[0] == "raw":
                    continue
                arr, scale = pub.debug_value(prog, terms[t["id"]].index, 3, b)
                w = want[1] if want[0] == "cipher" else want[1][None]
                assert np.array_equal(arr, w), (name, b, t)
        # concurrent plan replicas: chunks of 2 + 1 instances (fuse=2) and 3 x 1 (fuse=1)
        for fuse in (2, 1):
            pub.set_options(num_streams=4, use_graph=True, cache_constants=False, fuse=fuse)
            outs3 = pub.execute_batch(prog, vals)
            for b in range(3):
                for oname in d["outputs"]:
                    assert np.array_equal(outs3[b].get(oname)[1], outs[b].get(oname)[1])


def test_replicas_bounded_by_memory(monkeypatch):
    """execute_batch with fewer plan replicas than chunks (what happens when the arenas do not all fit in
    device memory): replicas are reused round-robin after their previous chunk completed; same results"""
    from eva_b200 import b200
    d = gl.load_json("polynomial")
    prog, params, sig, terms = gl.build_program(d)
    N = d["poly_modulus_degree"]
    orc = o.Oracle(N, d["prime_bits"]).keygen(3)
    op = OracleProgram(d, orc)
    op.prepare_keys()
    pub = b200.context_from_raw_keys(N, orc.primes, op.rk, {int(e): k for e, k in op.gks.items()})
    rng = np.random.default_rng(9)
    vals, want = [], []
    for b in range(7):
        val, io_ = b200.B200Valuation(), {}
        for name_, info in d["signature"].items():
            x = rng.uniform(-1, 1, d["vec_size"])
            ct = orc.encrypt(orc.encode(x, 2.0 ** info["scale"], orc.k - 1 - info["level"]), seed=200 + b)
            val.set_cipher(name_, ct, 2.0 ** info["scale"]); io_[name_] = ("cipher", ct, 2.0 ** info["scale"])
        vals.append(val); want.append(op.run(io_))
    for cap, fuse in (("2", 1), ("3", 2), ("1", 1)):
        monkeypatch.setenv("EVAB_MAX_REPLICAS", cap)
        pub.set_options(fuse=fuse)
        outs = pub.execute_batch(prog, vals)
        for b in range(7):
            for oname, oid in d["outputs"].items():
                assert np.array_equal(outs[b].get(oname)[1], want[b][oid][1]), (cap, fuse, b)


def test_zero_digit_falls_back_to_the_exact_path():
    """The shared mod-up of a rotation group assumes no digit coefficient is zero (negate(0) = 0 carries no q_J).  An input
    whose c1 has a zero coefficient raises the flag; execute() redoes the call on plans without the shared mod-up and still
    returns the oracle's bits."""
    from eva_b200 import b200
    d = gl.load_json("sobel")
    prog, params, sig, terms = gl.build_program(d)
    N = d["poly_modulus_degree"]
    orc = o.Oracle(N, d["prime_bits"]).keygen(1)
    op = OracleProgram(d, orc)
    op.prepare_keys()
    pub = b200.context_from_raw_keys(N, orc.primes, op.rk, {int(e): k for e, k in op.gks.items()})
    rng = np.random.default_rng(4)
    (name_, info), = d["signature"].items()
    ell = orc.k - 1 - info["level"]
    ct = orc.encrypt(orc.encode(rng.uniform(0, 1, d["vec_size"]), 2.0 ** info["scale"], ell), seed=3)
    c1 = np.stack([orc.ntt_inv(ct[1, i], i) for i in range(ell)])
    c1[1, 7] = 0                                     # one zero coefficient in digit 1
    ct[1] = np.stack([orc.ntt_fwd(c1[i], i) for i in range(ell)])
    val = b200.B200Valuation()
    val.set_cipher(name_, ct, 2.0 ** info["scale"])
    V = op.run({name_: ("cipher", ct, 2.0 ** info["scale"])})
    for _ in range(2):                               # first call falls back, second runs on the rebuilt plan
        out = pub.execute(prog, val)
        for oname, oid in d["outputs"].items():
            assert np.array_equal(out.get(oname)[1], V[oid][1])
    # and an ordinary input still matches on the same context
    ct2 = orc.encrypt(orc.encode(rng.uniform(0, 1, d["vec_size"]), 2.0 ** info["scale"], ell), seed=5)
    val2 = b200.B200Valuation()
    val2.set_cipher(name_, ct2, 2.0 ** info["scale"])
    V2 = op.run({name_: ("cipher", ct2, 2.0 ** info["scale"])})
    out2 = pub.execute(prog, val2)
    for oname, oid in d["outputs"].items():
        assert np.array_equal(out2.get(oname)[1], V2[oid][1])
