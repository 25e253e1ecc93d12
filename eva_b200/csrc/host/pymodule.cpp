// pymodule.cpp -- pybind11 module `eva_b200._eva_b200`: the Python-facing
// surface of the backend, name-compatible with the reference's `_eva` module
// (python/eva/wrapper.cpp:26-246) for Program/Term/Op/Type/evaluate/
// CKKSParameters/CKKSSignature/CKKSEncodingInfo/generate_keys/*Valuation/
// *Public.encrypt/execute/*Secret.decrypt, plus a few test/benchmark hooks
// (raw key/ciphertext injection, device-resident execution).
#include "backend.hpp"
#include "compiler.hpp"
#include "reference_eval.hpp"
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

namespace py = pybind11;
using namespace evab;
typedef py::array_t<std::uint64_t, py::array::c_style | py::array::forcecast> u64arr;

static std::vector<u64> toVec(const u64arr &a) { return std::vector<u64>(a.data(), a.data() + a.size()); }
static HostBuf toHostBuf(const u64arr &a) { HostBuf b(a.size()); std::memcpy(b.data(), a.data(), a.size() * 8); return b; }

// key material loaded from files / callers: shapes are checked against (k, N) before anything reaches the device
static void requireShape(const u64arr &a, std::initializer_list<std::size_t> shape, const char *what) {
  bool ok = (std::size_t)a.ndim() == shape.size();
  std::size_t i = 0;
  for (std::size_t d : shape) { if (ok && (std::size_t)a.shape(i) != d) ok = false; i++; }
  if (!ok) throw std::runtime_error(std::string(what) + " does not match the parameters (modulus count / degree)");
}

// execute_batch_async handle: the batch plus the Python objects it reads from.  Destroyed with the GIL held (the batch first:
// it waits for its copies before the valuations may go)
struct PyPending {
  std::shared_ptr<evab::PendingBatch> batch;
  py::object program, inputs;
  ~PyPending() { batch.reset(); }
};

PYBIND11_MODULE(_eva_b200, m) {
  m.doc() = "B200-native EVA backend";

  py::enum_<Op>(m, "Op")
      .value("Undef", Op::Undef).value("Input", Op::Input).value("Output", Op::Output).value("Constant", Op::Constant)
      .value("Negate", Op::Negate).value("Add", Op::Add).value("Sub", Op::Sub).value("Mul", Op::Mul)
      .value("RotateLeftConst", Op::RotateLeftConst).value("RotateRightConst", Op::RotateRightConst)
      .value("Relinearize", Op::Relinearize).value("ModSwitch", Op::ModSwitch).value("Rescale", Op::Rescale).value("Encode", Op::Encode);
  py::enum_<Type>(m, "Type").value("Undef", Type::Undef).value("Cipher", Type::Cipher).value("Raw", Type::Raw).value("Plain", Type::Plain);

  py::class_<Term, std::shared_ptr<Term>>(m, "Term", "EVA's native Term class")
      .def_readonly("op", &Term::op, "The operation performed by this term")
      .def_readonly("index", &Term::index)
      .def_property_readonly("operands", &Term::getOperands)
      .def_property_readonly("attributes", [](const Term &t) {
        py::dict d;
        if (t.rescaleDivisor) d["RescaleDivisorAttribute"] = *t.rescaleDivisor;
        if (t.rotation) d["RotationAttribute"] = *t.rotation;
        if (t.type) d["TypeAttribute"] = *t.type;
        if (t.range) d["RangeAttribute"] = *t.range;
        if (t.encodeAtScale) d["EncodeAtScaleAttribute"] = *t.encodeAtScale;
        if (t.encodeAtLevel) d["EncodeAtLevelAttribute"] = *t.encodeAtLevel;
        if (t.constant) d["ConstantValueAttribute"] = t.constant->values();
        return d;
      })
      .def("_set_attributes", [](Term &t, const py::dict &d) {  // fixture loading / tests
        for (auto item : d) {
          const std::string k = py::cast<std::string>(item.first);
          if (k == "RescaleDivisorAttribute") t.rescaleDivisor = py::cast<std::uint32_t>(item.second);
          else if (k == "RotationAttribute") t.rotation = py::cast<std::int32_t>(item.second);
          else if (k == "TypeAttribute") t.type = py::cast<Type>(item.second);
          else if (k == "RangeAttribute") t.range = py::cast<std::uint32_t>(item.second);
          else if (k == "EncodeAtScaleAttribute") t.encodeAtScale = py::cast<std::uint32_t>(item.second);
          else if (k == "EncodeAtLevelAttribute") t.encodeAtLevel = py::cast<std::uint32_t>(item.second);
          else throw std::runtime_error("unknown attribute " + k);
        }
      });

  py::class_<Program>(m, "Program", "EVA's native Program class")
      .def(py::init<std::string, std::uint64_t>(), py::arg("name"), py::arg("vec_size"))
      .def_property("name", &Program::getName, &Program::setName, "The name of this program")
      .def_property_readonly("vec_size", &Program::getVecSize, "The number of elements for all vectors in this program")
      .def_property_readonly("inputs", &Program::getInputs, "A dictionary from input names to terms")
      .def_property_readonly("outputs", &Program::getOutputs, "A dictionary from output names to terms")
      .def("set_output_ranges", [](const Program &p, std::uint32_t range) { for (auto &e : p.getOutputs()) e.second->range = range; }, py::arg("range"))
      .def("set_input_scales", [](const Program &p, std::uint32_t scale) { for (auto &s : p.getSources()) s->encodeAtScale = scale; }, py::arg("scale"))
      .def("to_DOT", &Program::toDOT)
      .def("terms", &Program::toposort, "All live terms in topological order")
      .def("_make_term", &Program::makeTerm, py::keep_alive<0, 1>())
      .def("_make_left_rotation", &Program::makeLeftRotation, py::keep_alive<0, 1>())
      .def("_make_right_rotation", &Program::makeRightRotation, py::keep_alive<0, 1>())
      .def("_make_dense_constant", &Program::makeDenseConstant, py::keep_alive<0, 1>())
      .def("_make_uniform_constant", &Program::makeUniformConstant, py::keep_alive<0, 1>())
      .def("_make_input", &Program::makeInput, py::keep_alive<0, 1>())
      .def("_make_output", &Program::makeOutput, py::keep_alive<0, 1>());

  m.def("evaluate", &evaluate, py::arg("program"), py::arg("inputs"), "Evaluate the program without homomorphic encryption");
  m.def("set_num_threads", [](int) {}, py::arg("num_threads"),
        "Kept for API compatibility: the DAG is scheduled on CUDA streams, not CPU threads.");
  py::class_<int>(m, "_GaloisGuard").def(py::init());

  py::module mckks = m.def_submodule("_ckks", "CKKS compiler types");
  py::class_<CKKSCompiler>(mckks, "CKKSCompiler")
      .def(py::init(), "Create a compiler with the default config")
      .def(py::init<std::unordered_map<std::string, std::string>>(), py::arg("config"), "Create a compiler with a custom config (dict from strings to strings)")
      .def("compile", &CKKSCompiler::compile, py::arg("program"),
           "Compile a program for CKKS; returns (compiled Program, CKKSParameters, CKKSSignature)");
  py::class_<CKKSParameters>(mckks, "CKKSParameters", "Abstract encryption parameters for CKKS")
      .def(py::init([](std::vector<std::uint32_t> bits, std::set<int> rots, std::uint32_t n) { CKKSParameters p; p.primeBits = bits; p.rotations = rots; p.polyModulusDegree = n; return p; }),
           py::arg("prime_bits"), py::arg("rotations"), py::arg("poly_modulus_degree"))
      .def_readonly("prime_bits", &CKKSParameters::primeBits)
      .def_readonly("rotations", &CKKSParameters::rotations)
      .def_readonly("poly_modulus_degree", &CKKSParameters::polyModulusDegree);
  py::class_<CKKSEncodingInfo>(mckks, "CKKSEncodingInfo")
      .def(py::init<Type, int, int>(), py::arg("input_type"), py::arg("scale"), py::arg("level"))
      .def_readonly("input_type", &CKKSEncodingInfo::inputType)
      .def_readonly("scale", &CKKSEncodingInfo::scale)
      .def_readonly("level", &CKKSEncodingInfo::level);
  py::class_<CKKSSignature>(mckks, "CKKSSignature")
      .def(py::init<int, std::map<std::string, CKKSEncodingInfo>>(), py::arg("vec_size"), py::arg("inputs"))
      .def_readonly("vec_size", &CKKSSignature::vecSize)
      .def_readonly("inputs", &CKKSSignature::inputs);

  py::module mb = m.def_submodule("_b200", "B200 CUDA backend (replaces the reference's _seal submodule)");
  py::class_<B200Valuation>(mb, "B200Valuation", "A valuation for inputs or outputs holding encrypted values")
      .def(py::init<>())
      .def("names", [](const B200Valuation &v) { std::vector<std::string> n; for (auto &e : v) n.push_back(e.first); return n; })
      .def("set_cipher", [](B200Valuation &v, const std::string &name, const u64arr &data, double scale) {
        if (data.ndim() != 3) throw std::runtime_error("ciphertext array must be [size][ell][N]");
        HostCipher h; h.data = toHostBuf(data); h.size = (int)data.shape(0); h.ell = (int)data.shape(1); h.scale = scale;
        v[name] = std::move(h);
      }, "test/benchmark hook: inject a raw ciphertext [size][ell][N]")
      .def("set_plain", [](B200Valuation &v, const std::string &name, const u64arr &data, double scale) {
        if (data.ndim() != 2) throw std::runtime_error("plaintext array must be [ell][N]");
        HostPlain h; h.data = toHostBuf(data); h.ell = (int)data.shape(0); h.scale = scale; v[name] = std::move(h);
      })
      .def("set_raw", [](B200Valuation &v, const std::string &name, const std::vector<double> &x) { v[name] = std::make_shared<ConstantValue>(x.size(), x); })
      .def("get", [](const B200Valuation &v, const std::string &name) -> py::object {
        const SchemeValue &sv = v.at(name);
        if (auto *c = std::get_if<HostCipher>(&sv)) {
          const std::size_t N = c->data.size() / ((std::size_t)c->size * c->ell);
          u64arr a({(std::size_t)c->size, (std::size_t)c->ell, N});
          std::memcpy(a.mutable_data(), c->data.data(), c->data.size() * 8);
          return py::make_tuple("cipher", a, c->scale);
        }
        if (auto *p = std::get_if<HostPlain>(&sv)) {
          const std::size_t N = p->data.size() / (std::size_t)p->ell;
          u64arr a({(std::size_t)p->ell, N});
          std::memcpy(a.mutable_data(), p->data.data(), p->data.size() * 8);
          return py::make_tuple("plain", a, p->scale);
        }
        return py::make_tuple("raw", std::get<std::shared_ptr<ConstantValue>>(sv)->values(), 0.0);
      });

  py::class_<PyPending, std::shared_ptr<PyPending>>(mb, "PendingBatch", "a submitted, not yet collected execute_batch");
  py::class_<B200Public>(mb, "B200Public", "The public part of the context: encryption and execution on the GPU")
      .def("encrypt", &B200Public::encrypt, py::arg("inputs"), py::arg("signature"))
      .def("execute", &B200Public::execute, py::arg("program"), py::arg("inputs"), py::call_guard<py::gil_scoped_release>())
      // the list is converted to pointers to the caller's valuations: no copy of the host ciphertexts
      .def("execute_batch", [](B200Public &p, Program &prog, const std::vector<const B200Valuation *> &in) { return p.executeMany(prog, in); },
           py::arg("program"), py::arg("inputs"), py::call_guard<py::gil_scoped_release>(),
           "Execute one compiled program on a list of valuations with batched kernels; returns a list of valuations")
      // pipelined execute_batch: the handle keeps the program and the caller's valuations alive until it is collected
      .def("execute_batch_async", [](B200Public &p, py::object program, py::list inputs, int slot) {
             Program &prog = program.cast<Program &>();
             std::vector<const B200Valuation *> in;
             for (auto h : inputs) in.push_back(h.cast<const B200Valuation *>());
             auto out = std::make_shared<PyPending>();
             { py::gil_scoped_release rel; out->batch = p.submitMany(prog, in, slot); }
             out->program = program; out->inputs = inputs;
             return out;
           }, py::arg("program"), py::arg("inputs"), py::arg("slot") = 0,
           "Enqueue execute_batch (H2D, plan replays, D2H) and return a handle; collect with execute_batch_result(handle).  Batches in different slots (0..7) overlap")
      .def("execute_batch_result", [](B200Public &p, std::shared_ptr<PyPending> h) {
             std::vector<B200Valuation> outs;
             { py::gil_scoped_release rel; outs = p.collect(*h->batch); }
             return outs;
           }, py::arg("handle"))
      .def("set_options", [](B200Public &p, int streams, bool graph, bool cache, bool dedup, int fuse, bool fuseSums, bool hoist, bool uniformEncode, bool dedupTerms, bool hoistModUp, bool approxHoist, int rotationChunk) {
             p.options.numStreams = streams; p.options.useGraph = graph; p.options.cacheConstants = cache; p.options.dedupConstants = dedup;
             p.options.fuse = fuse; p.options.fuseSums = fuseSums; p.options.hoistRotations = hoist; p.options.uniformEncode = uniformEncode;
             p.options.dedupTerms = dedupTerms; p.options.hoistModUp = hoistModUp; p.options.approxHoist = approxHoist; p.options.rotationChunk = rotationChunk;
           },
           py::arg("num_streams") = 8, py::arg("use_graph") = true, py::arg("cache_constants") = true, py::arg("dedup_constants") = true,
           py::arg("fuse") = 1, py::arg("fuse_sums") = true, py::arg("hoist_rotations") = true, py::arg("uniform_encode") = true, py::arg("dedup_terms") = true, py::arg("hoist_mod_up") = true, py::arg("approx_hoist") = false, py::arg("rotation_chunk") = 16)
      .def("set_input_sizes", [](B200Public &p, const std::map<std::string, int> &sizes) { p.options.inputSizes = sizes; },
           "ciphertext inputs that are not size 2 (name -> polynomials); applies to plans built afterwards")
      .def("drop_plan", &B200Public::dropExecutor, py::arg("program"), py::arg("batch") = 1, py::arg("replica") = 0)
      .def("plan_stats", [](B200Public &p, Program &prog, int batch, int replica) { return p.executorFor(prog, batch, replica).planStats(); },
           py::arg("program"), py::arg("batch") = 1, py::arg("replica") = 0, "shape of the execution plan: steps, hoist groups, rotation chunks, lazy rotation sums")
      .def("cipher_op_count", [](B200Public &p, Program &prog) { return p.executorFor(prog).cipherOpCount(); })
      // serialization (eva_b200/serialization.py): public key material only (never the secret key)
      .def("_export", [](B200Public &x) {
        auto s = x.shared();
        const std::size_t N = s->dev->N(), k = s->dev->k();
        py::dict d;
        d["N"] = N; d["primes"] = s->dev->primes();
        s->dev->sync();
        auto down = [&](const DBuf &b, std::vector<std::size_t> shape) {
          u64arr a(shape);
          s->dev->download(a.mutable_data(), b.get(), a.size() * 8);
          s->dev->sync();
          return a;
        };
        if (s->keys.pk) d["public_key"] = down(s->keys.pk, {2, k, N});
        if (s->keys.relin) d["relin_key"] = down(s->keys.relin, {k - 1, 2, k, N});
        py::dict g;
        for (auto &kv : s->keys.galois) g[py::int_(kv.first)] = down(kv.second, {k - 1, 2, k, N});
        d["galois_keys"] = g;
        return d;
      })
      .def("launch_count", [](B200Public &p) { return evab_launch_count(p.shared()->dev->ctx()); })
      .def("primes", [](B200Public &p) { return p.shared()->dev->primes(); })
      // ---- benchmark hooks: device-resident execution on a caller-provided stream
      .def("stage_inputs", [](B200Public &p, Program &prog, const std::vector<const B200Valuation *> &in, std::uintptr_t stream, int replica) {
        Executor &ex = p.executorFor(prog, (int)in.size(), replica);
        for (std::size_t b = 0; b < in.size(); b++) p.stageInputs(ex, prog, *in[b], (void *)stream, (int)b);
      }, py::arg("program"), py::arg("inputs"), py::arg("stream"), py::arg("replica") = 0)
      .def("run_resident", [](B200Public &p, Program &prog, std::uintptr_t stream, int batch, int replica) { p.executorFor(prog, batch, replica).run((void *)stream); },
           py::arg("program"), py::arg("stream"), py::arg("batch") = 1, py::arg("replica") = 0, py::call_guard<py::gil_scoped_release>())
      .def("sync", [](B200Public &p, std::uintptr_t stream) { p.shared()->dev->sync((void *)stream); }, py::call_guard<py::gil_scoped_release>())
      // ---- device-resident plumbing of the DAG-sharded mode (eva_b200/shard.py): where a plan keeps its named
      // inputs and outputs in device memory, so that NCCL can gather partial ciphertexts straight from one
      // rank's arena into another's (no host staging), and a download of the outputs of a resident run
      .def("io_pointers", [](B200Public &p, Program &prog, int batch, int replica) {
        Executor &ex = p.executorFor(prog, batch, replica);
        const std::size_t N = p.shared()->dev->N();
        auto describe = [&](const Term::Ptr &t) -> py::object {
          const ValueInfo &vi = ex.info(t);
          if (vi.kind != Kind::Cipher && vi.kind != Kind::Plain) return py::none();
          const std::size_t polys = vi.kind == Kind::Cipher ? (std::size_t)vi.size : 1;
          py::dict d;
          d["ptr"] = (std::uintptr_t)ex.valuePtr(t, 0); d["bytes"] = polys * (std::size_t)vi.ell * N * 8;
          d["size"] = (int)polys; d["ell"] = vi.ell; d["scale"] = vi.scale; d["cipher"] = vi.kind == Kind::Cipher;
          return d;
        };
        py::dict ins, outs;
        for (auto &in : prog.getInputs()) ins[py::str(in.first)] = describe(in.second);
        for (auto &o : prog.getOutputs()) outs[py::str(o.first)] = describe(o.second);
        py::dict r; r["inputs"] = ins; r["outputs"] = outs;
        return r;
      }, py::arg("program"), py::arg("batch") = 1, py::arg("replica") = 0)
      .def("download_outputs", [](B200Public &p, Program &prog, std::uintptr_t stream, int replica) {
        Executor &ex = p.executorFor(prog, 1, replica);
        auto dev = p.shared()->dev;
        const std::size_t N = dev->N();
        B200Valuation out;
        for (auto &o : prog.getOutputs()) {
          const ValueInfo &vi = ex.info(o.second);
          if (vi.kind == Kind::Cipher) {
            HostCipher h; h.size = vi.size; h.ell = vi.ell; h.scale = vi.scale; h.data.resize((std::size_t)vi.size * vi.ell * N);
            dev->download(h.data.data(), ex.valuePtr(o.second, 0), h.data.size() * 8, (void *)stream);
            out[o.first] = std::move(h);
          } else if (vi.kind == Kind::Plain) {
            HostPlain h; h.ell = vi.ell; h.scale = vi.scale; h.data.resize((std::size_t)vi.ell * N);
            dev->download(h.data.data(), ex.valuePtr(o.second, 0), h.data.size() * 8, (void *)stream);
            out[o.first] = std::move(h);
          } else {
            out[o.first] = std::make_shared<ConstantValue>(prog.getVecSize(), ex.rawValue(o.second->index, 0));
          }
        }
        dev->sync((void *)stream);
        if (ex.flagsRaised())
          throw std::runtime_error("a digit of a rotated ciphertext held a zero coefficient: rerun with set_options(hoist_mod_up=False) (execute() does so by itself)");
        return out;
      }, py::arg("program"), py::arg("stream"), py::arg("replica") = 0, py::call_guard<py::gil_scoped_release>())
      // ---- test hooks
      .def("debug_value", [](B200Public &p, Program &prog, std::uint64_t index, int batch, int b) -> py::object {
        Executor &ex = p.executorFor(prog, batch);
        const ValueInfo &vi = ex.info(index);
        auto dev = p.shared()->dev;
        if (vi.kind == Kind::Raw) return py::cast(ex.rawValue(index, b));
        if (vi.kind == Kind::None || vi.fused) return py::none();   // fused away: never materialised
        const std::size_t polys = vi.kind == Kind::Cipher ? vi.size : 1;
        u64arr a({polys, (std::size_t)vi.ell, (std::size_t)dev->N()});
        dev->sync();
        dev->download(a.mutable_data(), ex.valuePtr(index, b), a.size() * 8);
        dev->sync();
        return py::make_tuple(a, vi.scale);
      }, py::arg("program"), py::arg("index"), py::arg("batch") = 1, py::arg("instance") = 0)
      .def("encode", [](B200Public &p, const std::vector<double> &values, double scale, int ell, bool host) {
        auto dev = p.shared()->dev;
        DBuf pt(dev, (std::size_t)ell * dev->N());
        std::vector<double> rep;
        const std::size_t slots = dev->N() / 2;
        for (std::size_t r = slots / values.size(); r > 0; --r) rep.insert(rep.end(), values.begin(), values.end());
        if (host) p.shared()->client->encoder().encodeHost(rep, scale, ell, pt.get());
        else p.shared()->client->encoder().encode(values, scale, ell, pt.get());
        u64arr a({(std::size_t)ell, (std::size_t)dev->N()});
        dev->download(a.mutable_data(), pt.get(), a.size() * 8);
        dev->sync();
        return a;
      }, py::arg("values"), py::arg("scale"), py::arg("ell"), py::arg("host") = false)
      // test hook: Encryptor::encrypt of a plaintext polynomial [ell][N] with EXPLICIT randomness -> ciphertext [2][ell][N]
      .def("encrypt_poly", [](B200Public &p, const u64arr &pt, const std::vector<int> &u, const std::vector<int> &e0, const std::vector<int> &e1) {
        auto s = p.shared();
        if (!s->keys.pk) throw std::runtime_error("this context has no public key");
        if (pt.ndim() != 2 || (std::size_t)pt.shape(1) != s->dev->N()) throw std::runtime_error("plaintext array must be [ell][N]");
        const int ell = (int)pt.shape(0);
        DBuf d(s->dev, pt.size());
        s->dev->upload(d.get(), pt.data(), pt.size() * 8);
        DBuf ct = s->client->encryptWith(s->keys, d.get(), ell, u, e0, e1);
        u64arr a({(std::size_t)2, (std::size_t)ell, (std::size_t)s->dev->N()});
        s->dev->download(a.mutable_data(), ct.get(), a.size() * 8);
        s->dev->sync();
        return a;
      }, py::arg("plaintext"), py::arg("u"), py::arg("e0"), py::arg("e1"))
      .def("decode", [](B200Public &p, const u64arr &pt, double scale) {
        auto dev = p.shared()->dev;
        DBuf d(dev, pt.size());
        dev->upload(d.get(), pt.data(), pt.size() * 8);
        return p.shared()->client->encoder().decode(d.get(), (int)pt.shape(0), scale);
      });
  py::class_<B200Secret>(mb, "B200Secret", "The secret part of the context: decryption")
      .def("decrypt", &B200Secret::decrypt, py::arg("enc_outputs"), py::arg("signature"))
      // test hook: Decryptor::decrypt of a raw ciphertext [size][ell][N] -> plaintext polynomial [ell][N] (before decoding)
      .def("decrypt_poly", [](B200Secret &x, const u64arr &ct) {
        auto s = x.shared();
        if (ct.ndim() != 3 || (std::size_t)ct.shape(2) != s->dev->N()) throw std::runtime_error("ciphertext array must be [size][ell][N]");
        DBuf d(s->dev, ct.size());
        s->dev->upload(d.get(), ct.data(), ct.size() * 8);
        DBuf pt = s->client->decrypt(s->keys, d.get(), (int)ct.shape(0), (int)ct.shape(1));
        u64arr a({(std::size_t)ct.shape(1), (std::size_t)s->dev->N()});
        s->dev->download(a.mutable_data(), pt.get(), a.size() * 8);
        s->dev->sync();
        return a;
      }, py::arg("ciphertext"))
      .def("decode", [](B200Secret &x, const u64arr &pt, double scale) {
        auto s = x.shared();
        DBuf d(s->dev, pt.size());
        s->dev->upload(d.get(), pt.data(), pt.size() * 8);
        return s->client->encoder().decode(d.get(), (int)pt.shape(0), scale);
      }, py::arg("plaintext"), py::arg("scale"))
      // serialization (eva_b200/serialization.py): secret key image [k][N] + the modulus chain
      .def("_export", [](B200Secret &x) {
        auto s = x.shared();
        py::dict d;
        d["N"] = s->dev->N(); d["primes"] = s->dev->primes();
        u64arr a({(std::size_t)s->dev->k(), (std::size_t)s->dev->N()});
        s->dev->sync();
        s->dev->download(a.mutable_data(), s->keys.sk.get(), a.size() * 8);
        s->dev->sync();
        d["secret_key"] = a;
        return d;
      });

  mb.def("generate_keys", [](const CKKSParameters &p, int device, std::uint64_t seed) { return generateKeys(p, device, seed); },
         py::arg("abstract_params"), py::arg("device") = 0, py::arg("seed") = 0);
  // test/benchmark hook: evaluation context from externally supplied key material
  mb.def("context_from_raw_keys", [](std::uint64_t N, const std::vector<u64> &primes, const u64arr &relin, const std::map<u64, u64arr> &galois, int device) {
    auto s = std::make_shared<Shared>();
    s->dev = std::make_shared<Device>(N, primes, device);
    s->client = std::make_unique<CkksClient>(s->dev, 0);   // evaluation context; any encryption through it draws OS entropy
    const std::size_t kk = primes.size();
    requireShape(relin, {kk - 1, 2, kk, (std::size_t)N}, "relinearization key");
    for (auto &g : galois) requireShape(g.second, {kk - 1, 2, kk, (std::size_t)N}, "Galois key");
    s->keys.relin = DBuf(s->dev, relin.size());
    s->dev->upload(s->keys.relin.get(), relin.data(), relin.size() * 8);
    for (auto &g : galois) {
      check(evab_galois_prepare(s->dev->ctx(), g.first));
      DBuf d(s->dev, g.second.size());
      s->dev->upload(d.get(), g.second.data(), g.second.size() * 8);
      s->keys.galois.emplace(g.first, std::move(d));
    }
    s->dev->sync();
    return std::make_unique<B200Public>(s);
  }, py::arg("N"), py::arg("primes"), py::arg("relin_key"), py::arg("galois_keys"), py::arg("device") = 0);
  // serialization: contexts rebuilt from saved key material
  mb.def("public_from_raw", [](std::uint64_t N, const std::vector<u64> &primes, py::object pk, py::object relin, const std::map<u64, u64arr> &galois, int device) {
    auto s = std::make_shared<Shared>();
    s->dev = std::make_shared<Device>(N, primes, device);
    s->client = std::make_unique<CkksClient>(s->dev, 0);   // OS entropy (csprng.hpp)
    auto up = [&](DBuf &dst, const u64arr &a) { dst = DBuf(s->dev, a.size()); s->dev->upload(dst.get(), a.data(), a.size() * 8); s->dev->sync(); };
    const std::size_t kk = primes.size();
    if (!pk.is_none()) { requireShape(py::cast<u64arr>(pk), {2, kk, (std::size_t)N}, "public key"); up(s->keys.pk, py::cast<u64arr>(pk)); }
    if (!relin.is_none()) { requireShape(py::cast<u64arr>(relin), {kk - 1, 2, kk, (std::size_t)N}, "relinearization key"); up(s->keys.relin, py::cast<u64arr>(relin)); }
    for (auto &g : galois) requireShape(g.second, {kk - 1, 2, kk, (std::size_t)N}, "Galois key");
    for (auto &g : galois) {
      check(evab_galois_prepare(s->dev->ctx(), g.first));
      DBuf d;
      up(d, g.second);
      s->keys.galois.emplace(g.first, std::move(d));
    }
    return std::make_unique<B200Public>(s);
  }, py::arg("N"), py::arg("primes"), py::arg("public_key"), py::arg("relin_key"), py::arg("galois_keys"), py::arg("device") = 0);
  mb.def("secret_from_raw", [](std::uint64_t N, const std::vector<u64> &primes, const u64arr &sk, int device) {
    auto s = std::make_shared<Shared>();
    s->dev = std::make_shared<Device>(N, primes, device);
    s->client = std::make_unique<CkksClient>(s->dev, 0);
    requireShape(sk, {primes.size(), (std::size_t)N}, "secret key");
    s->keys.sk = DBuf(s->dev, sk.size());
    s->dev->upload(s->keys.sk.get(), sk.data(), sk.size() * 8);
    s->dev->sync();
    return std::make_unique<B200Secret>(s);
  }, py::arg("N"), py::arg("primes"), py::arg("secret_key"), py::arg("device") = 0);
  mb.def("create_coeff_modulus", [](std::uint64_t N, const std::vector<int> &bits) { return hmod::createCoeffModulus(N, bits); });
}
