// backend.hpp -- the B200 counterpart of the reference's SEAL backend API
// (eva/seal/seal.h:24-97, eva/seal/seal.cpp): generateKeys, B200Public::{encrypt,
// execute}, B200Secret::decrypt, B200Valuation.  Same argument meaning, ownership
// and error behaviour (C++ exceptions -> Python exceptions through pybind11).
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "executor.hpp"
#include <atomic>
#include <mutex>
#include <tuple>
#include <variant>

namespace evab {

// compiler outputs (reference eva/ckks/ckks_parameters.h:14-18, ckks_signature.h:16-36)
struct CKKSParameters {
  std::vector<std::uint32_t> primeBits;
  std::set<int> rotations;
  std::uint32_t polyModulusDegree = 0;
};
struct CKKSEncodingInfo {
  Type inputType; int scale; int level;
  CKKSEncodingInfo(Type t, int s, int l) : inputType(t), scale(s), level(l) {}
};
struct CKKSSignature {
  int vecSize = 0;
  std::map<std::string, CKKSEncodingInfo> inputs;
  CKKSSignature() {}
  CKKSSignature(int v, std::map<std::string, CKKSEncodingInfo> in) : vecSize(v), inputs(std::move(in)) {}
};
using Valuation = std::map<std::string, std::vector<double>>;

// host-resident encrypted values (the reference keeps seal::Ciphertext objects
// in host memory inside SEALValuation, eva/seal/seal.h:21-41)
// host images of ciphertexts / plaintexts live in page-locked memory (asynchronous, full-rate H2D / D2H)
struct HostCipher { HostBuf data; int size = 0, ell = 0; double scale = 0; };
struct HostPlain { HostBuf data; int ell = 0; double scale = 0; };
using SchemeValue = std::variant<HostCipher, HostPlain, std::shared_ptr<ConstantValue>>;

class B200Valuation {
public:
  SchemeValue &operator[](const std::string &name) { return values[name]; }
  const SchemeValue &at(const std::string &name) const { return values.at(name); }
  auto begin() const { return values.begin(); }
  auto end() const { return values.end(); }
  std::size_t size() const { return values.size(); }
  std::map<std::string, SchemeValue> values;
};

struct Shared {
  std::shared_ptr<Device> dev;
  std::unique_ptr<CkksClient> client;
  KeySet keys;
  std::mutex execMutex;   // serialises execute / executeMany callers of one context (see executeMany)
};

// the cached plan of one (program, context) pair; owned by the Program
struct PlanHolder {
  std::shared_ptr<Shared> keepAlive;  // the plan references the context's encoder and keys
  std::unique_ptr<Executor> exec;
  std::uint64_t termCount = 0;
  ExecOptions opt;
};

// a batch that has been enqueued (H2D, plan replays, D2H) and not yet collected: B200Public::submitMany / collect
struct PendingBatch {
  Program *program = nullptr;
  std::vector<const B200Valuation *> inputs;   // the caller keeps them alive until collect()
  std::vector<B200Valuation> outs;
  std::vector<void *> streams;
  std::vector<std::shared_ptr<PlanHolder>> plans;   // the plans it runs on stay alive until it is collected (set_options may replace them meanwhile)
  int slot = 0;
  bool collected = false;
  std::shared_ptr<Shared> keep;
  // dropped without collect(): the copies into `outs` must not outlive their buffers
  ~PendingBatch() { if (!collected && keep) for (void *st : streams) keep->dev->sync(st); }
};

class B200Public {
public:
  explicit B200Public(std::shared_ptr<Shared> s) : s_(std::move(s)) {
    static std::atomic<std::uint64_t> next{1};
    id_ = next.fetch_add(1);
  }

  // SEALPublic::encrypt -- reference eva/seal/seal.cpp:24-102
  B200Valuation encrypt(const Valuation &inputs, const CKKSSignature &sig) {
    auto &enc = s_->client->encoder();
    const std::size_t slots = enc.slotCount();
    if (slots < (std::size_t)sig.vecSize) throw std::runtime_error("Vector size cannot be larger than slot count");
    if (slots % sig.vecSize) throw std::runtime_error("Vector size must exactly divide the slot count");
    B200Valuation out;
    const u64 N = s_->dev->N();
    for (auto &in : inputs) {
      const auto &v = in.second;
      if (v.size() != (std::size_t)sig.vecSize) throw std::runtime_error("Input size does not match program vector size");
      const CKKSEncodingInfo &info = sig.inputs.at(in.first);
      if (info.inputType == Type::Raw) { out[in.first] = std::make_shared<ConstantValue>(sig.vecSize, v); continue; }
      const int ell = s_->dev->k() - 1 - info.level;
      if (ell < 1) throw std::runtime_error("input level exceeds the modulus chain");
      std::vector<double> rep;
      rep.reserve(slots);
      for (std::size_t r = slots / v.size(); r > 0; --r) rep.insert(rep.end(), v.begin(), v.end());
      const double scale = std::ldexp(1.0, info.scale);
      DBuf pt(s_->dev, (std::size_t)ell * N);
      enc.encode(rep, scale, ell, pt.get());
      if (info.inputType == Type::Cipher) {
        DBuf ct = s_->client->encrypt(s_->keys, pt.get(), ell);
        HostCipher h; h.size = 2; h.ell = ell; h.scale = scale; h.data.resize((std::size_t)2 * ell * N);
        s_->dev->download(h.data.data(), ct.get(), h.data.size() * 8); s_->dev->sync();
        out[in.first] = std::move(h);
      } else {
        HostPlain h; h.ell = ell; h.scale = scale; h.data.resize((std::size_t)ell * N);
        s_->dev->download(h.data.data(), pt.get(), h.data.size() * 8); s_->dev->sync();
        out[in.first] = std::move(h);
      }
    }
    return out;
  }

  // one plan (arena, streams, captured graph) per (context, batch, replica); replicas of the same
  // batch size run concurrently in executeMany
  std::uint64_t planKey(int batch, int replica) const { return (id_ << 32) | ((std::uint64_t)replica << 16) | (std::uint64_t)batch; }
  Executor &executorFor(Program &program, int batch = 1, int replica = 0) { return *holderFor(program, batch, replica)->exec; }
  std::shared_ptr<PlanHolder> holderFor(Program &program, int batch = 1, int replica = 0) {
    if (batch < 1 || batch > 65535 || replica < 0 || replica > 65535) throw std::invalid_argument("batch / replica out of range");
    const std::uint64_t key = planKey(batch, replica);
    auto h = std::static_pointer_cast<PlanHolder>(program.attachment(key));
    if (!h || h->termCount != program.termCount() || h->opt.hoistModUp != options.hoistModUp || h->opt.approxHoist != options.approxHoist || h->opt.rotationChunk != options.rotationChunk) {   // stale plan: rebuilt
      h = std::make_shared<PlanHolder>();
      h->keepAlive = s_;
      ExecOptions o = options;
      o.batch = batch;
      h->opt = o;   // as requested (what the staleness test above compares)
      h->exec = std::make_unique<Executor>(s_->dev, s_->client->encoder(), s_->keys, program, o);
      h->termCount = program.termCount();
      program.attach(key, h);
    }
    return h;
  }
  void dropExecutor(Program &program, int batch = 1, int replica = 0) { program.attach(planKey(batch, replica), nullptr); }

  // upload host inputs into the executor's arena (H2D on `stream`).  Client-supplied valuations are untrusted
  // (they may come from a file): every buffer is checked against the shape the plan expects before any copy.
  void stageInputs(Executor &ex, Program &program, const B200Valuation &inputs, void *stream, int b = 0) {
    const u64 N = s_->dev->N();
    for (auto &in : inputs) {
      auto term = program.getInput(in.first);
      const ValueInfo &vi = ex.info(term);
      if (auto *c = std::get_if<HostCipher>(&in.second)) {
        if (vi.kind != Kind::Cipher || c->ell != vi.ell || c->size != vi.size || c->data.size() != (std::size_t)vi.size * vi.ell * N)
          throw std::runtime_error("input " + in.first + ": ciphertext does not match the program signature");
        s_->dev->upload(ex.valuePtr(term, b), c->data.data(), c->data.size() * 8, stream);
      } else if (auto *p = std::get_if<HostPlain>(&in.second)) {
        if (vi.kind != Kind::Plain || p->ell != vi.ell || p->data.size() != (std::size_t)vi.ell * N)
          throw std::runtime_error("input " + in.first + ": plaintext does not match the program signature");
        s_->dev->upload(ex.valuePtr(term, b), p->data.data(), (std::size_t)p->ell * N * 8, stream);
      } else {
        if (vi.kind != Kind::Raw) throw std::runtime_error("input " + in.first + ": raw vector does not match the program signature");
        auto &cv = std::get<std::shared_ptr<ConstantValue>>(in.second);
        std::vector<double> x;
        cv->expandTo(x, program.getVecSize());
        ex.setRawInput(in.first, x, b);
      }
    }
  }
  // plans and arenas are cached on the Program: a call that omits an input would silently compute on the previous
  // call's ciphertext.  The reference builds a fresh executor per call and fails on the missing value
  // (seal_executor.h:264); so does this.
  static void requireAllInputs(Program &program, const B200Valuation &inputs) {
    for (auto &in : program.getInputs())
      if (inputs.values.find(in.first) == inputs.values.end()) throw std::runtime_error("Missing input value: " + in.first);
  }
  // SEALPublic::execute -- reference eva/seal/seal.cpp:104-122.  Host buffers in,
  // host buffers out: H2D of the inputs, the DAG on the GPU, D2H of the outputs.
  B200Valuation execute(Program &program, const B200Valuation &inputs) {
    std::vector<const B200Valuation *> in{&inputs};
    return std::move(executeMany(program, in)[0]);
  }
  // Batched execute: N valuations of the same program run as ONE plan replay whose kernels
  // each cover all instances (horizontal fusion).  Results are identical
  // to N separate execute() calls.
  std::vector<B200Valuation> executeBatch(Program &program, const std::vector<B200Valuation> &inputs) {
    std::vector<const B200Valuation *> in;
    for (auto &v : inputs) in.push_back(&v);
    return executeMany(program, in);
  }
  // B valuations are split into chunks of options.fuse instances; every chunk has its own plan
  // replica (arena + captured graph) and all replicas are in flight at once: H2D, graph and D2H of
  // different chunks overlap, and 148 SMs are filled by concurrency rather than by launch width.
  std::vector<B200Valuation> executeMany(Program &program, const std::vector<const B200Valuation *> &inputs) {
    auto p = submitMany(program, inputs, 0);
    return collect(*p);
  }
  // Pipelined form of executeMany: submitMany enqueues a batch (H2D of its inputs, the plan replays, D2H of its outputs) and returns;
  // collect waits for it and hands out the results.  Batches submitted with different `slot` numbers (0..7) use disjoint plan
  // replicas, so the head of one overlaps the tail of the other (each batch alone ramps up replica by replica and drains the same way).
  std::shared_ptr<PendingBatch> submitMany(Program &program, const std::vector<const B200Valuation *> &inputs, int slot = 0) {
    if (inputs.empty()) throw std::invalid_argument("execute needs at least one valuation");
    if (slot < 0 || slot > 7) throw std::invalid_argument("slot out of range (0..7)");
    for (auto *v : inputs) requireAllInputs(program, *v);
    // one caller at a time per context: plan replicas own single arenas and raw-input buffers (the reference's
    // execute is re-entrant because it builds a new executor per call; concurrent callers are serialised here)
    std::lock_guard<std::mutex> guard(s_->execMutex);
    if (auto prev = inFlight_[slot].lock(); prev && !prev->collected) throw std::runtime_error("slot in use: collect the batch submitted before");
    auto p = std::make_shared<PendingBatch>();
    p->program = &program; p->inputs = inputs; p->slot = slot; p->keep = s_;
    enqueueLocked(*p);
    inFlight_[slot] = p;
    return p;
  }
  std::vector<B200Valuation> collect(PendingBatch &p) {
    if (p.collected) throw std::runtime_error("this batch has been collected already");
    std::lock_guard<std::mutex> guard(s_->execMutex);
    p.collected = true;
    for (void *st : p.streams) s_->dev->sync(st);
    bool redo = false;
    for (auto &h : p.plans) if (h->exec->flagsRaised()) redo = true;
    if (redo) {
      // a digit of a rotated ciphertext held a zero coefficient: the shared mod-up of that rotation group is not SEAL's value
      // (ops_impl.hpp hoisted_modup).  Redo the call on plans without it -- exact, and from now on for this context.
      if (verbosity() >= 1) std::fprintf(stderr, "EVA: zero digit coefficient met; rotation groups of this context no longer share their mod-up\n");
      options.hoistModUp = false;      // executorFor rebuilds every plan whose option differs
      p.outs.clear(); p.streams.clear(); p.plans.clear();
      enqueueLocked(p);
      for (void *st : p.streams) s_->dev->sync(st);
    }
    return std::move(p.outs);
  }
  static constexpr int kSlotReplicas = 4096;   // replica numbers of slot s: s * kSlotReplicas + r
  void enqueueLocked(PendingBatch &p) {
    Program &program = *p.program;
    const std::vector<const B200Valuation *> &inputs = p.inputs;
    const int B = (int)inputs.size();
    const int F = std::max(1, std::min(options.fuse, B));
    const int base = p.slot * kSlotReplicas;
    std::vector<B200Valuation> &outs = p.outs;
    outs.assign(B, B200Valuation());
    std::vector<void *> &streams = p.streams;
    const u64 N = s_->dev->N();
    static const bool trace = std::getenv("EVAB_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    double tPlan = 0, tStage = 0, tRun = 0;
    auto t00 = now();
    // replicas are bounded by device memory (an arena each): with R replicas chunk g runs on replica g % R
    // after that replica's previous chunk has completed (its outputs are already on their way to the host)
    int R = std::min((B + F - 1) / F, kSlotReplicas);
    {
      Executor &first = executorFor(program, std::min(F, B), base);
      std::size_t have = 1;   // the slot's first replica exists; count the ones already built for this program
      while ((int)have < R && program.attachment(planKey(std::min(F, B), base + (int)have))) have++;
      if ((int)have < R) {    // new arenas are needed: how many fit?  (cudaMemGetInfo is slow: never on the steady path)
        std::size_t freeB = 0, totalB = 0;
        check(evab_mem_info(s_->dev->ctx(), &freeB, &totalB));
        const std::size_t per = std::max<std::size_t>(first.arenaBytes(), 1);
        const std::size_t extra = (std::size_t)(0.8 * (double)freeB) / per;
        R = (int)std::max<std::size_t>(1, std::min<std::size_t>((std::size_t)R, have + extra));
      }
      if (const char *cap = std::getenv("EVAB_MAX_REPLICAS")) R = std::max(1, std::min(R, std::atoi(cap)));   // tests / tuning
    }
    std::vector<char> busy(R, 0);
    for (int b0 = 0, g = 0; b0 < B; b0 += F, g++) {
      const int nb = std::min(F, B - b0);
      const int r = g % R;
      auto t0 = now();
      std::shared_ptr<PlanHolder> holder = holderFor(program, nb, base + r);
      Executor &ex = *holder->exec;
      void *st = ex.mainStream();
      if (busy[r]) s_->dev->sync(st);   // the replica's arena is reused: wait for its previous chunk
      busy[r] = 1;
      if (std::find(streams.begin(), streams.end(), st) == streams.end()) streams.push_back(st);
      if (std::find(p.plans.begin(), p.plans.end(), holder) == p.plans.end()) p.plans.push_back(holder);
      auto t1 = now();
      for (int b = 0; b < nb; b++) stageInputs(ex, program, *inputs[b0 + b], st, b);
      auto t2 = now();
      ex.run(st);
      auto t3 = now();
      tPlan += std::chrono::duration<double>(t1 - t0).count(); tStage += std::chrono::duration<double>(t2 - t1).count();
      tRun += std::chrono::duration<double>(t3 - t2).count();
      for (int b = 0; b < nb; b++)
        for (auto &o : program.getOutputs()) {
          const ValueInfo &vi = ex.info(o.second);
          B200Valuation &out = outs[b0 + b];
          if (vi.kind == Kind::Cipher) {
            HostCipher h; h.size = vi.size; h.ell = vi.ell; h.scale = vi.scale; h.data.resize((std::size_t)vi.size * vi.ell * N);
            s_->dev->download(h.data.data(), ex.valuePtr(o.second, b), h.data.size() * 8, st);
            out[o.first] = std::move(h);
          } else if (vi.kind == Kind::Plain) {
            HostPlain h; h.ell = vi.ell; h.scale = vi.scale; h.data.resize((std::size_t)vi.ell * N);
            s_->dev->download(h.data.data(), ex.valuePtr(o.second, b), h.data.size() * 8, st);
            out[o.first] = std::move(h);
          } else {
            out[o.first] = std::make_shared<ConstantValue>(program.getVecSize(), ex.rawValue(o.second->index, b));
          }
        }
    }
    if (trace)
      std::fprintf(stderr, "[evab] enqueue B=%d F=%d slot=%d: plan %.3f stage %.3f run %.3f enqueue-total %.3f ms\n", B, F, p.slot, tPlan * 1e3, tStage * 1e3, tRun * 1e3,
                   std::chrono::duration<double>(now() - t00).count() * 1e3);
  }
  std::shared_ptr<Shared> shared() const { return s_; }
  ExecOptions options;

private:
  std::shared_ptr<Shared> s_;
  std::uint64_t id_ = 0;
  std::weak_ptr<PendingBatch> inFlight_[8];
};

class B200Secret {
public:
  explicit B200Secret(std::shared_ptr<Shared> s) : s_(std::move(s)) {}
  // SEALSecret::decrypt -- reference eva/seal/seal.cpp:124-146
  Valuation decrypt(const B200Valuation &enc, const CKKSSignature &sig) {
    Valuation out;
    const u64 N = s_->dev->N();
    auto &encoder = s_->client->encoder();
    for (auto &e : enc) {
      if (auto *c = std::get_if<HostCipher>(&e.second)) {
        DBuf ct(s_->dev, c->data.size());
        s_->dev->upload(ct.get(), c->data.data(), c->data.size() * 8);
        DBuf pt = s_->client->decrypt(s_->keys, ct.get(), c->size, c->ell);
        out[e.first] = encoder.decode(pt.get(), c->ell, c->scale);
      } else if (auto *p = std::get_if<HostPlain>(&e.second)) {
        DBuf pt(s_->dev, (std::size_t)p->ell * N);
        s_->dev->upload(pt.get(), p->data.data(), (std::size_t)p->ell * N * 8);
        out[e.first] = encoder.decode(pt.get(), p->ell, p->scale);
      } else {
        std::get<std::shared_ptr<ConstantValue>>(e.second)->expandTo(out[e.first], sig.vecSize);
      }
      out.at(e.first).resize(sig.vecSize);
    }
    return out;
  }
  std::shared_ptr<Shared> shared() const { return s_; }
private:
  std::shared_ptr<Shared> s_;
};

// generateKeys -- reference eva/seal/seal.cpp:174-203
inline std::tuple<std::unique_ptr<B200Public>, std::unique_ptr<B200Secret>>
generateKeys(const CKKSParameters &params, int device = 0, std::uint64_t seed = 0) {
  std::vector<int> bits(params.primeBits.begin(), params.primeBits.end());
  auto primes = hmod::createCoeffModulus(params.polyModulusDegree, bits);
  auto s = std::make_shared<Shared>();
  s->dev = std::make_shared<Device>(params.polyModulusDegree, primes, device);
  // seed == 0 (the default): keys and encryption randomness from a CSPRNG keyed with 256 bits of OS entropy;
  // a non-zero seed is the deterministic test path (host/csprng.hpp)
  s->client = std::make_unique<CkksClient>(s->dev, seed);
  std::vector<int> rots(params.rotations.begin(), params.rotations.end());
  s->client->keygen(s->keys, rots);
  return std::make_tuple(std::make_unique<B200Public>(s), std::make_unique<B200Secret>(s));
}

}  // namespace evab
