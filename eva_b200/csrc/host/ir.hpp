// ir.hpp -- term-DAG intermediate representation of an EVA program (host side).
//
// API-compatible re-implementation of the reference IR surface that the Python
// DSL, the CKKS compiler and the executor rely on (reference eva/ir/program.h,
// term.h, ops.h, types.h, attributes.h, constant_value.h): same op / type codes,
// same seven attributes, Program::make* factory names.  Defs are shared-owned by
// their uses (a term nobody uses and no output names disappears), uses are
// tracked as plain back-pointers; sources/sinks are maintained incrementally so
// traversals can start from either end while passes rewrite the graph.
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

namespace evab {

enum class Op : int {
  Undef = 0, Input = 1, Output = 2, Constant = 3, Negate = 10, Add = 11, Sub = 12, Mul = 13,
  RotateLeftConst = 14, RotateRightConst = 15, Relinearize = 20, ModSwitch = 21, Rescale = 22, Encode = 23
};
enum class Type : int { Undef = 0, Cipher = 1, Raw = 2, Plain = 3 };

inline const char *opName(Op op) {
  switch (op) {
    case Op::Undef: return "Undef"; case Op::Input: return "Input"; case Op::Output: return "Output";
    case Op::Constant: return "Constant"; case Op::Negate: return "Negate"; case Op::Add: return "Add";
    case Op::Sub: return "Sub"; case Op::Mul: return "Mul"; case Op::RotateLeftConst: return "RotateLeftConst";
    case Op::RotateRightConst: return "RotateRightConst"; case Op::Relinearize: return "Relinearize";
    case Op::ModSwitch: return "ModSwitch"; case Op::Rescale: return "Rescale"; case Op::Encode: return "Encode";
  }
  throw std::runtime_error("Invalid op");
}
inline const char *typeName(Type t) {
  switch (t) {
    case Type::Undef: return "Undef"; case Type::Cipher: return "Cipher"; case Type::Raw: return "Raw"; case Type::Plain: return "Plain";
  }
  throw std::runtime_error("Invalid type");
}

// A constant vector of `size` logical slots given by `values` repeated
// (values.size() divides size); a single value is a uniform constant.
class ConstantValue {
public:
  ConstantValue(std::size_t size, std::vector<double> values) : size_(size), values_(std::move(values)) {
    if (values_.empty() || size_ % values_.size() != 0) throw std::runtime_error("DenseConstantValue size must exactly divide size");
  }
  void expandTo(std::vector<double> &out, std::size_t slots) const {
    if (slots < size_) throw std::runtime_error("Slots must be at least size of constant");
    if (slots % size_ != 0) throw std::runtime_error("Size must exactly divide slots");
    out.clear();
    out.reserve(slots);
    for (std::size_t r = slots / values_.size(); r > 0; --r) out.insert(out.end(), values_.begin(), values_.end());
  }
  bool isZero() const { return std::all_of(values_.begin(), values_.end(), [](double v) { return v == 0; }); }
  const std::vector<double> &values() const { return values_; }
  std::size_t size() const { return size_; }
private:
  std::size_t size_;
  std::vector<double> values_;
};

class Program;

class Term : public std::enable_shared_from_this<Term> {
public:
  using Ptr = std::shared_ptr<Term>;
  Term(Op op, Program &program);
  ~Term();
  Term(const Term &) = delete;

  const Op op;
  Program &program;
  std::uint64_t index;
  // false once the owning Program is gone (a Python-held Term may outlive it): the destructor then
  // must not touch the program's source / sink sets
  std::shared_ptr<bool> programAlive;

  // attributes (reference eva/ir/attributes.h:11-18)
  std::optional<std::uint32_t> rescaleDivisor, range, encodeAtScale, encodeAtLevel;
  std::optional<std::int32_t> rotation;
  std::optional<Type> type;
  std::shared_ptr<ConstantValue> constant;
  void copyAttributesFrom(const Term &o) {
    rescaleDivisor = o.rescaleDivisor; range = o.range; encodeAtScale = o.encodeAtScale; encodeAtLevel = o.encodeAtLevel;
    rotation = o.rotation; type = o.type; constant = o.constant;
  }

  const std::vector<Ptr> &getOperands() const { return operands_; }
  std::size_t numOperands() const { return operands_.size(); }
  Ptr operandAt(std::size_t i) const { return operands_.at(i); }
  void addOperand(const Ptr &t);
  void setOperands(std::vector<Ptr> o);
  bool eraseOperand(const Ptr &t);
  bool replaceOperand(const Ptr &oldT, const Ptr &newT);

  std::size_t numUses() const { return uses_.size(); }
  std::vector<Ptr> getUses() const {
    std::vector<Ptr> r;
    for (Term *u : uses_) r.push_back(u->shared_from_this());
    return r;
  }
  void replaceUsesWithIf(const Ptr &t, const std::function<bool(const Ptr &)> &pred);
  void replaceAllUsesWith(const Ptr &t) { replaceUsesWithIf(t, [](const Ptr &) { return true; }); }
  void replaceOtherUsesWith(const Ptr &t) { replaceUsesWithIf(t, [&](const Ptr &u) { return u != t; }); }
  bool isInternal() const { return !operands_.empty() && !uses_.empty(); }

private:
  std::vector<Ptr> operands_;   // use -> def (owning)
  std::vector<Term *> uses_;    // def -> use (back pointers, one entry per operand slot)
  void addUse(Term *u);
  bool eraseUse(Term *u);
};

// index-addressed side tables that follow the program as it grows
class TermMapBase {
public:
  virtual ~TermMapBase() {}
  virtual void resize(std::size_t n) = 0;
};

class Program {
public:
  Program(std::string name, std::uint64_t vecSize) : name_(std::move(name)), vecSize_((std::uint32_t)vecSize) {
    if (vecSize == 0) throw std::runtime_error("Vector size must be non-zero");
    if (vecSize & (vecSize - 1)) throw std::runtime_error("Vector size must be a power-of-two");
  }
  Program(const Program &) = delete;
  ~Program() {
    attachments_.clear();  // backend plans hold raw Term pointers: they die with the program
    // release the roots first; Terms deregister themselves from sources_/sinks_
    outputs_.clear();
    inputs_.clear();
    *alive_ = false;   // terms still referenced from outside are orphaned
  }
  // opaque per-backend state (e.g. a cached execution plan) that must not outlive the program
  std::shared_ptr<void> attachment(std::uint64_t key) const { auto it = attachments_.find(key); return it == attachments_.end() ? nullptr : it->second; }
  void attach(std::uint64_t key, std::shared_ptr<void> v) { if (v) attachments_[key] = std::move(v); else attachments_.erase(key); }

  Term::Ptr makeTerm(Op op, const std::vector<Term::Ptr> &operands = {}) {
    auto t = std::make_shared<Term>(op, *this);
    if (!operands.empty()) t->setOperands(operands);
    return t;
  }
  Term::Ptr makeConstant(std::shared_ptr<ConstantValue> v) { auto t = makeTerm(Op::Constant); t->constant = std::move(v); return t; }
  Term::Ptr makeDenseConstant(std::vector<double> values) { return makeConstant(std::make_shared<ConstantValue>(vecSize_, std::move(values))); }
  Term::Ptr makeUniformConstant(double value) { return makeDenseConstant({value}); }
  Term::Ptr makeInput(const std::string &name, Type type = Type::Cipher) {
    auto t = makeTerm(Op::Input); t->type = type; inputs_.emplace(name, t); return t;
  }
  Term::Ptr makeOutput(const std::string &name, const Term::Ptr &term) {
    auto t = makeTerm(Op::Output, {term}); outputs_.emplace(name, t); return t;
  }
  Term::Ptr makeLeftRotation(const Term::Ptr &term, std::int32_t slots) { auto t = makeTerm(Op::RotateLeftConst, {term}); t->rotation = slots; return t; }
  Term::Ptr makeRightRotation(const Term::Ptr &term, std::int32_t slots) { auto t = makeTerm(Op::RotateRightConst, {term}); t->rotation = slots; return t; }
  Term::Ptr makeRescale(const Term::Ptr &term, std::uint32_t by) { auto t = makeTerm(Op::Rescale, {term}); t->rescaleDivisor = by; return t; }

  Term::Ptr getInput(const std::string &name) const {
    auto it = inputs_.find(name);
    if (it == inputs_.end()) throw std::out_of_range("No input named " + name);
    return it->second;
  }
  const std::map<std::string, Term::Ptr> &getInputs() const { return inputs_; }
  const std::map<std::string, Term::Ptr> &getOutputs() const { return outputs_; }
  const std::string &getName() const { return name_; }
  void setName(std::string n) { name_ = std::move(n); }
  std::uint32_t getVecSize() const { return vecSize_; }
  std::uint64_t termCount() const { return nextIndex_; }

  // terms without operands / without uses, ordered by index (deterministic)
  std::vector<Term::Ptr> getSources() const { return sorted(sources_); }
  std::vector<Term::Ptr> getSinks() const { return sorted(sinks_); }

  // every live term in a topological order (operands before uses), ties by index
  std::vector<Term::Ptr> toposort() const;
  std::unique_ptr<Program> deepCopy() const;
  std::string toDOT() const;

  void registerMap(TermMapBase *m) { maps_.push_back(m); m->resize(nextIndex_); }
  void unregisterMap(TermMapBase *m) { maps_.erase(std::remove(maps_.begin(), maps_.end(), m), maps_.end()); }

private:
  friend class Term;
  std::uint64_t allocateIndex() {
    std::uint64_t i = nextIndex_++;
    for (auto *m : maps_) m->resize(nextIndex_);
    return i;
  }
  static std::vector<Term::Ptr> sorted(const std::unordered_set<Term *> &s) {
    std::vector<Term::Ptr> r;
    for (Term *t : s) r.push_back(t->shared_from_this());
    std::sort(r.begin(), r.end(), [](const Term::Ptr &a, const Term::Ptr &b) { return a->index < b->index; });
    return r;
  }
  std::string name_;
  std::uint32_t vecSize_;
  std::uint64_t nextIndex_ = 0;
  std::unordered_set<Term *> sources_, sinks_;
  std::shared_ptr<bool> alive_ = std::make_shared<bool>(true);
  std::vector<TermMapBase *> maps_;
  std::map<std::uint64_t, std::shared_ptr<void>> attachments_;
  // roots last: their destruction tears the graph down while the sets above are alive
  std::map<std::string, Term::Ptr> outputs_, inputs_;
};

template <class T> class TermMap : public TermMapBase {
public:
  explicit TermMap(Program &p) : p_(p) { p_.registerMap(this); }
  ~TermMap() override { p_.unregisterMap(this); }
  TermMap(const TermMap &o) : p_(o.p_), v_(o.v_) { p_.registerMap(this); }
  void resize(std::size_t n) override { if (v_.size() < n) v_.resize(n); }
  T &operator[](const Term::Ptr &t) { return v_.at(t->index); }
  T &operator[](const Term *t) { return v_.at(t->index); }
  const T &at(const Term::Ptr &t) const { return v_.at(t->index); }
  void clear() { std::fill(v_.begin(), v_.end(), T{}); }
private:
  Program &p_;
  std::vector<T> v_;
};
template <> class TermMap<bool> : public TermMapBase {
public:
  explicit TermMap(Program &p) : p_(p) { p_.registerMap(this); }
  ~TermMap() override { p_.unregisterMap(this); }
  void resize(std::size_t n) override { if (v_.size() < n) v_.resize(n, 0); }
  char &operator[](const Term::Ptr &t) { return v_.at(t->index); }
  void clear() { std::fill(v_.begin(), v_.end(), 0); }
private:
  Program &p_;
  std::vector<char> v_;
};

// ---------------------------------------------------------------------------
inline Term::Term(Op o, Program &p) : op(o), program(p), index(p.allocateIndex()), programAlive(p.alive_) {
  p.sources_.insert(this);
  p.sinks_.insert(this);
}
inline Term::~Term() {
  if (!*programAlive) return;   // operands outlive or die with this term; nothing to deregister from
  for (auto &d : operands_) d->eraseUse(this);
  if (operands_.empty()) program.sources_.erase(this);
  if (uses_.empty()) program.sinks_.erase(this);
}
inline void Term::addUse(Term *u) {
  if (uses_.empty()) program.sinks_.erase(this);
  uses_.push_back(u);
}
inline bool Term::eraseUse(Term *u) {
  auto it = std::find(uses_.begin(), uses_.end(), u);
  if (it == uses_.end()) return false;
  uses_.erase(it);
  if (uses_.empty()) program.sinks_.insert(this);
  return true;
}
inline void Term::addOperand(const Ptr &t) {
  if (operands_.empty()) program.sources_.erase(this);
  operands_.push_back(t);
  t->addUse(this);
}
inline void Term::setOperands(std::vector<Ptr> o) {
  if (operands_.empty() && !o.empty()) program.sources_.erase(this);
  for (auto &d : operands_) d->eraseUse(this);
  operands_ = std::move(o);
  for (auto &d : operands_) d->addUse(this);
  if (operands_.empty()) program.sources_.insert(this);
}
inline bool Term::eraseOperand(const Ptr &t) {
  auto it = std::find(operands_.begin(), operands_.end(), t);
  if (it == operands_.end()) return false;
  Ptr keep = *it;
  operands_.erase(it);
  keep->eraseUse(this);
  if (operands_.empty()) program.sources_.insert(this);
  return true;
}
inline bool Term::replaceOperand(const Ptr &oldT, const Ptr &newT) {
  bool any = false;
  for (auto &d : operands_)
    if (d == oldT) {
      Ptr keep = d;
      d = newT;
      newT->addUse(this);
      keep->eraseUse(this);
      any = true;
    }
  return any;
}
inline void Term::replaceUsesWithIf(const Ptr &t, const std::function<bool(const Ptr &)> &pred) {
  Ptr self = shared_from_this();
  for (auto &u : getUses())
    if (pred(u)) u->replaceOperand(self, t);
}

inline std::vector<Term::Ptr> Program::toposort() const {
  // Kahn over the live graph; ready set ordered by index for determinism
  std::vector<Term::Ptr> order;
  std::map<std::uint64_t, Term::Ptr> ready;
  std::map<std::uint64_t, std::size_t> pending;
  for (auto &s : getSources()) ready[s->index] = s;
  while (!ready.empty()) {
    auto it = ready.begin();
    Term::Ptr t = it->second;
    ready.erase(it);
    order.push_back(t);
    std::unordered_set<Term *> seen;
    for (auto &u : t->getUses()) {
      if (!seen.insert(u.get()).second) continue;
      auto pit = pending.find(u->index);
      if (pit == pending.end()) pit = pending.emplace(u->index, u->numOperands()).first;
      std::size_t cnt = 0;
      for (auto &o : u->getOperands()) if (o == t) cnt++;
      pit->second -= cnt;
      if (pit->second == 0) ready[u->index] = u;
    }
  }
  return order;
}
inline std::unique_ptr<Program> Program::deepCopy() const {
  auto np = std::make_unique<Program>(name_, vecSize_);
  std::map<std::uint64_t, Term::Ptr> m;
  for (auto &t : toposort()) {
    std::vector<Term::Ptr> ops;
    for (auto &o : t->getOperands()) ops.push_back(m.at(o->index));
    auto nt = np->makeTerm(t->op, ops);
    nt->copyAttributesFrom(*t);
    m[t->index] = nt;
  }
  for (auto &e : inputs_) np->inputs_.emplace(e.first, m.at(e.second->index));
  for (auto &e : outputs_) np->outputs_.emplace(e.first, m.at(e.second->index));
  return np;
}
inline std::string Program::toDOT() const {
  std::string s = "digraph \"" + name_ + "\" {\n";
  for (auto &t : toposort()) {
    s += "t" + std::to_string(t->index) + " [label=\"" + opName(t->op);
    if (t->rotation) s += "(" + std::to_string(*t->rotation) + ")";
    if (t->rescaleDivisor) s += "(" + std::to_string(*t->rescaleDivisor) + ")";
    s += "\"];\n";
    for (auto &o : t->getOperands()) s += "t" + std::to_string(o->index) + " -> t" + std::to_string(t->index) + ";\n";
  }
  s += "}\n";
  return s;
}

}  // namespace evab
