// reference_eval.hpp -- plaintext ("reference") semantics of an EVA program on
// vectors of doubles: evaluate(program, inputs).  Mirrors the behaviour of the
// reference's ReferenceExecutor (eva/common/reference_executor.cpp:60-115,
// eva/eva.cpp:11-21): HE-only ops (Encode, Relinearize, ModSwitch, Rescale,
// Output) are identities.
#pragma once
#include "backend.hpp"

namespace evab {

inline Valuation evaluate(Program &program, const Valuation &inputs) {
  std::vector<std::vector<double>> val(program.termCount());
  const std::size_t vs = program.getVecSize();
  for (auto &in : inputs) {
    auto t = program.getInput(in.first);  // throws out_of_range for unknown names
    ConstantValue(vs, in.second).expandTo(val[t->index], vs);
  }
  for (auto &t : program.toposort()) {
    auto &out = val[t->index];
    auto A = [&](int i) -> const std::vector<double> & { return val[t->operandAt(i)->index]; };
    switch (t->op) {
      case Op::Input:
        if (out.empty()) throw std::runtime_error("Missing input value");
        break;
      case Op::Constant: t->constant->expandTo(out, vs); break;
      case Op::Add: case Op::Sub: case Op::Mul: {
        const auto &x = A(0), &y = A(1);
        out.resize(x.size());
        for (std::size_t i = 0; i < x.size(); i++) out[i] = t->op == Op::Add ? x[i] + y[i] : t->op == Op::Sub ? x[i] - y[i] : x[i] * y[i];
      } break;
      case Op::Negate: { const auto &x = A(0); out.resize(x.size()); for (std::size_t i = 0; i < x.size(); i++) out[i] = -x[i]; } break;
      case Op::RotateLeftConst: case Op::RotateRightConst: {
        const auto &x = A(0);
        const long long n = (long long)x.size();
        long long sh = *t->rotation;
        if (t->op == Op::RotateRightConst) sh = -sh;
        sh %= n; if (sh < 0) sh += n;
        out.resize(x.size());
        for (long long i = 0; i < n; i++) out[i] = x[(i + sh) % n];
      } break;
      case Op::Encode: case Op::Output: case Op::Relinearize: case Op::ModSwitch: case Op::Rescale:
        out = A(0);
        break;
      default: throw std::runtime_error(std::string("Unhandled op ") + opName(t->op));
    }
  }
  Valuation res;
  for (auto &o : program.getOutputs()) res[o.first] = val[o.second->index];
  return res;
}

}  // namespace evab
