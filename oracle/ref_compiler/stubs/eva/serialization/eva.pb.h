// Stand-in for the protoc-generated header (absent: no protoc here); just
// enough surface for the inline uses in /root/reference/eva/ir/constant_value.h.
#pragma once
#include <cstdint>
#include <vector>
namespace eva { namespace msg {
struct RepeatedDouble { void Reserve(int) {} void Add(double) {} };
struct ConstantValue {
  void set_size(std::uint32_t) {}
  RepeatedDouble *mutable_values() { return &v; }
  void add_sparse_indices(std::uint32_t) {}
  void add_values(double) {}
  RepeatedDouble v;
};
struct Attribute {};
struct Program {};
} }
