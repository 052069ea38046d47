// miplib/util.hpp - small helpers around Var (reference miplib/util.hpp:15-97): VarPair,
// the Vars<> container factory and the PartialSolution warm-start type.  No Boost.
#pragma once

#include <algorithm>
#include <functional>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <vector>

#include "var.hpp"

namespace miplib {

using VarPair = std::pair<Var, Var>;

// Vars(solver, Var::Type::Binary).as_vector(10) makes ten fresh variables with the same arguments.
template<class... VarArgs>
struct Vars
{
  explicit Vars(VarArgs const&... args) : m_var_args(args...) {}

  std::vector<Var> as_vector(std::size_t count) const
  {
    std::vector<Var> out;
    out.reserve(count);
    while (out.size() < count) out.push_back(std::make_from_tuple<Var>(m_var_args));
    return out;
  }

  template<class Keys>
  auto as_unordered_map_values(Keys const& keys) const
  {
    using Key = std::decay_t<decltype(*std::begin(keys))>;
    std::unordered_map<Key, Var> out;
    for (auto const& k : keys) out.emplace(k, std::make_from_tuple<Var>(m_var_args));
    return out;
  }

  template<class Keys>
  auto as_unordered_map_values(
    Keys const& keys, std::function<std::string(typename Keys::value_type const&)> const& naming_fn) const
  {
    using Key = std::decay_t<decltype(*std::begin(keys))>;
    std::unordered_map<Key, Var> out;
    for (auto const& k : keys)
    {
      Var v = std::make_from_tuple<Var>(m_var_args);
      v.set_name(naming_fn(k));
      out.emplace(k, v);
    }
    return out;
  }

  std::tuple<VarArgs...> m_var_args;
};

// values for some variables; used as a start point (Solver::set_warm_start)
using PartialSolution = std::unordered_map<Var, double>;

}  // namespace miplib

namespace std {

template<> struct hash<miplib::VarPair>
{
  size_t operator()(miplib::VarPair const& p) const noexcept
  {
    size_t a = hash<miplib::Var>()(p.first), b = hash<miplib::Var>()(p.second);
    return a ^ (b + 0x9e3779b97f4a7c15ull + (a << 6) + (a >> 2));
  }
};

template<> struct equal_to<miplib::VarPair>
{
  bool operator()(miplib::VarPair const& a, miplib::VarPair const& b) const noexcept
  {
    return a.first.is_same(b.first) && a.second.is_same(b.second);
  }
};

}  // namespace std
