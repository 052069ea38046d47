// tests/cpp/harness.hpp - a ~60 line stand-in for Catch2 2.13.3 (reference test/CMakeLists.txt:18),
// which cannot be fetched offline.  REQUIRE / REQUIRE_THROWS_AS / REQUIRE_NOTHROW keep the
// reference's spelling so the re-hosted assertions read like test/unit/*.cpp.
#pragma once

#include <functional>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace harness {

struct Failure : std::runtime_error { using std::runtime_error::runtime_error; };

struct Case
{
  std::string name;
  bool needs_gpu;
  std::function<void()> body;
};

inline std::vector<Case>& registry()
{
  static std::vector<Case> cases;
  return cases;
}

struct Registrar
{
  Registrar(std::string name, bool needs_gpu, std::function<void()> body)
  {
    registry().push_back({std::move(name), needs_gpu, std::move(body)});
  }
};

template<class T>
std::string str(T const& v)
{
  std::ostringstream s;
  s << v;
  return s.str();
}

}  // namespace harness

#define HARNESS_CAT2(a, b) a##b
#define HARNESS_CAT(a, b) HARNESS_CAT2(a, b)
#define TEST_CASE_IMPL(fn, name, gpu)                                        \
  static void fn();                                                          \
  static harness::Registrar HARNESS_CAT(fn, _reg)(name, gpu, fn);            \
  static void fn()
#define CPU_TEST(name) TEST_CASE_IMPL(HARNESS_CAT(test_fn_, __LINE__), name, false)
#define GPU_TEST(name) TEST_CASE_IMPL(HARNESS_CAT(test_fn_, __LINE__), name, true)

#define REQUIRE(cond)                                                                              \
  do {                                                                                             \
    if (!(cond))                                                                                   \
      throw harness::Failure(std::string(__FILE__) + ":" + std::to_string(__LINE__) +             \
                             ": REQUIRE(" #cond ") failed");                                       \
  } while (0)

#define REQUIRE_EQ_STR(expr, expected)                                                             \
  do {                                                                                             \
    std::string got_ = harness::str(expr);                                                         \
    if (got_ != (expected))                                                                        \
      throw harness::Failure(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": \"" +    \
                             got_ + "\" != \"" + (expected) + "\"");                               \
  } while (0)

#define REQUIRE_THROWS_AS(expr, type)                                                              \
  do {                                                                                             \
    bool thrown_ = false;                                                                          \
    try { (void)(expr); } catch (type const&) { thrown_ = true; }                                  \
    if (!thrown_)                                                                                  \
      throw harness::Failure(std::string(__FILE__) + ":" + std::to_string(__LINE__) +             \
                             ": " #expr " did not throw " #type);                                  \
  } while (0)

#define REQUIRE_NOTHROW(expr)                                                                      \
  do {                                                                                             \
    try { (void)(expr); } catch (std::exception const& e_) {                                       \
      throw harness::Failure(std::string(__FILE__) + ":" + std::to_string(__LINE__) +             \
                             ": " #expr " threw: " + e_.what());                                   \
    }                                                                                              \
  } while (0)
