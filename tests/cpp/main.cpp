// tests/cpp/main.cpp - runner: `miplib_unit_test [--cpu-only] [--filter substr]`.
// Exit code 0 iff every selected case passed.  GPU cases are skipped (and reported as such) when
// Solver::backend_is_available(Backend::Cuda) is false, the reference's own convention
// (test/unit/expr.cpp:21-25).
#include <cstring>
#include <iostream>

#include <miplib/solver.hpp>

#include "harness.hpp"

int main(int argc, char** argv)
{
  bool cpu_only = false;
  std::string filter;
  for (int i = 1; i < argc; ++i)
  {
    if (!std::strcmp(argv[i], "--cpu-only")) cpu_only = true;
    else if (!std::strcmp(argv[i], "--filter") && i + 1 < argc) filter = argv[++i];
  }
  bool const gpu = miplib::Solver::backend_is_available(miplib::Solver::Backend::Cuda);
  int passed = 0, failed = 0, skipped = 0;
  for (auto const& c : harness::registry())
  {
    if (!filter.empty() && c.name.find(filter) == std::string::npos) continue;
    if (c.needs_gpu && (cpu_only || !gpu))
    {
      std::cout << "SKIP " << c.name << " (needs a GPU)\n";
      ++skipped;
      continue;
    }
    try
    {
      c.body();
      std::cout << "PASS " << c.name << "\n";
      ++passed;
    }
    catch (std::exception const& e)
    {
      std::cout << "FAIL " << c.name << ": " << e.what() << "\n";
      ++failed;
    }
  }
  std::cout << passed << " passed, " << failed << " failed, " << skipped << " skipped\n";
  return failed ? 1 : 0;
}
