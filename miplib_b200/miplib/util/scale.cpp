// miplib/util/scale.cpp - geometric-mean row scaling to a power of two (Tomlin 1975), the host
// side helper behind Solver::add(c, scale = true) (reference miplib/util/scale.cpp:52-112).  It
// is NOT the Ruiz / Pock-Chambolle equilibration of the Cuda backend, which runs on the device.
#include "scale.hpp"

#include <cmath>
#include <iostream>
#include <stdexcept>

#include "../constr.hpp"
#include "../solver.hpp"

namespace miplib {
namespace detail {

namespace {

// below 1: the smallest power of two that is not smaller; from 1 up: the nearest power of two
double friendly_power_of_two(double x)
{
  double const e = std::log2(x);
  return std::exp2(e < 0 ? std::ceil(e) : std::round(e));
}

}  // namespace

Constr scale_gm(Constr const& constr, double skip_lb, double skip_ub, bool ignore_inf_var_bounds)
{
  auto [low, high] = constr.expr().numerical_range(ignore_inf_var_bounds);
  double const inf = constr.expr().solver().infinity();
  if (low == inf || high == inf)
    throw std::logic_error("All variables must have defined domains for scaling constraint.");
  if (low >= skip_lb && high <= skip_ub) return constr;

  // divisor that centres [low, high] (together with 1, the scale of a unit coefficient)
  // geometrically around 1
  double divisor;
  if (high <= 1) divisor = std::sqrt(low);
  else if (low >= 1) divisor = std::sqrt(high);
  else divisor = std::sqrt(low * high);

  double const new_high = std::max(1 / divisor, high / divisor);
  double const new_low = std::min(1 / divisor, low / divisor);
  if (!ignore_inf_var_bounds && new_high / new_low >= 1e8)
    std::cerr << "[miplib] warning: constraint terms differ more than 1e8 times - expect numerical"
                 " issues!: | " << constr.expr() << " | in [" << new_low << ", " << new_high << "]\n";

  double const factor = friendly_power_of_two(1 / divisor);
  if (constr.type() == Constr::Equal) return factor * constr.expr() == 0;
  return factor * constr.expr() <= 0;
}

}  // namespace detail
}  // namespace miplib
