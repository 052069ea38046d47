// miplib/util/scale.hpp - optional per-row power-of-two scaling (reference miplib/util/scale.hpp).
#pragma once

namespace miplib {
struct Constr;
namespace detail {
Constr scale_gm(Constr const& constr, double skip_lb, double skip_ub, bool ignore_inf_var_bounds);
}
}  // namespace miplib
