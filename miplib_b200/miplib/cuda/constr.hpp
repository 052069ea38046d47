// miplib/cuda/constr.hpp - constraint handle of the Cuda backend: the front-end object plus the
// row it was lowered to (cf. reference miplib/lpsolve/constr.hpp, miplib/scip/constr.hpp:9-35).
#pragma once

#include <cstdint>

#include "../constr.hpp"

namespace miplib {

struct CudaConstr : detail::IConstr
{
  using detail::IConstr::IConstr;
  mutable std::int64_t m_row = -1;   // -1: not posted
};

}  // namespace miplib
