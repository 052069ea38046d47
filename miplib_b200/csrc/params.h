// params.h - algorithm constants.  oracle/pdhg_ref.py carries the same names and values.
#pragma once

namespace mipcu {

constexpr int    RUIZ_PASSES      = 10;      // inf-norm equilibration sweeps
constexpr double STEP_SAFETY      = 0.998;   // eta = STEP_SAFETY / sigma_max(K)
constexpr int    POWER_BATCH      = 10;      // power iterations between convergence tests
constexpr int    POWER_MAX        = 400;
constexpr double POWER_TOL        = 1e-6;
constexpr double BETA_SUFFICIENT  = 0.2;     // restart if fp <= 0.2 fp0
constexpr double BETA_NECESSARY   = 0.8;     // restart if fp <= 0.8 fp0 and fp stopped decreasing
constexpr double BETA_ARTIFICIAL  = 0.36;    // restart if k >= 0.36 * total iterations
constexpr double WEIGHT_SMOOTHING = 0.5;     // theta of the primal-weight update
constexpr double WEIGHT_EPS       = 1e-10;
constexpr double RAY_TOL          = 1e-8;    // certificate violation allowed relative to its value
constexpr double kSellMaxAvgRow   = 96.0;    // SELL (one lane per row) only for matrices with short rows
constexpr double INF_THRESHOLD    = 1e20;    // |bound| >= this means "no bound" (MIPCU_INF)

}  // namespace mipcu
