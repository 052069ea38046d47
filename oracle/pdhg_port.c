/* pdhg_port.c - plain C (OpenMP) restatement of the two fused per-iteration kernels and the
 * SpMV, used (a) as a second, independent check of the NumPy oracle and (b) as the multi-core
 * CPU baseline that bench.py times beside the GPU ("cpu_baseline.kind = port").
 *
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product path never does.
 *
 * What it restates: the reference's solve is one opaque third-party call (SCIPsolve, reference
 * miplib/scip/solver.cpp:370; ::solve(lprec*), miplib/lpsolve/solver.cpp:156) with no first-party
 * arithmetic, so this follows the published method the new backend implements (restarted
 * reflected-Halpern PDHG), operation for operation like oracle/pdhg_ref.py row_step / col_step.
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC -o oracle/_build/libpdhg_port.so oracle/pdhg_port.c -lm
 */
#include <math.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

int port_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1: the baseline asks for every host thread explicitly */
void port_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* NUMA first touch: the destination arrays come from np.empty (untouched pages); every thread
 * writes the rows / entries it will later stream under the same static schedule, so on a
 * multi-socket host each page lands on the socket that uses it. */
void port_touch_copy_csr(int64_t nrows, const int64_t* ptr, const int32_t* idx, const double* val,
                         int64_t* ptr_out, int32_t* idx_out, double* val_out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nrows; ++i) {
    ptr_out[i] = ptr[i];
    if (i == nrows - 1) ptr_out[nrows] = ptr[nrows];
    for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k) { idx_out[k] = idx[k]; val_out[k] = val[k]; }
  }
}

void port_touch_copy_vec(int64_t n, const double* src, double* dst) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) dst[i] = src[i];
}

void port_spmv(int64_t nrows, const int64_t* ptr, const int32_t* idx, const double* val,
               const double* v, double* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nrows; ++i) {
    double acc = 0.0;
    for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k) acc += val[k] * v[idx[k]];
    out[i] = acc;
  }
}

/* kx = K xbar ; yhat = sigma (clamp(t, lo, hi) - t), t = kx - y/sigma ;
 * y <- a (2 yhat - y) + (1 - a) y0.   Returns sum (yhat - y_old)^2. */
double port_row_step(int64_t m, const int64_t* ptr, const int32_t* idx, const double* val,
                     const double* xbar, const double* lo, const double* hi, const double* y0,
                     double* y, double* yhat, double sigma, double a) {
  double dy2 = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : dy2)
  for (int64_t i = 0; i < m; ++i) {
    double kx = 0.0;
    for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k) kx += val[k] * xbar[idx[k]];
    const double yi = y[i];
    const double t = kx - yi / sigma;
    const double yh = sigma * (clampd(t, lo[i], hi[i]) - t);
    yhat[i] = yh;
    y[i] = a * (2.0 * yh - yi) + (1.0 - a) * y0[i];
    dy2 += (yh - yi) * (yh - yi);
  }
  return dy2;
}

/* atyh = K' yhat ; x <- a xbar + (1-a) x0 ; aty <- a (2 atyh - aty) + (1-a) aty0 ;
 * xhat = clamp(x - tau (c - aty), l, u) ; xbar <- 2 xhat - x.  xhat_out may be NULL. */
void port_col_step(int64_t n, const int64_t* ptr, const int32_t* idx, const double* val,
                   const double* yhat, const double* x0, const double* aty0, const double* c,
                   const double* l, const double* u, double* xbar, double* aty, double* xhat_out,
                   double tau, double a) {
#pragma omp parallel for schedule(static)
  for (int64_t j = 0; j < n; ++j) {
    double atyh = 0.0;
    for (int64_t k = ptr[j]; k < ptr[j + 1]; ++k) atyh += val[k] * yhat[idx[k]];
    const double xn = a * xbar[j] + (1.0 - a) * x0[j];
    const double atn = a * (2.0 * atyh - aty[j]) + (1.0 - a) * aty0[j];
    const double xh = clampd(xn - tau * (c[j] - atn), l[j], u[j]);
    xbar[j] = 2.0 * xh - xn;
    aty[j] = atn;
    if (xhat_out) xhat_out[j] = xh;
  }
}
