/*
 * Oracle-side BLAS shim.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference C kernels (/root/reference/src/fqe/lib/fqe_data.c) receive
 * zaxpy/zscal through a `struct blasfunctions` whose pointers the reference
 * fills from scipy.linalg.cython_blas inside a Cython shim
 * (lib/_blas_helpers.h:21-36, lib/_fqe_data.pyx:83-87).  To call the compiled
 * reference kernels through plain ctypes we provide the two routines here with
 * the same Fortran-style by-reference signature.
 */
#include <complex.h>
#include <stddef.h>

typedef void (*zaxpy_func)(const int *n, const double complex *alpha,
                           const double complex *x, const int *incx,
                           double complex *y, const int *incy);
typedef void (*zscal_func)(const int *n, const double complex *alpha,
                           double complex *x, const int *incx);
struct blasfunctions {
  zaxpy_func zaxpy;
  zscal_func zscal;
};

static void shim_zaxpy(const int *n, const double complex *alpha,
                       const double complex *x, const int *incx,
                       double complex *y, const int *incy) {
  const int len = *n, ix = *incx, iy = *incy;
  const double complex a = *alpha;
  if (a == 0.0) return;
  if (ix == 1 && iy == 1) {
    for (int k = 0; k < len; ++k) y[k] += a * x[k];
  } else {
    for (int k = 0; k < len; ++k) y[(ptrdiff_t)k * iy] += a * x[(ptrdiff_t)k * ix];
  }
}

static void shim_zscal(const int *n, const double complex *alpha,
                       double complex *x, const int *incx) {
  const int len = *n, ix = *incx;
  const double complex a = *alpha;
  for (int k = 0; k < len; ++k) x[(ptrdiff_t)k * ix] *= a;
}

static struct blasfunctions g_blas = {shim_zaxpy, shim_zscal};

const struct blasfunctions *fqe_oracle_blas(void) { return &g_blas; }
