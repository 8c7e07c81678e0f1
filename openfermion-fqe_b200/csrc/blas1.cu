// BLAS-1 on coefficient vectors: the HBM-bound glue of the Taylor / Chebyshev
// recurrences.  Replaces FqeData.ax_plus_y / scale / norm and util.vdot
// (reference src/fqe/fqe_data.py:2620-2632, 2745-2751, 2701-2707; util.py:506-530),
// which are numpy temporaries on the CPU.  Each kernel streams complex128 as
// double2 (16-byte accesses, fully coalesced), grid-strided over a grid of
// 8 CTAs per SM.  Reductions are two-pass and deterministic.
#include "fqeb_common.cuh"

namespace fqeb {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 148 * 8 * 2;  // scratch sizing bound

static inline int grid_for(int64_t n) {
  int64_t want = (n + kThreads - 1) / kThreads;
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__global__ void k_zaxpy(int64_t n, double2 a, const double2 *__restrict__ x,
                        double2 *__restrict__ y) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double2 xv = x[i];
    double2 yv = y[i];
    yv.x += a.x * xv.x - a.y * xv.y;
    yv.y += a.x * xv.y + a.y * xv.x;
    y[i] = yv;
  }
}

__global__ void k_zaxpby(int64_t n, double2 a, const double2 *__restrict__ x, double2 b,
                         double2 *__restrict__ y) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double2 ax = cmul(a, x[i]);
    const double2 by = cmul(b, y[i]);
    y[i] = make_double2(ax.x + by.x, ax.y + by.y);
  }
}

__global__ void k_zscal(int64_t n, double2 a, double2 *__restrict__ x) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    x[i] = cmul(a, x[i]);
  }
}

// block-level sum of a double2 via warp shuffles; result valid in thread 0
__device__ __forceinline__ double2 block_sum(double2 v) {
  __shared__ double2 warp_part[kThreads / 32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    v.x += __shfl_down_sync(0xffffffffu, v.x, off);
    v.y += __shfl_down_sync(0xffffffffu, v.y, off);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_part[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < kThreads / 32 ? warp_part[lane] : make_double2(0.0, 0.0);
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
      v.x += __shfl_down_sync(0xffffffffu, v.x, off);
      v.y += __shfl_down_sync(0xffffffffu, v.y, off);
    }
  }
  return v;
}

// MODE 0: sum |x|^2 ; MODE 1: sum conj(x)*y ; MODE 2: y += a*x and sum |x|^2
template <int MODE>
__global__ void k_reduce_pass1(int64_t n, double2 a, const double2 *__restrict__ x,
                               double2 *__restrict__ y, double2 *__restrict__ partial) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double2 acc = make_double2(0.0, 0.0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double2 xv = x[i];
    if (MODE == 0) {
      acc.x += xv.x * xv.x + xv.y * xv.y;
    } else if (MODE == 1) {
      const double2 yv = y[i];
      acc.x += xv.x * yv.x + xv.y * yv.y;
      acc.y += xv.x * yv.y - xv.y * yv.x;
    } else {
      double2 yv = y[i];
      yv.x += a.x * xv.x - a.y * xv.y;
      yv.y += a.x * xv.y + a.y * xv.x;
      y[i] = yv;
      acc.x += xv.x * xv.x + xv.y * xv.y;
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void k_reduce_pass2(int nparts, const double2 *__restrict__ partial,
                               double2 *__restrict__ out) {
  double2 acc = make_double2(0.0, 0.0);
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    acc.x += partial[i].x;
    acc.y += partial[i].y;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) out[0] = acc;
}

template <int MODE>
static int reduce(int64_t n, double2 a, const double *d_x, double *d_y, void *d_scratch,
                  double *h_out, int nout, cudaStream_t st) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(n >= 0 && d_scratch && h_out, "reduce: bad argument");
  double2 *partial = (double2 *)d_scratch;
  double2 *result = partial + kMaxBlocks;
  const int grid = grid_for(n);
  k_reduce_pass1<MODE><<<grid, kThreads, 0, st>>>(n, a, (const double2 *)d_x, (double2 *)d_y,
                                                  partial);
  FQEB_CHECK_LAUNCH();
  k_reduce_pass2<<<1, kThreads, 0, st>>>(grid, partial, result);
  FQEB_CHECK_LAUNCH();
  double2 host;
  FQEB_CUDA(cudaMemcpyAsync(&host, result, sizeof(double2), cudaMemcpyDeviceToHost, st));
  FQEB_CUDA(cudaStreamSynchronize(st));
  h_out[0] = host.x;
  if (nout > 1) h_out[1] = host.y;
  return FQEB_OK;
}

}  // namespace fqeb

using namespace fqeb;

extern "C" size_t fqeb_reduce_scratch_bytes(void) { return sizeof(double2) * (kMaxBlocks + 8); }

extern "C" int fqeb_zaxpy(int64_t n, double ar, double ai, const double *d_x, double *d_y,
                          void *stream) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(n >= 0 && d_x && d_y, "fqeb_zaxpy: bad argument");
  if (n == 0) return FQEB_OK;
  k_zaxpy<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(n, make_double2(ar, ai),
                                                              (const double2 *)d_x, (double2 *)d_y);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

extern "C" int fqeb_zaxpby(int64_t n, double ar, double ai, const double *d_x, double br,
                           double bi, double *d_y, void *stream) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(n >= 0 && d_x && d_y, "fqeb_zaxpby: bad argument");
  if (n == 0) return FQEB_OK;
  k_zaxpby<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(
      n, make_double2(ar, ai), (const double2 *)d_x, make_double2(br, bi), (double2 *)d_y);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

extern "C" int fqeb_zscal(int64_t n, double ar, double ai, double *d_x, void *stream) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(n >= 0 && d_x, "fqeb_zscal: bad argument");
  if (n == 0) return FQEB_OK;
  k_zscal<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(n, make_double2(ar, ai),
                                                              (double2 *)d_x);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

extern "C" int fqeb_znorm2(int64_t n, const double *d_x, void *d_scratch, double *h_out,
                           void *stream) {
  return reduce<0>(n, make_double2(0, 0), d_x, nullptr, d_scratch, h_out, 1, (cudaStream_t)stream);
}

extern "C" int fqeb_zdotc(int64_t n, const double *d_x, const double *d_y, void *d_scratch,
                          double *h_out, void *stream) {
  return reduce<1>(n, make_double2(0, 0), d_x, (double *)d_y, d_scratch, h_out, 2,
                   (cudaStream_t)stream);
}

extern "C" int fqeb_axpy_norm2(int64_t n, double cr, double ci, const double *d_work,
                               double *d_evol, void *d_scratch, double *h_out, void *stream) {
  return reduce<2>(n, make_double2(cr, ci), d_work, d_evol, d_scratch, h_out, 1,
                   (cudaStream_t)stream);
}
