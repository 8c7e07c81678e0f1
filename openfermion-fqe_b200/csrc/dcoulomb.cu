// Diagonal-Coulomb apply and evolve: one fused streaming pass over C[a,b].
//
// Replaces zdiagonal_coulomb_apply / zdiagonal_coulomb (reference
// src/fqe/lib/fqe_data.c:455-524 / 526-602), called from
// FqeData.apply_diagonal_coulomb / evolve_diagonal_coulomb (fqe_data.py:263-402).
//
//   apply :  C[a,b] *= ( sum_{j in b} S_a[j] + B[b] + A[a] ),  S_a[j] = sum_{i in a} v[i,j]+v[j,i]
//   evolve:  C[a,b] *= ( prod_{j in b} P_a[j] )^2 * B[b] * A[a],  P_a[j] = prod_{i in a} e^{v[i,j]}
// with A/B the same-spin terms (sum resp. product over diag[i] and v[i,j], i,j in the
// string).  The two alpha-beta conventions differ on purpose; each
// mirrors the reference so that non-symmetric v reproduces its goldens.
//
// HBM roofline: 32 bytes per determinant (read + write C); strings and per-string
// terms (24 B per beta string) are re-read per alpha row from L2.  One CTA owns an
// alpha row at a time (grid-strided), threads sweep the beta index with 16-byte
// coalesced accesses; the norb-long S_a / P_a vector lives in shared memory.
#include "fqeb_common.cuh"

namespace fqeb {

__device__ __forceinline__ double2 zmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 zadd(double2 a, double2 b) {
  return make_double2(a.x + b.x, a.y + b.y);
}

__global__ void k_cexp(int n, const double2 *__restrict__ in, double2 *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 z = in[i];
  double s, c;
  sincos(z.y, &s, &c);
  const double m = exp(z.x);
  out[i] = make_double2(m * c, m * s);
}

// same-spin term of every string (lib/fqe_data.c:408-453)
template <bool EVOLVE>
__global__ void k_dc_string_terms(int norb, int64_t len, const uint64_t *__restrict__ str,
                                  const double2 *__restrict__ diag,
                                  const double2 *__restrict__ v, double2 *__restrict__ out) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= len) return;
  const uint64_t s = str[x];
  double2 acc = EVOLVE ? make_double2(1.0, 0.0) : make_double2(0.0, 0.0);
  uint64_t si = s;
  while (si) {
    const int i = __ffsll((long long)si) - 1;
    si &= si - 1;
    acc = EVOLVE ? zmul(acc, diag[i]) : zadd(acc, diag[i]);
    uint64_t sj = s;
    while (sj) {
      const int j = __ffsll((long long)sj) - 1;
      sj &= sj - 1;
      acc = EVOLVE ? zmul(acc, v[i * norb + j]) : zadd(acc, v[i * norb + j]);
    }
  }
  out[x] = acc;
}

template <bool EVOLVE>
__global__ void __launch_bounds__(256)
k_dc_main(int norb, int64_t lena, int64_t lenb, const uint64_t *__restrict__ astr,
          const uint64_t *__restrict__ bstr, const double2 *__restrict__ v,
          const double2 *__restrict__ aterm, const double2 *__restrict__ bterm,
          double2 *__restrict__ coeff) {
  __shared__ double2 cross[64];
  for (int64_t a = blockIdx.x; a < lena; a += gridDim.x) {
    const uint64_t sa = astr[a];
    if (threadIdx.x < norb) {
      const int j = threadIdx.x;
      double2 acc = EVOLVE ? make_double2(1.0, 0.0) : make_double2(0.0, 0.0);
      uint64_t si = sa;
      while (si) {
        const int i = __ffsll((long long)si) - 1;
        si &= si - 1;
        if (EVOLVE) {
          acc = zmul(acc, v[i * norb + j]);
        } else {
          acc = zadd(acc, v[i * norb + j]);
          acc = zadd(acc, v[j * norb + i]);
        }
      }
      cross[j] = acc;
    }
    __syncthreads();
    const double2 at = aterm[a];
    double2 *row = coeff + a * lenb;
    for (int64_t b = threadIdx.x; b < lenb; b += blockDim.x) {
      uint64_t sb = bstr[b];
      double2 x = EVOLVE ? make_double2(1.0, 0.0) : make_double2(0.0, 0.0);
      while (sb) {
        const int j = __ffsll((long long)sb) - 1;
        sb &= sb - 1;
        x = EVOLVE ? zmul(x, cross[j]) : zadd(x, cross[j]);
      }
      double2 f;
      if (EVOLVE) {
        f = zmul(zmul(zmul(x, x), bterm[b]), at);
      } else {
        f = zadd(zadd(x, bterm[b]), at);
      }
      row[b] = zmul(row[b], f);
    }
    __syncthreads();
  }
}

template <bool EVOLVE>
static int dc_run(const fqeb_graph *g, const double *h_diag, const double *h_array,
                  double *d_coeff, cudaStream_t st) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(g && h_diag && h_array && d_coeff, "fqeb_dc: NULL argument");
  const int norb = g->norb;
  if (norb == 0) return FQEB_OK;
  const int nd = norb, nv = norb * norb;
  double2 *d_diag = (double2 *)g->d_small;
  double2 *d_v = d_diag + 64;
  double2 *d_diag_e = d_v + 64 * 64;
  double2 *d_v_e = d_diag_e + 64;
  FQEB_CUDA(cudaMemcpyAsync(d_diag, h_diag, sizeof(double2) * nd, cudaMemcpyHostToDevice, st));
  FQEB_CUDA(cudaMemcpyAsync(d_v, h_array, sizeof(double2) * nv, cudaMemcpyHostToDevice, st));
  const double2 *use_diag = d_diag, *use_v = d_v;
  if (EVOLVE) {
    k_cexp<<<1, 64, 0, st>>>(nd, d_diag, d_diag_e);
    FQEB_CHECK_LAUNCH();
    k_cexp<<<(nv + 255) / 256, 256, 0, st>>>(nv, d_v, d_v_e);
    FQEB_CHECK_LAUNCH();
    use_diag = d_diag_e;
    use_v = d_v_e;
  }
  const int nspin = g->shared_spin ? 1 : 2;
  for (int s = 0; s < nspin; ++s) {
    const unsigned blocks = (unsigned)((g->len[s] + 255) / 256);
    k_dc_string_terms<EVOLVE><<<blocks, 256, 0, st>>>(norb, g->len[s], g->d_str[s], use_diag,
                                                      use_v, (double2 *)g->d_sterm[s]);
    FQEB_CHECK_LAUNCH();
  }
  const double2 *aterm = (const double2 *)g->d_sterm[0];
  const double2 *bterm = (const double2 *)g->d_sterm[g->shared_spin ? 0 : 1];
  int64_t grid = (int64_t)sm_count() * 8;
  if (grid > g->len[0]) grid = g->len[0];
  k_dc_main<EVOLVE><<<(unsigned)grid, 256, 0, st>>>(norb, g->len[0], g->len[1], g->d_str[0],
                                                    g->d_str[1], use_v, aterm, bterm,
                                                    (double2 *)d_coeff);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

}  // namespace fqeb

extern "C" int fqeb_dc_apply(const fqeb_graph *g, const double *h_diag, const double *h_array,
                             double *d_coeff, void *stream) {
  return fqeb::dc_run<false>(g, h_diag, h_array, d_coeff, (cudaStream_t)stream);
}

extern "C" int fqeb_dc_evolve(const fqeb_graph *g, const double *h_diag, const double *h_array,
                              double *d_coeff, void *stream) {
  return fqeb::dc_run<true>(g, h_diag, h_array, d_coeff, (cudaStream_t)stream);
}
