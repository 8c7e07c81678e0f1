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
// coalesced accesses; the cross term comes from byte-indexed shared-memory tables.
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

// One CTA owns an alpha row at a time.  The alpha-beta cross term of element (a, b) is
// sum (apply) / product (evolve) over the occupied orbitals j of b of cross_a[j]; it is
// evaluated with byte-indexed lookup tables built once per row in shared memory
// (tab[k][m] = combination of cross_a[8k + bit] over the bits of m), so an element costs
// ceil(norb/8) shared-memory lookups instead of a loop over its nbeta set bits.
#ifndef FQEB_DC_MINB
#define FQEB_DC_MINB 4
#endif
template <bool EVOLVE>
__global__ void __launch_bounds__(256, FQEB_DC_MINB)
k_dc_main(int norb, int64_t lena, int64_t lenb, const uint64_t *__restrict__ astr,
          const uint64_t *__restrict__ bstr, const double2 *__restrict__ v,
          const double2 *__restrict__ aterm, const double2 *__restrict__ bterm,
          double2 *__restrict__ coeff) {
  extern __shared__ double2 s_tab[];  // [nbytes][256]
  __shared__ double2 cross[64];
  const int nbytes = (norb + 7) / 8;
  const double2 ident = EVOLVE ? make_double2(1.0, 0.0) : make_double2(0.0, 0.0);
  for (int64_t a = blockIdx.x; a < lena; a += gridDim.x) {
    const uint64_t sa = astr[a];
    if (threadIdx.x < 64) {
      const int j = threadIdx.x;
      double2 acc = ident;
      if (j < norb) {
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
      }
      cross[j] = acc;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nbytes * 256; e += blockDim.x) {
      const int k = e >> 8;
      unsigned m = e & 255;
      double2 acc = ident;
      while (m) {
        const int bit = __ffs((int)m) - 1;
        m &= m - 1;
        const double2 cj = cross[8 * k + bit];
        acc = EVOLVE ? zmul(acc, cj) : zadd(acc, cj);
      }
      s_tab[e] = acc;
    }
    __syncthreads();
    const double2 at = aterm[a];
    double2 *__restrict__ row = coeff + a * lenb;
    // four elements per trip: all loads first (memory-level parallelism), then math + stores
#ifndef FQEB_DC_U
#define FQEB_DC_U 4
#endif
    constexpr int U = FQEB_DC_U;
    for (int64_t b0 = threadIdx.x; b0 < lenb; b0 += (int64_t)U * blockDim.x) {
      uint64_t sb[U];
      double2 bt[U], cv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t b = b0 + (int64_t)u * blockDim.x;
        const bool on = b < lenb;
        sb[u] = on ? __ldg(bstr + b) : 0ull;
        bt[u] = on ? __ldg(bterm + b) : ident;
#ifdef FQEB_DC_STREAM
        cv[u] = on ? __ldcs(row + b) : make_double2(0.0, 0.0);
#else
        cv[u] = on ? row[b] : make_double2(0.0, 0.0);
#endif
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t b = b0 + (int64_t)u * blockDim.x;
        double2 x = ident;
        uint64_t bits = sb[u];
        for (int k = 0; k < nbytes; ++k) {
          const double2 tk = s_tab[k * 256 + (int)(bits & 255)];
          bits >>= 8;
          x = EVOLVE ? zmul(x, tk) : zadd(x, tk);
        }
        double2 f;
        if (EVOLVE) {
          f = zmul(zmul(zmul(x, x), bt[u]), at);
        } else {
          f = zadd(zadd(x, bt[u]), at);
        }
#ifdef FQEB_DC_STREAM
        if (b < lenb) __stcs(row + b, zmul(cv[u], f));
#else
        if (b < lenb) row[b] = zmul(cv[u], f);
#endif
      }
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
  cudaStream_t st_key = st;
  GraphScratch sc;
  rc = graph_scratch(g, st_key, &sc);
  if (rc != FQEB_OK) return rc;
  double2 *d_diag = (double2 *)sc.small;
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
                                                      use_v, (double2 *)sc.sterm[s]);
    FQEB_CHECK_LAUNCH();
  }
  const double2 *aterm = (const double2 *)sc.sterm[0];
  const double2 *bterm = (const double2 *)sc.sterm[g->shared_spin ? 0 : 1];
  const size_t tab_bytes = sizeof(double2) * 256 * (size_t)((norb + 7) / 8);
  // persistent grid: exactly the CTAs that are resident at once, so that every CTA walks the
  // same number of rows (8 CTAs per SM with only 3 resident left a 2/3-empty last wave)
  int per_sm = 0;
  FQEB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dc_main<EVOLVE>, 256,
                                                          tab_bytes));
  if (getenv("FQEB_DC_CTAS_PER_SM")) per_sm = atoi(getenv("FQEB_DC_CTAS_PER_SM"));
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > g->len[0]) grid = g->len[0];
  k_dc_main<EVOLVE><<<(unsigned)grid, 256, tab_bytes, st>>>(norb, g->len[0], g->len[1], g->d_str[0],
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
