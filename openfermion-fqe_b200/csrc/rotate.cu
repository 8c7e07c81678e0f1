// Orbital rotation by column operators, and one-body diagonal apply / evolve.
//
// Replace lm_apply_array1_column_alpha (reference src/fqe/lib/fqe_data.c:305-347) as driven by
// FqeData._apply_columns_recursive_alpha / apply_columns_recursive_inplace
// (fqe_data.py:1476-1535), and apply_diagonal_inplace / evolve_diagonal_inplace
// (lib/fqe_data.c:1320-1383; fqe_data.py:153-261).  Together they are
// Wavefunction.transform and the quadratic branch of Wavefunction.time_evolve
// (wavefunction.py:813-959, 1013-1034): SURVEY 8f rank 1, the orbital-rotation half of a
// double-factorised Trotter step.
//
// One column operator is  C <- (1 + sum_i m[i] a^+_i a_icol) C  on one spin, in place:
//   targets  (strings with icol empty)    C[t] += sum_{i in t} sign * m[i] * C[s_i],
//                                         s_i = t with the electron moved from i to icol
//   sources  (strings with icol occupied) C[s] *= 1 + m[icol]
// The two sets are disjoint, so the accumulation pass reads only rows / columns that the
// scaling pass (launched after it) has not touched yet: two launches per column, each over a
// precomputed ordered list of its strings.  The reference transposes C to reuse its alpha
// routine for beta; here the beta pass works on columns directly (in-row gathers).
//
// Roofline: HBM-bound.  Per column the accumulation reads (nele + 1) and writes 1 of half the
// rows of C and the scaling reads and writes the other half: 16 * L^2 * (nele/2 + 2) bytes.
#include "fqeb_common.cuh"

#include <vector>

namespace fqeb {

constexpr int kRB = 256;

// ordered list of the strings whose bit `icol` equals `want`, for every icol (one CTA each)
__global__ void k_occupancy_lists(int norb, int64_t len, const uint64_t *__restrict__ str,
                                  int want, int64_t cap, int32_t *__restrict__ list) {
  const int icol = blockIdx.x;
  __shared__ int s_wcount[kRB / 32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int64_t x0 = 0; x0 < len; x0 += kRB) {
    const int64_t x = x0 + threadIdx.x;
    const bool hit = x < len && (int)((str[x] >> icol) & 1ull) == want;
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_wcount[warp] = __popc(ballot);
    __syncthreads();
    int offset = s_base;
    for (int w = 0; w < warp; ++w) offset += s_wcount[w];
    if (hit) list[icol * cap + offset + __popc(ballot & ((1u << lane) - 1u))] = (int32_t)x;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < kRB / 32; ++w) tot += s_wcount[w];
      s_base += tot;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// alpha accumulation: one CTA = one target row x 256 beta strings
__global__ void __launch_bounds__(kRB)
k_column_alpha_acc(int norb, int icol, int64_t lenb, const int32_t *__restrict__ targets,
                   const int32_t *__restrict__ amapT, const double2 *__restrict__ col, int nbt,
                   double2 *__restrict__ coeff) {
  __shared__ int s_src[64];
  __shared__ double2 s_fac[64];
  __shared__ int s_n;
  const int64_t tile = blockIdx.x;
  const int64_t t = targets[tile / nbt];
  if (threadIdx.x == 0) {
    // sources of target t: amap[i*norb+icol][t] = sign*(s+1) with a^+_icol a_i |t> = sign |s>
    int n = 0;
    for (int i = 0; i < norb; ++i) {
      const int e = amapT[t * (int64_t)(norb * norb) + i * norb + icol];
      if (e != 0) {
        const double2 m = col[i];
        s_src[n] = abs(e) - 1;
        s_fac[n] = e > 0 ? m : make_double2(-m.x, -m.y);
        ++n;
      }
    }
    s_n = n;
  }
  __syncthreads();
  const int64_t b = (tile % nbt) * kRB + threadIdx.x;
  if (b >= lenb) return;
  double2 acc = coeff[t * lenb + b];
  const int n = s_n;
  for (int k = 0; k < n; ++k) {
    const double2 v = cmul(s_fac[k], coeff[(int64_t)s_src[k] * lenb + b]);
    acc.x += v.x;
    acc.y += v.y;
  }
  coeff[t * lenb + b] = acc;
}

// alpha scaling: rows with icol occupied
__global__ void __launch_bounds__(kRB)
k_column_alpha_scale(int64_t lenb, const int32_t *__restrict__ rows, double2 fac, int nbt,
                     double2 *__restrict__ coeff) {
  const int64_t tile = blockIdx.x;
  const int64_t r = rows[tile / nbt];
  const int64_t b = (tile % nbt) * kRB + threadIdx.x;
  if (b >= lenb) return;
  coeff[r * lenb + b] = cmul(fac, coeff[r * lenb + b]);
}

// beta accumulation / scaling: one CTA = kRows alpha rows x 256 listed beta strings
constexpr int kRows = 8;
template <bool ACC>
__global__ void __launch_bounds__(kRB)
k_column_beta(int norb, int icol, int64_t lena, int64_t lenb, const int32_t *__restrict__ cols,
              int64_t ncols, const int32_t *__restrict__ amap, const double2 *__restrict__ col,
              int nct, double2 *__restrict__ coeff) {
  const int64_t tile = blockIdx.x;
  const int64_t k = (tile % nct) * kRB + threadIdx.x;
  if (k >= ncols) return;
  const int64_t b = cols[k];
  const int64_t a0 = (tile / nct) * kRows;
  const int64_t a1 = a0 + kRows < lena ? a0 + kRows : lena;
  if constexpr (!ACC) {
    const double2 m = col[icol];
    const double2 fac = make_double2(1.0 + m.x, m.y);
    for (int64_t a = a0; a < a1; ++a) coeff[a * lenb + b] = cmul(fac, coeff[a * lenb + b]);
  } else {
  double2 acc[kRows];
#pragma unroll
  for (int r = 0; r < kRows; ++r)
    acc[r] = (a0 + r < a1) ? coeff[(a0 + r) * lenb + b] : make_double2(0.0, 0.0);
  for (int i = 0; i < norb; ++i) {
    const int e = amap[(int64_t)(i * norb + icol) * lenb + b];
    if (e == 0) continue;
    const double2 m = col[i];
    const double2 fac = e > 0 ? m : make_double2(-m.x, -m.y);
    const int64_t s = abs(e) - 1;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      if (a0 + r < a1) {
        const double2 v = cmul(fac, coeff[(a0 + r) * lenb + s]);
        acc[r].x += v.x;
        acc[r].y += v.y;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r)
    if (a0 + r < a1) coeff[(a0 + r) * lenb + b] = acc[r];
  }
}

// per-string one-body diagonal term: sum of arr over the occupied orbitals (EXP: its exp)
template <bool EXP>
__global__ void k_string_diag(int norb, int64_t len, const uint64_t *__restrict__ str,
                              const double2 *__restrict__ arr, double2 *__restrict__ out) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= len) return;
  const uint64_t s = str[x];
  double2 acc = make_double2(0.0, 0.0);
  for (int i = 0; i < norb; ++i)
    if ((s >> i) & 1ull) {
      acc.x += arr[i].x;
      acc.y += arr[i].y;
    }
  if (EXP) {
    double sn, cs;
    sincos(acc.y, &sn, &cs);
    const double m = exp(acc.x);
    acc = make_double2(m * cs, m * sn);
  }
  out[x] = acc;
}

// C[a,b] *= fa[a] * fb[b]  (PROD)   or   C[a,b] *= fa[a] + fb[b]
template <bool PROD>
__global__ void __launch_bounds__(kRB)
k_outer_scale(int64_t lenb, const double2 *__restrict__ fa, const double2 *__restrict__ fb, int nbt,
              double2 *__restrict__ coeff) {
  const int64_t tile = blockIdx.x;
  const int64_t a = tile / nbt;
  const int64_t b = (tile % nbt) * kRB + threadIdx.x;
  if (b >= lenb) return;
  const double2 x = fa[a], y = fb[b];
  const double2 f = PROD ? cmul(x, y) : make_double2(x.x + y.x, x.y + y.y);
  coeff[a * lenb + b] = cmul(f, coeff[a * lenb + b]);
}

static int64_t binom_i64(int n, int k) {
  if (k < 0 || k > n) return 0;
  int64_t r = 1;
  for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i;
  return r;
}

// occupancy lists of one spin, built on first use:
// d_occ[spin][icol][.] strings with icol occupied, d_unocc[spin][icol][.] with icol empty
static int ensure_lists_locked(const fqeb_graph *cg, int spin);
static int ensure_lists(const fqeb_graph *cg, int spin) {
  GraphLock lock(cg);   // lazily built, shared by every sector of this graph
  return ensure_lists_locked(cg, spin);
}
static int ensure_lists_locked(const fqeb_graph *cg, int spin) {
  fqeb_graph *g = const_cast<fqeb_graph *>(cg);
  if (g->d_occ[spin]) return FQEB_OK;
  if (spin == 1 && g->shared_spin) {
    int rc = ensure_lists_locked(cg, 0);
    if (rc != FQEB_OK) return rc;
    g->d_occ[1] = g->d_occ[0];
    g->d_unocc[1] = g->d_unocc[0];
    return FQEB_OK;
  }
  const int norb = g->norb, nele = g->nele[spin];
  const int64_t len = g->len[spin];
  const int64_t nocc = binom_i64(norb - 1, nele - 1), nun = binom_i64(norb - 1, nele);
  int32_t *occ = nullptr, *un = nullptr;
  FQEB_CUDA(cudaMalloc(&occ, sizeof(int32_t) * (size_t)norb * (nocc > 0 ? nocc : 1)));
  FQEB_CUDA(cudaMalloc(&un, sizeof(int32_t) * (size_t)norb * (nun > 0 ? nun : 1)));
  k_occupancy_lists<<<norb, kRB>>>(norb, len, g->d_str[spin], 1, nocc, occ);
  FQEB_CHECK_LAUNCH();
  k_occupancy_lists<<<norb, kRB>>>(norb, len, g->d_str[spin], 0, nun, un);
  FQEB_CHECK_LAUNCH();
  FQEB_CUDA(cudaDeviceSynchronize());
  g->d_occ[spin] = occ;
  g->d_unocc[spin] = un;
  return FQEB_OK;
}

}  // namespace fqeb

using namespace fqeb;

extern "C" int fqeb_apply_columns(const fqeb_graph *g, int spin, const double *h_mat,
                                  double *d_coeff, void *stream) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(g && h_mat && d_coeff, "fqeb_apply_columns: NULL argument");
  FQEB_REQUIRE(spin == 0 || spin == 1, "fqeb_apply_columns: spin must be 0 (alpha) or 1 (beta)");
  const int norb = g->norb, nele = g->nele[spin];
  const int64_t lena = g->len[0], lenb = g->len[1];
  if (norb == 0 || nele == 0) return FQEB_OK;  // no electron of this spin: identity
  FQEB_REQUIRE(sizeof(double) * 2 * (size_t)norb * norb <= g->small_bytes,
               "fqeb_apply_columns: norb too large for the operator scratch");
  rc = ensure_lists(g, spin);
  if (rc != FQEB_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  GraphScratch sc;
  rc = graph_scratch(g, st, &sc);
  if (rc != FQEB_OK) return rc;
  // device copy of the matrix, transposed so that column icol is contiguous
  std::vector<double> mt(2 * (size_t)norb * norb);
  for (int i = 0; i < norb; ++i)
    for (int c = 0; c < norb; ++c) {
      mt[2 * ((size_t)c * norb + i)] = h_mat[2 * ((size_t)i * norb + c)];
      mt[2 * ((size_t)c * norb + i) + 1] = h_mat[2 * ((size_t)i * norb + c) + 1];
    }
  FQEB_CUDA(cudaMemcpyAsync(sc.small, mt.data(), sizeof(double) * mt.size(),
                            cudaMemcpyHostToDevice, st));
  FQEB_CUDA(cudaStreamSynchronize(st));  // mt is a stack-lifetime staging buffer
  const double2 *d_mt = (const double2 *)sc.small;
  double2 *c = (double2 *)d_coeff;
  const int64_t nocc = binom_i64(norb - 1, nele - 1), nun = binom_i64(norb - 1, nele);
  const int nbt = (int)((lenb + kRB - 1) / kRB);
  for (int icol = 0; icol < norb; ++icol) {
    const double2 *col = d_mt + (size_t)icol * norb;
    const double2 fac = make_double2(1.0 + mt[2 * ((size_t)icol * norb + icol)],
                                     mt[2 * ((size_t)icol * norb + icol) + 1]);
    if (spin == 0) {
      if (nun > 0) {
        FQEB_REQUIRE(nun * nbt < (1ll << 31), "fqeb_apply_columns: problem too large");
        k_column_alpha_acc<<<(unsigned)(nun * nbt), kRB, 0, st>>>(
            norb, icol, lenb, g->d_unocc[0] + (size_t)icol * nun, g->d_amapT[0], col, nbt, c);
        FQEB_CHECK_LAUNCH();
      }
      if (nocc > 0) {
        k_column_alpha_scale<<<(unsigned)(nocc * nbt), kRB, 0, st>>>(
            lenb, g->d_occ[0] + (size_t)icol * nocc, fac, nbt, c);
        FQEB_CHECK_LAUNCH();
      }
    } else {
      const int64_t nrt = (lena + kRows - 1) / kRows;
      if (nun > 0) {
        const int nct = (int)((nun + kRB - 1) / kRB);
        FQEB_REQUIRE(nrt * nct < (1ll << 31), "fqeb_apply_columns: problem too large");
        k_column_beta<true><<<(unsigned)(nrt * nct), kRB, 0, st>>>(
            norb, icol, lena, lenb, g->d_unocc[1] + (size_t)icol * nun, nun, g->d_amap[1], col,
            nct, c);
        FQEB_CHECK_LAUNCH();
      }
      if (nocc > 0) {
        const int nct = (int)((nocc + kRB - 1) / kRB);
        k_column_beta<false><<<(unsigned)(nrt * nct), kRB, 0, st>>>(
            norb, icol, lena, lenb, g->d_occ[1] + (size_t)icol * nocc, nocc, g->d_amap[1], col,
            nct, c);
        FQEB_CHECK_LAUNCH();
      }
    }
  }
  return FQEB_OK;
}

static int diagonal_common(const fqeb_graph *g, const double *h_aarray, const double *h_barray,
                           double *d_coeff, bool evolve, cudaStream_t st) {
  const int norb = g->norb;
  const int64_t lena = g->len[0], lenb = g->len[1];
  FQEB_REQUIRE(sizeof(double) * 4 * (size_t)norb <= g->small_bytes,
               "diagonal: norb too large for the operator scratch");
  if (norb == 0) return FQEB_OK;
  std::vector<double> both(4 * (size_t)norb);
  for (int i = 0; i < 2 * norb; ++i) {
    both[i] = h_aarray[i];
    both[2 * norb + i] = h_barray[i];
  }
  GraphScratch sc;
  int rc = graph_scratch(g, st, &sc);
  if (rc != FQEB_OK) return rc;
  FQEB_CUDA(cudaMemcpyAsync(sc.small, both.data(), sizeof(double) * both.size(),
                            cudaMemcpyHostToDevice, st));
  FQEB_CUDA(cudaStreamSynchronize(st));
  const double2 *da = (const double2 *)sc.small, *db = da + norb;
  double2 *fa = (double2 *)sc.sterm[0], *fb = (double2 *)sc.sterm[1];
  const int threads = 256;
  if (evolve) {
    k_string_diag<true><<<(unsigned)((lena + threads - 1) / threads), threads, 0, st>>>(
        norb, lena, g->d_str[0], da, fa);
    k_string_diag<true><<<(unsigned)((lenb + threads - 1) / threads), threads, 0, st>>>(
        norb, lenb, g->d_str[1], db, fb);
  } else {
    k_string_diag<false><<<(unsigned)((lena + threads - 1) / threads), threads, 0, st>>>(
        norb, lena, g->d_str[0], da, fa);
    k_string_diag<false><<<(unsigned)((lenb + threads - 1) / threads), threads, 0, st>>>(
        norb, lenb, g->d_str[1], db, fb);
  }
  FQEB_CHECK_LAUNCH();
  const int nbt = (int)((lenb + kRB - 1) / kRB);
  FQEB_REQUIRE(lena * nbt < (1ll << 31), "diagonal: problem too large for one launch");
  if (evolve)
    k_outer_scale<true><<<(unsigned)(lena * nbt), kRB, 0, st>>>(lenb, fa, fb, nbt,
                                                                (double2 *)d_coeff);
  else
    k_outer_scale<false><<<(unsigned)(lena * nbt), kRB, 0, st>>>(lenb, fa, fb, nbt,
                                                                 (double2 *)d_coeff);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

extern "C" int fqeb_apply_diagonal(const fqeb_graph *g, const double *h_aarray,
                                   const double *h_barray, double *d_coeff, void *stream) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(g && h_aarray && h_barray && d_coeff, "fqeb_apply_diagonal: NULL argument");
  return diagonal_common(g, h_aarray, h_barray, d_coeff, false, (cudaStream_t)stream);
}

extern "C" int fqeb_evolve_diagonal(const fqeb_graph *g, const double *h_aarray,
                                    const double *h_barray, double *d_coeff, void *stream) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(g && h_aarray && h_barray && d_coeff, "fqeb_evolve_diagonal: NULL argument");
  return diagonal_common(g, h_aarray, h_barray, d_coeff, true, (cudaStream_t)stream);
}
