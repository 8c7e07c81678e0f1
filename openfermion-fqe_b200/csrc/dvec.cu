// Gather (make_dvec) and scatter (make_coeff) over the C[Ia, Ib] matrix.
//
// Replace zdvec_make / zcoeff_make (reference src/fqe/lib/fqe_data.c:350-406) as
// driven by FqeData._calculate_dvec_spatial_with_coeff (fqe_data.py:2209-2234) and
// FqeData._calculate_coeff_spatial_with_dvec (fqe_data.py:2309-2334).
//
// The reference walks (source, target, sign) triples and issues one zaxpy per
// triple: alpha triples move whole rows, beta triples move strided columns.  On
// the GPU both directions are written as BY-TARGET gathers driven by one signed
// "adjoint map" per spin,  amap[ij][x] = sign*(y+1)  with  a^+_j a_i |x> = sign|y>:
//
//     D[ij, a, b]  =  sgn(amap_a[ij][a]) * C[y_a, b]  +  sgn(amap_b[ij][b]) * C[a, y_b]
//     out[a, b]   +=  sum_ij sgn(amap_a[ij][a]) * E[ij, y_a, b] + sgn(amap_b[ij][b]) * E[ij, a, y_b]
//
// so every output element is produced by exactly one thread (no atomics, run-to-run
// deterministic), alpha terms are coalesced row reads, and beta terms are gathers
// inside one row that stay in L1/L2.  Writes are 16-byte, fully coalesced.
//
// Work is restricted to a chunk of alpha rows [row0, row0+nrows) and (for the
// gather) a slice of pairs [ij0, ij1); that is what lets the 678 GB norb=16 D
// tensor be streamed through a workspace, and what the multi-GPU shards select.
//
// Roofline: gather moves 16*(npair+1) bytes per determinant (write D
// once, read C once), scatter the same in the other direction: both HBM-bound.
#include "fqeb_common.cuh"

namespace fqeb {

constexpr int kTB = 256;  // beta strings per CTA

__device__ __forceinline__ void axpy_sign(double2 &acc, int t, const double2 v) {
  if (t > 0) {
    acc.x += v.x;
    acc.y += v.y;
  } else {
    acc.x -= v.x;
    acc.y -= v.y;
  }
}

// Pair list: D row c (c in [c0, c1)) is the sum of the excitation pairs
// pairs[2c], pairs[2c+1] (second = -1 if absent).  The plain operator lists every
// ij once; an operator that is symmetric under i<->j lists (ij, ji) so that the
// compressed tensor D_c[i>=j] = D[ij] + D[ji] of the reference's real-integral
// branch (fqe_data.py:2336-2353) is produced directly.
//
// WRITE_D: store D;  H1: accumulate sum_ij h1[ij]*D[ij] into sig (one-body term,
// fqe_data.py:655 `einsum("ij,ijkl->kl", h1e, dvec)`; h1 needs no symmetry).
template <bool WRITE_D, bool H1>
__global__ void __launch_bounds__(kTB)
k_make_dvec(int npair_total, int64_t lena, int64_t lenb, const int32_t *__restrict__ amapT_a,
            const int32_t *__restrict__ amap_b, const double2 *__restrict__ coeff,
            double2 *__restrict__ dvec, int64_t ldd, int64_t row0, int nbt, int c0, int c1,
            const int32_t *__restrict__ pairs, const double2 *__restrict__ h1,
            double2 *__restrict__ sig) {
  const int64_t tile = blockIdx.x;
  const int64_t r = tile / nbt;
  const int64_t b = (tile % nbt) * kTB + threadIdx.x;
  if (b >= lenb) return;
  const int64_t a = row0 + r;
  const int32_t *__restrict__ ta_row = amapT_a + a * (int64_t)npair_total;
  const double2 *__restrict__ crow = coeff + a * lenb;
  double2 *__restrict__ dout = dvec + r * lenb + b;
  double2 acc = make_double2(0.0, 0.0);
#pragma unroll 2
  for (int c = c0; c < c1; ++c) {
    double2 tot = make_double2(0.0, 0.0);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ij = pairs[2 * c + h];  // warp-uniform
      if (ij < 0) continue;
      const int ta = ta_row[ij];                         // warp-uniform
      const int tb = amap_b[(int64_t)ij * lenb + b];     // coalesced
      double2 val = make_double2(0.0, 0.0);
      if (ta != 0) axpy_sign(val, ta, coeff[(int64_t)(abs(ta) - 1) * lenb + b]);
      if (tb != 0) axpy_sign(val, tb, crow[abs(tb) - 1]);
      tot.x += val.x;
      tot.y += val.y;
      if (H1) {
        const double2 hh = h1[ij];
        acc.x += hh.x * val.x - hh.y * val.y;
        acc.y += hh.x * val.y + hh.y * val.x;
      }
    }
    if (WRITE_D) dout[(int64_t)(c - c0) * ldd] = tot;
  }
  if (H1) {
    double2 s = sig[a * lenb + b];
    s.x += acc.x;
    s.y += acc.y;
    sig[a * lenb + b] = s;
  }
}

__global__ void __launch_bounds__(kTB)
k_make_coeff(int npair, int64_t lena, int64_t lenb, const int32_t *__restrict__ amapT_a,
             const int32_t *__restrict__ amap_b, const int32_t *__restrict__ rowmap,
             const double2 *__restrict__ evec, int64_t lde, int64_t row0, int64_t nrows, int nbt,
             double2 z, double2 *__restrict__ out) {
  const int64_t tile = blockIdx.x;
  const int64_t x = tile / nbt;
  const int64_t b = (tile % nbt) * kTB + threadIdx.x;
  if (b >= lenb) return;
  const int32_t *__restrict__ ta_row = amapT_a + x * (int64_t)npair;
  double2 acc = make_double2(0.0, 0.0);
  bool touched = false;
  // alpha: rows of E that live in this chunk and map onto row x
#pragma unroll 4
  for (int kl = 0; kl < npair; ++kl) {
    const int ta = ta_row[kl];  // warp-uniform
    if (ta != 0) {
      const int64_t y = (int64_t)(abs(ta) - 1) - row0;
      if (y >= 0 && y < nrows) {
        touched = true;
        axpy_sign(acc, ta, evec[(int64_t)rowmap[kl] * lde + y * lenb + b]);
      }
    }
  }
  // beta: only rows of the chunk itself
  const int64_t xr = x - row0;
  if (xr >= 0 && xr < nrows) {
    touched = true;
    const double2 *__restrict__ erow = evec + xr * lenb;
#pragma unroll 4
    for (int kl = 0; kl < npair; ++kl) {
      const int tb = amap_b[(int64_t)kl * lenb + b];
      if (tb != 0) axpy_sign(acc, tb, erow[(int64_t)rowmap[kl] * lde + (abs(tb) - 1)]);
    }
  }
  if (touched) {
    double2 s = out[x * lenb + b];
    s.x += z.x * acc.x - z.y * acc.y;
    s.y += z.x * acc.y + z.y * acc.x;
    out[x * lenb + b] = s;
  }
}

// pairs == nullptr: identity pair list of the graph (D row c <-> pair c), np_eff = norb^2
int launch_make_dvec(const fqeb_graph *g, const double *d_coeff, double *d_dvec, int64_t ldd,
                     int64_t row0, int64_t nrows, int ij0, int ij1, const int32_t *d_pairs,
                     int np_eff, const double *d_h1, double *d_sig, cudaStream_t st) {
  const int npair_full = g->norb * g->norb;
  const int npair = d_pairs ? np_eff : npair_full;
  if (!d_pairs) d_pairs = g->d_pairs_id;
  const int64_t lena = g->len[0], lenb = g->len[1];
  FQEB_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= lena,
               "make_dvec: rows [%lld,+%lld) outside [0,%lld)", (long long)row0, (long long)nrows,
               (long long)lena);
  FQEB_REQUIRE(ij0 >= 0 && ij0 <= ij1 && ij1 <= npair, "make_dvec: pair slice [%d,%d) invalid", ij0,
               ij1);
  FQEB_REQUIRE(d_dvec == nullptr || ldd >= nrows * lenb, "make_dvec: ldd=%lld < nrows*lenb",
               (long long)ldd);
  if (nrows == 0 || ij0 == ij1) return FQEB_OK;
  const int nbt = (int)((lenb + kTB - 1) / kTB);
  const int64_t tiles = nrows * nbt;
  FQEB_REQUIRE(tiles < (1ll << 31), "make_dvec: chunk too large for one launch");
  const double2 *c = (const double2 *)d_coeff;
  double2 *d = (double2 *)d_dvec;
  const double2 *h1 = (const double2 *)d_h1;
  double2 *sig = (double2 *)d_sig;
#define FQEB_LAUNCH_DVEC(WD, HH)                                                              \
  k_make_dvec<WD, HH><<<(unsigned)tiles, kTB, 0, st>>>(npair_full, lena, lenb, g->d_amapT[0], \
                                                       g->d_amap[1], c, d, ldd, row0, nbt,    \
                                                       ij0, ij1, d_pairs, h1, sig)
  if (d && h1) FQEB_LAUNCH_DVEC(true, true);
  else if (d) FQEB_LAUNCH_DVEC(true, false);
  else if (h1) FQEB_LAUNCH_DVEC(false, true);
  else return FQEB_OK;
#undef FQEB_LAUNCH_DVEC
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

// rowmap == nullptr: identity (E row kl <-> pair kl)
int launch_make_coeff(const fqeb_graph *g, const double *d_evec, int64_t lde, int64_t row0,
                      int64_t nrows, const int32_t *d_rowmap, double zr, double zi,
                      double *d_out, cudaStream_t st) {
  if (!d_rowmap) d_rowmap = g->d_rowmap_id;
  const int npair = g->norb * g->norb;
  const int64_t lena = g->len[0], lenb = g->len[1];
  FQEB_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= lena,
               "make_coeff: rows [%lld,+%lld) outside [0,%lld)", (long long)row0,
               (long long)nrows, (long long)lena);
  FQEB_REQUIRE(lde >= nrows * lenb, "make_coeff: lde=%lld < nrows*lenb", (long long)lde);
  if (nrows == 0 || npair == 0) return FQEB_OK;
  const int nbt = (int)((lenb + kTB - 1) / kTB);
  const int64_t tiles = lena * nbt;
  FQEB_REQUIRE(tiles < (1ll << 31), "make_coeff: problem too large for one launch");
  k_make_coeff<<<(unsigned)tiles, kTB, 0, st>>>(npair, lena, lenb, g->d_amapT[0], g->d_amap[1],
                                                d_rowmap, (const double2 *)d_evec, lde, row0,
                                                nrows, nbt, make_double2(zr, zi),
                                                (double2 *)d_out);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

}  // namespace fqeb

extern "C" int fqeb_make_dvec(const fqeb_graph *g, const double *d_coeff, double *d_dvec,
                              int64_t ldd, int64_t row0, int64_t nrows, int ij0, int ij1,
                              void *stream) {
  int rc = fqeb::require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(g && d_coeff && d_dvec, "fqeb_make_dvec: NULL argument");
  return fqeb::launch_make_dvec(g, d_coeff, d_dvec, ldd, row0, nrows, ij0, ij1, nullptr, 0,
                                nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int fqeb_make_coeff(const fqeb_graph *g, const double *d_evec, int64_t lde,
                               int64_t row0, int64_t nrows, double zr, double zi, double *d_out,
                               void *stream) {
  int rc = fqeb::require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(g && d_evec && d_out, "fqeb_make_coeff: NULL argument");
  return fqeb::launch_make_coeff(g, d_evec, lde, row0, nrows, nullptr, zr, zi, d_out,
                                 (cudaStream_t)stream);
}
