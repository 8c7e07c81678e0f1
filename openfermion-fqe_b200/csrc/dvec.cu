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

// Loads as volatile asm: the compiler keeps them in program order, so a batch of
// independent loads written back-to-back really is issued back-to-back (the plain
// C++ form gets re-serialised load->use->load by the scheduler to save registers,
// which leaves these latency-bound kernels with ~2 loads in flight per thread).
__device__ __forceinline__ double2 ldg_c128(const double2 *p) {
  double2 v;
  asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ldg_c128_if(bool pred, const double2 *p) {
  double2 v = make_double2(0.0, 0.0);
  asm volatile(
      "{\n .reg .pred q;\n setp.ne.b32 q, %3, 0;\n @q ld.global.nc.v2.f64 {%0,%1}, [%2];\n}"
      : "+d"(v.x), "+d"(v.y)
      : "l"(p), "r"((int)pred));
  return v;
}
__device__ __forceinline__ int2 ldg_int2(const int2 *p) {
  int2 v;
  asm volatile("ld.global.nc.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ int ldg_int(const int *p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

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
// One CTA = one alpha row x 256 beta strings.  Everything that depends only on the
// row (pair ids and the alpha sources/signs of every pair) is resolved once per CTA
// into shared memory, so the per-element loop has a single level of dependent loads
// (beta map entry -> C element) and can be unrolled for memory-level parallelism.
//
// WRITE_D: store D;  H1: accumulate sum_ij h1[ij]*D[ij] into sig (one-body term,
// fqe_data.py:655 `einsum("ij,ijkl->kl", h1e, dvec)`; h1 needs no symmetry).
template <bool WRITE_D, bool H1>
__global__ void __launch_bounds__(kTB, 3)
k_make_dvec(int npair_total, int64_t lena, int64_t lenb, const int32_t *__restrict__ amapT_a,
            const int32_t *__restrict__ amap_b, const double2 *__restrict__ coeff,
            double2 *__restrict__ dvec, int64_t ldd, int64_t row0, int nbt, int c0, int c1,
            const int32_t *__restrict__ pairs, const double2 *__restrict__ h1,
            double2 *__restrict__ sig) {
  extern __shared__ int4 s_info[];  // (ij1, ij2, ta1, ta2) per pair-space index
  const int64_t tile = blockIdx.x;
  const int64_t r = tile / nbt;
  const int64_t a = row0 + r;
  {
    const int32_t *__restrict__ ta_row = amapT_a + a * (int64_t)npair_total;
    for (int c = c0 + threadIdx.x; c < c1; c += kTB) {
      const int ij1 = pairs[2 * c], ij2 = pairs[2 * c + 1];
      s_info[c - c0] = make_int4(ij1, ij2, ta_row[ij1], ij2 >= 0 ? ta_row[ij2] : 0);
    }
  }
  __syncthreads();
  const int64_t b = (tile % nbt) * kTB + threadIdx.x;
  if (b >= lenb) return;
  const double2 *__restrict__ crow = coeff + a * lenb;
  const double2 *__restrict__ ccol = coeff + b;
  const int32_t *__restrict__ mb = amap_b + b;
  double2 *__restrict__ dout = dvec + r * lenb + b;
  double2 acc = make_double2(0.0, 0.0);
  // Two pair-space rows per trip.  Software-pipelined: the beta-map entries of trip n+1
  // are requested before the C elements of trip n are consumed, so a trip costs one
  // global-memory round trip instead of two dependent ones.
  // (absent pairs are masked, which keeps the batch branch-free)
  auto load_tb = [&](int c, int4 &j0, int4 &j1, int &t00, int &t01, int &t10, int &t11) {
    const bool one = (c < c1), two = (c + 1 < c1);
    j0 = one ? s_info[c - c0] : make_int4(0, -1, 0, 0);
    j1 = two ? s_info[c + 1 - c0] : make_int4(0, -1, 0, 0);
    t00 = one ? ldg_int(mb + (int64_t)j0.x * lenb) : 0;
    t01 = (j0.y >= 0) ? ldg_int(mb + (int64_t)j0.y * lenb) : 0;
    t10 = two ? ldg_int(mb + (int64_t)j1.x * lenb) : 0;
    t11 = (j1.y >= 0) ? ldg_int(mb + (int64_t)j1.y * lenb) : 0;
    if (!one) j0.z = j0.w = 0;
    if (!two) j1.z = j1.w = 0;
  };
  int4 i0, i1;
  int tb00, tb01, tb10, tb11;
  load_tb(c0, i0, i1, tb00, tb01, tb10, tb11);
  for (int c = c0; c < c1; c += 2) {
    const bool two = (c + 1 < c1);
    const int ta00 = i0.z, ta01 = i0.w, ta10 = i1.z, ta11 = i1.w;
    // C elements of this trip (predicated loads)
    const double2 va00 = ldg_c128_if(ta00 != 0, ccol + (int64_t)(abs(ta00) - 1) * lenb);
    const double2 va01 = ldg_c128_if(ta01 != 0, ccol + (int64_t)(abs(ta01) - 1) * lenb);
    const double2 va10 = ldg_c128_if(ta10 != 0, ccol + (int64_t)(abs(ta10) - 1) * lenb);
    const double2 va11 = ldg_c128_if(ta11 != 0, ccol + (int64_t)(abs(ta11) - 1) * lenb);
    const double2 vb00 = ldg_c128_if(tb00 != 0, crow + (abs(tb00) - 1));
    const double2 vb01 = ldg_c128_if(tb01 != 0, crow + (abs(tb01) - 1));
    const double2 vb10 = ldg_c128_if(tb10 != 0, crow + (abs(tb10) - 1));
    const double2 vb11 = ldg_c128_if(tb11 != 0, crow + (abs(tb11) - 1));
    // map entries of the next trip
    int4 n0, n1;
    int nb00, nb01, nb10, nb11;
    load_tb(c + 2, n0, n1, nb00, nb01, nb10, nb11);
    // phase 3: signed sums (multiplying by +-1.0 is exact)
    auto sgn = [](int t) { return t < 0 ? -1.0 : 1.0; };
    const double2 d00 = make_double2(sgn(ta00) * va00.x + sgn(tb00) * vb00.x,
                                     sgn(ta00) * va00.y + sgn(tb00) * vb00.y);
    const double2 d01 = make_double2(sgn(ta01) * va01.x + sgn(tb01) * vb01.x,
                                     sgn(ta01) * va01.y + sgn(tb01) * vb01.y);
    const double2 d10 = make_double2(sgn(ta10) * va10.x + sgn(tb10) * vb10.x,
                                     sgn(ta10) * va10.y + sgn(tb10) * vb10.y);
    const double2 d11 = make_double2(sgn(ta11) * va11.x + sgn(tb11) * vb11.x,
                                     sgn(ta11) * va11.y + sgn(tb11) * vb11.y);
    if (WRITE_D) {
      // streaming stores: D is far larger than L2 and is read back only by the next kernel;
      // evict-first keeps C rows and the map tables resident instead
      __stcs(dout + (int64_t)(c - c0) * ldd, make_double2(d00.x + d01.x, d00.y + d01.y));
      if (two)
        __stcs(dout + (int64_t)(c + 1 - c0) * ldd, make_double2(d10.x + d11.x, d10.y + d11.y));
    }
    if (H1) {
      const double2 h00 = h1[i0.x];
      acc.x += h00.x * d00.x - h00.y * d00.y;
      acc.y += h00.x * d00.y + h00.y * d00.x;
      if (i0.y >= 0) {
        const double2 h01 = h1[i0.y];
        acc.x += h01.x * d01.x - h01.y * d01.y;
        acc.y += h01.x * d01.y + h01.y * d01.x;
      }
      if (two) {
        const double2 h10 = h1[i1.x];
        acc.x += h10.x * d10.x - h10.y * d10.y;
        acc.y += h10.x * d10.y + h10.y * d10.x;
        if (i1.y >= 0) {
          const double2 h11 = h1[i1.y];
          acc.x += h11.x * d11.x - h11.y * d11.y;
          acc.y += h11.x * d11.y + h11.y * d11.x;
        }
      }
    }
    i0 = n0;
    i1 = n1;
    tb00 = nb00;
    tb01 = nb01;
    tb10 = nb10;
    tb11 = nb11;
  }
  if (H1) {
    double2 s = sig[a * lenb + b];
    s.x += acc.x;
    s.y += acc.y;
    sig[a * lenb + b] = s;
  }
}

// exact sign flip: XOR the sign bit of the map entry into the double
__device__ __forceinline__ double flip_sign(double v, int t) {
  return __hiloint2double(__double2hiint(v) ^ (t & (int)0x80000000), __double2loint(v));
}

// Single-source gather: D row (c - c0) of table row c is
//     D[a, b] = sgn(ta) C[|ta|-1, b] + sgn(tb) C[a, |tb|-1],   ta = mapT_a[a][c], tb = map_b[c][b]
// with either the plain adjoint map (ntab = norb^2, one row per ij) or the merged map of
// the compressed pair space (ntab = norb(norb+1)/2, fqeb_graph::d_smap), in which
// D_c = D[ij] + D[ji] has at most one alpha and one beta source per element.
// One CTA = one alpha row x 256 beta strings; ta is resolved once per CTA into shared memory.
// U table rows per trip; the beta-map entries of trip n+1 are requested before the C elements
// of trip n are consumed (one global round trip per trip instead of two dependent ones).
template <int U>
__global__ void __launch_bounds__(kTB, 3)
k_gather(int ntab, int64_t lenb, const int32_t *__restrict__ mapT_a,
         const int32_t *__restrict__ map_b, const double2 *__restrict__ coeff,
         double2 *__restrict__ dvec, int64_t ldd, int64_t row0, int nbt, int c0, int c1,
         int group, int64_t nrows) {
  extern __shared__ int s_ta[];
  // tile order: groups of `group` consecutive alpha rows, beta-tile-major inside a group, so
  // that the CTAs in flight together cover the same beta tile of neighbouring rows (which
  // share part of their alpha sources: better L2 reuse of the C re-reads)
  const int64_t tile = blockIdx.x;
  const int64_t per_group = (int64_t)group * nbt;
  const int64_t gidx = tile / per_group;
  const int64_t rem = tile - gidx * per_group;
  const int64_t g_rows = (gidx + 1) * group <= nrows ? group : nrows - gidx * group;
  const int64_t r = gidx * group + rem % g_rows;
  const int64_t btile = rem / g_rows;
  const int64_t a = row0 + r;
  const int n = c1 - c0;
  for (int c = threadIdx.x; c < n; c += kTB) s_ta[c] = mapT_a[a * (int64_t)ntab + c0 + c];
  __syncthreads();
  const int64_t b = btile * kTB + threadIdx.x;
  if (b >= lenb) return;
  const double2 *__restrict__ crow = coeff + a * lenb;
  const double2 *__restrict__ ccol = coeff + b;
  const int32_t *__restrict__ mb = map_b + (int64_t)c0 * lenb + b;
  double2 *__restrict__ dout = dvec + r * lenb + b;
  const int nfull = n - n % U;
  int tb[U];
#pragma unroll
  for (int u = 0; u < U; ++u) tb[u] = (u < nfull) ? ldg_int(mb + (int64_t)u * lenb) : 0;
  for (int c = 0; c < nfull; c += U) {
    int ta[U];
    double2 va[U], vb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ta[u] = s_ta[c + u];
      va[u] = ldg_c128_if(ta[u] != 0, ccol + (int64_t)(abs(ta[u]) - 1) * lenb);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) vb[u] = ldg_c128_if(tb[u] != 0, crow + (abs(tb[u]) - 1));
    int nb[U];
    const bool more = (c + U < nfull);
#pragma unroll
    for (int u = 0; u < U; ++u) nb[u] = more ? ldg_int(mb + (int64_t)(c + U + u) * lenb) : 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // streaming stores: D is far larger than L2 and is read back only by the next kernel
      __stcs(dout + (int64_t)(c + u) * ldd,
             make_double2(flip_sign(va[u].x, ta[u]) + flip_sign(vb[u].x, tb[u]),
                          flip_sign(va[u].y, ta[u]) + flip_sign(vb[u].y, tb[u])));
      tb[u] = nb[u];
    }
  }
  for (int c = nfull; c < n; ++c) {
    const int ta = s_ta[c], t = mb[(int64_t)c * lenb];
    const double2 va = ldg_c128_if(ta != 0, ccol + (int64_t)(abs(ta) - 1) * lenb);
    const double2 vb = ldg_c128_if(t != 0, crow + (abs(t) - 1));
    __stcs(dout + (int64_t)c * ldd, make_double2(flip_sign(va.x, ta) + flip_sign(vb.x, t),
                                                 flip_sign(va.y, ta) + flip_sign(vb.y, t)));
  }
}

// Scatter, by target.  One CTA = one target alpha row x 256 beta strings.
//   alpha part: the lk_a non-vanishing excitations of row x are filtered ONCE per CTA
//               against the chunk's row range (ordered, deterministic compaction into
//               shared memory); the element loop then only visits real hits.
//   beta part : compact per-column excitation lists clist_b[slot][b] (coalesced) give
//               exactly lk_b gathers per element instead of norb^2 table probes.
__global__ void __launch_bounds__(kTB, 3)
k_make_coeff(int npair, int64_t lena, int64_t lenb, int lk_a, int lk_b,
             const int2 *__restrict__ clistT_a, const int2 *__restrict__ clist_b,
             const int32_t *__restrict__ rowmap, const double2 *__restrict__ evec, int64_t lde,
             int64_t pitch, int64_t row0, int64_t nrows, int nbt, int64_t x0, double2 z,
             double2 *__restrict__ out) {
  extern __shared__ int s_mem[];
  int *s_rowmap = s_mem;                                   // [npair]
  int2 *s_hits = reinterpret_cast<int2 *>(s_mem + ((npair + 1) & ~1));  // [lk_a]
  __shared__ int s_wcount[kTB / 32];
  __shared__ int s_nhit;
  const int64_t tile = blockIdx.x;
  const int64_t x = x0 + tile / nbt;   // target row
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int p = threadIdx.x; p < npair; p += kTB) s_rowmap[p] = rowmap[p];
  if (threadIdx.x == 0) s_nhit = 0;
  __syncthreads();
  // ordered compaction of the alpha hits of row x that fall into [row0, row0+nrows)
  for (int base = 0; base < lk_a; base += kTB) {
    const int slot = base + threadIdx.x;
    int2 e = make_int2(0, 0);
    bool hit = false;
    if (slot < lk_a) {
      e = clistT_a[x * (int64_t)lk_a + slot];
      const int64_t y = (int64_t)(abs(e.y) - 1) - row0;
      hit = (y >= 0 && y < nrows);
      if (hit) e = make_int2(s_rowmap[e.x], e.y > 0 ? (int)(y + 1) : -(int)(y + 1));
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_wcount[warp] = __popc(ballot);
    __syncthreads();
    int offset = s_nhit;
    for (int w = 0; w < warp; ++w) offset += s_wcount[w];
    if (hit) s_hits[offset + __popc(ballot & ((1u << lane) - 1u))] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < kTB / 32; ++w) tot += s_wcount[w];
      s_nhit += tot;
    }
    __syncthreads();
  }
  const int nhit = s_nhit;
  const int64_t xr = x - row0;
  const bool in_chunk = (xr >= 0 && xr < nrows);
  if (nhit == 0 && !in_chunk) return;
  const int64_t b = (tile % nbt) * kTB + threadIdx.x;
  if (b >= lenb) return;
  double2 acc = make_double2(0.0, 0.0);
  const double2 *__restrict__ ecol = evec + b;
  // batches: all loads of a batch are issued before any is consumed (memory-level
  // parallelism); signs are applied as exact multiplications by +-1.0
  constexpr int UA = 4, UB = 8;
  int h = 0;
  for (; h + UA <= nhit; h += UA) {
    double2 v[UA];
    double sg[UA];
#pragma unroll
    for (int u = 0; u < UA; ++u) {
      const int2 e = s_hits[h + u];
      v[u] = ldg_c128(ecol + (int64_t)e.x * lde + (int64_t)(abs(e.y) - 1) * pitch);
      sg[u] = e.y < 0 ? -1.0 : 1.0;
    }
#pragma unroll
    for (int u = 0; u < UA; ++u) {
      acc.x += sg[u] * v[u].x;
      acc.y += sg[u] * v[u].y;
    }
  }
  for (; h < nhit; ++h) {
    const int2 e = s_hits[h];
    axpy_sign(acc, e.y, ecol[(int64_t)e.x * lde + (int64_t)(abs(e.y) - 1) * pitch]);
  }
  if (in_chunk) {
    const double2 *__restrict__ erow = evec + xr * pitch;
    const int2 *__restrict__ cl = clist_b + b;
    int slot = 0;
    for (; slot + UB <= lk_b; slot += UB) {
      int2 e[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) e[u] = ldg_int2(cl + (int64_t)(slot + u) * lenb);
      double2 v[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u)
        v[u] = ldg_c128(erow + (int64_t)s_rowmap[e[u].x] * lde + (abs(e[u].y) - 1));
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const double sg = e[u].y < 0 ? -1.0 : 1.0;
        acc.x += sg * v[u].x;
        acc.y += sg * v[u].y;
      }
    }
    for (; slot < lk_b; ++slot) {
      const int2 e = cl[(int64_t)slot * lenb];
      axpy_sign(acc, e.y, erow[(int64_t)s_rowmap[e.x] * lde + (abs(e.y) - 1)]);
    }
  }
  double2 s = out[x * lenb + b];
  s.x += z.x * acc.x - z.y * acc.y;
  s.y += z.x * acc.y + z.y * acc.x;
  out[x * lenb + b] = s;
}

// One-body operator, sigma[a,b] = sum_ij h1[ij] * D[ij,a,b], without ever forming D:
// the compact per-string excitation lists give exactly lk_a coalesced row reads and lk_b
// in-row gathers per element (instead of probing all norb^2 pairs).  This is
// FqeData._apply_array_spatial1 (fqe_data.py:477-530) and the inner step of the Taylor
// series for quadratic Hamiltonians / orbital rotations.
__global__ void __launch_bounds__(kTB, 3)
k_apply_one_body(int npair, int64_t lena, int64_t lenb, int lk_a, int lk_b,
                 const int2 *__restrict__ clistT_a, const int2 *__restrict__ clist_b,
                 const double2 *__restrict__ h1, const double2 *__restrict__ coeff, int64_t row0,
                 int nbt, double2 *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char s_raw1[];
  double2 *s_h1 = reinterpret_cast<double2 *>(s_raw1);            // [npair]
  int2 *s_alpha = reinterpret_cast<int2 *>(s_h1 + npair);         // [lk_a]
  const int64_t tile = blockIdx.x;
  const int64_t a = row0 + tile / nbt;
  for (int p = threadIdx.x; p < npair; p += kTB) s_h1[p] = h1[p];
  for (int sl = threadIdx.x; sl < lk_a; sl += kTB) s_alpha[sl] = clistT_a[a * (int64_t)lk_a + sl];
  __syncthreads();
  const int64_t b = (tile % nbt) * kTB + threadIdx.x;
  if (b >= lenb) return;
  const double2 *__restrict__ ccol = coeff + b;
  const double2 *__restrict__ crow = coeff + a * lenb;
  double2 acc = make_double2(0.0, 0.0);
  constexpr int U = 8;
  auto fma_c = [&](double2 h, int sy, double2 v) {
    const double sg = sy < 0 ? -1.0 : 1.0;
    acc.x += sg * (h.x * v.x - h.y * v.y);
    acc.y += sg * (h.x * v.y + h.y * v.x);
  };
  int sl = 0;
  for (; sl + U <= lk_a; sl += U) {
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      v[u] = ldg_c128(ccol + (int64_t)(abs(s_alpha[sl + u].y) - 1) * lenb);
#pragma unroll
    for (int u = 0; u < U; ++u) fma_c(s_h1[s_alpha[sl + u].x], s_alpha[sl + u].y, v[u]);
  }
  for (; sl < lk_a; ++sl)
    fma_c(s_h1[s_alpha[sl].x], s_alpha[sl].y, ccol[(int64_t)(abs(s_alpha[sl].y) - 1) * lenb]);
  const int2 *__restrict__ cl = clist_b + b;
  sl = 0;
  for (; sl + U <= lk_b; sl += U) {
    int2 e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = ldg_int2(cl + (int64_t)(sl + u) * lenb);
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldg_c128(crow + (abs(e[u].y) - 1));
#pragma unroll
    for (int u = 0; u < U; ++u) fma_c(s_h1[e[u].x], e[u].y, v[u]);
  }
  for (; sl < lk_b; ++sl) {
    const int2 e = cl[(int64_t)sl * lenb];
    fma_c(s_h1[e.x], e.y, crow[abs(e.y) - 1]);
  }
  double2 sv = out[a * lenb + b];
  sv.x += acc.x;
  sv.y += acc.y;
  out[a * lenb + b] = sv;
}

// sigma rows [row0, row0+nrows) += h1 applied to C (all pairs)
int launch_one_body(const fqeb_graph *g, const double *d_coeff, const double *d_h1, int64_t row0,
                    int64_t nrows, double *d_out, cudaStream_t st) {
  const int npair = g->norb * g->norb;
  const int64_t lena = g->len[0], lenb = g->len[1];
  FQEB_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= lena, "one_body: bad row range");
  if (nrows == 0 || npair == 0) return FQEB_OK;
  const int nbt = (int)((lenb + kTB - 1) / kTB);
  const int64_t tiles = nrows * nbt;
  FQEB_REQUIRE(tiles < (1ll << 31), "one_body: problem too large for one launch");
  const size_t smem = sizeof(double2) * (size_t)npair + sizeof(int2) * (size_t)g->lk[0];
  if (smem > 48 * 1024)
    FQEB_CUDA(cudaFuncSetAttribute(k_apply_one_body, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  k_apply_one_body<<<(unsigned)tiles, kTB, smem, st>>>(
      npair, lena, lenb, g->lk[0], g->lk[1], g->d_clistT[0], g->d_clist[1],
      (const double2 *)d_h1, (const double2 *)d_coeff, row0, nbt, (double2 *)d_out);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

// D rows for the table rows [c0, c1): the plain pair space (sym = false, table row = ij) or
// the compressed one (sym = true, table row = i(i+1)/2 + j)
int launch_gather(const fqeb_graph *g, bool sym, const double *d_coeff, double *d_dvec,
                  int64_t ldd, int64_t row0, int64_t nrows, int c0, int c1, cudaStream_t st) {
  const int ntab = sym ? g->norb * (g->norb + 1) / 2 : g->norb * g->norb;
  const int64_t lena = g->len[0], lenb = g->len[1];
  FQEB_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= lena,
               "make_dvec: rows [%lld,+%lld) outside [0,%lld)", (long long)row0, (long long)nrows,
               (long long)lena);
  FQEB_REQUIRE(c0 >= 0 && c0 <= c1 && c1 <= ntab, "make_dvec: pair slice [%d,%d) invalid", c0, c1);
  FQEB_REQUIRE(ldd >= nrows * lenb, "make_dvec: ldd=%lld < nrows*lenb", (long long)ldd);
  if (nrows == 0 || c0 == c1) return FQEB_OK;
  const int nbt = (int)((lenb + kTB - 1) / kTB);
  const int64_t tiles = nrows * nbt;
  FQEB_REQUIRE(tiles < (1ll << 31), "make_dvec: chunk too large for one launch");
  const size_t smem = sizeof(int) * (size_t)(c1 - c0);
  // measured at norb=16: 85.0 ms per sigma with row-major tiles, 79.9 with groups of 128 rows
  static const int group = getenv("FQEB_GATHER_GROUP") ? atoi(getenv("FQEB_GATHER_GROUP")) : 128;
  k_gather<4><<<(unsigned)tiles, kTB, smem, st>>>(
      ntab, lenb, sym ? g->d_smapT[0] : g->d_amapT[0], sym ? g->d_smap[1] : g->d_amap[1],
      (const double2 *)d_coeff, (double2 *)d_dvec, ldd, row0, nbt, c0, c1,
      group < 1 ? 1 : group, nrows);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
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
  const size_t smem = sizeof(int4) * (size_t)(ij1 - ij0);
#define FQEB_LAUNCH_DVEC(WD, HH)                                                               \
  do {                                                                                         \
    if (smem > 48 * 1024)                                                                      \
      FQEB_CUDA(cudaFuncSetAttribute(k_make_dvec<WD, HH>,                                      \
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_make_dvec<WD, HH><<<(unsigned)tiles, kTB, smem, st>>>(                                   \
        npair_full, lena, lenb, g->d_amapT[0], g->d_amap[1], c, d, ldd, row0, nbt, ij0, ij1,   \
        d_pairs, h1, sig);                                                                     \
  } while (0)
  // D itself is produced by launch_gather; this launcher only serves the one-body
  // accumulation over a pair slice
  FQEB_REQUIRE(d == nullptr && h1 != nullptr && sig != nullptr,
               "make_dvec: one-body accumulation needs h1 and sigma, and no D buffer");
  FQEB_LAUNCH_DVEC(false, true);
#undef FQEB_LAUNCH_DVEC
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

// rowmap == nullptr: identity (E row kl <-> pair kl)
// pitch: distance (complex elements) between consecutive alpha rows inside an E row
// (lenb for the plain layout, a multiple of 64 for the fused kernel's layout)
// x0 <= x < x1: target rows to complete (the whole sigma: 0, lena)
int launch_make_coeff(const fqeb_graph *g, const double *d_evec, int64_t lde, int64_t pitch,
                      int64_t row0, int64_t nrows, const int32_t *d_rowmap, double zr, double zi,
                      double *d_out, cudaStream_t st, int64_t x0, int64_t x1) {
  if (!d_rowmap) d_rowmap = g->d_rowmap_id;
  const int npair = g->norb * g->norb;
  const int64_t lena = g->len[0], lenb = g->len[1];
  FQEB_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= lena,
               "make_coeff: rows [%lld,+%lld) outside [0,%lld)", (long long)row0,
               (long long)nrows, (long long)lena);
  FQEB_REQUIRE(pitch >= lenb && lde >= nrows * pitch, "make_coeff: lde=%lld < nrows*pitch",
               (long long)lde);
  if (x1 < 0) x1 = lena;
  FQEB_REQUIRE(x0 >= 0 && x0 <= x1 && x1 <= lena, "make_coeff: target rows [%lld,%lld) invalid",
               (long long)x0, (long long)x1);
  if (nrows == 0 || npair == 0 || x0 == x1) return FQEB_OK;
  const int nbt = (int)((lenb + kTB - 1) / kTB);
  const int64_t tiles = (x1 - x0) * nbt;
  FQEB_REQUIRE(tiles < (1ll << 31), "make_coeff: problem too large for one launch");
  size_t smem = sizeof(int) * (size_t)((npair + 1) & ~1) + sizeof(int2) * (size_t)g->lk[0];
  {
    // experiment knob: FQEB_SCATTER_CTAS_PER_SM=k pads the dynamic shared memory so that only k
    // CTAs fit on an SM (fewer target rows in flight -> smaller beta-way footprint in L2)
    static const int cap = getenv("FQEB_SCATTER_CTAS_PER_SM") ? atoi(getenv("FQEB_SCATTER_CTAS_PER_SM")) : 0;
    if (cap >= 1 && cap <= 8) {
      const size_t pad = (size_t)(227 * 1024) / cap - 2048;
      if (pad > smem) smem = pad;
    }
  }
  static PerDeviceSize attr_dev;
  int attr_dev_id = 0;
  if (smem > 48 * 1024 && attr_dev.needs(smem, &attr_dev_id)) {
    FQEB_CUDA(cudaFuncSetAttribute(k_make_coeff, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    attr_dev.record(attr_dev_id, smem);
  }
  k_make_coeff<<<(unsigned)tiles, kTB, smem, st>>>(
      npair, lena, lenb, g->lk[0], g->lk[1], g->d_clistT[0], g->d_clist[1], d_rowmap,
      (const double2 *)d_evec, lde, pitch, row0, nrows, nbt, x0, make_double2(zr, zi),
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
  return fqeb::launch_gather(g, false, d_coeff, d_dvec, ldd, row0, nrows, ij0, ij1,
                             (cudaStream_t)stream);
}

extern "C" int fqeb_make_coeff(const fqeb_graph *g, const double *d_evec, int64_t lde,
                               int64_t row0, int64_t nrows, double zr, double zi, double *d_out,
                               void *stream) {
  int rc = fqeb::require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(g && d_evec && d_out, "fqeb_make_coeff: NULL argument");
  return fqeb::launch_make_coeff(g, d_evec, lde, g->len[1], row0, nrows, nullptr, zr, zi, d_out,
                                 (cudaStream_t)stream, 0, -1);
}
