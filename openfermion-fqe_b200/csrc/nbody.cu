// Individual n-body operators: apply, and the masked scalings of their exact evolution.
//
// Replace make_mapping_each (reference src/fqe/lib/fci_graph.c:223-264),
// apply_individual_nbody1_accumulate (lib/fqe_data.c:1157-1182), evaluate_map_each
// (lib/fqe_data.c:1417-1440) and sparse_scale (lib/fqe_data.c:1265-1278) as driven by
// FqeData.apply_individual_nbody_accumulate (fqe_data.py:1590-1653), apply_cos_inplace,
// evolve_individual_nbody_nontrivial and evolve_inplace_individual_nbody_trivial
// (fqe_data.py:2385-2580): SURVEY 8f rank 2.
//
// An individual operator is  z * prod a^+_{dag} prod a_{undag}  on the alpha strings times the
// same on the beta strings.  Each spin part is a signed partial permutation of the string
// space.  The reference materialises it as a (source, target, parity) list on the host; here
// one kernel writes the INVERSE of both spin parts as by-target tables inv[t] = sign * (s + 1)
// (0: no source), so that the accumulation
//       out[ta, tb] += z * pa * pb * in[sa, sb]
// is a coalesced by-target gather over the (few) target rows: no host round trip, no atomics.
//
// Roofline: HBM-bound, 3 * 16 bytes per touched determinant (read in, read + write out).
#include "fqeb_common.cuh"

namespace fqeb {

constexpr int kNB = 256;
constexpr int kMaxOps = 8;  // ladder operators per spin and kind (the reference goes to 4-body)

struct OpList {
  int ndag, nundag;
  int dag[kMaxOps], undag[kMaxOps];
};

__device__ __forceinline__ int bits_above(uint64_t s, int i) {
  return __popcll(s >> (i + 1));
}

// By-target table of both spin parts in one launch (blockIdx.y = spin):
//   inv[t] = sign * (s + 1)  with  A |s> = sign |t>,   0 if no string maps onto t.
// The source of a target is found by walking the ADJOINT operator
//   A^+ = a^+_{u_k} ... a^+_{u_1} a_{d_k} ... a_{d_1}      (A = a^+_{d_1}..a^+_{d_k} a_{u_1}..a_{u_k})
// on t, rightmost operator first; the sign of <t|A|s> equals the sign of <s|A^+|t>, and each
// step's sign is (-1)^(occupied orbitals above the one acted on), as in the reference's
// make_mapping_each (lib/fci_graph.c:248-257).  Every entry is written, so no memset is needed.
__global__ void k_nbody_invmap(int norb, int64_t lena, int64_t lenb,
                               const uint64_t *__restrict__ astr, const uint64_t *__restrict__ bstr,
                               const int32_t *__restrict__ za, const int32_t *__restrict__ zb,
                               OpList opa, OpList opb, int32_t *__restrict__ inva,
                               int32_t *__restrict__ invb) {
  const bool beta = blockIdx.y != 0;
  const int64_t len = beta ? lenb : lena;
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= len) return;
  const OpList &ops = beta ? opb : opa;
  uint64_t cur = (beta ? bstr : astr)[x];
  int parity = 0;
  bool ok = true;
  for (int j = 0; j < ops.ndag && ok; ++j) {      // a_{d_1} first
    const int o = ops.dag[j];
    ok = (cur >> o) & 1ull;
    parity += bits_above(cur, o);
    cur &= ~(1ull << o);
  }
  for (int j = 0; j < ops.nundag && ok; ++j) {    // then a^+_{u_1}
    const int o = ops.undag[j];
    ok = !((cur >> o) & 1ull);
    parity += bits_above(cur, o);
    cur |= (1ull << o);
  }
  int32_t val = 0;
  if (ok) {
    const int s = fqeb_string_address(cur, beta ? zb : za, norb);
    val = (parity & 1) ? -(int32_t)(s + 1) : (int32_t)(s + 1);
  }
  (beta ? invb : inva)[x] = val;
}

__global__ void __launch_bounds__(kNB)
k_nbody_accumulate(int64_t lenb, const int32_t *__restrict__ inva,
                   const int32_t *__restrict__ invb, double2 zc, int nbt,
                   const double2 *__restrict__ in, double2 *__restrict__ out) {
  const int64_t tile = blockIdx.x;
  const int64_t ta = tile / nbt;
  const int ea = inva[ta];
  if (ea == 0) return;
  const int64_t tb = (tile % nbt) * kNB + threadIdx.x;
  if (tb >= lenb) return;
  const int eb = invb[tb];
  if (eb == 0) return;
  const double2 v = in[(int64_t)(abs(ea) - 1) * lenb + (abs(eb) - 1)];
  const double sg = ((ea < 0) != (eb < 0)) ? -1.0 : 1.0;
  double2 o = out[ta * lenb + tb];
  o.x += sg * (zc.x * v.x - zc.y * v.y);
  o.y += sg * (zc.x * v.y + zc.y * v.x);
  out[ta * lenb + tb] = o;
}

// C[a,b] *= f on the strings with all `occ` orbitals occupied and all `emp` orbitals empty
__global__ void __launch_bounds__(kNB)
k_sparse_scale(int64_t lenb, const uint64_t *__restrict__ astr, const uint64_t *__restrict__ bstr,
               uint64_t a_occ, uint64_t a_emp, uint64_t b_occ, uint64_t b_emp, double2 f, int nbt,
               double2 *__restrict__ coeff) {
  const int64_t tile = blockIdx.x;
  const int64_t a = tile / nbt;
  const uint64_t sa = astr[a];
  if ((sa & a_occ) != a_occ || (sa & a_emp) != 0) return;
  const int64_t b = (tile % nbt) * kNB + threadIdx.x;
  if (b >= lenb) return;
  const uint64_t sb = bstr[b];
  if ((sb & b_occ) != b_occ || (sb & b_emp) != 0) return;
  const double2 v = coeff[a * lenb + b];
  coeff[a * lenb + b] = make_double2(f.x * v.x - f.y * v.y, f.x * v.y + f.y * v.x);
}

static int fill_ops(OpList *ops, const int *dag, const int *undag, int n, int norb) {
  FQEB_REQUIRE(n >= 0 && n <= kMaxOps, "nbody: %d operators per spin exceeds %d", n, kMaxOps);
  ops->ndag = ops->nundag = n;
  for (int k = 0; k < n; ++k) {
    FQEB_REQUIRE(dag[k] >= 0 && dag[k] < norb && undag[k] >= 0 && undag[k] < norb,
                 "nbody: orbital index outside [0,%d)", norb);
    ops->dag[k] = dag[k];
    ops->undag[k] = undag[k];
  }
  return FQEB_OK;
}

}  // namespace fqeb

using namespace fqeb;

extern "C" int fqeb_nbody_accumulate(const fqeb_graph *g, double zr, double zi, const int *daga,
                                     const int *undaga, int na, const int *dagb,
                                     const int *undagb, int nb, const double *d_in,
                                     double *d_out, void *stream) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(g && d_in && d_out && d_in != d_out, "fqeb_nbody_accumulate: NULL or aliased");
  FQEB_REQUIRE((na == 0 || (daga && undaga)) && (nb == 0 || (dagb && undagb)),
               "fqeb_nbody_accumulate: NULL operator list");
  OpList oa, ob;
  rc = fill_ops(&oa, daga, undaga, na, g->norb);
  if (rc != FQEB_OK) return rc;
  rc = fill_ops(&ob, dagb, undagb, nb, g->norb);
  if (rc != FQEB_OK) return rc;
  const int64_t lena = g->len[0], lenb = g->len[1];
  cudaStream_t st = (cudaStream_t)stream;
  // by-target tables live in this stream's per-string scratch of the graph (16 bytes per string)
  GraphScratch sc;
  rc = graph_scratch(g, st, &sc);
  if (rc != FQEB_OK) return rc;
  int32_t *inva = (int32_t *)sc.sterm[0], *invb = (int32_t *)sc.sterm[1];
  const int64_t lmax = lena > lenb ? lena : lenb;
  k_nbody_invmap<<<dim3((unsigned)((lmax + kNB - 1) / kNB), 2), kNB, 0, st>>>(
      g->norb, lena, lenb, g->d_str[0], g->d_str[1], g->d_Z[0], g->d_Z[1], oa, ob, inva, invb);
  FQEB_CHECK_LAUNCH();
  const int nbt = (int)((lenb + kNB - 1) / kNB);
  FQEB_REQUIRE(lena * nbt < (1ll << 31), "fqeb_nbody_accumulate: problem too large");
  k_nbody_accumulate<<<(unsigned)(lena * nbt), kNB, 0, st>>>(
      lenb, inva, invb, make_double2(zr, zi), nbt, (const double2 *)d_in, (double2 *)d_out);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

extern "C" int fqeb_sparse_scale(const fqeb_graph *g, uint64_t a_occ, uint64_t a_emp,
                                 uint64_t b_occ, uint64_t b_emp, double fr, double fi,
                                 double *d_coeff, void *stream) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(g && d_coeff, "fqeb_sparse_scale: NULL argument");
  const int64_t lena = g->len[0], lenb = g->len[1];
  const int nbt = (int)((lenb + kNB - 1) / kNB);
  FQEB_REQUIRE(lena * nbt < (1ll << 31), "fqeb_sparse_scale: problem too large");
  k_sparse_scale<<<(unsigned)(lena * nbt), kNB, 0, (cudaStream_t)stream>>>(
      lenb, g->d_str[0], g->d_str[1], a_occ, a_emp, b_occ, b_emp, make_double2(fr, fi), nbt,
      (double2 *)d_coeff);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}
