// Two-electron contraction  E[kl, det] = sum_ij h2'[kl, ij] * D[ij, det]
// as an FP64 tensor-core (DMMA) GEMM for sm_100a.
//
// Replaces numpy.einsum("ijkl,klmn->ijmn", h2e, dvec) (reference
// src/fqe/fqe_data.py:656), i.e. a [norb^2 x norb^2] by [norb^2 x ndet] complex
// GEMM with 8*norb^4 real flops per determinant: the only compute-bound step of
// the sigma build (arithmetic intensity norb^2/4 flop/B).
//
// sm_100a has no tcgen05 / TMEM path for FP64: the FP64 tensor op is the warp-level
// mma.sync.m8n8k4.f64 (SASS DMMA.8).  The kernel is therefore a classic
// cp.async multi-stage, warp-tiled MMA kernel, specialised for this problem:
//
//   * the small operator A (<= 4 MB) is pre-expanded ON THE HOST into the real
//     matrix the tensor cores consume, stays L2-resident, and is streamed in
//     128x16 tiles;
//   * the big operand D is consumed in its natural complex128 (interleaved)
//     layout - no planar copy.  Two modes share one kernel template:
//       REAL  (h2' real, or purely imaginary = i*real as in Taylor's -i*t*H):
//             D is viewed as a real [K x 2*ndet] matrix, E likewise: ONE real GEMM
//             of M=norb^2, K=norb^2, N=2*ndet   (4*norb^4 flop/det, half the work);
//       CPLX  (general complex h2'): the real 2M x 2K operator
//             [[Ar,-Ai],[Ai,Ar]] is applied with the real/imag parts of D taken as
//             separate k-indices straight out of the interleaved shared-memory
//             tile; A's rows are ordered in groups of 16 (8 real-part rows then the
//             8 imaginary-part rows of the same kl) so that each thread ends up
//             holding (re, im) of the same output element and the epilogue writes
//             interleaved complex128 with 16-byte stores.
//   * CTA tile 128 (real rows) x 128 (real cols / dets), 8 warps as 2(M) x 4(N),
//     warp tile 64x32 = 8x4 DMMA tiles, 64 FP64 accumulators per thread,
//     4-stage cp.async pipeline (146 KB shared memory), k-step 16.
//   * shared-memory strides (20 / 132 / 264 doubles) make every fragment load
//     bank-conflict free for 64-bit accesses.
//   * tiles are ordered M-fastest so the CTAs sharing a D tile run together and
//     D is read from HBM once.
#include "fqeb_common.cuh"

#include <stdlib.h>
#include <string.h>
#include <vector>

namespace fqeb {

constexpr int BM = 128;       // real rows per CTA
constexpr int BNR = 128;      // real columns per CTA
constexpr int KSTEP = 16;     // real k per pipeline stage
constexpr int STAGES = 4;
constexpr int A_STRIDE = KSTEP + 4;   // doubles
constexpr int B_STRIDE_R = BNR + 4;   // REAL: [16][132]
constexpr int B_STRIDE_C = 2 * BNR + 8;  // CPLX: [8][264]
constexpr int A_TILE = BM * A_STRIDE;            // doubles
constexpr int B_TILE = 16 * B_STRIDE_R;          // == 8 * B_STRIDE_C
static_assert(16 * B_STRIDE_R == 8 * B_STRIDE_C, "B tile size mismatch");
constexpr int STAGE_DOUBLES = A_TILE + B_TILE;
constexpr size_t GEMM_SMEM = sizeof(double) * STAGE_DOUBLES * STAGES;
constexpr int COL_ALIGN = 128;  // leading dimensions (complex elements) must be multiples

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// A: real row-major, leading dimension lda (doubles); a_col0 = first real k of the slice.
// B: complex [.][ldb] (D rows of the slice start at row 0).  E: complex [.][lde].
// m_valid: number of real rows of A that carry data (multiple of 8; of 16 for CPLX).
// nk: number of KSTEP iterations.  npair: number of valid complex output rows.
template <bool CPLX>
__global__ void __launch_bounds__(256, 1)
k_dgemm(const double *__restrict__ A, int lda, int a_col0, const double2 *__restrict__ B,
        int64_t ldb, double2 *__restrict__ E, int64_t lde, int m_valid, int npair, int nk,
        int nmb) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int wm0 = (warp >> 2) * 64;
  const int wn0 = (warp & 3) * 32;
  const int mb = blockIdx.x % nmb;
  const int64_t nb = blockIdx.x / nmb;
  const int m0 = mb * BM;
  // first determinant (complex column) of this tile
  const int64_t n0 = nb * (CPLX ? BNR : BNR / 2);

  int mt_active = (m_valid - (m0 + wm0)) / 8;
  mt_active = mt_active < 0 ? 0 : (mt_active > 8 ? 8 : mt_active);

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const double *Ag = A + (int64_t)m0 * lda + a_col0;
  const double2 *Bg = B + n0;

  auto load_stage = [&](int stage, int kt) {
    double *As = smem + stage * STAGE_DOUBLES;
    double *Bs = As + A_TILE;
    const int k0 = kt * KSTEP;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = tid + i * 256;
      const int row = c >> 3, cc = c & 7;
      cp_async16(As + row * A_STRIDE + cc * 2, Ag + (int64_t)row * lda + k0 + cc * 2);
    }
    if (CPLX) {
      const int r0 = k0 >> 1;  // 8 complex rows of D per stage
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = tid + i * 256;
        const int row = c >> 7, cc = c & 127;
        cp_async16(Bs + row * B_STRIDE_C + cc * 2, Bg + (int64_t)(r0 + row) * ldb + cc);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = tid + i * 256;
        const int row = c >> 6, cc = c & 63;
        cp_async16(Bs + row * B_STRIDE_R + cc * 2, Bg + (int64_t)(k0 + row) * ldb + cc);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kt + STAGES - 1;
      if (nxt < nk) load_stage(nxt % STAGES, nxt);
      cp_async_commit();
    }
    const double *As = smem + (kt % STAGES) * STAGE_DOUBLES;
    const double *Bs = As + A_TILE;
#pragma unroll
    for (int kk = 0; kk < KSTEP / 4; ++kk) {
      double bf[4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = wn0 + nt * 8 + g;
        if (CPLX) {
          bf[nt] = Bs[(kk * 2 + (tg >> 1)) * B_STRIDE_C + n * 2 + (tg & 1)];
        } else {
          bf[nt] = Bs[(kk * 4 + tg) * B_STRIDE_R + n];
        }
      }
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) {
        if (mt < mt_active) {
          const double af = As[(wm0 + mt * 8 + g) * A_STRIDE + kk * 4 + tg];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af, bf[nt]);
        }
      }
    }
  }
  cp_async_wait<0>();

  // ---- epilogue: interleaved complex128, 16-byte stores --------------------------------
  if (CPLX) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      if (2 * p < mt_active) {
        const int kl = ((m0 + wm0) >> 1) + p * 8 + g;
        if (kl < npair) {
          double2 *erow = E + (int64_t)kl * lde + n0 + wn0 + 2 * tg;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            erow[nt * 8] = make_double2(acc[2 * p][nt][0], acc[2 * p + 1][nt][0]);
            erow[nt * 8 + 1] = make_double2(acc[2 * p][nt][1], acc[2 * p + 1][nt][1]);
          }
        }
      }
    }
  } else {
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      if (mt < mt_active) {
        const int kl = m0 + wm0 + mt * 8 + g;
        if (kl < npair) {
          double2 *erow = E + (int64_t)kl * lde + n0 + (wn0 >> 1) + tg;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
            erow[nt * 4] = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        }
      }
    }
  }
}

static bool g_attr_set[2] = {false, false};

int launch_contract(const fqeb_op *op, const double *d_dvec, int64_t ldd, double *d_evec,
                    int64_t lde, int64_t ncols, int ij0, int ij1, cudaStream_t st) {
  const int npair = op->norb * op->norb;
  FQEB_REQUIRE(op->has_h2, "contract: operator has no two-body part");
  FQEB_REQUIRE(ij0 >= 0 && ij0 < ij1 && ij1 <= npair, "contract: pair slice [%d,%d) invalid", ij0,
               ij1);
  FQEB_REQUIRE(ldd % COL_ALIGN == 0 && lde % COL_ALIGN == 0,
               "contract: leading dimensions must be multiples of %d", COL_ALIGN);
  FQEB_REQUIRE(ncols >= 0 && ncols <= ldd && ncols <= lde, "contract: ncols exceeds ld");
  FQEB_REQUIRE((ij0 & 1) == 0, "contract: pair slice must start at an even index");
  if (ncols == 0) return FQEB_OK;
  const bool cplx = op->kind == FQEB_OP_COMPLEX;
  const int nij = ij1 - ij0;
  const int kslice = cplx ? 2 * nij : nij;
  const int nk = (kslice + KSTEP - 1) / KSTEP;
  const int a_col0 = cplx ? 2 * ij0 : ij0;
  FQEB_REQUIRE(a_col0 + nk * KSTEP <= op->Kp, "contract: operator padding too small");
  const int m_valid = cplx ? 2 * (int)round_up(npair, 8) : (int)round_up(npair, 8);
  const int nmb = (m_valid + BM - 1) / BM;
  const int64_t cols_pad = round_up(ncols, COL_ALIGN);
  const int64_t nnb = cols_pad / (cplx ? BNR : BNR / 2);
  const int64_t tiles = nnb * nmb;
  FQEB_REQUIRE(tiles < (1ll << 31), "contract: too many tiles for one launch");
  const int idx = cplx ? 1 : 0;
  if (!g_attr_set[idx]) {
    if (cplx) {
      FQEB_CUDA(cudaFuncSetAttribute(k_dgemm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)GEMM_SMEM));
    } else {
      FQEB_CUDA(cudaFuncSetAttribute(k_dgemm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)GEMM_SMEM));
    }
    g_attr_set[idx] = true;
  }
  if (cplx) {
    k_dgemm<true><<<(unsigned)tiles, 256, GEMM_SMEM, st>>>(
        op->d_A, op->Kp, a_col0, (const double2 *)d_dvec, ldd, (double2 *)d_evec, lde, m_valid,
        npair, nk, nmb);
  } else {
    k_dgemm<false><<<(unsigned)tiles, 256, GEMM_SMEM, st>>>(
        op->d_A, op->Kp, a_col0, (const double2 *)d_dvec, ldd, (double2 *)d_evec, lde, m_valid,
        npair, nk, nmb);
  }
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

// rows of D a caller must provide (zero-filled beyond the slice) for a slice of nij pairs
int dvec_rows_padded(const fqeb_op *op, int nij) {
  const bool cplx = op->kind == FQEB_OP_COMPLEX;
  return (int)round_up(nij, cplx ? KSTEP / 2 : KSTEP);
}

}  // namespace fqeb

using namespace fqeb;

extern "C" int fqeb_gemm_col_align(void) { return COL_ALIGN; }

extern "C" int fqeb_contract_dvec_rows(const fqeb_op *op, int nij) {
  if (!op || nij < 0) return -1;
  return dvec_rows_padded(op, nij);
}

extern "C" int fqeb_op_create(int norb, const double *h_h1p, const double *h_h2p, fqeb_op **out) {
  FQEB_REQUIRE(out != nullptr, "fqeb_op_create: out is NULL");
  *out = nullptr;
  FQEB_REQUIRE(norb >= 1 && norb <= kMaxOrb, "fqeb_op_create: norb=%d outside [1,%d]", norb,
               kMaxOrb);
  FQEB_REQUIRE(h_h1p != nullptr, "fqeb_op_create: h1p is NULL");
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  const int npair = norb * norb;
  fqeb_op *op = (fqeb_op *)calloc(1, sizeof(fqeb_op));
  if (!op) {
    set_error("fqeb_op_create: host allocation failed");
    return FQEB_ERR_NOMEM;
  }
  op->norb = norb;
  op->has_h2 = h_h2p != nullptr;
  op->kind = FQEB_OP_REAL;
  op->zr = 1.0;
  op->zi = 0.0;
  cudaGetDevice(&op->device);
  auto fail = [&](int code) {
    fqeb_op_destroy(op);
    return code;
  };
  if (cudaMalloc(&op->d_h1, sizeof(double) * 2 * npair) != cudaSuccess ||
      cudaMemcpy(op->d_h1, h_h1p, sizeof(double) * 2 * npair, cudaMemcpyHostToDevice) !=
          cudaSuccess) {
    set_error("fqeb_op_create: cannot upload h1: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(FQEB_ERR_CUDA);
  }
  if (op->has_h2) {
    const size_t n2 = (size_t)npair * npair;
    bool any_re = false, any_im = false;
    for (size_t k = 0; k < n2; ++k) {
      any_re |= (h_h2p[2 * k] != 0.0);
      any_im |= (h_h2p[2 * k + 1] != 0.0);
    }
    if (any_im && any_re) op->kind = FQEB_OP_COMPLEX;
    else if (any_im) {
      op->kind = FQEB_OP_IMAG;
      op->zr = 0.0;
      op->zi = 1.0;
    }
    std::vector<double> a;
    if (op->kind == FQEB_OP_COMPLEX) {
      const int np8 = (int)round_up(npair, 8);
      op->Mp = (int)round_up(2 * np8, BM);
      op->Kp = (int)round_up(2 * npair, KSTEP) + KSTEP;  // slack for ragged slices
      a.assign((size_t)op->Mp * op->Kp, 0.0);
      for (int kl = 0; kl < npair; ++kl) {
        const int rre = (kl / 8) * 16 + (kl % 8), rim = rre + 8;
        for (int ij = 0; ij < npair; ++ij) {
          const double re = h_h2p[2 * ((size_t)kl * npair + ij)];
          const double im = h_h2p[2 * ((size_t)kl * npair + ij) + 1];
          a[(size_t)rre * op->Kp + 2 * ij] = re;
          a[(size_t)rre * op->Kp + 2 * ij + 1] = -im;
          a[(size_t)rim * op->Kp + 2 * ij] = im;
          a[(size_t)rim * op->Kp + 2 * ij + 1] = re;
        }
      }
    } else {
      const int off = op->kind == FQEB_OP_IMAG ? 1 : 0;
      op->Mp = (int)round_up(round_up(npair, 8), BM);
      op->Kp = (int)round_up(npair, KSTEP) + KSTEP;
      a.assign((size_t)op->Mp * op->Kp, 0.0);
      for (int kl = 0; kl < npair; ++kl)
        for (int ij = 0; ij < npair; ++ij)
          a[(size_t)kl * op->Kp + ij] = h_h2p[2 * ((size_t)kl * npair + ij) + off];
    }
    if (cudaMalloc(&op->d_A, sizeof(double) * a.size()) != cudaSuccess ||
        cudaMemcpy(op->d_A, a.data(), sizeof(double) * a.size(), cudaMemcpyHostToDevice) !=
            cudaSuccess) {
      set_error("fqeb_op_create: cannot upload h2 operand: %s",
                cudaGetErrorString(cudaGetLastError()));
      return fail(FQEB_ERR_CUDA);
    }
  }
  *out = op;
  return FQEB_OK;
}

extern "C" int fqeb_op_destroy(fqeb_op *op) {
  if (!op) return FQEB_OK;
  if (op->d_A) cudaFree(op->d_A);
  if (op->d_h1) cudaFree(op->d_h1);
  free(op);
  return FQEB_OK;
}

extern "C" int fqeb_op_kind(const fqeb_op *op, int *kind) {
  FQEB_REQUIRE(op && kind, "fqeb_op_kind: NULL argument");
  *kind = op->kind;
  return FQEB_OK;
}

extern "C" int fqeb_contract(const fqeb_op *op, const double *d_dvec, int64_t ldd,
                             double *d_evec, int64_t lde, int64_t ncols, int ij0, int ij1,
                             void *stream) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(op && d_dvec && d_evec, "fqeb_contract: NULL argument");
  return launch_contract(op, d_dvec, ldd, d_evec, lde, ncols, ij0, ij1, (cudaStream_t)stream);
}
