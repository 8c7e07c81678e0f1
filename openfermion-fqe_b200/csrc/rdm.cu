// Contraction over the determinant index for reduced density matrices:
//
//     G[m, n] += sum_c conj(bra[m, c]) * ket[n, c],     m < M, n < N, c < ncols
//
// with bra = D_bra[ij] = E_ij |bra> and ket = D_ket[kl] (+ one extra row holding C itself, which
// gives the one-particle part <D_bra[ij] | C> in the same pass).  Replaces the reductions of the
// reference's rdm12 (src/fqe/fqe_data.py:1726-1838: `numpy.tensordot(dvec2.conj(), dvec, ...)`
// and the 1-RDM einsum of :1668-1724).
//
// Shape: M, N <= norb^2 + 1 (a few hundred) against a contraction index of up to 1e8 columns, so
// the kernel is a split-K Gram product: each CTA owns one 64 x 64 block of G and one slab of
// columns, streams the two 64-row panels of its slab through a 3-stage cp.async ring and keeps
// the block's real and imaginary accumulators in registers.  The FP64 tensor instruction is the
// warp-level mma.sync.m8n8k4.f64 (SASS DMMA; sm_100a has no tcgen05 path for FP64).
// Both panels are consumed in their natural interleaved complex layout, viewed as real rows of
// 2 ncols doubles with k = (column, re|im):
//
//     Re G = sum_k bra[m, k] * ket[n, k]
//     Im G = sum_k bra[m, k] * ket'[n, k],   ket'[(c, re)] = ket[(c, im)],  ket'[(c, im)] = -ket[(c, re)]
//
// i.e. the imaginary part reuses the same shared-memory tile, read at k ^ 1 with a sign.
// Partial blocks go to a scratch buffer [slab][M][N]; a second kernel adds them to G in slab
// order, so the result does not depend on scheduling (no atomics).
#include "fqeb_common.cuh"

namespace fqeb {

constexpr int GR_TM = 64, GR_TN = 64;        // complex block of G per CTA
constexpr int GR_KS = 16;                    // complex columns per stage (32 doubles)
constexpr int GR_STAGES = 3;
constexpr int GR_LD = 2 * GR_KS + 4;         // doubles between rows of a shared tile: conflict-free
constexpr int GR_THREADS = 256;              // 8 warps: 4 along m x 2 along n, warp tile 16 x 32
constexpr int GR_PANEL = GR_TM * GR_LD;      // doubles per panel per stage

__device__ __forceinline__ void gr_cp_async16(void *smem, const void *gmem, bool on) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = on ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void gr_dmma(double &c0, double &c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(GR_THREADS, 2)
k_gram(int M, int N, int64_t ncols, int64_t slab, const double2 *__restrict__ bra, int64_t ldb,
       const double2 *__restrict__ ket, int64_t ldk, const double2 *__restrict__ ket_extra,
       double2 *__restrict__ part, int upper_only) {
  extern __shared__ __align__(16) double gr_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_n = (N + GR_TN - 1) / GR_TN;
  const int m0 = (blockIdx.x / tiles_n) * GR_TM, n0 = (blockIdx.x % tiles_n) * GR_TN;
  // bra == ket: G is Hermitian; blocks strictly below the diagonal are left to the caller's
  // mirror (their partial sums stay zero: the scratch buffer is cleared first)
  if (upper_only && n0 + GR_TN <= m0) return;
  const int64_t c_begin = (int64_t)blockIdx.y * slab;
  const int64_t c_end = c_begin + slab < ncols ? c_begin + slab : ncols;
  const int nstage = c_end > c_begin ? (int)((c_end - c_begin + GR_KS - 1) / GR_KS) : 0;

  // loader: a panel row of a stage is 16 chunks of 16 bytes (one complex column each);
  // thread t copies chunk t % 16 of rows t / 16 + {0, 16, 32, 48} of both panels
  const int chunk = tid & 15, lrow = tid >> 4;
  auto issue = [&](int s) {
    double *sa = gr_smem + (size_t)(s % GR_STAGES) * 2 * GR_PANEL, *sb = sa + GR_PANEL;
    const int64_t c = c_begin + (int64_t)s * GR_KS + chunk;
    const bool col_on = s < nstage && c < c_end;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = lrow + 16 * r;
      const int m = m0 + row, n = n0 + row;
      const bool a_on = col_on && m < M;
      gr_cp_async16(sa + row * GR_LD + 2 * chunk, bra + (a_on ? (int64_t)m * ldb + c : 0), a_on);
      // row N - 1 may live in a separate buffer (the coefficient block itself)
      const bool b_on = col_on && n < N;
      const double2 *src = ket != nullptr ? ket : ket_extra;
      if (b_on) src = (ket_extra != nullptr && n == N - 1) ? ket_extra + c : ket + (int64_t)n * ldk + c;
      gr_cp_async16(sb + row * GR_LD + 2 * chunk, src, b_on);
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };

  const int wm = warp >> 1, wn = warp & 1;        // warp tile origin: (16 wm, 32 wn)
  const int fr = lane >> 2, fk = lane & 3;        // fragment row and k of this lane
  double cre[2][4][2], cim[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0;

  for (int s = 0; s < GR_STAGES - 1; ++s) issue(s);
  for (int s = 0; s < nstage; ++s) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(GR_STAGES - 2));
    __syncthreads();
    issue(s + GR_STAGES - 1);   // refills the slot consumed in iteration s - 1
    const double *sa = gr_smem + (size_t)(s % GR_STAGES) * 2 * GR_PANEL, *sb = sa + GR_PANEL;
    const double *pa = sa + (16 * wm + fr) * GR_LD + fk;
    const double *pb = sb + (32 * wn + fr) * GR_LD + fk;
    const double *pb_x = sb + (32 * wn + fr) * GR_LD + (fk ^ 1);
    const unsigned long long flip = (fk & 1) ? 0x8000000000000000ull : 0ull;   // sign of ket'
#pragma unroll
    for (int k4 = 0; k4 < 2 * GR_KS / 4; ++k4) {
      double a[2], b[4], bx[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = pa[i * 8 * GR_LD + 4 * k4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        b[j] = pb[j * 8 * GR_LD + 4 * k4];
        bx[j] = __longlong_as_double(__double_as_longlong(pb_x[j * 8 * GR_LD + 4 * k4]) ^ flip);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          gr_dmma(cre[i][j][0], cre[i][j][1], a[i], b[j]);
          gr_dmma(cim[i][j][0], cim[i][j][1], a[i], bx[j]);
        }
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::);

  // accumulator fragment: lane holds (row fr, columns 2 fk, 2 fk + 1) of each 8 x 8 block
  double2 *dst = part + (size_t)blockIdx.y * M * N;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + 16 * wm + 8 * i + fr;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = n0 + 32 * wn + 8 * j + 2 * fk + e;
        if (n < N) dst[(size_t)m * N + n] = make_double2(cre[i][j][e], cim[i][j][e]);
      }
  }
}

// out[m][n] += sum over slabs of part[slab][m][n] in slab order (bitwise reproducible).  With
// herm_tile > 0 the blocks strictly below the block diagonal were not computed: those elements
// take the conjugate of the mirrored partial sums instead.
__global__ void k_gram_reduce(int64_t n, int nslab, int M, int N, int herm_tile,
                              const double2 *__restrict__ part, double2 *__restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t src = i;
  double conj = 1.0;
  if (herm_tile > 0) {
    const int m = (int)(i / N), c = (int)(i % N);
    if ((c / herm_tile + 1) * herm_tile <= (m / herm_tile) * herm_tile) {   // block not computed
      src = (int64_t)c * N + m;   // c < m < M: a valid row
      conj = -1.0;
    }
  }
  double2 acc = out[i];
  for (int s = 0; s < nslab; ++s) {   // fixed order
    const double2 v = part[(size_t)s * n + src];
    acc.x += v.x;
    acc.y += conj * v.y;
  }
  out[i] = acc;
}

}  // namespace fqeb

using namespace fqeb;

// G[m, n] += sum_c conj(bra[m, c]) ket[n, c]; G is row-major [M][N] complex on the device.
// bra: [M][ldb], ket: [N][ldk] complex (interleaved doubles); when d_ket_last is not NULL, row
// N - 1 of ket is read from there instead (ncols contiguous complex numbers).
extern "C" int fqeb_gram_accumulate(int M, int N, int64_t ncols, const double *d_bra, int64_t ldb,
                                    const double *d_ket, int64_t ldk, const double *d_ket_last,
                                    double *d_G, void *stream) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(M >= 0 && N >= 0 && ncols >= 0, "gram: negative size");
  if (M == 0 || N == 0 || ncols == 0) return FQEB_OK;
  FQEB_REQUIRE(d_bra && d_G && (d_ket || (N == 1 && d_ket_last)), "gram: NULL argument");
  FQEB_REQUIRE(ldb >= ncols && (N == 1 && d_ket_last ? true : ldk >= ncols),
               "gram: leading dimension smaller than the column count");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_m = (M + GR_TM - 1) / GR_TM, tiles_n = (N + GR_TN - 1) / GR_TN;
  const int tiles = tiles_m * tiles_n;
  // Hermitian case (bra and ket are the same rows): only the blocks on and above the diagonal
  // are computed; k_gram_reduce fills G[n][m] = conj(G[m][n]) for the rest
  const bool herm = d_bra == d_ket && ldb == ldk && (d_ket_last ? N == M + 1 : N == M);
  int active = tiles;
  if (herm) {
    active = 0;
    for (int tm = 0; tm < tiles_m; ++tm)
      for (int tn = 0; tn < tiles_n; ++tn)
        if (!((tn + 1) * GR_TN <= tm * GR_TM)) ++active;
  }
  // ONE wave: as many column slabs as fit two CTAs per SM with the blocks that do work
  // (one CTA more than fits costs a whole second wave), whole stages per slab
  int64_t nslab = (2 * (int64_t)sm_count()) / active;
  if (nslab < 1) nslab = 1;
  const int64_t stages = (ncols + GR_KS - 1) / GR_KS;
  if (nslab > stages) nslab = stages;
  if (nslab > 65535) nslab = 65535;
  const int64_t slab = (stages + nslab - 1) / nslab * GR_KS;
  nslab = (ncols + slab - 1) / slab;
  double2 *part = nullptr;
  FQEB_CUDA(cudaMallocAsync((void **)&part, sizeof(double2) * (size_t)nslab * M * N, st));
  if (herm) FQEB_CUDA(cudaMemsetAsync(part, 0, sizeof(double2) * (size_t)nslab * M * N, st));
  const size_t smem = sizeof(double) * (size_t)GR_STAGES * 2 * GR_PANEL;
  static PerDeviceSize attr_dev;
  int attr_dev_id = 0;
  if (attr_dev.needs(smem, &attr_dev_id)) {
    FQEB_CUDA(cudaFuncSetAttribute(k_gram, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // two CTAs of 108 KB per SM: ask for the largest shared-memory carve-out
    FQEB_CUDA(cudaFuncSetAttribute(k_gram, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_dev.record(attr_dev_id, smem);
  }
  k_gram<<<dim3((unsigned)tiles, (unsigned)nslab), GR_THREADS, smem, st>>>(
      M, N, ncols, slab, (const double2 *)d_bra, ldb, (const double2 *)d_ket, ldk,
      (const double2 *)d_ket_last, part, herm ? 1 : 0);
  FQEB_CHECK_LAUNCH();
  const int64_t n = (int64_t)M * N;
  k_gram_reduce<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, (int)nslab, M, N,
                                                             herm ? GR_TM : 0, part,
                                                             (double2 *)d_G);
  FQEB_CHECK_LAUNCH();
  FQEB_CUDA(cudaFreeAsync(part, st));
  return FQEB_OK;
}
