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
//   * CTA tile BM (real rows) x 128 (real cols), 8 warps, 4-stage cp.async pipeline,
//     k-step 16.  "wide" shape: 128 x 128, warps 2(M) x 4(N), warp tile 8x4 DMMA tiles
//     (64 accumulators per thread, 146 KB smem).  "tall" shapes: 8w x 128 with every
//     warp spanning all rows (w x 2 DMMA tiles), w picked from a small menu so that
//     the operator's row count is covered without a ragged block.
//   * shared-memory strides (20 / 132 / 264 doubles) make every fragment load
//     bank-conflict free for 64-bit accesses.
//   * persistent grid (one CTA per SM) with the pipeline running across tile
//     boundaries; tiles are ordered M-fastest so the CTAs sharing a D tile run
//     together and D is read from HBM once.
//   * pair-symmetric operators (real-orbital integrals) are contracted in the
//     compressed i>=j pair space: norb(norb+1)/2 instead of norb^2 on both M and K.
#include "fqeb_common.cuh"

#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>
#include <type_traits>
#include <vector>

namespace fqeb {

constexpr int BNR = 128;      // real columns per CTA tile
constexpr int KSTEP = 16;     // real k per pipeline stage (default geometry)
constexpr int KSTEP_MAX = 32; // largest k-step of any instantiated pipeline (padding)
constexpr int STAGES = 4;
constexpr int A_STRIDE = KSTEP + 4;   // doubles
constexpr int B_STRIDE_R = BNR + 4;   // REAL: [16][132]
constexpr int B_STRIDE_C = 2 * BNR + 8;  // CPLX: [8][264]
constexpr int B_TILE = 16 * B_STRIDE_R;          // == 8 * B_STRIDE_C
static_assert(16 * B_STRIDE_R == 8 * B_STRIDE_C, "B tile size mismatch");
constexpr int COL_ALIGN = 128;  // leading dimensions (complex elements) must be multiples
#ifndef FQEB_FRAG_PREFETCH
#define FQEB_FRAG_PREFETCH 3
#endif
constexpr int MAX_BM = 144;     // tallest CTA tile of the menu below (operand row padding)

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

__device__ __forceinline__ int ldg_int_g(const int *p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ldg_c128_if_g(bool pred, const double2 *p) {
  double2 v = make_double2(0.0, 0.0);
  asm volatile(
      "{\n .reg .pred q;\n setp.ne.b32 q, %3, 0;\n @q ld.global.nc.v2.f64 {%0,%1}, [%2];\n}"
      : "+d"(v.x), "+d"(v.y)
      : "l"(p), "r"((int)pred));
  return v;
}

// exact sign flip: XOR the sign bit of the map entry into the double
__device__ __forceinline__ double flip_sign_g(double v, int t) {
  return __hiloint2double(__double2hiint(v) ^ (t & (int)0x80000000), __double2loint(v));
}

// ---- fragment loads with compile-time offsets ------------------------------------------
// Written as asm with an immediate offset from ONE base register per operand: left to
// itself the compiler, short of registers next to 136 accumulator registers,
// rematerialises every fragment address from (stage, row, lane) with 3 dependent IMADs
// in front of each LDS (seen in SASS / ncu source view), which sits on the critical
// path of the DMMA stream.
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F &&f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

template <int OFF>
__device__ __forceinline__ double lds_f64(unsigned base) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(base), "n"(OFF));
  return v;
}

// All DMMAs of one pipeline stage for one warp.  a_base / b_base: shared-memory byte
// addresses of this lane's first A / B fragment element in the stage.
//
// The A fragments rotate through PF register slots and the B fragments through two sets, so
// that the LDS of the fragment PF-1 steps ahead (and of the next k4 step's B fragments) is in
// flight while the DMMAs of the current one issue.  With a single A register (the obvious
// loop) every DMMA pair waits a full shared-memory latency (~30 cycles against 31 cycles of
// tensor-pipe time for the pair), which the second warp of the sub-partition hides only when
// the two stay perfectly out of phase: measured 31.1 -> see DESIGN.md.  Loads past kk_count
// stay inside the stage buffer and are never consumed.
template <bool CPLX, int WM, int WN, bool RAGGED, int KS = KSTEP>
__device__ __forceinline__ void mma_stage(double (&acc)[WM][WN][2], unsigned a_base,
                                          unsigned b_base, int kk_count, int mt_active) {
  constexpr int A_STRIDE = KS + 4;  // shadows the KSTEP=16 constant: row stride of this KS
  constexpr int NKK = KS / 4;
  if constexpr (RAGGED) {
    static_for<0, NKK>([&](auto kk_c) {
      constexpr int kk = decltype(kk_c)::value;
      if (kk < kk_count) {
        double bf[WN];
        static_for<0, WN>([&](auto nt_c) {
          constexpr int nt = decltype(nt_c)::value;
          constexpr int off = CPLX ? (kk * 2 * B_STRIDE_C + nt * 16) * 8
                                   : (kk * 4 * B_STRIDE_R + nt * 8) * 8;
          bf[nt] = lds_f64<off>(b_base);
        });
        static_for<0, WM>([&](auto mt_c) {
          constexpr int mt = decltype(mt_c)::value;
          if (mt < mt_active) {
            const double af = lds_f64<(mt * 8 * A_STRIDE + kk * 4) * 8>(a_base);
#pragma unroll
            for (int nt = 0; nt < WN; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af, bf[nt]);
          }
        });
      }
    });
  } else {
    // A-fragment slots in flight (one less next to 144 accumulator registers)
    constexpr int PF = (WM * WN * 4 > 140 && FQEB_FRAG_PREFETCH > 2) ? 2 : FQEB_FRAG_PREFETCH;
    static_assert(PF >= 1 && PF <= WM, "prefetch depth");
    double af[PF];
    double bf[2][WN];
    auto load_b = [&](auto kk_c, auto set_c) {
      constexpr int kk = decltype(kk_c)::value;
      constexpr int set = decltype(set_c)::value;
      static_for<0, WN>([&](auto nt_c) {
        constexpr int nt = decltype(nt_c)::value;
        constexpr int off = CPLX ? (kk * 2 * B_STRIDE_C + nt * 16) * 8
                                 : (kk * 4 * B_STRIDE_R + nt * 8) * 8;
        bf[set][nt] = lds_f64<off>(b_base);
      });
    };
    // prologue: B fragments of k4 step 0 and the first PF A fragments
    load_b(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
    static_for<0, PF>([&](auto i_c) {
      constexpr int i = decltype(i_c)::value;
      af[i] = lds_f64<(i * 8 * A_STRIDE) * 8>(a_base);
    });
    static_for<0, NKK>([&](auto kk_c) {
      constexpr int kk = decltype(kk_c)::value;
      if (kk < kk_count) {
        if constexpr (kk + 1 < NKK)
          load_b(std::integral_constant<int, kk + 1>{}, std::integral_constant<int, (kk + 1) & 1>{});
        static_for<0, WM>([&](auto mt_c) {
          constexpr int mt = decltype(mt_c)::value;
          constexpr int slot = (kk * WM + mt) % PF;
#pragma unroll
          for (int nt = 0; nt < WN; ++nt)
            dmma884(acc[mt][nt][0], acc[mt][nt][1], af[slot], bf[kk & 1][nt]);
          // refill the slot with the fragment PF steps ahead (next k4 step past the last row)
          constexpr int nxt = mt + PF;
          if constexpr (nxt < WM)
            af[slot] = lds_f64<(nxt * 8 * A_STRIDE + kk * 4) * 8>(a_base);
          else if constexpr (kk + 1 < NKK)
            af[slot] = lds_f64<((nxt - WM) * 8 * A_STRIDE + (kk + 1) * 4) * 8>(a_base);
        });
      }
    });
  }
}

// CTA tile = BM x 128 real elements, 8 warps arranged WARPS_M x (8/WARPS_M); each
// warp owns WM x WN DMMA (8x8) tiles.  Two shapes are instantiated:
//   wide  WARPS_M=2, WM=8, WN=4 : 128 x 128, for row counts that are multiples of 128
//   tall  WARPS_M=1, WM=w, WN=2 : 8w x 128, every warp spans all rows; w is chosen so
//         that the operator's row count is covered without a ragged last block
//         (e.g. the 136 compressed pairs of norb=16 -> w=17, one block).
//
// A: real row-major, leading dimension lda (doubles); a_col0 = first real k of the slice.
// B: complex [.][ldb] (D rows of the slice start at row 0).  E: complex [.][lde].
// m_valid: real rows of A that carry data (multiple of 8; of 16 for CPLX).
// k_valid: real k extent of the slice (the last stage may be partial: only whole k4
//          steps that contain data are issued).  nrows_out: valid complex output rows.
//
// The kernel is grid-strided over tiles (M-fastest order, so the CTAs that share a D
// tile run at the same time) and the cp.async pipeline runs CONTINUOUSLY across tile
// boundaries: launched with one CTA per SM it is persistent - the first stages of
// the next tile are in flight while the last stages of the current one are in the
// tensor cores and the epilogue's stores overlap those loads; launched with one CTA
// per tile it degenerates to the classic kernel.
template <bool CPLX, int WARPS_M, int WM, int WN, bool RAGGED>
__global__ void __launch_bounds__(256, 1)
k_dgemm(const double *__restrict__ A, int lda, int a_col0, const double2 *__restrict__ B,
        int64_t ldb, double2 *__restrict__ E, int64_t lde, int m_valid, int nrows_out, int k_valid,
        int nmb, int64_t ntiles) {
  constexpr int BM = WARPS_M * WM * 8;
  constexpr int WARPS_N = 8 / WARPS_M;
  static_assert(WARPS_N * WN * 8 == BNR, "warp layout must span 128 columns");
  static_assert(!CPLX || (WM % 2 == 0), "complex mode pairs re/im row tiles");
  constexpr int A_TILE = BM * A_STRIDE;
  constexpr int STAGE_DOUBLES = A_TILE + B_TILE;
  constexpr int TILE_DETS = CPLX ? BNR : BNR / 2;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int wm0 = (warp / WARPS_N) * (WM * 8);
  const int wn0 = (warp % WARPS_N) * (WN * 8);
  const int nk = (k_valid + KSTEP - 1) / KSTEP;

  if ((int64_t)blockIdx.x >= ntiles) return;
  const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t total_it = my_tiles * nk;

  double acc[WM][WN][2];
#pragma unroll
  for (int i = 0; i < WM; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // tile walk without divisions: tile -> (mb, nb), advanced by (mb_step, nb_step)
  const int mb_step = (int)(gridDim.x % nmb);
  const int64_t nb_step = gridDim.x / nmb;

  // ---- producer: per-thread source/destination of its cp.async slots -----------------
  // slot i of A covers rows i*32 + tid/8 (8 x 16-byte chunks per row); slot i of B covers
  // rows i*B_ROWS + tid/B_CHUNKS.  Only one pointer per operand is thread-dependent.
  constexpr int A_SLOTS = (BM * 8 + 255) / 256;
  constexpr int B_CHUNKS = CPLX ? 128 : 64;     // 16-byte chunks per B row
  constexpr int B_ROWS = 256 / B_CHUNKS;        // B rows covered by one slot
  constexpr int B_STRIDE = CPLX ? B_STRIDE_C : B_STRIDE_R;
  constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
  const unsigned smem_u32 = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned a_dst0 = smem_u32 + ((tid >> 3) * A_STRIDE + (tid & 7) * 2) * 8;
  const unsigned b_dst0 =
      smem_u32 + (A_TILE + (tid / B_CHUNKS) * B_STRIDE + (tid % B_CHUNKS) * 2) * 8;
  const double *a_thr = A + a_col0 + (int64_t)(tid >> 3) * lda + (tid & 7) * 2;
  const double2 *b_thr = B + (int64_t)(tid / B_CHUNKS) * ldb + (tid % B_CHUNKS);
  const int64_t b_stage_stride = (int64_t)(CPLX ? KSTEP / 2 : KSTEP) * ldb;

  int ld_mb = (int)(blockIdx.x % nmb);
  int64_t ld_nb = blockIdx.x / nmb;
  int ld_kt = 0, ld_stage = 0;
  const double *a_src = a_thr + (int64_t)(ld_mb * BM) * lda;
  const double2 *b_src = b_thr + ld_nb * TILE_DETS;
  auto issue_load = [&]() {
    const unsigned sa = a_dst0 + ld_stage * STAGE_BYTES;
    const unsigned sb = b_dst0 + ld_stage * STAGE_BYTES;
#pragma unroll
    for (int i = 0; i < A_SLOTS; ++i) {
      if ((BM * 8) % 256 == 0 || i < A_SLOTS - 1 || (tid >> 3) + i * 32 < BM)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                         sa + i * 32 * A_STRIDE * 8),
                     "l"(a_src + (int64_t)(i * 32) * lda));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                       sb + i * B_ROWS * B_STRIDE * 8),
                   "l"(b_src + (int64_t)(i * B_ROWS) * ldb));
    }
    ld_stage = (ld_stage + 1 == STAGES) ? 0 : ld_stage + 1;
    a_src += KSTEP;
    b_src += b_stage_stride;
    if (++ld_kt == nk) {
      ld_kt = 0;
      ld_mb += mb_step;
      ld_nb += nb_step;
      if (ld_mb >= nmb) {
        ld_mb -= nmb;
        ld_nb += 1;
      }
      a_src = a_thr + (int64_t)(ld_mb * BM) * lda;
      b_src = b_thr + ld_nb * TILE_DETS;
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < total_it) issue_load();
    cp_async_commit();
  }

  // ---- consumer state ------------------------------------------------------------------
  const unsigned a_frag_off = ((wm0 + g) * A_STRIDE + tg) * 8;
  const unsigned b_frag_off =
      (A_TILE + (CPLX ? (tg >> 1) * B_STRIDE_C + (wn0 + g) * 2 + (tg & 1)
                      : tg * B_STRIDE_R + wn0 + g)) * 8;
  int mb = (int)(blockIdx.x % nmb);
  int64_t nb = blockIdx.x / nmb;
  int kt = 0, stage = 0;
  int m0 = mb * BM;
  int mt_active = (m_valid - (m0 + wm0)) / 8;
  mt_active = mt_active < 0 ? 0 : (mt_active > WM ? WM : mt_active);

  for (int64_t it = 0; it < total_it; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (it + STAGES - 1 < total_it) issue_load();
    cp_async_commit();
    const unsigned stage_u32 = smem_u32 + stage * STAGE_BYTES;
    stage = (stage + 1 == STAGES) ? 0 : stage + 1;
    // whole k4 steps of this stage that contain data
    int kk_count = (k_valid - kt * KSTEP + 3) / 4;
    kk_count = kk_count > KSTEP / 4 ? KSTEP / 4 : kk_count;
    mma_stage<CPLX, WM, WN, RAGGED>(acc, stage_u32 + a_frag_off, stage_u32 + b_frag_off, kk_count,
                                    mt_active);

    if (++kt == nk) {
      // ---- epilogue of this tile: interleaved complex128, 16-byte stores -----------
      const int64_t n0 = nb * TILE_DETS;
      if (CPLX) {
#pragma unroll
        for (int p = 0; p < WM / 2; ++p) {
          if (!RAGGED || 2 * p < mt_active) {
            const int kl = ((m0 + wm0) >> 1) + p * 8 + g;
            if (kl < nrows_out) {
              double2 *erow = E + (int64_t)kl * lde + n0 + wn0 + 2 * tg;
#pragma unroll
              for (int nt = 0; nt < WN; ++nt) {
                __stcs(erow + nt * 8, make_double2(acc[2 * p][nt][0], acc[2 * p + 1][nt][0]));
                __stcs(erow + nt * 8 + 1, make_double2(acc[2 * p][nt][1], acc[2 * p + 1][nt][1]));
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int mt = 0; mt < WM; ++mt) {
          if (!RAGGED || mt < mt_active) {
            const int kl = m0 + wm0 + mt * 8 + g;
            if (kl < nrows_out) {
              double2 *erow = E + (int64_t)kl * lde + n0 + (wn0 >> 1) + tg;
#pragma unroll
              for (int nt = 0; nt < WN; ++nt)
                __stcs(erow + nt * 4, make_double2(acc[mt][nt][0], acc[mt][nt][1]));
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      kt = 0;
      mb += mb_step;
      nb += nb_step;
      if (mb >= nmb) {
        mb -= nmb;
        nb += 1;
      }
      m0 = mb * BM;
      mt_active = (m_valid - (m0 + wm0)) / 8;
      mt_active = mt_active < 0 ? 0 : (mt_active > WM ? WM : mt_active);
    }
  }
  cp_async_wait<0>();
}

// epilogue store of one complex element of E (streaming: E is read back once, by the scatter)
__device__ __forceinline__ void store_e(double2 *p, double2 v) {
#if defined(FQEB_EXP_NOSTORE)
  if (v.x == 1.2345e300) __stcs(p, v);   // experiment: measure the epilogue's cost
#elif defined(FQEB_EXP_PLAINSTORE)
  *p = v;
#else
  __stcs(p, v);
#endif
}

// ---- warp-specialised variant --------------------------------------------------------
// Same tiles, same math, different plumbing: one PRODUCER warpgroup (two warps stream A
// tiles, two stream D tiles, cp.async) feeds eight CONSUMER warps (two warpgroups) through a
// ring of STAGES
// shared-memory slots guarded by mbarriers (full[s]: data landed, empty[s]: all
// consumer warps are done with slot s).  There is no CTA-wide barrier in the main
// loop: consumer warps drift apart instead of hitting their non-tensor phases (fragment
// loads, epilogue) in lock-step, and they issue nothing but LDS + DMMA.
__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count));
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned addr) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(addr) : "memory");
}

template <bool CPLX, int WARPS_M, int WM, int WN, bool RAGGED, int KS, int NST>
__global__ void __launch_bounds__(384, 1)
k_dgemm_ws(const double *__restrict__ A, int lda, int a_col0, const double2 *__restrict__ B,
           int64_t ldb, double2 *__restrict__ E, int64_t lde, int m_valid, int nrows_out,
           int k_valid, int nmb, int64_t ntiles) {
  // pipeline geometry of this instantiation (shadow the KSTEP=16 / STAGES=4 constants)
  constexpr int KSTEP = KS, STAGES = NST;
  constexpr int A_STRIDE = KS + 4;
  constexpr int B_TILE = KS * B_STRIDE_R;  // == (KS/2) * B_STRIDE_C
  constexpr int BM = WARPS_M * WM * 8;
  constexpr int WARPS_N = 8 / WARPS_M;
  static_assert(WARPS_N * WN * 8 == BNR, "warp layout must span 128 columns");
  static_assert(!CPLX || (WM % 2 == 0), "complex mode pairs re/im row tiles");
  constexpr int A_TILE = BM * A_STRIDE;
  constexpr int STAGE_DOUBLES = A_TILE + B_TILE;
  constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
  constexpr int TILE_DETS = CPLX ? BNR : BNR / 2;
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int nk = (k_valid + KSTEP - 1) / KSTEP;
  const unsigned smem_u32 = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned bar_full = smem_u32 + STAGES * STAGE_BYTES;   // STAGES x 8 bytes
  const unsigned bar_empty = bar_full + STAGES * 8;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 128);  // 4 producer warps, one arrive per lane
      mbar_init(bar_empty + s * 8, 8);   // one arrive per consumer warp
    }
  }
  __syncthreads();
  if ((int64_t)blockIdx.x >= ntiles) return;
  const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int64_t total_it = my_tiles * nk;
  const int mb_step = (int)(gridDim.x % nmb);
  const int64_t nb_step = gridDim.x / nmb;
  int mb = (int)(blockIdx.x % nmb);
  int64_t nb = blockIdx.x / nmb;

  if (warp < 4) {
    // =========================== producers (warpgroup 0) =============================
    // give the registers this warpgroup does not need to the two consumer warpgroups
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;\n");
    const bool loads_a = (warp < 2);
    const int half = warp & 1;                  // two warps share each operand
    constexpr int B_CHUNKS = CPLX ? 128 : 64;   // 16-byte chunks per B row
    constexpr int B_STRIDE = CPLX ? B_STRIDE_C : B_STRIDE_R;
    constexpr int B_SUB = B_CHUNKS / 32;        // lane-strided pieces per B row
    constexpr int B_ROWS_STAGE = CPLX ? KSTEP / 2 : KSTEP;
    // A: a row of the tile is KSTEP/2 16-byte chunks; one pass of the warp covers
    // A_RPP = 32 / (KSTEP/2) rows: chunk id = lane + 32*i -> row = A_RPP*i + lane/A_CPR
    constexpr int A_CPR = KSTEP / 2;      // chunks per row
    constexpr int A_RPP = 32 / A_CPR;     // rows per pass
    static_assert(32 % A_CPR == 0 && BM % A_RPP == 0, "operand tile / warp mismatch");
    const unsigned a_dst0 = smem_u32 + ((lane / A_CPR) * A_STRIDE + (lane % A_CPR) * 2) * 8;
    const double *a_thr = A + a_col0 + (int64_t)(lane / A_CPR) * lda + (lane % A_CPR) * 2;
    // B: piece j of row r -> chunk lane + 32*j
    const unsigned b_dst0 = smem_u32 + (A_TILE + lane * 2) * 8;
    const double2 *b_thr = B + lane;
    int kt = 0, stage = 0;
    unsigned round_parity = 1;  // parity of (round-1) with round = 0 -> no wait first time
    bool first_round = true;
    const double *a_src = a_thr + (int64_t)(mb * BM) * lda;
    const double2 *b_src = b_thr + nb * TILE_DETS;
    for (int64_t it = 0; it < total_it; ++it) {
      if (!first_round) mbar_wait(bar_empty + stage * 8, round_parity);
      const unsigned sbase = stage * STAGE_BYTES;
      if (loads_a) {
#pragma unroll 8
        for (int i = half; i < BM / A_RPP; i += 2)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                           a_dst0 + sbase + i * A_RPP * A_STRIDE * 8),
                       "l"(a_src + (int64_t)(i * A_RPP) * lda));
      } else {
#pragma unroll 8
        for (int i = half; i < B_ROWS_STAGE * B_SUB; i += 2) {
          const int r = i / B_SUB, j = i % B_SUB;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                           b_dst0 + sbase + (r * B_STRIDE + j * 64) * 8),
                       "l"(b_src + (int64_t)r * ldb + j * 32));
        }
      }
      cp_async_mbar_arrive(bar_full + stage * 8);
      a_src += KSTEP;
      b_src += (int64_t)B_ROWS_STAGE * ldb;
      if (++stage == STAGES) {
        stage = 0;
        round_parity ^= 1;
        first_round = false;
      }
      if (++kt == nk) {
        kt = 0;
        mb += mb_step;
        nb += nb_step;
        if (mb >= nmb) {
          mb -= nmb;
          nb += 1;
        }
        a_src = a_thr + (int64_t)(mb * BM) * lda;
        b_src = b_thr + nb * TILE_DETS;
      }
    }
    asm volatile("cp.async.wait_all;\n" ::);
    return;
  }

  // ============================= consumers (warpgroups 1, 2) ============================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 208;\n");
  const int cw = warp - 4;
  const int g = lane >> 2, tg = lane & 3;
  const int wm0 = (cw / WARPS_N) * (WM * 8);
  const int wn0 = (cw % WARPS_N) * (WN * 8);
  const unsigned a_frag_off = ((wm0 + g) * A_STRIDE + tg) * 8;
  const unsigned b_frag_off =
      (A_TILE + (CPLX ? (tg >> 1) * B_STRIDE_C + (wn0 + g) * 2 + (tg & 1)
                      : tg * B_STRIDE_R + wn0 + g)) * 8;
  double acc[WM][WN][2];
#pragma unroll
  for (int i = 0; i < WM; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  int kt = 0, stage = 0;
  unsigned parity = 0;
  int m0 = mb * BM;
  int mt_active = (m_valid - (m0 + wm0)) / 8;
  mt_active = mt_active < 0 ? 0 : (mt_active > WM ? WM : mt_active);

  for (int64_t it = 0; it < total_it; ++it) {
    mbar_wait(bar_full + stage * 8, parity);
    const unsigned stage_u32 = smem_u32 + stage * STAGE_BYTES;
    int kk_count = (k_valid - kt * KSTEP + 3) / 4;
    kk_count = kk_count > KSTEP / 4 ? KSTEP / 4 : kk_count;
    mma_stage<CPLX, WM, WN, RAGGED, KS>(acc, stage_u32 + a_frag_off, stage_u32 + b_frag_off, kk_count,
                                    mt_active);
    // release the slot: every lane's fragment loads have been consumed by its DMMAs
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + stage * 8);
    if (++stage == STAGES) {
      stage = 0;
      parity ^= 1;
    }

    if (++kt == nk) {
      const int64_t n0 = nb * TILE_DETS;
      if (CPLX) {
#pragma unroll
        for (int p = 0; p < WM / 2; ++p) {
          if (!RAGGED || 2 * p < mt_active) {
            const int kl = ((m0 + wm0) >> 1) + p * 8 + g;
            if (kl < nrows_out) {
              double2 *erow = E + (int64_t)kl * lde + n0 + wn0 + 2 * tg;
#pragma unroll
              for (int nt = 0; nt < WN; ++nt) {
                store_e(erow + nt * 8, make_double2(acc[2 * p][nt][0], acc[2 * p + 1][nt][0]));
                store_e(erow + nt * 8 + 1, make_double2(acc[2 * p][nt][1], acc[2 * p + 1][nt][1]));
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int mt = 0; mt < WM; ++mt) {
          if (!RAGGED || mt < mt_active) {
            const int kl = m0 + wm0 + mt * 8 + g;
            if (kl < nrows_out) {
              double2 *erow = E + (int64_t)kl * lde + n0 + (wn0 >> 1) + tg;
#pragma unroll
              for (int nt = 0; nt < WN; ++nt)
                store_e(erow + nt * 4, make_double2(acc[mt][nt][0], acc[mt][nt][1]));
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      kt = 0;
      mb += mb_step;
      nb += nb_step;
      if (mb >= nmb) {
        mb -= nmb;
        nb += 1;
      }
      m0 = mb * BM;
      mt_active = (m_valid - (m0 + wm0)) / 8;
      mt_active = mt_active < 0 ? 0 : (mt_active > WM ? WM : mt_active);
    }
  }
}

// ---- tile-shape menu ---------------------------------------------------------------
struct GemmShape {
  int warps_m, wm;
  int bm() const { return warps_m * wm * 8; }
};
static const GemmShape kShapes[] = {{2, 8}, {1, 18}, {1, 17}, {1, 14}, {1, 13}, {1, 10}};

template <bool CPLX, int WARPS_M, int WM, int WN, int KS, int NST>
static int launch_ws(const fqeb_op *op, const double *d_A, int a_col0, const double *d_dvec, int64_t ldd,
                     double *d_evec, int64_t lde, int m_valid, int k_valid, int64_t nnb,
                     cudaStream_t st) {
  constexpr int BM = WARPS_M * WM * 8;
  constexpr size_t SMEM =
      sizeof(double) * (size_t)(BM * (KS + 4) + KS * B_STRIDE_R) * NST + 16 * NST;
  static_assert(SMEM <= 227 * 1024, "pipeline does not fit in shared memory");
  const bool ragged = (m_valid % BM) != 0;
  auto kern = ragged ? k_dgemm_ws<CPLX, WARPS_M, WM, WN, true, KS, NST>
                     : k_dgemm_ws<CPLX, WARPS_M, WM, WN, false, KS, NST>;
  static PerDeviceSize attr_dev;
  int attr_dev_id = 0;
  if (attr_dev.needs(1, &attr_dev_id)) {
    FQEB_CUDA(cudaFuncSetAttribute(k_dgemm_ws<CPLX, WARPS_M, WM, WN, true, KS, NST>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    FQEB_CUDA(cudaFuncSetAttribute(k_dgemm_ws<CPLX, WARPS_M, WM, WN, false, KS, NST>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    attr_dev.record(attr_dev_id, 1);
  }
  const int nmb = (m_valid + BM - 1) / BM;
  const int64_t tiles = nnb * nmb;
  int64_t grid = sm_count();
  if (grid > nmb) grid -= grid % nmb;
  if (grid > tiles) grid = tiles;
  kern<<<(unsigned)grid, 384, SMEM, st>>>(d_A, op->Kp, a_col0, (const double2 *)d_dvec, ldd,
                                          (double2 *)d_evec, lde, m_valid, op->np, k_valid, nmb,
                                          tiles);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

template <bool CPLX, int WARPS_M, int WM, int WN>
static int launch_shape(const fqeb_op *op, const double *d_A, int a_col0, const double *d_dvec, int64_t ldd,
                        double *d_evec, int64_t lde, int m_valid, int k_valid, int64_t nnb,
                        cudaStream_t st) {
  constexpr int BM = WARPS_M * WM * 8;
  // FQEB_GEMM_WS=0 selects the classic (all warps load and compute) variant;
  // FQEB_GEMM_PIPE picks the warp-specialised pipeline geometry: 1 = k-step 32 x 3 stages
  // (default: fewest stage hand-overs per DMMA; 31.8 / 35.0 TFLOP/s at norb=16),
  // 0 = k-step 16 x 4 stages (31.0 / 33.6), 2 = k-step 16 x 5 stages (30.8 / 33.3)
  static const bool ws = !(getenv("FQEB_GEMM_WS") && getenv("FQEB_GEMM_WS")[0] == '0');
  static const int pipe = getenv("FQEB_GEMM_PIPE") ? atoi(getenv("FQEB_GEMM_PIPE")) : 1;
  if (ws) {
    if (pipe == 1)
      return launch_ws<CPLX, WARPS_M, WM, WN, 32, 3>(op, d_A, a_col0, d_dvec, ldd, d_evec, lde,
                                                     m_valid, k_valid, nnb, st);
    if (pipe == 2)
      return launch_ws<CPLX, WARPS_M, WM, WN, 16, 5>(op, d_A, a_col0, d_dvec, ldd, d_evec, lde,
                                                     m_valid, k_valid, nnb, st);
    return launch_ws<CPLX, WARPS_M, WM, WN, 16, 4>(op, d_A, a_col0, d_dvec, ldd, d_evec, lde, m_valid,
                                                   k_valid, nnb, st);
  }
  constexpr size_t SMEM = sizeof(double) * (size_t)(BM * A_STRIDE + B_TILE) * STAGES;
  const bool ragged = (m_valid % BM) != 0;
  auto kern = ragged ? k_dgemm<CPLX, WARPS_M, WM, WN, true> : k_dgemm<CPLX, WARPS_M, WM, WN, false>;
  static PerDeviceSize attr_dev;
  int attr_dev_id = 0;
  if (attr_dev.needs(1, &attr_dev_id)) {
    FQEB_CUDA(cudaFuncSetAttribute(k_dgemm<CPLX, WARPS_M, WM, WN, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    FQEB_CUDA(cudaFuncSetAttribute(k_dgemm<CPLX, WARPS_M, WM, WN, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    attr_dev.record(attr_dev_id, 1);
  }
  const int nmb = (m_valid + BM - 1) / BM;
  const int64_t tiles = nnb * nmb;
  // persistent: one CTA per SM, rounded down to a multiple of nmb so that the CTAs
  // sharing a D tile stay aligned.  FQEB_GEMM_PERSISTENT=0 launches one CTA per tile.
  static const bool persistent = !(getenv("FQEB_GEMM_PERSISTENT") &&
                                   getenv("FQEB_GEMM_PERSISTENT")[0] == '0');
  int64_t grid = tiles;
  if (persistent) {
    grid = sm_count();
    if (grid > nmb) grid -= grid % nmb;
    if (grid > tiles) grid = tiles;
  }
  FQEB_REQUIRE(grid < (1ll << 31), "contract: too many tiles for one launch");
  kern<<<(unsigned)grid, 256, SMEM, st>>>(d_A, op->Kp, a_col0, (const double2 *)d_dvec, ldd,
                                          (double2 *)d_evec, lde, m_valid, op->np, k_valid, nmb,
                                          tiles);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

// rows covered per block and the resulting padded row count for a shape
static int padded_rows(const GemmShape &s, int m_valid) {
  const int bm = s.bm();
  return (m_valid + bm - 1) / bm * bm;
}

static GemmShape pick_shape(bool cplx, int m_valid) {
  GemmShape best = kShapes[0];
  int best_rows = padded_rows(best, m_valid);
  for (const GemmShape &s : kShapes) {
    if (cplx && (s.wm & 1)) continue;
    const int rows = padded_rows(s, m_valid);
    if (rows < best_rows) {
      best = s;
      best_rows = rows;
    }
  }
  return best;
}

// ---- fused gather + contraction ----------------------------------------------------------
// The D tensor is never written to HBM: two PRODUCER warpgroups (256 threads) build each
// 32-pair x 64-determinant tile of D directly in the shared-memory ring, as signed
// single-source gathers from C (exactly what k_gather computes, from the plain or the merged
// pair-space map), and stream the operand tile next to it with cp.async, while the two
// CONSUMER warpgroups run the DMMA stream of k_dgemm_ws on the previous stages.  setmaxnreg
// splits the register file 80 / 176 between the two roles.  Applies when one CTA covers all
// rows of the operator (pair space <= 144, real / imaginary class: norb=16 with real-orbital
// integrals), so that no D tile is gathered twice.  The one-body term is folded into the
// operand beforehand (absorbed_operand), so sigma gets everything through E.  Saves writing
// and re-reading D (2 x 360 GB per sigma at norb=16) and the whole gather launch.
//
// Column space: determinants of a chunk are laid out with a row pitch that is a multiple
// of 64, so a tile never straddles two alpha rows:  col = r*pitch + b.
// Producer thread p: determinant p & 63 of the tile, pair rows (p >> 6) + 4*j of every stage.
template <int WM, bool RAGGED>
__global__ void __launch_bounds__(512, 1)
k_sigma_fused(const double *__restrict__ A, int lda, int a_col0, int c_first, int k_valid,
              const int32_t *__restrict__ mapT_a, const int32_t *__restrict__ map_b, int ntab,
              const double2 *__restrict__ coeff, int64_t lenb, int64_t row0, int pitch,
              int tiles_per_row, double2 *__restrict__ E, int64_t lde, int m_valid,
              int nrows_out, int64_t ntiles) {
  constexpr int KS = 32, NST = 3, WN = 2;
  constexpr int BM = WM * 8;
  constexpr int A_STRIDE = KS + 4;
  constexpr int A_TILE = BM * A_STRIDE;
  constexpr int B_TILE = KS * B_STRIDE_R;
  constexpr int STAGE_BYTES = (A_TILE + B_TILE) * 8;
  constexpr int TILE_DETS = BNR / 2;  // 64
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int nk = (k_valid + KS - 1) / KS;
  const unsigned smem_u32 = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned bar_full = smem_u32 + NST * STAGE_BYTES;
  const unsigned bar_empty = bar_full + NST * 8;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_full + s * 8, 512);  // 256 producer threads x (cp.async arrive + arrive)
      mbar_init(bar_empty + s * 8, 8);   // one arrive per consumer warp
    }
  }
  __syncthreads();
  if ((int64_t)blockIdx.x >= ntiles) return;
  const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  int64_t nb = blockIdx.x;

  if (warp >= 8) {
    // =========================== producers (warpgroups 2, 3) =========================
// diagnostics: compile the alpha / beta source loads out to see what bounds the kernel
#ifndef FQEB_FUSED_DIAG_A
#define FQEB_FUSED_DIAG_A true
#endif
#ifndef FQEB_FUSED_DIAG_B
#define FQEB_FUSED_DIAG_B true
#endif
#ifndef FQEB_FUSED_PREG
#define FQEB_FUSED_PREG 80
#endif
#ifndef FQEB_FUSED_CREG
#define FQEB_FUSED_CREG 176   /* 256 - FQEB_FUSED_PREG */
#endif
#define FQEB_STR2(x) #x
#define FQEB_STR(x) FQEB_STR2(x)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 " FQEB_STR(FQEB_FUSED_PREG) ";\n");
    const int pt = tid - 256;
    const int det = pt & 63, phase = pt >> 6;
    int stage = 0;
    unsigned round_parity = 1;
    bool first_round = true;
    // Software pipeline over "trips" (4 pair rows of one determinant): the map entries of
    // trip n+1 are requested before the C elements of trip n are consumed, across stage and
    // tile boundaries, so a trip costs one global round trip instead of two dependent ones.
    struct Coords {
      const int32_t *ta_row, *mb;
      const double2 *crow, *ccol;
      bool valid;
    };
    auto tile_coords = [&](int64_t tile) {
      Coords c;
      const int r = (int)(tile / tiles_per_row);
      const int bt = (int)(tile - (int64_t)r * tiles_per_row);
      const int64_t a = row0 + r;
      const int64_t bb = (int64_t)bt * TILE_DETS + det;
      c.valid = bb < lenb;
      c.ta_row = mapT_a + a * (int64_t)ntab + c_first;
      c.crow = coeff + a * lenb;
      c.ccol = coeff + (c.valid ? bb : 0);
      c.mb = map_b + (int64_t)c_first * lenb + (c.valid ? bb : 0);
      return c;
    };
    auto load_maps = [&](const Coords &c, int kbase, int (&ta_)[4], int (&tb_)[4]) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = kbase + 4 * u;
        const bool on = k < k_valid;
        ta_[u] = on ? ldg_int_g(c.ta_row + k) : 0;
        tb_[u] = (on && c.valid) ? ldg_int_g(c.mb + (int64_t)k * lenb) : 0;
      }
    };
    Coords cur = tile_coords(nb);
    int ta[4], tb[4];
    load_maps(cur, phase, ta, tb);
    for (int64_t t = 0; t < my_tiles; ++t) {
      for (int kt = 0; kt < nk; ++kt) {
        if (!first_round) mbar_wait(bar_empty + stage * 8, round_parity);
        const unsigned sbase = smem_u32 + stage * STAGE_BYTES;
        // operand tile: BM rows x 16 chunks of 16 bytes, spread over the 256 producer threads
        {
          const double *a_src = A + a_col0 + kt * KS;
#pragma unroll 3
          for (int q = pt; q < BM * (KS / 2); q += 256) {
            const int row = q >> 4, cc = q & 15;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                             sbase + (row * A_STRIDE + cc * 2) * 8),
                         "l"(a_src + (int64_t)row * lda + cc * 2));
          }
          cp_async_mbar_arrive(bar_full + stage * 8);
        }
        // D tile: this thread's 8 pair rows of the stage, two trips of four
        const unsigned d_dst = sbase + (A_TILE + phase * B_STRIDE_R + 2 * det) * 8;
#pragma unroll
        for (int trip = 0; trip < 2; ++trip) {
          double2 va[4], vb[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            va[u] = ldg_c128_if_g(FQEB_FUSED_DIAG_A && ta[u] != 0 && cur.valid,
                                  cur.ccol + (int64_t)(abs(ta[u]) - 1) * lenb);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            vb[u] = ldg_c128_if_g(FQEB_FUSED_DIAG_B && tb[u] != 0, cur.crow + (abs(tb[u]) - 1));
          // maps of the next trip: second half of this stage, or the first half of the next
          // stage (possibly of the next tile)
          int nta[4], ntb[4];
          Coords nxt = cur;
          if (trip == 0) {
            load_maps(cur, kt * KS + phase + 16, nta, ntb);
          } else if (kt + 1 < nk) {
            load_maps(cur, (kt + 1) * KS + phase, nta, ntb);
          } else if (t + 1 < my_tiles) {
            nxt = tile_coords(nb + gridDim.x);
            load_maps(nxt, phase, nta, ntb);
          } else {
#pragma unroll
            for (int u = 0; u < 4; ++u) nta[u] = ntb[u] = 0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double re = flip_sign_g(va[u].x, ta[u]) + flip_sign_g(vb[u].x, tb[u]);
            const double im = flip_sign_g(va[u].y, ta[u]) + flip_sign_g(vb[u].y, tb[u]);
            asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(
                             d_dst + (4 * (4 * trip + u)) * B_STRIDE_R * 8),
                         "d"(re), "d"(im)
                         : "memory");
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            ta[u] = nta[u];
            tb[u] = ntb[u];
          }
          cur = nxt;
        }
        mbar_arrive(bar_full + stage * 8);
        if (++stage == NST) {
          stage = 0;
          round_parity ^= 1;
          first_round = false;
        }
      }
      nb += gridDim.x;
    }
    asm volatile("cp.async.wait_all;\n" ::);
    return;
  }

  // ============================= consumers (warpgroups 0, 1) ===========================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 " FQEB_STR(FQEB_FUSED_CREG) ";\n");
  const int g = lane >> 2, tg = lane & 3;
  const int wn0 = warp * (WN * 8);
  const unsigned a_frag_off = (g * A_STRIDE + tg) * 8;
  const unsigned b_frag_off = (A_TILE + tg * B_STRIDE_R + wn0 + g) * 8;
  double acc[WM][WN][2];
#pragma unroll
  for (int i = 0; i < WM; ++i)
#pragma unroll
    for (int jn = 0; jn < WN; ++jn) acc[i][jn][0] = acc[i][jn][1] = 0.0;
  int stage = 0;
  unsigned parity = 0;
  int mt_active = m_valid / 8;
  mt_active = mt_active > WM ? WM : mt_active;
  for (int64_t t = 0; t < my_tiles; ++t, nb += gridDim.x) {
    for (int kt = 0; kt < nk; ++kt) {
      mbar_wait(bar_full + stage * 8, parity);
      const unsigned stage_u32 = smem_u32 + stage * STAGE_BYTES;
      int kk_count = (k_valid - kt * KS + 3) / 4;
      kk_count = kk_count > KS / 4 ? KS / 4 : kk_count;
      mma_stage<false, WM, WN, RAGGED, KS>(acc, stage_u32 + a_frag_off, stage_u32 + b_frag_off,
                                           kk_count, mt_active);
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + stage * 8);
      if (++stage == NST) {
        stage = 0;
        parity ^= 1;
      }
    }
    const int64_t n0 = nb * TILE_DETS;
#pragma unroll
    for (int mt = 0; mt < WM; ++mt) {
      if (!RAGGED || mt < mt_active) {
        const int kl = mt * 8 + g;
        if (kl < nrows_out) {
          double2 *erow = E + (int64_t)kl * lde + n0 + (wn0 >> 1) + tg;
#pragma unroll
          for (int nt = 0; nt < WN; ++nt)
            store_e(erow + nt * 4, make_double2(acc[mt][nt][0], acc[mt][nt][1]));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int jn = 0; jn < WN; ++jn) acc[i][jn][0] = acc[i][jn][1] = 0.0;
  }
}

template <int WM>
static int launch_fused_wm(const double *d_A, int lda, int a_col0, int c_first, int k_valid,
                           const fqeb_graph *g, const fqeb_op *op, const double *d_coeff,
                           int64_t row0, int64_t nrows, int pitch, double *d_evec, int64_t lde,
                           int m_valid, cudaStream_t st) {
  constexpr int BM = WM * 8;
  const size_t smem = sizeof(double) * (size_t)(BM * (32 + 4) + 32 * B_STRIDE_R) * 3 + 16 * 3;
  const bool ragged = m_valid < BM;
  auto kern = ragged ? k_sigma_fused<WM, true> : k_sigma_fused<WM, false>;
  static PerDeviceSize attr_dev;
  int attr_dev_id = 0;
  if (attr_dev.needs(1, &attr_dev_id)) {
    FQEB_CUDA(cudaFuncSetAttribute(k_sigma_fused<WM, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    FQEB_CUDA(cudaFuncSetAttribute(k_sigma_fused<WM, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_dev.record(attr_dev_id, 1);
  }
  FQEB_REQUIRE(smem <= 227 * 1024, "fused sigma: shared memory budget exceeded");
  const int tiles_per_row = pitch / 64;
  const int64_t tiles = nrows * tiles_per_row;
  int64_t grid = sm_count();
  if (grid > tiles) grid = tiles;
  kern<<<(unsigned)grid, 512, smem, st>>>(
      d_A, lda, a_col0, c_first, k_valid, op->sym ? g->d_smapT[0] : g->d_amapT[0],
      op->sym ? g->d_smap[1] : g->d_amap[1], op->np, (const double2 *)d_coeff, g->len[1], row0,
      pitch, tiles_per_row, (double2 *)d_evec, lde, m_valid, op->np, tiles);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

// fused gather + contraction of alpha rows [row0, row0+nrows) for pairs [ij0, ij1).
// d_A: operand with the one-body term absorbed (absorbed_operand).
int launch_fused(const fqeb_graph *g, const fqeb_op *op, const double *d_A, const double *d_coeff,
                 int64_t row0, int64_t nrows, int pitch, double *d_evec, int64_t lde, int ij0,
                 int ij1, cudaStream_t st) {
  FQEB_REQUIRE(op->kind != FQEB_OP_COMPLEX, "fused sigma: real / imaginary operators only");
  FQEB_REQUIRE((ij0 & 1) == 0 && ij0 < ij1 && ij1 <= op->np, "fused sigma: bad pair slice");
  FQEB_REQUIRE(pitch % 64 == 0 && pitch >= g->len[1] && lde >= nrows * (int64_t)pitch,
               "fused sigma: bad column layout");
  const int m_valid = (int)round_up(op->np, 8);
  const int k_valid = ij1 - ij0;
  const GemmShape sh = pick_shape(false, m_valid);
  FQEB_REQUIRE(sh.warps_m == 1 && m_valid <= sh.bm(), "fused sigma: pair space too large");
#define FQEB_FUSED(WMV)                                                                        \
  if (sh.wm == WMV)                                                                            \
    return launch_fused_wm<WMV>(d_A, op->Kp, ij0, ij0, k_valid, g, op, d_coeff, row0, nrows,   \
                                pitch, d_evec, lde, m_valid, st);
  FQEB_FUSED(18)
  FQEB_FUSED(17)
  FQEB_FUSED(14)
  FQEB_FUSED(13)
  FQEB_FUSED(10)
#undef FQEB_FUSED
  set_error("fused sigma: no kernel for tile height %d", sh.wm);
  return FQEB_ERR_INVALID;
}

// can the fused kernel cover this operator's row space with one block?
bool fused_shape_ok(const fqeb_op *op) {
  if (!op->has_h2 || op->kind == FQEB_OP_COMPLEX) return false;
  const int m_valid = (int)round_up(op->np, 8);
  const GemmShape sh = pick_shape(false, m_valid);
  return sh.warps_m == 1 && m_valid <= sh.bm();
}

// d_A: operand override (absorbed_operand), or nullptr for the operator's own h2' operand
int launch_contract(const fqeb_op *op, const double *d_A, const double *d_dvec, int64_t ldd,
                    double *d_evec, int64_t lde, int64_t ncols, int ij0, int ij1,
                    cudaStream_t st) {
  const int np = op->np;
  if (!d_A) d_A = op->d_A;
  FQEB_REQUIRE(op->has_h2, "contract: operator has no two-body part");
  FQEB_REQUIRE(ij0 >= 0 && ij0 < ij1 && ij1 <= np, "contract: pair slice [%d,%d) invalid", ij0,
               ij1);
  FQEB_REQUIRE(ldd % COL_ALIGN == 0 && lde % COL_ALIGN == 0,
               "contract: leading dimensions must be multiples of %d", COL_ALIGN);
  FQEB_REQUIRE(ncols >= 0 && ncols <= ldd && ncols <= lde, "contract: ncols exceeds ld");
  FQEB_REQUIRE((ij0 & 1) == 0, "contract: pair slice must start at an even index");
  if (ncols == 0) return FQEB_OK;
  const bool cplx = op->kind == FQEB_OP_COMPLEX;
  const int nij = ij1 - ij0;
  const int k_valid = cplx ? 2 * nij : nij;
  const int a_col0 = cplx ? 2 * ij0 : ij0;
  FQEB_REQUIRE(a_col0 + round_up(k_valid, KSTEP_MAX) <= op->Kp,
               "contract: operator padding too small");
  const int m_valid = cplx ? 2 * (int)round_up(np, 8) : (int)round_up(np, 8);
  const int64_t cols_pad = round_up(ncols, COL_ALIGN);
  const int64_t nnb = cols_pad / (cplx ? BNR : BNR / 2);
  const GemmShape sh = pick_shape(cplx, m_valid);
  FQEB_REQUIRE(padded_rows(sh, m_valid) <= op->Mp, "contract: operator row padding too small");
#define FQEB_SHAPE_BOTH(WMV, WARPSM, WNV)                                                      \
  if (sh.warps_m == WARPSM && sh.wm == WMV) {                                                  \
    return cplx ? launch_shape<true, WARPSM, WMV, WNV>(op, d_A, a_col0, d_dvec, ldd, d_evec, lde,   \
                                                       m_valid, k_valid, nnb, st)              \
                : launch_shape<false, WARPSM, WMV, WNV>(op, d_A, a_col0, d_dvec, ldd, d_evec, lde,  \
                                                        m_valid, k_valid, nnb, st);            \
  }
#define FQEB_SHAPE_REAL(WMV, WARPSM, WNV)                                                      \
  if (!cplx && sh.warps_m == WARPSM && sh.wm == WMV) {                                         \
    return launch_shape<false, WARPSM, WMV, WNV>(op, d_A, a_col0, d_dvec, ldd, d_evec, lde,         \
                                                 m_valid, k_valid, nnb, st);                   \
  }
  FQEB_SHAPE_BOTH(8, 2, 4)
  FQEB_SHAPE_BOTH(18, 1, 2)
  FQEB_SHAPE_REAL(17, 1, 2)
  FQEB_SHAPE_BOTH(14, 1, 2)
  FQEB_SHAPE_REAL(13, 1, 2)
  FQEB_SHAPE_BOTH(10, 1, 2)
#undef FQEB_SHAPE_BOTH
#undef FQEB_SHAPE_REAL
  set_error("contract: no kernel instantiated for tile shape %dx%d", sh.warps_m, sh.wm);
  return FQEB_ERR_INVALID;
}

// rows of D a caller must allocate for a slice of nij pairs (whole pipeline stages) ...
int dvec_rows_padded(const fqeb_op *op, int nij) {
  const bool cplx = op->kind == FQEB_OP_COMPLEX;
  return (int)round_up(nij, cplx ? KSTEP_MAX / 2 : KSTEP_MAX);
}
// ... of which rows [nij, this) are read by a partial k4 step and must be zero
int dvec_rows_zeroed(const fqeb_op *op, int nij) {
  const bool cplx = op->kind == FQEB_OP_COMPLEX;
  return (int)round_up(nij, cplx ? 2 : 4);
}

}  // namespace fqeb

using namespace fqeb;

extern "C" int fqeb_gemm_col_align(void) { return COL_ALIGN; }

extern "C" int fqeb_contract_dvec_rows(const fqeb_op *op, int nij) {
  if (!op || nij < 0) return -1;
  return dvec_rows_padded(op, nij);
}

extern "C" int fqeb_op_create(int norb, const double *h_h1p, const double *h_h2p, fqeb_op **out) {
  return fqeb_op_create_ex(norb, h_h1p, h_h2p, 0, out);
}

extern "C" int fqeb_op_create_ex(int norb, const double *h_h1p, const double *h_h2p, int flags,
                                 fqeb_op **out) {
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
  op->sym = false;
  op->np = npair;
  cudaGetDevice(&op->device);
  auto fail = [&](int code) {
    fqeb_op_destroy(op);
    return code;
  };
  if (upload_alloc((void **)&op->d_h1, h_h1p, sizeof(double) * 2 * npair) != FQEB_OK)
    return fail(FQEB_ERR_CUDA);
  auto h2 = [&](int i, int j, int k, int l, int part) {
    return h_h2p[2 * ((((size_t)i * norb + j) * norb + k) * norb + l) + part];
  };
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
    // pair symmetry h2'[ij,kl] == h2'[ji,kl] == h2'[ij,lk] (exact): real-orbital
    // integrals.  Then E[ij] == E[ji] and only i>=j pairs are contracted, against
    // D[ij]+D[ji]: the compressed algorithm of the reference's real branch
    // (fqe_data.py:659-681), valid here for any coefficient class.
    bool sym = getenv("FQEB_NO_SYMMETRY") == nullptr && !(flags & FQEB_OP_FLAG_FULL_PAIR_SPACE);
    for (int i = 0; i < norb && sym; ++i)
      for (int j = 0; j < norb && sym; ++j)
        for (int k = 0; k < norb && sym; ++k)
          for (int l = 0; l < norb && sym; ++l)
            for (int part = 0; part < 2; ++part)
              if (h2(i, j, k, l, part) != h2(j, i, k, l, part) ||
                  h2(i, j, k, l, part) != h2(i, j, l, k, part))
                sym = false;
    op->sym = sym;
    op->np = sym ? norb * (norb + 1) / 2 : npair;
  }
  // pair tables
  {
    const int np = op->np;
    std::vector<int32_t> tab(2 * (size_t)np + npair);
    int32_t *pairs = tab.data(), *rowmap = tab.data() + 2 * np;
    if (op->sym) {
      for (int i = 0; i < norb; ++i)
        for (int j = 0; j <= i; ++j) {
          const int c = i * (i + 1) / 2 + j;
          pairs[2 * c] = i * norb + j;
          pairs[2 * c + 1] = (i == j) ? -1 : j * norb + i;
          rowmap[i * norb + j] = c;
          rowmap[j * norb + i] = c;
        }
    } else {
      for (int p = 0; p < npair; ++p) {
        pairs[2 * p] = p;
        pairs[2 * p + 1] = -1;
        rowmap[p] = p;
      }
    }
    if (upload_alloc((void **)&op->d_pairs, tab.data(), sizeof(int32_t) * tab.size()) != FQEB_OK ||
        upload_finish() != FQEB_OK)   // tab goes out of scope
      return fail(FQEB_ERR_CUDA);
    op->d_rowmap = op->d_pairs + 2 * np;
  }
  if (op->has_h2) {
    const int np = op->np;
    // (row c, col d) of the contraction operand: h2'[pair(c), pair(d)]
    auto elem = [&](int c, int d, int part) {
      int ij, kl;
      if (op->sym) {
        // invert c = i(i+1)/2 + j
        int i = 0;
        while ((i + 1) * (i + 2) / 2 <= c) ++i;
        ij = i * norb + (c - i * (i + 1) / 2);
        int k = 0;
        while ((k + 1) * (k + 2) / 2 <= d) ++k;
        kl = k * norb + (d - k * (k + 1) / 2);
      } else {
        ij = c;
        kl = d;
      }
      return h_h2p[2 * ((size_t)ij * npair + kl) + part];
    };
    std::vector<double> a;
    if (op->kind == FQEB_OP_COMPLEX) {
      const int np8 = (int)round_up(np, 8);
      op->Mp = (int)round_up(2 * np8, MAX_BM) + MAX_BM;
      op->Kp = (int)round_up(2 * np, KSTEP_MAX) + KSTEP_MAX;  // slack for ragged slices
      a.assign((size_t)op->Mp * op->Kp, 0.0);
      for (int c = 0; c < np; ++c) {
        const int rre = (c / 8) * 16 + (c % 8), rim = rre + 8;
        for (int d = 0; d < np; ++d) {
          const double re = elem(c, d, 0), im = elem(c, d, 1);
          a[(size_t)rre * op->Kp + 2 * d] = re;
          a[(size_t)rre * op->Kp + 2 * d + 1] = -im;
          a[(size_t)rim * op->Kp + 2 * d] = im;
          a[(size_t)rim * op->Kp + 2 * d + 1] = re;
        }
      }
    } else {
      const int off = op->kind == FQEB_OP_IMAG ? 1 : 0;
      op->Mp = (int)round_up(round_up(np, 8), MAX_BM) + MAX_BM;
      op->Kp = (int)round_up(np, KSTEP_MAX) + KSTEP_MAX;
      a.assign((size_t)op->Mp * op->Kp, 0.0);
      for (int c = 0; c < np; ++c)
        for (int d = 0; d < np; ++d) a[(size_t)c * op->Kp + d] = elem(c, d, off);
    }
    if (upload_alloc((void **)&op->d_A, a.data(), sizeof(double) * a.size()) != FQEB_OK ||
        upload_finish() != FQEB_OK)   // a goes out of scope
      return fail(FQEB_ERR_CUDA);
  }
  // Can the one-body term be absorbed into the contraction operand,
  //     A[c, d] += h1'[pair(c)] / n_elec   for every diagonal pair d = (k, k)
  // (exact on a fixed-particle-number sector because sum_k E_kk = n_elec)?  It must not
  // leave the operator's class (real / imaginary) nor, in the compressed pair space,
  // break the i<->j symmetry.  If not, sigma.cu adds the one-body term with its own kernel.
  op->absorb_ok = false;
  if (op->has_h2) {
    bool ok = true;
    if (op->kind != FQEB_OP_COMPLEX) {
      const int part_zero = op->kind == FQEB_OP_IMAG ? 0 : 1;  // component that must vanish
      for (int p = 0; p < npair && ok; ++p) ok = (h_h1p[2 * p + part_zero] == 0.0);
    }
    if (op->sym)
      for (int i = 0; i < norb && ok; ++i)
        for (int j = 0; j < i && ok; ++j)
          ok = h_h1p[2 * (i * norb + j)] == h_h1p[2 * (j * norb + i)] &&
               h_h1p[2 * (i * norb + j) + 1] == h_h1p[2 * (j * norb + i) + 1];
    op->absorb_ok = ok;
  }
  // fused gather+contraction (opt-in): additionally needs a single row block
  op->fuse_ok = op->absorb_ok && fused_shape_ok(op);
  if (op->absorb_ok) {
    op->h_h1p = (double *)malloc(sizeof(double) * 2 * npair);
    op->h_h2p = (double *)malloc(sizeof(double) * 2 * (size_t)npair * npair);
    if (!op->h_h1p || !op->h_h2p) {
      set_error("fqeb_op_create: host allocation failed");
      return fail(FQEB_ERR_NOMEM);
    }
    memcpy(op->h_h1p, h_h1p, sizeof(double) * 2 * npair);
    memcpy(op->h_h2p, h_h2p, sizeof(double) * 2 * (size_t)npair * npair);
    op->fused_cache = new std::map<int, double *>();
  }
  if (upload_finish() != FQEB_OK) return fail(FQEB_ERR_CUDA);
  *out = op;
  return FQEB_OK;
}

namespace fqeb {
// Contraction operand for a sector with n_elec electrons: h2' in the operator's pair space
// (same layout as fqeb_op::d_A) with the one-body term absorbed into the diagonal-pair
// columns.  Built on first use and cached per n_elec.
static std::mutex g_operand_cache_mu;   // guards every operator's fused_cache map
int absorbed_operand(const fqeb_op *op, int n_elec, const double **d_A) {
  FQEB_REQUIRE(op->absorb_ok && n_elec > 0, "absorbed_operand: one-body term not absorbable");
  std::lock_guard<std::mutex> lock(g_operand_cache_mu);
  auto *cache = static_cast<std::map<int, double *> *>(op->fused_cache);
  auto it = cache->find(n_elec);
  if (it != cache->end()) {
    *d_A = it->second;
    return FQEB_OK;
  }
  const int norb = op->norb, npair = norb * norb, np = op->np;
  auto pair_of = [&](int c) {
    if (!op->sym) return c;
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= c) ++i;
    return i * norb + (c - i * (i + 1) / 2);
  };
  auto elem = [&](int c, int d, int part) {
    const int ij = pair_of(c), kl = pair_of(d);
    double v = op->h_h2p[2 * ((size_t)ij * npair + kl) + part];
    if (kl / norb == kl % norb) v += op->h_h1p[2 * ij + part] / (double)n_elec;
    return v;
  };
  std::vector<double> a((size_t)op->Mp * op->Kp, 0.0);
  if (op->kind == FQEB_OP_COMPLEX) {
    for (int c = 0; c < np; ++c) {
      const int rre = (c / 8) * 16 + (c % 8), rim = rre + 8;
      for (int d = 0; d < np; ++d) {
        const double re = elem(c, d, 0), im = elem(c, d, 1);
        a[(size_t)rre * op->Kp + 2 * d] = re;
        a[(size_t)rre * op->Kp + 2 * d + 1] = -im;
        a[(size_t)rim * op->Kp + 2 * d] = im;
        a[(size_t)rim * op->Kp + 2 * d + 1] = re;
      }
    }
  } else {
    const int off = op->kind == FQEB_OP_IMAG ? 1 : 0;
    for (int c = 0; c < np; ++c)
      for (int d = 0; d < np; ++d) a[(size_t)c * op->Kp + d] = elem(c, d, off);
  }
  double *dev = nullptr;
  int rc = upload_alloc((void **)&dev, a.data(), sizeof(double) * a.size());
  if (rc == FQEB_OK) rc = upload_finish();
  if (rc != FQEB_OK) return rc;
  (*cache)[n_elec] = dev;
  *d_A = dev;
  return FQEB_OK;
}
}  // namespace fqeb

// Device buffers are returned with cudaFreeAsync on `stream`: the release is ordered after the
// work already enqueued there (the kernels that may still read the operator) and does not
// synchronise the device.
namespace fqeb {
void ozaki_forget(const fqeb_op *op, cudaStream_t st);
}

extern "C" int fqeb_op_destroy_async(fqeb_op *op, void *stream) {
  if (!op) return FQEB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  fqeb::ozaki_forget(op, st);
  if (op->fused_cache) {
    auto *cache = static_cast<std::map<int, double *> *>(op->fused_cache);
    for (auto &kv : *cache) cudaFreeAsync(kv.second, st);
    delete cache;
  }
  free(op->h_h1p);
  free(op->h_h2p);
  if (op->d_A) cudaFreeAsync(op->d_A, st);
  if (op->d_h1) cudaFreeAsync(op->d_h1, st);
  if (op->d_pairs) cudaFreeAsync(op->d_pairs, st);
  free(op);
  return FQEB_OK;
}

// Blocking variant: waits for the device, so the operator may be destroyed whatever stream
// used it last.
extern "C" int fqeb_op_destroy(fqeb_op *op) {
  if (!op) return FQEB_OK;
  cudaDeviceSynchronize();
  return fqeb_op_destroy_async(op, nullptr);
}

extern "C" int fqeb_op_kind(const fqeb_op *op, int *kind) {
  FQEB_REQUIRE(op && kind, "fqeb_op_kind: NULL argument");
  *kind = op->kind;
  return FQEB_OK;
}

extern "C" int fqeb_op_pair_space(const fqeb_op *op, int *npairs, int *symmetric) {
  FQEB_REQUIRE(op, "fqeb_op_pair_space: NULL argument");
  if (npairs) *npairs = op->np;
  if (symmetric) *symmetric = op->sym ? 1 : 0;
  return FQEB_OK;
}

extern "C" int fqeb_contract(const fqeb_op *op, const double *d_dvec, int64_t ldd,
                             double *d_evec, int64_t lde, int64_t ncols, int ij0, int ij1,
                             void *stream) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(op && d_dvec && d_evec, "fqeb_contract: NULL argument");
  return launch_contract(op, nullptr, d_dvec, ldd, d_evec, lde, ncols, ij0, ij1,
                         (cudaStream_t)stream);
}
