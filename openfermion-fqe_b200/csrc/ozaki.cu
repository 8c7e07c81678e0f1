// INT8-sliced two-electron contraction on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// Same mathematics as k_sigma_fused (dgemm.cu): for a chunk of alpha rows
//
//     E[kl, det] = sum_ij A[kl, ij] * D[ij, det],   D[ij, det] = +-C[y_a, b] +- C[a, y_b]
//
// (reference: numpy.einsum("ijkl,klmn->ijmn", h2e, dvec), src/fqe/fqe_data.py:656, on the dvec of
// fqe_data.py:2209-2234; D is gathered on the fly and never reaches HBM).  sm_100a has no
// tcgen05 path for FP64 and the DMMA contraction already runs at ~90 % of the
// 37 TFLOP/s FP64 issue limit, so this kernel takes the only route past that roofline: the FP64
// product is evaluated EXACTLY in integer arithmetic on digit slices (an Ozaki-type scheme):
//
//     C / S = sum_i c_i R^-(i+1),   A / T = sum_j a_j R^-(j+1),   R = 127, |c_i|, |a_j| <= 63
//     D digits = +-c_i(alpha source) +- c_i(beta source)           |.| <= 126: int8
//     E = S T sum_{i+j <= DMAX} R^-(i+j+2) (a_j . d_i)             int8 x int8 -> int32, exact
//
// with NS = 6 slices and DMAX = 5 (21 slice products).  The only errors are the two
// quantisations (relative 127^-6 / 2 = 1.2e-13 of max|C| and of max|A|) and the dropped products
// i + j > DMAX (same order); measured against the FP64 oracle: 1.3e-12 relative on sigma for
// uniform random states (profiles/microbench/ozaki_proto.py).  The state's scale is GLOBAL, so a
// state dominated by a few determinants loses relative accuracy on its small coefficients;
// sigma.cu therefore takes this path only when  127^-6 * max|C| * sqrt(ndet) / ||C||  is below
// a threshold and the FP64 DMMA kernels otherwise (both are parity-tested).
//
// Kernel anatomy (one persistent CTA per SM, 17 warps):
//   * the digit planes of the operand A (B operand of the MMA: N = pairs kl, K = pairs ij) are
//     copied once per CTA into shared memory in the canonical K-major no-swizzle UMMA layout;
//   * per tile (one alpha row x 64 beta strings = 128 real rows m = (det, re|im)), 16 worker
//     warps gather the 8-byte digit words of the alpha and the beta source of every (m, ij) from
//     the pre-sliced coefficient planes, add them as packed biased bytes, transpose 4 x 4 bytes
//     with PRMT and store 16-byte core-matrix rows of the D^T tile (A operand: M = 128, K = ij);
//   * ONE thread issues the 21 x K/32 tcgen05.mma kind::i8 instructions of the tile, diagonal by
//     diagonal (d = i + j) into three rotating 144-column accumulators in TMEM;
//   * the same 16 warps drain the accumulators two diagonals at a time (acc_d * 127 + acc_{d+1}
//     fits int32), convert to FP64, apply the weights and stream E out, overlapped with the MMAs
//     of the following diagonals.
#include "fqeb_common.cuh"

#include <math.h>
#include <string.h>
#include <map>
#include <mutex>
#include <vector>

namespace fqeb {

constexpr int OZ_NS = 6;           // digit slices of C and of the operand
constexpr int OZ_DMAX = 5;         // slice products (i, j) with i + j <= OZ_DMAX are kept
constexpr int OZ_RADIX = 127;
constexpr int OZ_NMAX = 136;       // largest pair space (MMA N and K) whose tile + operand image fit in shared memory
constexpr int OZ_SLOT = 144;       // TMEM columns per accumulator slot
constexpr int OZ_WORKERS = 512;    // 16 worker warps: producers, then epilogue
constexpr int OZ_THREADS = OZ_WORKERS + 128;   // + one warpgroup: MMA issuer (3 warps idle)
constexpr int OZ_TILE_DETS = 64;   // determinants per tile (128 real rows)
constexpr int OZ_NPROD = (OZ_DMAX + 1) * (OZ_DMAX + 2) / 2;   // 21 slice products
constexpr int OZ_MAX_MMAS = OZ_NPROD * 5;                     // x K steps (<= 5)

// ---------------------------------------------------------------------------------------
// coefficient digit planes
// ---------------------------------------------------------------------------------------
// planes[(sign * ndet + det) * 2 + part] : 8 bytes, byte s = biased digit (d_s + 64) of slice s of
// the real (part 0) / imaginary (part 1) part of +C (sign 0) or -C (sign 1); bytes 6, 7 = 64.
// (real and imaginary words of a determinant are adjacent: the two lanes of a determinant read
// one 16-byte element, and 16 consecutive beta strings touch 2.3 cache lines on average.)
// Biased digits of two sources add without carries between bytes (<= 254), and
// (sum ^ 0x80) is the two's-complement sum of the two signed digits.

__global__ void k_absmax_sumsq(int64_t n2, const double *__restrict__ x, double *__restrict__ part) {
  // part[2*block] = max |x|, part[2*block+1] = sum x^2 over this block's grid-stride share
  double mx = 0.0, ss = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double v = x[i];
    mx = fmax(mx, fabs(v));
    ss += v * v;
  }
  __shared__ double s_mx[32], s_ss[32];
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_mx[warp] = mx;
    s_ss[warp] = ss;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    mx = lane < nw ? s_mx[lane] : 0.0;
    ss = lane < nw ? s_ss[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (lane == 0) {
      part[2 * blockIdx.x] = mx;
      part[2 * blockIdx.x + 1] = ss;
    }
  }
}

__global__ void k_absmax_sumsq_final(int nblocks, const double *__restrict__ part,
                                     double *__restrict__ out) {
  // deterministic second pass: out[0] = max, out[1] = sum of squares, out[2] = scale S
  double mx = 0.0, ss = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 32) {
    mx = fmax(mx, part[2 * i]);
    ss += part[2 * i + 1];
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (threadIdx.x == 0) {
    out[0] = mx;
    out[1] = ss;
    out[2] = 2.0001 * mx;   // |x| / S < 0.5: inside the range of OZ_NS balanced digits
  }
}

__device__ __forceinline__ uint64_t balanced_digits(double x, double scale_pow) {
  // biased digit bytes (most significant slice in byte 0) of x, |x| < 0.5
  long long n = __double2ll_rn(x * scale_pow);
  uint64_t w = 0x4040000000000000ull;   // bytes 6, 7: digit 0
#pragma unroll
  for (int s = OZ_NS - 1; s >= 0; --s) {
    long long q = n / OZ_RADIX;
    long long d = n - q * OZ_RADIX;               // truncated remainder in (-127, 127)
    if (d > 63) {
      d -= OZ_RADIX;
      q += 1;
    } else if (d < -63) {
      d += OZ_RADIX;
      q -= 1;
    }
    n = q;
    w |= (uint64_t)(d + 64) << (8 * s);
  }
  return w;
}

// One block slices a 32 x 32 block of determinants and writes it twice: in the coefficient layout
// (planes, [a][b]) and transposed (planesT, [b][a]), both with coalesced 16-byte stores.
__global__ void __launch_bounds__(256) k_slice_coeff(int64_t lena, int64_t lenb,
                                                      const double2 *__restrict__ coeff,
                                                      const double *__restrict__ stats,
                                                      ulonglong2 *__restrict__ planes,
                                                      ulonglong2 *__restrict__ planesT) {
  __shared__ ulonglong2 tile[32][33];
  const double s = stats[2];
  const double inv = s > 0.0 ? 1.0 / s : 0.0;
  double pw = 1.0;
#pragma unroll
  for (int i = 0; i < OZ_NS; ++i) pw *= (double)OZ_RADIX;
  const int64_t ndet = lena * lenb;
  const int64_t nbt = (lenb + 31) / 32;
  const int64_t a0 = (blockIdx.x / nbt) * 32, b0 = (blockIdx.x % nbt) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const uint64_t k128 = 0x8080808080808080ull;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t a = a0 + ty + 8 * r, b = b0 + tx;
    if (a < lena && b < lenb) {
      const double2 c = coeff[a * lenb + b];
      const uint64_t re = balanced_digits(c.x * inv, pw), im = balanced_digits(c.y * inv, pw);
      // digits of -x: 128 - u per byte, no borrow between bytes (every byte <= 127)
      planes[a * lenb + b] = make_ulonglong2(re, im);
      planes[ndet + a * lenb + b] = make_ulonglong2(k128 - re, k128 - im);
      tile[ty + 8 * r][tx] = make_ulonglong2(re, im);
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int64_t a = a0 + tx, b = b0 + ty + 8 * r;
    if (a < lena && b < lenb) {
      const ulonglong2 v = tile[tx][ty + 8 * r];
      planesT[b * lena + a] = v;
      planesT[ndet + b * lena + a] = make_ulonglong2(k128 - v.x, k128 - v.y);
    }
  }
}

// zero-padded copy of a by-string map: dst[x][0..kpad) = src[x][0..np), 0 beyond
__global__ void k_pad_map(int64_t len, int np, int kpad, const int32_t *__restrict__ src,
                          int32_t *__restrict__ dst) {
  const int64_t n = len * kpad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t x = i / kpad;
    const int k = (int)(i - x * kpad);
    dst[i] = k < np ? src[x * np + k] : 0;
  }
}

// ---------------------------------------------------------------------------------------
// operand digit planes (host side, cached per operator and electron count)
// ---------------------------------------------------------------------------------------
struct OzOperand {
  int8_t *d_img;     // shared-memory image: [slice][kcol][row group][8][16] + zero block
  size_t img_bytes;
  double scale;      // T
  int np, kc, ng, n_mma;
};

static std::mutex g_oz_mu;
static std::map<std::pair<const fqeb_op *, int>, OzOperand> g_oz_cache;

int absorbed_operand(const fqeb_op *op, int n_elec, const double **d_A);

static void host_digits(double x, double scale_pow, int8_t *out) {
  long long n = llrint(x * scale_pow);
  for (int s = OZ_NS - 1; s >= 0; --s) {
    long long q = n / OZ_RADIX, d = n - q * OZ_RADIX;
    if (d > 63) {
      d -= OZ_RADIX;
      q += 1;
    } else if (d < -63) {
      d += OZ_RADIX;
      q -= 1;
    }
    n = q;
    out[s] = (int8_t)d;
  }
}

bool ozaki_shape_ok(const fqeb_op *op) {
  return op->has_h2 && op->kind != FQEB_OP_COMPLEX && op->absorb_ok && op->np >= 1 &&
         op->np <= OZ_NMAX;
}

// digit image of the contraction operand with the one-body term absorbed (same matrix as
// absorbed_operand builds for the DMMA kernels)
static int ozaki_operand(const fqeb_op *op, int n_elec, OzOperand *out) {
  std::lock_guard<std::mutex> lock(g_oz_mu);
  auto key = std::make_pair(op, n_elec);
  auto it = g_oz_cache.find(key);
  if (it != g_oz_cache.end()) {
    *out = it->second;
    return FQEB_OK;
  }
  const int norb = op->norb, npair = norb * norb, np = op->np;
  auto pair_of = [&](int c) {
    if (!op->sym) return c;
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= c) ++i;
    return i * norb + (c - i * (i + 1) / 2);
  };
  const int part = op->kind == FQEB_OP_IMAG ? 1 : 0;
  std::vector<double> a((size_t)np * np);
  double amax = 0.0;
  for (int c = 0; c < np; ++c)
    for (int d = 0; d < np; ++d) {
      const int ij = pair_of(c), kl = pair_of(d);
      double v = op->h_h2p[2 * ((size_t)ij * npair + kl) + part];
      if (kl / norb == kl % norb) v += op->h_h1p[2 * ij + part] / (double)n_elec;
      a[(size_t)c * np + d] = v;
      amax = fmax(amax, fabs(v));
    }
  OzOperand o;
  o.np = np;
  o.kc = (np + 15) / 16;
  o.ng = (np + 7) / 8;
  o.n_mma = (np + 15) / 16 * 16;
  o.scale = 2.0001 * amax;
  const size_t plane = (size_t)o.kc * o.ng * 128;
  const size_t zero_block = (size_t)(o.n_mma / 8) * 128 + 128;
  o.img_bytes = OZ_NS * plane + zero_block;
  std::vector<int8_t> img(o.img_bytes, 0);
  if (amax > 0.0) {
    double pw = 1.0;
    for (int i = 0; i < OZ_NS; ++i) pw *= (double)OZ_RADIX;
    int8_t dg[OZ_NS];
    for (int n = 0; n < np; ++n)        // row of the operand = output pair kl
      for (int k = 0; k < np; ++k) {    // contraction index ij
        host_digits(a[(size_t)n * np + k] / o.scale, pw, dg);
        const size_t off = ((size_t)(k >> 4) * o.ng + (n >> 3)) * 128 + (n & 7) * 16 + (k & 15);
        for (int s = 0; s < OZ_NS; ++s) img[s * plane + off] = dg[s];
      }
  }
  void *dev = nullptr;
  int rc = upload_alloc(&dev, img.data(), img.size());
  if (rc == FQEB_OK) rc = upload_finish();
  if (rc != FQEB_OK) return rc;
  o.d_img = (int8_t *)dev;
  g_oz_cache[key] = o;
  *out = o;
  return FQEB_OK;
}

// called from fqeb_op_destroy_async: drop the cached images of this operator
void ozaki_forget(const fqeb_op *op, cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_oz_mu);
  for (auto it = g_oz_cache.begin(); it != g_oz_cache.end();) {
    if (it->first.first == op) {
      cudaFreeAsync(it->second.d_img, st);
      it = g_oz_cache.erase(it);
    } else {
      ++it;
    }
  }
}

// ---------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t oz_smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void oz_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void oz_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nOZ_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra OZ_DONE;\nbra OZ_WAIT;\nOZ_DONE:\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void oz_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor: K-major, no swizzle, version 1 (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void oz_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void oz_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ uint64_t oz_ldg64(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.global.nc.b64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
// predicated 8-byte load (straight-line code: no branch per source), `dflt` when off
__device__ __forceinline__ uint64_t oz_ldg64_if(bool pred, const uint64_t *p, uint64_t dflt) {
  uint64_t v = dflt;
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q ld.global.nc.b64 %0, [%1];\n}"
               : "+l"(v)
               : "l"(p), "r"((int)pred));
  return v;
}
// predicated 16-byte load of four map entries, zeros when off
__device__ __forceinline__ void oz_ldg128_if(bool pred, const int *p, int &x, int &y, int &z,
                                             int &w) {
  x = y = z = w = 0;
  asm volatile(
      "{\n .reg .pred q;\n setp.ne.b32 q, %5, 0;\n @q ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];\n}"
      : "+r"(x), "+r"(y), "+r"(z), "+r"(w)
      : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ int oz_ldg32_if(bool pred, const int *p) {
  int v = 0;
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n @q ld.global.nc.s32 %0, [%1];\n}"
               : "+r"(v)
               : "l"(p), "r"((int)pred));
  return v;
}
__device__ __forceinline__ int oz_ldg32(const int *p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
#define OZ_TMEM_LD16(r, addr)                                                                  \
  asm volatile(                                                                                \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13," \
      "%14,%15}, [%16];"                                                                       \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),    \
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),             \
        "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                                                  \
      : "r"(addr))
#define OZ_TMEM_LD8(r, addr)                                                                \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"      \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]),      \
                 "=r"(r[6]), "=r"(r[7])                                                       \
               : "r"(addr))
#define OZ_TMEM_LD4(r, addr)                                                      \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"      \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])                   \
               : "r"(addr))

// ---------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------
struct OzParams {
  const int8_t *img;        // operand digit image (global)
  int img_bytes;
  int np, kc, ng, n_mma;    // pair space, 16-byte K columns, 8-row groups, MMA N
  const uint64_t *planes;   // coefficient digit planes [sign][a][b][part]
  const uint64_t *planesT;  // the same, transposed:     [sign][b][a][part]
  int64_t ndet;
  const int32_t *mapT_a;    // [lena][kpad]  alpha adjoint map by string, zero-padded to kpad = 16 kc
  const int32_t *mapT_b;    // [lenb][kpad]  beta adjoint map by string
  int kpad;
  int64_t lena, lenb, row0, nrows;
  int pitch, tiles_per_row;
  int64_t ntiles;
  double2 *E;
  int64_t lde;
  const double *stats;      // stats[2] = S (scale of the coefficient digits)
  double op_scale;          // T
  unsigned long long *prof; // optional [8] cycle counters summed over CTAs (FQEB_OZAKI_PROF=1)
};

__global__ void __launch_bounds__(OZ_THREADS, 1) k_sigma_ozaki(const OzParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_bar[8];
  __shared__ uint32_t s_tmem;
  // (A descriptor, B descriptor) of every MMA of a tile, in issue order: loop-invariant, so the
  // single issuing thread only loads 16 bytes and fires (a dependent ALU chain per MMA was
  // measured at 175 cycles per MMA, 2.4x the 72-cycle tensor time of a 128 x 144 x 32 step)
  __shared__ __align__(16) uint4 s_desc[OZ_MAX_MMAS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kc = p.kc, ng = p.ng;
  const uint32_t d_bytes = (uint32_t)OZ_NS * kc * 2048;   // D^T tile: [slice][kcol][16 groups][128]
  uint8_t *s_d = smem;
  uint8_t *s_b = smem + d_bytes;                           // operand image follows the tile
  const uint32_t bar_dfull = oz_smem_u32(&s_bar[0]), bar_dfree = oz_smem_u32(&s_bar[1]);
  const uint32_t bar_sfull = oz_smem_u32(&s_bar[2]), bar_sfree = oz_smem_u32(&s_bar[5]);

  // one-time setup: operand image -> shared memory, barriers, TMEM
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(p.img);
    uint4 *dst = reinterpret_cast<uint4 *>(s_b);
    for (int i = tid; i < p.img_bytes / 16; i += OZ_THREADS) dst[i] = src[i];
  }
  if (tid == 0) {
    oz_mbar_init(bar_dfull, OZ_WORKERS);
    oz_mbar_init(bar_dfree, 1);
    for (int s = 0; s < 3; ++s) {
      oz_mbar_init(bar_sfull + 8 * s, 1);
      oz_mbar_init(bar_sfree + 8 * s, OZ_WORKERS / 32);
    }
  }
  {
    const uint32_t d_base = oz_smem_u32(s_d), b_base = oz_smem_u32(s_b);
    const uint32_t b_plane = (uint32_t)kc * ng * 128;
    const uint32_t zero_base = b_base + OZ_NS * b_plane;
    const int ksteps = (kc + 1) / 2;
    for (int e = tid; e < OZ_NPROD * ksteps; e += OZ_THREADS) {
      // entry e = (product pr, K step ks); products ordered by diagonal d = i + j, then i
      const int pr = e / ksteps, ks = e - pr * ksteps;
      int d = 0, rem = pr;
      while (rem > d) {
        rem -= d + 1;
        ++d;
      }
      const int i = rem, j = d - rem;   // D slice i, operand slice j
      const uint32_t a0 = d_base + (uint32_t)(i * kc + 2 * ks) * 2048;
      const uint32_t b0 = b_base + (uint32_t)j * b_plane + (uint32_t)(2 * ks) * ng * 128;
      // second 16-byte K column of the step: the next column, or (odd column count) the zero
      // block on the operand side, which cancels whatever the tile side reads there
      const bool tail = (2 * ks + 1 >= kc);
      const uint64_t ad = oz_desc(a0, 2048, 128);
      const uint64_t bd = oz_desc(b0, tail ? zero_base - b0 : (uint32_t)ng * 128, 128);
      s_desc[e] = make_uint4((uint32_t)ad, (uint32_t)(ad >> 32), (uint32_t)bd, (uint32_t)(bd >> 32));
    }
  }
  if (warp == 16) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        oz_smem_u32(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const int64_t my_tiles =
      (int64_t)blockIdx.x < p.ntiles ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp >= 16) {
    // ================================ MMA issuer ======================================
    // this warpgroup hands its registers to the workers (5 warps per scheduler at launch
    // leave 96 registers per thread; the workers need ~110 for 36 running FP64 sums)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n");
    if (warp == 16 && lane == 0) {
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_mma >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      const int ksteps = (kc + 1) / 2;
      long long c_wait_tile = 0, c_wait_slot = 0, c_total = clock64();
      for (int64_t it = 0; it < my_tiles; ++it) {
        long long c0 = clock64();
        oz_mbar_wait(bar_dfull, (uint32_t)(it & 1));
        c_wait_tile += clock64() - c0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        int e = 0;
#pragma unroll 1
        for (int d = 0; d <= OZ_DMAX; ++d) {
          const int slot = d % 3;
          const int64_t use = 2 * it + d / 3;      // how often this slot has been filled before
          if (use >= 1) {
            const long long c1 = clock64();
            oz_mbar_wait(bar_sfree + 8 * slot, (uint32_t)((use - 1) & 1));
            c_wait_slot += clock64() - c1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          const uint32_t acc = tmem + (uint32_t)(slot * OZ_SLOT);
          const int e_end = e + (d + 1) * ksteps;
          uint4 cur = s_desc[e];
          uint32_t first = 0;
#pragma unroll 1
          for (; e < e_end; ++e) {
            const uint4 nxt = s_desc[e + 1 < OZ_MAX_MMAS ? e + 1 : e];   // prefetch
            oz_mma(acc, (uint64_t)cur.x | ((uint64_t)cur.y << 32),
                   (uint64_t)cur.z | ((uint64_t)cur.w << 32), idesc, first);
            first = 1;
            cur = nxt;
          }
          oz_commit(bar_sfull + 8 * slot);
        }
        oz_commit(bar_dfree);
      }
      if (p.prof) {   // issuer: total, waiting for the tile, waiting for an accumulator slot
        atomicAdd(p.prof + 0, (unsigned long long)(clock64() - c_total));
        atomicAdd(p.prof + 1, (unsigned long long)c_wait_tile);
        atomicAdd(p.prof + 2, (unsigned long long)c_wait_slot);
        atomicAdd(p.prof + 7, 1ull);
      }
    }
  } else {
    // ====================== workers: produce the tile, then drain it ===================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;\n");
    // Row m of the tile = (determinant, part) of an 8 x 8 block of determinants (8 alpha rows x
    // 8 beta columns); the 32 rows of a lane quarter are a 4 x 4 sub-block, so that one warp-wide
    // gather of alpha sources (rows of `planes`, contiguous in b) and one of beta sources (rows of
    // `planesT`, contiguous in a) each touch four 64-byte segments instead of 16 scattered lines
    const int m = tid & 127;
    const int h = tid >> 7;              // K-column phase: columns h, h+4, h+8
    const int part = m & 1;
    const int ar = 4 * (m >> 6) + ((m >> 3) & 3), bc = 4 * ((m >> 5) & 1) + ((m >> 1) & 3);
    const uint64_t *pl = p.planes + part, *plT = p.planesT + part;   // element stride: 2 words
    const int64_t neg_off = 2 * p.ndet;
    const uint64_t ZERO = 0x4040404040404040ull;
    // epilogue geometry: lane quarter q (TMEM lanes 32q..32q+31) and column block cb
    const int q = warp & 3, cblk = warp >> 2;
    const int cpb = (p.np + 3) / 4;                 // columns per block (<= 36)
    const int col0 = cblk * cpb;
    const int erow = q * 32 + lane;                 // accumulator row handled in the epilogue
    const int e_ar = 4 * (erow >> 6) + ((erow >> 3) & 3), e_bc = 4 * ((erow >> 5) & 1) + ((erow >> 1) & 3);
    const double st = p.stats[2] * p.op_scale;      // S * T
    double w[3];
    {
      const double r = 1.0 / (double)OZ_RADIX;
      w[0] = st * r * r * r;                        // pair (d0,d1): R^-3 (acc0 R + acc1)
      w[1] = w[0] * r * r;
      w[2] = w[1] * r * r;
    }
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int r = (int)(tile / p.tiles_per_row);
      const int bt = (int)(tile - (int64_t)r * p.tiles_per_row);
      // ---------------- produce ----------------
      // Octets of 8 pair indices: the 16 digit-word loads of octet o are in flight while the map
      // entries of octet o+1 are fetched (two 16-byte loads per map), so a thread pays about one
      // memory round trip per octet.
      long long c_p0 = clock64();
      if (it >= 1) oz_mbar_wait(bar_dfree, (uint32_t)((it - 1) & 1));
      const long long c_p1 = clock64();
      {
        const int64_t a_loc = 8 * (int64_t)r + ar, b = 8 * (int64_t)bt + bc;
        const bool valid = a_loc < p.nrows && b < p.lenb;
        const int64_t a = p.row0 + (valid ? a_loc : 0), bb = valid ? b : 0;
        const int32_t *ta_row = p.mapT_a + a * p.kpad, *tb_row = p.mapT_b + bb * p.kpad;
        const uint64_t *src_a = pl + 2 * bb, *src_b = plT + 2 * a;
        const int ngroups = h < kc ? (kc - h + 3) / 4 : 0;     // groups h, h+4, h+8 < kc
        auto load_maps = [&](int o, int (&ta_)[8], int (&tb_)[8]) {
          const int k0 = 16 * (h + 4 * (o >> 1)) + 8 * (o & 1);
          oz_ldg128_if(valid, ta_row + k0, ta_[0], ta_[1], ta_[2], ta_[3]);
          oz_ldg128_if(valid, ta_row + k0 + 4, ta_[4], ta_[5], ta_[6], ta_[7]);
          oz_ldg128_if(valid, tb_row + k0, tb_[0], tb_[1], tb_[2], tb_[3]);
          oz_ldg128_if(valid, tb_row + k0 + 4, tb_[4], tb_[5], tb_[6], tb_[7]);
        };
        int ta[8], tb[8];
        if (ngroups > 0) load_maps(0, ta, tb);
        uint32_t out[OZ_NS][4];
        for (int o = 0; o < 2 * ngroups; ++o) {
          uint64_t va[8], vb[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            va[u] = oz_ldg64_if(ta[u] != 0, src_a + (ta[u] < 0 ? neg_off : 0) +
                                                2 * (int64_t)(abs(ta[u]) - 1) * p.lenb, ZERO);
            vb[u] = oz_ldg64_if(tb[u] != 0, src_b + (tb[u] < 0 ? neg_off : 0) +
                                                2 * (int64_t)(abs(tb[u]) - 1) * p.lena, ZERO);
          }
          int nta[8], ntb[8];
          if (o + 1 < 2 * ngroups) {
            load_maps(o + 1, nta, ntb);
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) nta[u] = ntb[u] = 0;
          }
          const int half = o & 1;
#pragma unroll
          for (int qd = 0; qd < 2; ++qd) {
            uint64_t e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)   // signed digit sums, per byte
              e[u] = (va[4 * qd + u] + vb[4 * qd + u]) ^ 0x8080808080808080ull;
            // 4 x 4 byte transposes: word s of `out` = slice s of the four k of this quad
            const uint32_t l0 = (uint32_t)e[0], l1 = (uint32_t)e[1], l2 = (uint32_t)e[2],
                           l3 = (uint32_t)e[3];
            const uint32_t h0 = (uint32_t)(e[0] >> 32), h1 = (uint32_t)(e[1] >> 32),
                           h2 = (uint32_t)(e[2] >> 32), h3 = (uint32_t)(e[3] >> 32);
            const uint32_t t0 = __byte_perm(l0, l1, 0x5140), t1 = __byte_perm(l2, l3, 0x5140);
            const uint32_t t2 = __byte_perm(l0, l1, 0x7362), t3 = __byte_perm(l2, l3, 0x7362);
            const uint32_t t4 = __byte_perm(h0, h1, 0x5140), t5 = __byte_perm(h2, h3, 0x5140);
            const uint32_t w0 = __byte_perm(t0, t1, 0x5410), w1 = __byte_perm(t0, t1, 0x7632);
            const uint32_t w2 = __byte_perm(t2, t3, 0x5410), w3 = __byte_perm(t2, t3, 0x7632);
            const uint32_t w4 = __byte_perm(t4, t5, 0x5410), w5 = __byte_perm(t4, t5, 0x7632);
            if (half == 0) {
              out[0][qd] = w0; out[1][qd] = w1; out[2][qd] = w2;
              out[3][qd] = w3; out[4][qd] = w4; out[5][qd] = w5;
            } else {
              out[0][2 + qd] = w0; out[1][2 + qd] = w1; out[2][2 + qd] = w2;
              out[3][2 + qd] = w3; out[4][2 + qd] = w4; out[5][2 + qd] = w5;
            }
          }
          if (half == 1) {
            // 16-byte core-matrix rows: conflict-free (32 consecutive rows = 512 contiguous bytes)
            const int g = h + 4 * (o >> 1);
            const uint32_t dst =
                oz_smem_u32(s_d) + (uint32_t)(g * 16 + (m >> 3)) * 128 + (m & 7) * 16;
#pragma unroll
            for (int sl = 0; sl < OZ_NS; ++sl)
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(
                               dst + (uint32_t)(sl * kc) * 2048),
                           "r"(out[sl][0]), "r"(out[sl][1]), "r"(out[sl][2]), "r"(out[sl][3])
                           : "memory");
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            ta[u] = nta[u];
            tb[u] = ntb[u];
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      oz_mbar_arrive(bar_dfull);
      const long long c_p2 = clock64();
      // ---------------- drain ----------------
      double run[36];
#pragma unroll
      for (int c = 0; c < 36; ++c) run[c] = 0.0;
#pragma unroll
      for (int pr = 0; pr < 3; ++pr) {
        const int sa = (2 * pr) % 3, sb = (2 * pr + 1) % 3;
        const int64_t ua = 2 * it + (2 * pr) / 3, ub = 2 * it + (2 * pr + 1) / 3;
        oz_mbar_wait(bar_sfull + 8 * sa, (uint32_t)(ua & 1));
        oz_mbar_wait(bar_sfull + 8 * sb, (uint32_t)(ub & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        // 36 columns per thread in rounds of 8, 8, 8, 8, 4 (small register footprint next to
        // the 36 running sums)
#pragma unroll
        for (int rd = 0; rd < 5; ++rd) {
          uint32_t ra[8], rb[8];
          const uint32_t ca = tmem + lane_addr + (uint32_t)(sa * OZ_SLOT + col0 + 8 * rd);
          const uint32_t cbb = tmem + lane_addr + (uint32_t)(sb * OZ_SLOT + col0 + 8 * rd);
          if (rd < 4) {
            OZ_TMEM_LD8(ra, ca);
            OZ_TMEM_LD8(rb, cbb);
          } else {
            OZ_TMEM_LD4(ra, ca);
            OZ_TMEM_LD4(rb, cbb);
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int c = 0; c < (rd < 4 ? 8 : 4); ++c) {
            // int32 -> double on the FP64 pipe (exact): 2^52 + 2^31 + comb, minus the offset
            // (I2F.F64 runs on the quarter-rate conversion unit)
            const int comb = (int)ra[c] * OZ_RADIX + (int)rb[c];
            const double cd = __hiloint2double(0x43300000, comb ^ (int)0x80000000) -
                              4503601774854144.0;
            run[8 * rd + c] = fma(w[pr], cd, run[8 * rd + c]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          oz_mbar_arrive(bar_sfree + 8 * sa);
          oz_mbar_arrive(bar_sfree + 8 * sb);
        }
      }
      const long long c_p3 = clock64();
      // E[kl][r*pitch + b].{re,im}: row erow = (det, part); consecutive lanes -> consecutive doubles
      {
        const int epart = erow & 1;
        const int64_t a_loc = 8 * (int64_t)r + e_ar, b = 8 * (int64_t)bt + e_bc;
        if (a_loc < p.nrows && b < p.lenb) {
          double *base = reinterpret_cast<double *>(p.E + (a_loc * p.pitch + b)) + epart;
#pragma unroll
          for (int c = 0; c < 36; ++c) {
            const int kl = col0 + c;
            if (c < cpb && kl < p.np) __stcs(base + 2 * (int64_t)kl * p.lde, run[c]);
          }
        }
      }
      if (p.prof && tid == 0) {   // worker 0: wait for the tile buffer, produce, drain, store
        atomicAdd(p.prof + 3, (unsigned long long)(c_p1 - c_p0));
        atomicAdd(p.prof + 4, (unsigned long long)(c_p2 - c_p1));
        atomicAdd(p.prof + 5, (unsigned long long)(c_p3 - c_p2));
        atomicAdd(p.prof + 6, (unsigned long long)(clock64() - c_p3));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 16)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
constexpr int OZ_STATS_DOUBLES = 4 + 2 * 1024;
// workspace bytes of the sliced path: digit planes + statistics block
static size_t oz_planes_bytes(const fqeb_graph *g) {
  // [layout: by row, transposed][sign][det][part] digit words
  return (size_t)round_up(sizeof(uint64_t) * 8 * (size_t)g->len[0] * g->len[1] + 64, 256);
}
size_t ozaki_workspace_bytes(const fqeb_graph *g) {
  return oz_planes_bytes(g) +
         (size_t)round_up(sizeof(double) * OZ_STATS_DOUBLES, 256);
}
// the statistics block sits behind the planes
double *ozaki_stats_ptr(const fqeb_graph *g, void *d_oz) {
  return (double *)((char *)d_oz + oz_planes_bytes(g));
}

// d_stats: device buffer of OZ_STATS_DOUBLES doubles.  Returns max |Re/Im C| and ||C||^2 on the
// host (one small synchronising copy: the caller decides between this path and the DMMA path).
int ozaki_stats(const fqeb_graph *g, const double *d_coeff, double *d_stats, double *h_absmax,
                double *h_sumsq, cudaStream_t st) {
  const int64_t ndet = g->len[0] * g->len[1];
  const int nblocks = 1024;
  k_absmax_sumsq<<<nblocks, 256, 0, st>>>(2 * ndet, d_coeff, d_stats + 4);
  FQEB_CHECK_LAUNCH();
  k_absmax_sumsq_final<<<1, 32, 0, st>>>(nblocks, d_stats + 4, d_stats);
  FQEB_CHECK_LAUNCH();
  double h[2];
  FQEB_CUDA(cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, st));
  FQEB_CUDA(cudaStreamSynchronize(st));
  *h_absmax = h[0];
  *h_sumsq = h[1];
  return FQEB_OK;
}

// digit planes of the coefficients (scale = d_stats[2], written by ozaki_stats)
int ozaki_slice(const fqeb_graph *g, const double *d_coeff, const double *d_stats, void *d_planes,
                cudaStream_t st) {
  const int64_t lena = g->len[0], lenb = g->len[1], ndet = lena * lenb;
  const int64_t blocks = ((lena + 31) / 32) * ((lenb + 31) / 32);
  if (blocks < 1) return FQEB_OK;
  FQEB_REQUIRE(blocks < (1ll << 31), "ozaki: sector too large for the slicing grid");
  ulonglong2 *planes = (ulonglong2 *)d_planes;
  k_slice_coeff<<<(unsigned)blocks, 256, 0, st>>>(lena, lenb, (const double2 *)d_coeff, d_stats,
                                                   planes, planes + 2 * ndet);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}

// estimated relative error of sigma from the global-scale quantisation of C
double ozaki_error_estimate(const fqeb_graph *g, double absmax, double sumsq) {
  if (!(sumsq > 0.0)) return 0.0;
  const double ndet = (double)g->len[0] * (double)g->len[1];
  double q = 2.0001 * absmax * 0.5 / sqrt(3.0);   // rms rounding error in units of the last digit
  for (int i = 0; i < OZ_NS; ++i) q /= (double)OZ_RADIX;
  return q * sqrt(2.0 * ndet) / sqrt(sumsq);
}

static unsigned long long *g_oz_prof = nullptr;

// by-string adjoint maps zero-padded to kpad columns (built once per graph and pair-space kind)
static int ozaki_maps(const fqeb_graph *g, bool sym, int np, int kpad, const int32_t **ma,
                      const int32_t **mb) {
  GraphLock lock(g);
  fqeb_graph *gm = const_cast<fqeb_graph *>(g);
  const int nspin = g->shared_spin ? 1 : 2;
  for (int sp = 0; sp < nspin; ++sp) {
    if (gm->d_ozmapT[sym][sp]) continue;
    const int64_t len = g->len[sp];
    int32_t *dst = nullptr;
    FQEB_CUDA(cudaMalloc(&dst, sizeof(int32_t) * (size_t)len * kpad));
    int64_t blocks = (len * kpad + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_pad_map<<<(unsigned)blocks, 256>>>(len, np, kpad, sym ? g->d_smapT[sp] : g->d_amapT[sp], dst);
    FQEB_CHECK_LAUNCH();
    FQEB_CUDA(cudaDeviceSynchronize());
    gm->d_ozmapT[sym][sp] = dst;
  }
  *ma = gm->d_ozmapT[sym][0];
  *mb = gm->d_ozmapT[sym][g->shared_spin ? 0 : 1];
  return FQEB_OK;
}

int launch_ozaki(const fqeb_graph *g, const fqeb_op *op, const void *d_planes,
                 const double *d_stats, int64_t row0, int64_t nrows, int pitch, double *d_evec,
                 int64_t lde, cudaStream_t st) {
  FQEB_REQUIRE(ozaki_shape_ok(op), "ozaki: operator not supported by the sliced contraction");
  FQEB_REQUIRE(pitch >= g->len[1] && lde >= nrows * (int64_t)pitch, "ozaki: bad column layout");
  OzOperand o;
  int rc = ozaki_operand(op, g->nele[0] + g->nele[1], &o);
  if (rc != FQEB_OK) return rc;
  OzParams p;
  p.img = o.d_img;
  p.img_bytes = (int)o.img_bytes;
  p.np = o.np;
  p.kc = o.kc;
  p.ng = o.ng;
  p.n_mma = o.n_mma;
  p.ndet = g->len[0] * g->len[1];
  p.planes = (const uint64_t *)d_planes;
  p.planesT = p.planes + 4 * p.ndet;
  p.kpad = 16 * o.kc;
  rc = ozaki_maps(g, op->sym, op->np, p.kpad, &p.mapT_a, &p.mapT_b);
  if (rc != FQEB_OK) return rc;
  p.lena = g->len[0];
  p.lenb = g->len[1];
  p.row0 = row0;
  p.nrows = nrows;
  p.pitch = pitch;
  p.tiles_per_row = (int)((g->len[1] + 7) / 8);
  p.ntiles = ((nrows + 7) / 8) * p.tiles_per_row;
  p.E = (double2 *)d_evec;
  p.lde = lde;
  p.stats = d_stats;
  p.op_scale = o.scale;
  p.prof = nullptr;
  {
    static const bool prof_on = getenv("FQEB_OZAKI_PROF") && getenv("FQEB_OZAKI_PROF")[0] == '1';
    if (prof_on) {
      if (!g_oz_prof) {
        FQEB_CUDA(cudaMalloc(&g_oz_prof, 8 * sizeof(unsigned long long)));
        FQEB_CUDA(cudaMemset(g_oz_prof, 0, 8 * sizeof(unsigned long long)));
      }
      p.prof = g_oz_prof;
    }
  }
  // D^T tile + operand image (its zero block is also what the row groups past the pair space
  // and the last slice's odd K column read)
  // (the odd K column of the last slice reads one 2048-byte column past the tile, i.e. the start
  // of the image: keep at least that much behind the tile for tiny pair spaces)
  const size_t smem = (size_t)OZ_NS * o.kc * 2048 + (o.img_bytes > 2048 ? o.img_bytes : 2048);
  FQEB_REQUIRE(smem + 1856 <= 227 * 1024, "ozaki: shared memory budget exceeded (%zu bytes)", smem);
  static size_t attr_bytes = 0;   // dynamic + static shared memory must stay within 227 KB
  if (smem > attr_bytes) {
    FQEB_CUDA(cudaFuncSetAttribute(k_sigma_ozaki, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    attr_bytes = smem;
  }
  int64_t grid = sm_count();
  if (grid > p.ntiles) grid = p.ntiles;
  if (grid < 1) return FQEB_OK;
  k_sigma_ozaki<<<(unsigned)grid, OZ_THREADS, smem, st>>>(p);
  FQEB_CHECK_LAUNCH();
  return FQEB_OK;
}


// ---------------------------------------------------------------------------------------
// INT8 tensor-core rate of this GPU, measured: the roofline denominator of k_sigma_ozaki.
// Every SM issues chains of tcgen05.mma kind::i8 (M = 128, N = 256, K = 32) on resident
// operands; the tensor time of such an MMA is 128 cycles, so the single issuing thread is
// not the limit.  (MEASURED_PEAKS.json carries bf16 only; INT8 is nominally twice that.)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_i8_rate(int iters, int32_t *__restrict__ sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tmem_s;
  constexpr int N = 256, KB = 160;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + N) * KB; i += blockDim.x) smem[i] = (uint8_t)((i * 37 + 11) & 0x3f);
  if (tid == 0) {
    oz_mbar_init(oz_smem_u32(&bar[0]), 1);
    oz_mbar_init(oz_smem_u32(&bar[1]), 1);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        oz_smem_u32(&tmem_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_s;
  if (tid == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    uint64_t ad[5], bd[5];
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) {
      ad[ks] = oz_desc(oz_smem_u32(smem) + ks * 2 * 16 * 128, 16 * 128, 128);
      bd[ks] = oz_desc(oz_smem_u32(smem) + 128 * KB + ks * 2 * (N / 8) * 128, (N / 8) * 128, 128);
    }
    for (int it = 0; it < iters; ++it) {     // two accumulators, two chains in flight
      const int slot = it & 1;
      if (it >= 2) oz_mbar_wait(oz_smem_u32(&bar[slot]), (uint32_t)((it / 2 - 1) & 1));
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) oz_mma(tmem + slot * N, ad[ks], bd[ks], idesc, (c | ks) != 0);
      }
      oz_commit(oz_smem_u32(&bar[slot]));
    }
    for (int it = (iters > 2 ? iters - 2 : 0); it < iters; ++it)
      oz_mbar_wait(oz_smem_u32(&bar[it & 1]), (uint32_t)((it / 2) & 1));
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t r0[4];
    OZ_TMEM_LD4(r0, tmem + ((uint32_t)(warp * 32) << 16));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (r0[0] == 0x12345678u) sink[tid] = (int32_t)r0[1];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

}  // namespace fqeb

// Cycle counters of k_sigma_ozaki since the last call (FQEB_OZAKI_PROF=1), summed over CTAs:
// [0] issuer total, [1] issuer waiting for a tile, [2] issuer waiting for an accumulator slot,
// [3] worker 0 waiting for the tile buffer, [4] producing, [5] draining, [6] storing E,
// [7] number of CTA launches.  Synchronises the device.
extern "C" int fqeb_ozaki_profile(uint64_t *h_out) {
  using namespace fqeb;
  FQEB_REQUIRE(h_out != nullptr, "fqeb_ozaki_profile: NULL argument");
  for (int i = 0; i < 8; ++i) h_out[i] = 0;
  if (!g_oz_prof) return FQEB_OK;
  FQEB_CUDA(cudaDeviceSynchronize());
  FQEB_CUDA(cudaMemcpy(h_out, g_oz_prof, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  FQEB_CUDA(cudaMemset(g_oz_prof, 0, 8 * sizeof(uint64_t)));
  return FQEB_OK;
}

// dense INT8 tensor-core throughput (tera-operations per second, 2 ops per multiply-add) of the
// current device, measured with tcgen05.mma kind::i8 on all SMs; CUDA-event timed, synchronous
extern "C" int fqeb_i8_tensor_peak(double *tops) {
  using namespace fqeb;
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(tops != nullptr, "fqeb_i8_tensor_peak: NULL argument");
  const int iters = 600, mmas_per_iter = 20;
  const size_t smem = (size_t)(128 + 256) * 160;
  FQEB_CUDA(cudaFuncSetAttribute(k_i8_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int32_t *d_sink = nullptr;
  FQEB_CUDA(cudaMalloc(&d_sink, sizeof(int32_t) * 128));
  cudaEvent_t e0, e1;
  FQEB_CUDA(cudaEventCreate(&e0));
  FQEB_CUDA(cudaEventCreate(&e1));
  const int sms = sm_count();
  double best_ms = 1e30;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, nullptr);
    k_i8_rate<<<sms, 128, smem>>>(iters, d_sink);
    cudaEventRecord(e1, nullptr);
    if (cudaEventSynchronize(e1) != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_ms) best_ms = ms;
  }
  cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_sink);
  if (err != cudaSuccess || best_ms > 1e29) {
    set_error("fqeb_i8_tensor_peak: %s", cudaGetErrorString(err));
    return FQEB_ERR_CUDA;
  }
  const double ops = 2.0 * 128 * 256 * 32 * (double)mmas_per_iter * iters * sms;
  *tops = ops / (best_ms * 1e-3) / 1e12;
  count_launch(4);
  return FQEB_OK;
}
