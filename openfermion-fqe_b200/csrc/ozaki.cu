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
// Kernel anatomy (one persistent CTA per SM, role-specialised warps; see k_sigma_ozaki2):
//   * the digit planes of the operand A (B operand of the MMA: N = pairs kl, K = pairs ij) are
//     copied once per CTA into shared memory in the canonical K-major no-swizzle UMMA layout;
//   * per tile (8 alpha rows x 8 beta strings = 128 real rows m = (det, re|im)), 12 producer
//     warps gather the 8-byte digit words of the alpha and the beta source of every (m, ij) from
//     the pre-sliced coefficient planes (stored twice, [a][b] and [b][a], so that both gathers of
//     a warp are short contiguous segments), add them as packed biased bytes, transpose 4 x 4
//     bytes with PRMT and store 16-byte core-matrix rows of the D^T tile into a staging buffer;
//   * ONE thread copies the staging buffer into TMEM (tcgen05.cp) and issues the 21 x K/32
//     tcgen05.mma kind::i8 instructions of the tile per 48-column block, A operand from TMEM,
//     diagonal by diagonal (d = i + j), one accumulator slot per diagonal in TMEM;
//   * 8 drain warps fold the accumulators two diagonals at a time (acc_d * 127 + acc_{d+1} fits
//     int32), convert to FP64, apply the weights and stream E out.
// Producers, tensor core and drain work on different tiles at the same time.
#include "fqeb_common.cuh"

#include <math.h>
#include <string.h>
#include <map>
#include <mutex>
#include <type_traits>
#include <vector>

namespace fqeb {

constexpr int OZ_NS = 6;           // digit slices of C and of the operand
constexpr int OZ_DMAX = 5;         // slice products (i, j) with i + j <= OZ_DMAX are kept
constexpr int OZ_RADIX = 127;
constexpr int OZ_NMAX = 136;       // largest pair space (MMA N and K) whose tile + operand image fit in shared memory
constexpr int OZ_NPROD = (OZ_DMAX + 1) * (OZ_DMAX + 2) / 2;   // 21 slice products
constexpr uint32_t OZ_NONE = 0xFFFFFFFFu;                     // source table entry: no source

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
  // part[3*block] = max |x|, [3*block+1] = sum x^2, [3*block+2] = number of non-zero x over this
  // block's grid-stride share
  double mx = 0.0, ss = 0.0, nz = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double v = x[i];
    mx = fmax(mx, fabs(v));
    ss += v * v;
    nz += v != 0.0 ? 1.0 : 0.0;
  }
  __shared__ double s_mx[32], s_ss[32], s_nz[32];
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    nz += __shfl_xor_sync(0xffffffffu, nz, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    s_mx[warp] = mx;
    s_ss[warp] = ss;
    s_nz[warp] = nz;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    mx = lane < nw ? s_mx[lane] : 0.0;
    ss = lane < nw ? s_ss[lane] : 0.0;
    nz = lane < nw ? s_nz[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
      nz += __shfl_xor_sync(0xffffffffu, nz, o);
    }
    if (lane == 0) {
      part[3 * blockIdx.x] = mx;
      part[3 * blockIdx.x + 1] = ss;
      part[3 * blockIdx.x + 2] = nz;
    }
  }
}

__global__ void k_absmax_sumsq_final(int nblocks, const double *__restrict__ part,
                                     double *__restrict__ out) {
  // deterministic second pass: out[0] = max, out[1] = sum of squares, out[2] = scale S,
  // out[3] = number of non-zero real / imaginary parts
  double mx = 0.0, ss = 0.0, nz = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 32) {
    mx = fmax(mx, part[3 * i]);
    ss += part[3 * i + 1];
    nz += part[3 * i + 2];
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    nz += __shfl_xor_sync(0xffffffffu, nz, o);
  }
  if (threadIdx.x == 0) {
    out[0] = mx;
    out[1] = ss;
    out[2] = 2.0001 * mx;   // |x| / S < 0.5: inside the range of OZ_NS balanced digits
    out[3] = nz;
  }
}

__device__ __forceinline__ uint64_t balanced_digits(double x, double scale_pow) {
  // biased digit bytes (most significant slice in byte 0) of x, |x| < 0.5.
  // n = rint(x 127^6) is an integer below 2^42, so everything stays exact in FP64: the nearest
  // integer to n / 127 is q (no ties: 127 is odd; the product n * (1/127) is off by far less than
  // the 1/254 that separates n / 127 from a half-integer) and d = n - 127 q is the balanced digit
  // in [-63, 63] - the same digits as 64-bit integer division gives, at a tenth of the
  // instructions (the integer version made this kernel ALU-bound: 3.7 ms for 13 GB).
  double n = rint(x * scale_pow);
  uint64_t w = 0x4040000000000000ull;   // bytes 6, 7: digit 0
#pragma unroll
  for (int s = OZ_NS - 1; s >= 0; --s) {
    const double q = rint(n * (1.0 / (double)OZ_RADIX));
    const int d = __double2int_rn(fma(-(double)OZ_RADIX, q, n));
    n = q;
    w |= (uint64_t)(uint32_t)(d + 64) << (8 * s);
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

// source table of a by-string map: the entry of (string x, pair k) = element index (in 16-byte
// elements, the negated planes behind the plain ones) of the first element of the source ROW of
// pair k acting on x, or OZ_NONE; the kernel adds the column and loads.  Layout: octets of 8
// consecutive k, the octets of 4 consecutive strings adjacent,
//     dst[(((x >> 2) * (kpad / 8) + (k >> 3)) * 4 + (x & 3)) * 8 + (k & 7)],
// so that the 16-byte loads of a warp (4 strings, one octet) fall into one 128-byte line.
__host__ __device__ inline int64_t oz_table_offset(int64_t x, int k, int kpad) {
  return (((x >> 2) * (kpad >> 3) + (k >> 3)) * 4 + (x & 3)) * 8 + (k & 7);
}
__global__ void k_source_table(int64_t len, int np, int kpad, uint32_t row_len, uint32_t ndet,
                               const int32_t *__restrict__ src, uint32_t *__restrict__ dst) {
  const int64_t len4 = (len + 3) / 4 * 4, n = len4 * kpad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t x = i / kpad;
    const int k = (int)(i - x * kpad);
    const int t = (x < len && k < np) ? src[x * np + k] : 0;
    dst[oz_table_offset(x, k, kpad)] =
        t == 0 ? OZ_NONE : (t < 0 ? ndet : 0u) + (uint32_t)(abs(t) - 1) * row_len;
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
// wait for a phase; the suspend-time hint lets the thread sleep in hardware until the phase
// completes instead of polling (polling warps were 30 % of the producers' issue slots)
__device__ __forceinline__ void oz_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nOZ_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
      "@P1 bra OZ_DONE;\nbra OZ_WAIT;\nOZ_DONE:\n}" ::"r"(bar),
      "r"(parity), "r"(1000000u)
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
// predicated 16-byte load of four source-table entries, OZ_NONE when off
__device__ __forceinline__ void oz_ldg128_if(bool pred, const uint32_t *p, uint32_t &x,
                                             uint32_t &y, uint32_t &z, uint32_t &w) {
  x = y = z = w = OZ_NONE;
  asm volatile(
      "{\n .reg .pred q;\n setp.ne.b32 q, %5, 0;\n @q ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];\n}"
      : "+r"(x), "+r"(y), "+r"(z), "+r"(w)
      : "l"(p), "r"((int)pred));
}
__device__ __forceinline__ void oz_prefetch_l2_if(bool pred, const void *p) {
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %1, 0;\n @q prefetch.global.L2 [%0];\n}" ::"l"(p),
               "r"((int)pred));
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
// kernel parameters
// ---------------------------------------------------------------------------------------
struct OzParams {
  const int8_t *img;        // operand digit image (global)
  int img_bytes;
  int np, kc, ng, n_mma;    // pair space, 16-byte K columns, 8-row groups, MMA N
  const uint64_t *planes;   // coefficient digit planes [sign][a][b][part]
  const uint64_t *planesT;  // the same, transposed:     [sign][b][a][part]
  int64_t ndet;
  const uint32_t *srcT_a;   // alpha source table by string (k_source_table layout), kpad = 16 kc
  const uint32_t *srcT_b;   // beta source table by string
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

// ---------------------------------------------------------------------------------------
// the kernel, second generation: role-specialised warps and the tile's MMA copy in TMEM
// ---------------------------------------------------------------------------------------
// The tile (110 KB) and the operand image (117 KB) fill shared memory, so the tile cannot be
// double-buffered there and the first-generation kernel above runs produce -> MMA -> drain ->
// store strictly one after the other (profiles/r02_ozaki_v1_cycle_profile.txt: 39 k cycles per
// tile for 8 k cycles of tensor work).  Here the shared-memory tile is only a STAGING buffer:
//   * 12 producer warps gather tile k+1 into it while the tensor core works on tile k;
//   * the MMA thread copies it into TMEM (54 x tcgen05.cp.128x128b, ordered with the MMAs in the
//     tensor pipe, so the write-after-read hazard on the TMEM copy needs no barrier), signals
//     "staging free" with a commit, and issues the 21 slice products with the A operand read
//     from TMEM (checked in profiles/microbench/i8_tmem_a.cu), per column block of N = 48,
//     diagonal d into accumulator slot d;
//   * 8 drain warps fold the accumulators two diagonals at a time into FP64 running sums (24 per
//     thread: the block width is set by the drain's register budget - it must not spill, see
//     below - and by TMEM, which holds six 48-column slots next to the tile) and stream E out;
//     a slot pair is released as soon as it has been read, so the MMAs of the next block overlap
//     the drain of this one.
// TMEM: columns [0, 224) tile, [224, 512) accumulator slots.
// Register budget: a CTA's setmaxnreg pool is what the launch allocated, and the allocation is
// per four warps - 24 warps x 80 registers = 61440.  The issuer warpgroup (one working warp)
// releases down to 24, producers take 88 and drain warps 96 (one octet of gathered digit words
// in flight per producer thread, source entries two octets ahead; 24 running FP64 sums per drain
// thread).  Neither role may spill: local memory shares the load/store queue with the producers'
// gathers, and a spill reload in the drain was measured waiting behind hundreds of them.
// Twelve producer warps: the SM's load queue returns in issue order, so what hides the latency of
// the gathers is the number of warps with an octet in flight, not the depth per warp (8 warps
// with two octets each were measured 2.2x slower than 12 warps with one).
constexpr int OZ2_PRODUCERS = 384;                 // warpgroups 0-2
constexpr int OZ2_DRAINERS = 256;                  // warpgroups 3-4
constexpr int OZ2_THREADS = OZ2_PRODUCERS + OZ2_DRAINERS + 128;   // + warpgroup 5: MMA issuer
constexpr int OZ2_W_DRAIN = OZ2_PRODUCERS / 32, OZ2_W_ISSUE = (OZ2_PRODUCERS + OZ2_DRAINERS) / 32;
constexpr int OZ2_A_COLS = 224;
constexpr int OZ2_NB = 48;                         // columns of a block = MMA N = slot width
constexpr int OZ2_NSLOT = OZ_DMAX + 1;             // one accumulator slot per diagonal

__device__ __forceinline__ bool oz_elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}"
               : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void oz_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void oz_cp_128x128b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

// PROF: in-kernel cycle counters (FQEB_OZAKI_PROF=1); compiled out of the production instance,
// where the 64-bit counters would cost the producers ten registers
template <bool PROF>
__device__ __forceinline__ long long oz_clock() {
  if constexpr (PROF) return clock64();
  return 0;
}
// KC: number of 16-byte K columns of the pair space (9 at norb = 16, 7 at norb = 14 with
// real-orbital integrals) as a compile-time constant, or 0 for "any".  With KC known the
// producers' octet pipeline is straight-line code; inside a run-time loop ptxas tracks every
// load that is in flight across the back edge with ONE scoreboard, so waiting for the oldest
// octet waits for everything and nothing overlaps (seen in the SASS control codes).
template <bool PROF, int KC>
__global__ void __launch_bounds__(OZ2_THREADS, 1) k_sigma_ozaki2(const OzParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  // 0: staging full, 1: staging free, then per accumulator slot: full (6), free (6)
  __shared__ __align__(8) uint64_t s_bar[2 + 2 * OZ2_NSLOT];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
  const int kc = KC > 0 ? KC : p.kc, ng = p.ng;
  const uint32_t d_bytes = (uint32_t)OZ_NS * kc * 2048;   // tile: [slice][kcol][16 groups][128]
  uint8_t *s_d = smem;
  uint8_t *s_b = smem + d_bytes;                           // operand image follows the tile
  const uint32_t bar_full = oz_smem_u32(&s_bar[0]), bar_free = oz_smem_u32(&s_bar[1]);
  const uint32_t bar_sfull = oz_smem_u32(&s_bar[2]), bar_sfree = oz_smem_u32(&s_bar[2 + OZ2_NSLOT]);

  // the operand image (117 KB) comes into shared memory as ONE bulk copy of the TMA unit,
  // completing on an mbarrier (s_bar_img) while the other barriers and TMEM are set up
  __shared__ __align__(8) uint64_t s_bar_img;
  const uint32_t bar_img = oz_smem_u32(&s_bar_img);
  if (tid == 0) {
    oz_mbar_init(bar_img, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_img),
                 "r"((uint32_t)p.img_bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            oz_smem_u32(s_b)),
        "l"(p.img), "r"((uint32_t)p.img_bytes), "r"(bar_img)
        : "memory");
  }
  if (tid == 0) {
    oz_mbar_init(bar_full, OZ2_PRODUCERS);
    oz_mbar_init(bar_free, 1);
    for (int s = 0; s < OZ2_NSLOT; ++s) {
      oz_mbar_init(bar_sfull + 8 * s, 1);
      oz_mbar_init(bar_sfree + 8 * s, OZ2_DRAINERS / 32);
    }
  }
  if (warp == OZ2_W_ISSUE) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        oz_smem_u32(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  oz_mbar_wait(bar_img, 0);   // operand image has landed (written by the async proxy)
  const uint32_t tmem = s_tmem;
  const int64_t my_tiles =
      (int64_t)blockIdx.x < p.ntiles ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  // column blocks of 48 pair-space columns (3 at norb = 16); diagonal d of every block
  // accumulates in slot d, so a slot is filled once per block
  const int nblk = (p.np + OZ2_NB - 1) / OZ2_NB;

  if (warp >= OZ2_W_ISSUE) {
    // ================================ MMA issuer ======================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 24;\n");
    if (warp == OZ2_W_ISSUE) {
      // the whole warp walks the loop (uniform descriptor arithmetic); one elected lane issues
      const bool leader = oz_elect_one();
      const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ2_NB >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      const uint32_t d_base = oz_smem_u32(s_d), b_base = oz_smem_u32(s_b);
      const uint32_t b_plane = (uint32_t)kc * ng * 128, b_kstep = 2u * ng * 128;
      const uint32_t zero_base = b_base + OZ_NS * b_plane;
      const int kfull = kc / 2, ncp = OZ_NS * kc;
      // descriptor words (K-major, no swizzle, version 1): low = address >> 4 | LBO >> 4 << 16,
      // high = SBO >> 4 | 1 << 14 with SBO = 128
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t lbo_full = (((uint32_t)ng * 128u) >> 4) << 16;
      const uint32_t d_lo0 = ((d_base & 0x3FFFF) >> 4) | ((2048u >> 4) << 16);
      long long c_wait_tile = 0, c_wait_slot = 0, c_total = oz_clock<PROF>();
      for (int64_t it = 0; it < my_tiles; ++it) {
        long long c0 = oz_clock<PROF>();
        oz_mbar_wait(bar_full, (uint32_t)(it & 1));
        c_wait_tile += oz_clock<PROF>() - c0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (leader) {
          // staging -> TMEM, one 16-byte K column (128 rows x 128 bits) per copy; descriptors are
          // advanced by plain adds (a dependent chain of descriptor arithmetic per instruction
          // was measured at ~80 cycles per MMA, twice the tensor time of an N = 80 step)
          uint32_t lo = d_lo0, ta = tmem;
#pragma unroll 2
          for (int c = 0; c < ncp; ++c) {
            oz_cp_128x128b(ta, ((uint64_t)desc_hi << 32) | lo);
            lo += 2048u >> 4;
            ta += 4;
          }
          oz_commit(bar_free);
        }
#pragma unroll 1
        for (int nb = 0; nb < nblk; ++nb) {
          const uint32_t b_blk = b_base + (uint32_t)nb * (OZ2_NB / 8) * 128;
#pragma unroll 1
          for (int d = 0; d <= OZ_DMAX; ++d) {
            const int slot = d;
            const int64_t use = nblk * it + nb;   // how often this slot has been filled before
            if (use >= 1) {
              const long long c1 = oz_clock<PROF>();
              oz_mbar_wait(bar_sfree + 8 * slot, (uint32_t)((use - 1) & 1));
              c_wait_slot += oz_clock<PROF>() - c1;
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            if (leader) {
              const uint32_t acc = tmem + (uint32_t)(OZ2_A_COLS + slot * OZ2_NB);
              uint32_t first = 0;
              uint32_t a_sl = tmem;                                   // D slice i = 0
              uint32_t b_sl = b_blk + (uint32_t)d * b_plane;          // operand slice j = d
#pragma unroll 1
              for (int i = 0; i <= d; ++i) {
                uint32_t ta = a_sl;
                uint32_t lo = ((b_sl & 0x3FFFF) >> 4) | lbo_full;
                if constexpr (KC > 0) {
#pragma unroll
                  for (int ks = 0; ks < KC / 2; ++ks) {
                    oz_mma_ts(acc, ta, ((uint64_t)desc_hi << 32) | lo, idesc, first);
                    first = 1;
                    ta += 8;
                    lo += b_kstep >> 4;
                  }
                } else {
#pragma unroll 1
                  for (int ks = 0; ks < kfull; ++ks) {
                    oz_mma_ts(acc, ta, ((uint64_t)desc_hi << 32) | lo, idesc, first);
                    first = 1;
                    ta += 8;
                    lo += b_kstep >> 4;
                  }
                }
                if (kc & 1) {
                  // odd column count: the second 16-byte K column of the last step is the zero
                  // block on the operand side, which cancels whatever the tile side holds there
                  const uint32_t b0 = b_sl + (uint32_t)kfull * b_kstep;
                  oz_mma_ts(acc, ta,
                            ((uint64_t)desc_hi << 32) | ((b0 & 0x3FFFF) >> 4) |
                                ((((zero_base - b0) >> 4) & 0x3FFF) << 16),
                            idesc, first);
                  first = 1;
                }
                a_sl += 4u * kc;
                b_sl -= b_plane;
              }
              oz_commit(bar_sfull + 8 * slot);
            }
            __syncwarp();
          }
        }
      }
      if (PROF && p.prof && leader) {
        atomicAdd(p.prof + 0, (unsigned long long)(oz_clock<PROF>() - c_total));
        atomicAdd(p.prof + 1, (unsigned long long)c_wait_tile);
        atomicAdd(p.prof + 2, (unsigned long long)c_wait_slot);
        atomicAdd(p.prof + 7, 1ull);
      }
    }
  } else if (warp >= OZ2_W_DRAIN) {
    // ================================== drain ==========================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 96;\n");
    
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch (warp % 4)
    const int ch = (warp - OZ2_W_DRAIN) >> 2;   // column half of a block
    const int erow = q * 32 + lane;         // accumulator row = (determinant of the 8 x 8 tile, part)
    // tile row = 16 ar + 2 bc + part: the 32 rows of a lane quarter are 2 alpha rows x 8 beta
    // columns x (re, im), so a warp-wide store of one pair row of E is two full 128-byte lines
    const int e_ar = erow >> 4, e_bc = (erow >> 1) & 7;
    const int epart = erow & 1;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int64_t e_step = 2 * p.lde;               // doubles between consecutive pair rows of E
    const double st = p.stats[2] * p.op_scale;      // S * T
    double w[3];
    {
      const double r = 1.0 / (double)OZ_RADIX;
      w[0] = st * r * r * r;                        // pair (d0,d1): R^-3 (acc0 R + acc1)
      w[1] = w[0] * r * r;
      w[2] = w[1] * r * r;
    }
    long long c_drain = 0, c_store = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int r = (int)(tile / p.tiles_per_row);
      const int bt = (int)(tile - (int64_t)r * p.tiles_per_row);
      const int64_t a_loc = 8 * (int64_t)r + e_ar, b = 8 * (int64_t)bt + e_bc;
      const bool live = a_loc < p.nrows && b < p.lenb;
      double *ebase = reinterpret_cast<double *>(p.E + (a_loc * p.pitch + b)) + epart;
      double *eptr = ebase + e_step * (OZ2_NB / 2) * ch;   // first column of this thread
      int nreal = p.np - (OZ2_NB / 2) * ch;                  // columns left in the pair space
#pragma unroll 1
      for (int nb = 0; nb < nblk; ++nb) {
        constexpr int NCOL = OZ2_NB / 2;                      // 24 columns per thread
        const long long c_d0 = oz_clock<PROF>();
        const uint32_t par = (uint32_t)((nblk * it + nb) & 1);
        double run[NCOL];
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
          const int sa = 2 * pr, sb = 2 * pr + 1;
          oz_mbar_wait(bar_sfull + 8 * sa, par);
          oz_mbar_wait(bar_sfull + 8 * sb, par);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int rd = 0; rd < NCOL / 8; ++rd) {
            uint32_t ra[8], rb[8];
            const uint32_t col = tmem + lane_addr + (uint32_t)(OZ2_A_COLS + NCOL * ch + 8 * rd);
            OZ_TMEM_LD8(ra, col + (uint32_t)(sa * OZ2_NB));
            OZ_TMEM_LD8(rb, col + (uint32_t)(sb * OZ2_NB));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              // int32 -> double on the FP64 pipe (exact): 2^52 + 2^31 + comb, minus the offset
              const int comb = (int)ra[c] * OZ_RADIX + (int)rb[c];
              const double cd = __hiloint2double(0x43300000, comb ^ (int)0x80000000) -
                                4503601774854144.0;
              run[8 * rd + c] = pr == 0 ? w[0] * cd : fma(w[pr], cd, run[8 * rd + c]);
            }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            oz_mbar_arrive(bar_sfree + 8 * sa);
            oz_mbar_arrive(bar_sfree + 8 * sb);
          }
        }
        const long long c_d1 = oz_clock<PROF>();
        // E[kl][a_loc * pitch + b].{re,im}: consecutive lanes of a row segment -> consecutive doubles
        if (live) {
          double *ptr = eptr;
#pragma unroll
          for (int c = 0; c < NCOL; ++c) {
#ifndef OZ_EXP_NOSTORE   // timing experiments only (results are wrong with any OZ_EXP_* macro)
            if (c < nreal) __stcs(ptr, run[c]);
#else
            if (c < nreal && run[c] == 1.2345e300) __stcs(ptr, run[c]);
#endif
            ptr += e_step;
          }
        }
        eptr += e_step * OZ2_NB;
        nreal -= OZ2_NB;
        c_drain += c_d1 - c_d0;
        c_store += oz_clock<PROF>() - c_d1;
      }
    }
    if (PROF && p.prof && tid == OZ2_PRODUCERS) {
      atomicAdd(p.prof + 5, (unsigned long long)c_drain);
      atomicAdd(p.prof + 6, (unsigned long long)c_store);
    }
  } else {
    // ================================= producers =======================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;\n");
    // A tile is an 8 x 8 block of determinants (8 alpha rows x 8 beta columns) x (re, im).  The 32
    // threads of a producer warp take a 4 x 4 sub-block, so that one warp-wide gather of alpha
    // sources (rows of `planes`, contiguous in b) and one of beta sources (rows of `planesT`,
    // contiguous in a) each touch four 64-byte segments; the tile ROW they fill is chosen for the
    // drain (16 ar + 2 bc + part, see there).
    const int mt = tid & 127;
    const int h = tid >> 7;              // which third of the tile's 2 kc octets (8 pair indices)
    const int part = mt & 1;
    const int ar = 4 * (mt >> 6) + ((mt >> 3) & 3), bc = 4 * ((mt >> 5) & 1) + ((mt >> 1) & 3);
    const int m = 16 * ar + 2 * bc + part;
    const uint64_t *pl = p.planes + part, *plT = p.planesT + part;   // element stride: 2 words
    const uint64_t ZERO = 0x4040404040404040ull;
    // octets of this thread: K column o >> 1, half o & 1
    const int o0 = (2 * kc * h) / 3, o1 = (2 * kc * (h + 1)) / 3;
    const uint32_t dst0 = oz_smem_u32(s_d) + (uint32_t)(m >> 3) * 128 + (m & 7) * 16;
    long long c_wait = 0, c_prod = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int r = (int)(tile / p.tiles_per_row);
      const int bt = (int)(tile - (int64_t)r * p.tiles_per_row);
      const int64_t a_loc = 8 * (int64_t)r + ar, b = 8 * (int64_t)bt + bc;
      const bool valid = a_loc < p.nrows && b < p.lenb;
      const uint32_t a = (uint32_t)(p.row0 + (valid ? a_loc : 0)), bb = valid ? (uint32_t)b : 0u;
      // octet o of string x: 8 entries at oz_table_offset(x, 8 o, kpad); 32 entries per octet step
      const uint32_t *ta_row = p.srcT_a + oz_table_offset(a, 0, p.kpad);
      const uint32_t *tb_row = p.srcT_b + oz_table_offset(bb, 0, p.kpad);
      // source-table entries of 8 pair indices (two 16-byte loads per spin); OZ_NONE when off
      auto load_srcs = [&](int k0, uint32_t (&ta_)[8], uint32_t (&tb_)[8]) {
        const int ofs = 4 * k0;   // (k0 / 8) octets x 32 entries
        oz_ldg128_if(valid, ta_row + ofs, ta_[0], ta_[1], ta_[2], ta_[3]);
        oz_ldg128_if(valid, ta_row + ofs + 4, ta_[4], ta_[5], ta_[6], ta_[7]);
        oz_ldg128_if(valid, tb_row + ofs, tb_[0], tb_[1], tb_[2], tb_[3]);
        oz_ldg128_if(valid, tb_row + ofs + 4, tb_[4], tb_[5], tb_[6], tb_[7]);
      };
      // the 16 digit words of an octet: alpha sources are rows of `planes` (+ column b), beta
      // sources rows of `planesT` (+ column a)
      auto load_digits = [&](const uint32_t (&ta_)[8], const uint32_t (&tb_)[8], uint64_t (&va)[8],
                             uint64_t (&vb)[8]) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#if defined(OZ_EXP_NOLOAD)
          va[u] = ZERO + ta_[u];
          vb[u] = ZERO + tb_[u];
#elif defined(OZ_EXP_NOALPHA)
          va[u] = ZERO + ta_[u];
          vb[u] = oz_ldg64_if(tb_[u] != OZ_NONE, plT + 2 * (uint64_t)(tb_[u] + a), ZERO);
#elif defined(OZ_EXP_NOBETA)
          va[u] = oz_ldg64_if(ta_[u] != OZ_NONE, pl + 2 * (uint64_t)(ta_[u] + bb), ZERO);
          vb[u] = ZERO + tb_[u];
#else
          va[u] = oz_ldg64_if(ta_[u] != OZ_NONE, pl + 2 * (uint64_t)(ta_[u] + bb), ZERO);
          vb[u] = oz_ldg64_if(tb_[u] != OZ_NONE, plT + 2 * (uint64_t)(tb_[u] + a), ZERO);
#endif
        }
      };
      // signed digit sums of an octet, 4 x 4 byte transposes (word s of a quad = slice s of its
      // four k), stored as the 8-byte half `half` of the 16-byte core-matrix rows of K column g
      auto combine_store = [&](const uint64_t (&va)[8], const uint64_t (&vb)[8], const int g,
                               const int half) {
        uint32_t out[OZ_NS][2];
#pragma unroll
        for (int qd = 0; qd < 2; ++qd) {
          uint64_t e[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            e[u] = (va[4 * qd + u] + vb[4 * qd + u]) ^ 0x8080808080808080ull;
          const uint32_t l0 = (uint32_t)e[0], l1 = (uint32_t)e[1], l2 = (uint32_t)e[2],
                         l3 = (uint32_t)e[3];
          const uint32_t h0 = (uint32_t)(e[0] >> 32), h1 = (uint32_t)(e[1] >> 32),
                         h2 = (uint32_t)(e[2] >> 32), h3 = (uint32_t)(e[3] >> 32);
          const uint32_t t0 = __byte_perm(l0, l1, 0x5140), t1 = __byte_perm(l2, l3, 0x5140);
          const uint32_t t2 = __byte_perm(l0, l1, 0x7362), t3 = __byte_perm(l2, l3, 0x7362);
          const uint32_t t4 = __byte_perm(h0, h1, 0x5140), t5 = __byte_perm(h2, h3, 0x5140);
          out[0][qd] = __byte_perm(t0, t1, 0x5410);
          out[1][qd] = __byte_perm(t0, t1, 0x7632);
          out[2][qd] = __byte_perm(t2, t3, 0x5410);
          out[3][qd] = __byte_perm(t2, t3, 0x7632);
          out[4][qd] = __byte_perm(t4, t5, 0x5410);
          out[5][qd] = __byte_perm(t4, t5, 0x7632);
        }
        const uint32_t dst = dst0 + (uint32_t)g * 2048 + 8u * half;
#pragma unroll
        for (int sl = 0; sl < OZ_NS; ++sl)
          asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(dst + (uint32_t)(sl * kc) * 2048),
                       "r"(out[sl][0]), "r"(out[sl][1])
                       : "memory");
      };
      // Two octets of digit words in flight and source entries fetched two octets ahead: while
      // octet o is being combined, the digits of o+1 and the sources of o+2 are on their way, so
      // neither the table nor the plane latency is exposed per octet.  The loads of the first
      // octets are issued BEFORE waiting for the staging buffer, so they overlap the copy of the
      // previous tile into TMEM.
      const long long c_p0 = oz_clock<PROF>();
      if constexpr (KC > 0) {
        constexpr int NOCT = (2 * KC + 2) / 3;   // octets per thread (the last one may be absent)
        uint32_t sa[2][8], sb[2][8];
        uint64_t va[8], vb[8];
        load_srcs(8 * o0, sa[0], sb[0]);
        if (NOCT > 1) load_srcs(8 * (o0 + 1), sa[1], sb[1]);
#pragma unroll
        for (int j = 0; j < NOCT; ++j) {
          // here: sa[j&1] = sources of octet j, sa[(j+1)&1] = sources of j+1 (on their way)
          const bool on = (2 * KC) % 3 == 0 || j + 1 < NOCT || o0 + j < o1;
          if (on) load_digits(sa[j & 1], sb[j & 1], va, vb);
          if (j + 2 < NOCT && o0 + j + 2 < o1) load_srcs(8 * (o0 + j + 2), sa[j & 1], sb[j & 1]);
          if (j == 0) {
            // first store of this tile: the staging buffer must have been copied out
            const long long c_w = oz_clock<PROF>();
            if (it >= 1) oz_mbar_wait(bar_free, (uint32_t)((it - 1) & 1));
            c_wait += oz_clock<PROF>() - c_w;
          }
          if (on) combine_store(va, vb, (o0 + j) >> 1, (o0 + j) & 1);
        }
      } else {
        // any K-column count: one octet in flight per thread, sources one octet ahead
        uint32_t ta[8], tb[8], sa[8], sb[8];
        uint64_t va[8], vb[8];
        load_srcs(8 * o0, ta, tb);
#pragma unroll 1
        for (int o = o0; o < o1; o += 2) {
          load_digits(ta, tb, va, vb);
          if (o + 1 < o1) load_srcs(8 * (o + 1), sa, sb);
          if (o == o0) {
            const long long c_w = oz_clock<PROF>();
            if (it >= 1) oz_mbar_wait(bar_free, (uint32_t)((it - 1) & 1));
            c_wait += oz_clock<PROF>() - c_w;
          }
          combine_store(va, vb, o >> 1, o & 1);
          if (o + 1 < o1) {
            load_digits(sa, sb, va, vb);
            if (o + 2 < o1) load_srcs(8 * (o + 2), ta, tb);
            combine_store(va, vb, (o + 1) >> 1, (o + 1) & 1);
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      oz_mbar_arrive(bar_full);
      c_prod += oz_clock<PROF>() - c_p0;
    }
    if (PROF && p.prof && tid == 0) {
      atomicAdd(p.prof + 3, (unsigned long long)c_wait);
      atomicAdd(p.prof + 4, (unsigned long long)(c_prod - c_wait));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == OZ2_W_ISSUE)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
constexpr int OZ_STATS_DOUBLES = 4 + 3 * 1024;
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

// d_stats: device buffer of OZ_STATS_DOUBLES doubles.  Returns max |Re/Im C|, ||C||^2 and the
// number of non-zero real / imaginary parts on the host (one small synchronising copy: the
// caller decides between this path and the DMMA path).
int ozaki_stats(const fqeb_graph *g, const double *d_coeff, double *d_stats, double *h_absmax,
                double *h_sumsq, double *h_nonzero, cudaStream_t st) {
  const int64_t ndet = g->len[0] * g->len[1];
  const int nblocks = 1024;
  k_absmax_sumsq<<<nblocks, 256, 0, st>>>(2 * ndet, d_coeff, d_stats + 4);
  FQEB_CHECK_LAUNCH();
  k_absmax_sumsq_final<<<1, 32, 0, st>>>(nblocks, d_stats + 4, d_stats);
  FQEB_CHECK_LAUNCH();
  double h[4];
  FQEB_CUDA(cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, st));
  FQEB_CUDA(cudaStreamSynchronize(st));
  *h_absmax = h[0];
  *h_sumsq = h[1];
  *h_nonzero = h[3];
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

// Estimated relative error of sigma from the global-scale quantisation of C: every NON-ZERO real
// or imaginary part carries a rounding error of rms u / sqrt(12), u = S 127^-6 the unit of the
// last digit (zeros are exact, and a part below u / 2 is replaced by zero, an error smaller than
// that), against ||C||.  A Hartree-Fock determinant or the first Taylor terms grown from it
// therefore pass (few non-zeros), a dense state with a handful of dominant coefficients does not.
double ozaki_error_estimate(double absmax, double sumsq, double nonzero) {
  if (!(sumsq > 0.0)) return 0.0;
  double q = 2.0001 * absmax * 0.5 / sqrt(3.0);   // u / sqrt(12) with u = S 127^-6
  for (int i = 0; i < OZ_NS; ++i) q /= (double)OZ_RADIX;
  return q * sqrt(nonzero) / sqrt(sumsq);
}

static unsigned long long *g_oz_prof = nullptr;

// source tables of both spins (built once per graph and pair-space kind): alpha sources are rows
// of `planes` (length lenb), beta sources rows of `planesT` (length lena)
static int ozaki_maps(const fqeb_graph *g, bool sym, int np, int kpad, const uint32_t **ma,
                      const uint32_t **mb) {
  GraphLock lock(g);
  fqeb_graph *gm = const_cast<fqeb_graph *>(g);
  const int64_t ndet = g->len[0] * g->len[1];
  FQEB_REQUIRE(2 * ndet + g->len[0] + g->len[1] < (1ll << 32),
               "ozaki: sector too large for 32-bit source indices");
  for (int sp = 0; sp < 2; ++sp) {
    if (gm->d_ozmapT[sym][sp]) continue;
    const int64_t len = g->len[sp];
    uint32_t *dst = nullptr;
    const int64_t len4 = (len + 3) / 4 * 4;
    FQEB_CUDA(cudaMalloc(&dst, sizeof(uint32_t) * (size_t)len4 * kpad));
    int64_t blocks = (len4 * kpad + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_source_table<<<(unsigned)blocks, 256>>>(len, np, kpad, (uint32_t)g->len[1 - sp],
                                              (uint32_t)ndet,
                                              sym ? g->d_smapT[sp] : g->d_amapT[sp], dst);
    FQEB_CHECK_LAUNCH();
    FQEB_CUDA(cudaDeviceSynchronize());
    gm->d_ozmapT[sym][sp] = (int32_t *)dst;
  }
  *ma = (const uint32_t *)gm->d_ozmapT[sym][0];
  *mb = (const uint32_t *)gm->d_ozmapT[sym][1];
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
  rc = ozaki_maps(g, op->sym, op->np, p.kpad, &p.srcT_a, &p.srcT_b);
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
  // (the kernel multiplies whole column blocks of 48 operand rows: for small pair spaces the
  // rows past the image must still be addressable shared memory)
  const size_t smem = (size_t)OZ_NS * o.kc * 2048 + (o.img_bytes > 2048 ? o.img_bytes : 2048) +
                      (o.np < 128 ? 4096 : 0);
  FQEB_REQUIRE(smem + 1856 <= 227 * 1024, "ozaki: shared memory budget exceeded (%zu bytes)", smem);
  static PerDeviceSize attr_dev;   // dynamic + static shared memory must stay within 227 KB
  int attr_dev_id = 0;
  if (attr_dev.needs(smem, &attr_dev_id)) {
    FQEB_CUDA(cudaFuncSetAttribute(k_sigma_ozaki2<false, 0>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FQEB_CUDA(cudaFuncSetAttribute(k_sigma_ozaki2<false, 7>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FQEB_CUDA(cudaFuncSetAttribute(k_sigma_ozaki2<false, 9>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FQEB_CUDA(cudaFuncSetAttribute(k_sigma_ozaki2<true, 0>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FQEB_CUDA(cudaFuncSetAttribute(k_sigma_ozaki2<true, 7>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FQEB_CUDA(cudaFuncSetAttribute(k_sigma_ozaki2<true, 9>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev.record(attr_dev_id, smem);
  }
  int64_t grid = sm_count();
  if (grid > p.ntiles) grid = p.ntiles;
  if (grid < 1) return FQEB_OK;
#define OZ_LAUNCH(PROF, KC) k_sigma_ozaki2<PROF, KC><<<(unsigned)grid, OZ2_THREADS, smem, st>>>(p)
  if (p.prof) {
    if (p.kc == 9) OZ_LAUNCH(true, 9);
    else if (p.kc == 7) OZ_LAUNCH(true, 7);
    else OZ_LAUNCH(true, 0);
  } else {
    if (p.kc == 9) OZ_LAUNCH(false, 9);
    else if (p.kc == 7) OZ_LAUNCH(false, 7);
    else OZ_LAUNCH(false, 0);
  }
#undef OZ_LAUNCH
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
