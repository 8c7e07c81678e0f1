// Microbenchmark 2 for the sliced contraction: tcgen05.mma kind::i8 with the A operand in TENSOR
// MEMORY, filled either by tcgen05.cp (shared memory -> TMEM, issued by the MMA thread, ordered
// with the MMAs in the tensor pipe) or by tcgen05.st from registers.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o i8_tmem_a i8_tmem_a.cu
//
// Why: the D^T tile of k_sigma_ozaki (110 KB) and the operand image (117 KB) fill shared memory,
// so the tile cannot be double-buffered there.  With the tile's MMA copy in TMEM (216 columns) the
// shared-memory buffer becomes a staging area that is free again as soon as the copy is done,
// i.e. producers can build tile k+1 while the tensor core works on tile k.
//
// (1) correctness of  D[128 x N] = A[128 x K] . B[N x K]^T  with A in TMEM at a column offset,
//     the accumulator behind it: A placed by (a) tcgen05.cp.128x128b per 16-byte K column from the
//     canonical K-major no-swizzle shared-memory layout, (b) tcgen05.st.32x32b.x4 (lane = row,
//     column c = K bytes 4c..4c+3);
// (2) rates: cycles per MMA for N = 72 / 144 with A in TMEM, one issuing thread; cycles for the
//     54 copies of one tile; TMEM read rate of 8 warps draining 2 x 36 columns each.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CHECK(x)                                                                       \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int M = 128;
constexpr int A_COL0 = 8;      // TMEM column where the A tile starts (not 0 on purpose)

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\nbra WAIT_LOOP;\nDONE:\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__host__ __device__ inline uint32_t make_idesc(int n, int a_signed, int b_signed) {
  return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_i8_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_cp_128x128b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__host__ __device__ inline int canon_offset(int rows, int row, int k) {
  return ((k >> 4) * (rows >> 3) + (row >> 3)) * 128 + (row & 7) * 16 + (k & 15);
}

// mode 0: A -> TMEM by tcgen05.cp.128x128b; mode 1: by tcgen05.st
__global__ void __launch_bounds__(128, 1)
k_check(const uint8_t *__restrict__ a, const int8_t *__restrict__ b, int n, int k, int a_signed,
        int mode, int32_t *__restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t *sa = smem, *sb = smem + M * k;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * k; i += blockDim.x) sa[canon_offset(M, i / k, i % k)] = a[i];
  for (int i = tid; i < n * k; i += blockDim.x) sb[canon_offset(n, i / k, i % k)] = (uint8_t)b[i];
  if (tid == 0) mbar_init(smem_u32(&bar), 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const int kcols = k / 16;
  const uint32_t acc = tmem + A_COL0 + 4 * kcols + 4;   // accumulator behind the A tile
  if (mode == 1) {
    // thread = row; 16 bytes of K per 4 columns
    const int row = tid;
    for (int c = 0; c < kcols; ++c) {
      uint32_t w[4];
      for (int j = 0; j < 4; ++j) {
        uint32_t v = 0;
        for (int bb = 0; bb < 4; ++bb) v |= (uint32_t)a[row * k + 16 * c + 4 * j + bb] << (8 * bb);
        w[j] = v;
      }
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + A_COL0 + 4 * c;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr),
                   "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (tid == 0) {
    if (mode == 0) {
      for (int c = 0; c < kcols; ++c)
        tmem_cp_128x128b(tmem + A_COL0 + 4 * c,
                         make_desc(smem_u32(sa) + c * (M / 8) * 128, (M / 8) * 128, 128));
    }
    const uint32_t idesc = make_idesc(n, a_signed, 1);
    for (int ks = 0; ks < k / 32; ++ks) {
      const uint32_t blk = (n / 8) * 128;
      const uint64_t bd = make_desc(smem_u32(sb) + ks * 2 * blk, blk, 128);
      umma_i8_ts(acc, tmem + A_COL0 + 8 * ks, bd, idesc, ks > 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < n; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = acc + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) out[row * n + c0 + j] = (int32_t)r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// rates.  what = 0: TS MMAs (A in TMEM), 1: SS MMAs, 2: 54 tcgen05.cp.128x128b + commit per
// iteration, 3: like 0 but every iteration is preceded by the 54 copies (the real per-tile order),
// 4: like 0 but consecutive MMAs alternate between two accumulators (is the fixed cost per MMA a
// read-after-write bubble on the accumulator?), 5: like 0 but TWO warps issue, each nchain MMAs per
// iteration into its own accumulators (is the fixed cost on the issuing side?)
__global__ void __launch_bounds__(128, 1)
k_rate(int what, int n, int nchain, int iters, long long *__restrict__ cycles,
       int32_t *__restrict__ sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) uint64_t s_bd[64];
  constexpr int KB = 160;   // 10 K columns of 16 bytes
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
  uint8_t *sa = smem, *sb = smem + 6 * M * KB;
  for (int i = tid; i < 6 * M * KB + n * KB; i += blockDim.x) smem[i] = (uint8_t)((i * 37 + 11) & 0x3f);
  if (tid == 0)
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1);
  if (tid < 5) s_bd[tid] = make_desc(smem_u32(sb) + tid * 2 * (n / 8) * 128, (n / 8) * 128, 128);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (warp == 1 || (what == 5 && warp == 2)) {
    const int w2 = warp - 1;   // what = 5: two issuing warps, each with its own accumulators
    // the whole warp walks the loop (descriptor arithmetic in the uniform datapath), one
    // elected lane issues: with a single-thread branch ptxas wraps every tcgen05 instruction
    // in R2UR moves and an ELECT loop (~117 cycles per MMA whatever the shape)
    uint32_t is_leader;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(is_leader));
    const bool leader = is_leader != 0;
    const uint32_t idesc = make_idesc(n, 0, 1);
    const uint32_t acc0 = tmem + 224;
    uint64_t adv[5];
#pragma unroll
    for (int ks = 0; ks < 5; ++ks)
      adv[ks] = make_desc(smem_u32(sa) + ks * 2 * (M / 8) * 128, (M / 8) * 128, 128);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int slot = (it & 1) + 2 * w2;
      if (it >= 2) mbar_wait(smem_u32(&bar[slot]), (uint32_t)((it / 2 - 1) & 1));
      if (leader) {
      if (what == 2 || what == 3) {
        for (int c = 0; c < 54; ++c)
          tmem_cp_128x128b(tmem + 4 * c, make_desc(smem_u32(sa) + c * (M / 8) * 128, (M / 8) * 128, 128));
      }
      if (what != 2) {
        // descriptors advance by plain adds in (uniform) registers: the issue loop must not be
        // the limit (a smem-table / ALU-chain loop was measured at 138 cycles per MMA)
        const uint32_t acc = acc0 + (uint32_t)(slot * n);
        const uint64_t bd0 = make_desc(smem_u32(sb), (n / 8) * 128, 128);
        const uint32_t bstep = (2u * (n / 8) * 128) >> 4;
        for (int c = 0; c < nchain; c += 5) {
          uint64_t bd = bd0;
          uint32_t ta = tmem + 40 * ((c / 5) % 5);
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) {
            if (what == 1) umma_i8_ss(acc, adv[ks], bd, idesc, (c + ks) > 0);
            else if (what == 4)   // consecutive MMAs alternate between two accumulators
              umma_i8_ts(acc0 + (uint32_t)(((c + ks) & 1) * n), ta, bd, idesc, (c + ks) > 1);
            else umma_i8_ts(acc, ta, bd, idesc, (c + ks) > 0);
            bd += bstep;
            ta += 8;
          }
        }
      }
      umma_commit(smem_u32(&bar[slot]));
      }
      __syncwarp();
    }
    for (int it = (iters > 2 ? iters - 2 : 0); it < iters; ++it)
      mbar_wait(smem_u32(&bar[(it & 1) + 2 * w2]), (uint32_t)((it / 2) & 1));
    if (leader) atomicMax((unsigned long long *)&cycles[blockIdx.x], (unsigned long long)(clock64() - t0));
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t r0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];"
                 : "=r"(r0)
                 : "r"(tmem + 230 + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (r0 == 0x12345678u) sink[tid] = (int32_t)r0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// TMEM read rate: `nwarps` warps (nwarps/4 per lane quarter) each read 2 x 36 columns per step
// (two accumulator slots) and fold them into 36 running FP64 sums, like the drain of the kernel
__global__ void __launch_bounds__(512, 1)
k_ldtm(int steps, long long *__restrict__ cycles, double *__restrict__ sink) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const int q = warp & 3, cb = (warp >> 2) & 1;
  const uint32_t base = tmem + ((uint32_t)(q * 32) << 16) + 224 + 36 * cb;
  double run[36];
#pragma unroll
  for (int c = 0; c < 36; ++c) run[c] = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
#pragma unroll
    for (int rd = 0; rd < 5; ++rd) {
      uint32_t ra[8], rb[8];
      if (rd < 4) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(ra[0]), "=r"(ra[1]), "=r"(ra[2]), "=r"(ra[3]), "=r"(ra[4]), "=r"(ra[5]),
                       "=r"(ra[6]), "=r"(ra[7])
                     : "r"(base + 8 * rd));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(rb[0]), "=r"(rb[1]), "=r"(rb[2]), "=r"(rb[3]), "=r"(rb[4]), "=r"(rb[5]),
                       "=r"(rb[6]), "=r"(rb[7])
                     : "r"(base + 72 + 8 * rd));
      } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(ra[0]), "=r"(ra[1]), "=r"(ra[2]), "=r"(ra[3])
                     : "r"(base + 32));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(rb[0]), "=r"(rb[1]), "=r"(rb[2]), "=r"(rb[3])
                     : "r"(base + 72 + 32));
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < (rd < 4 ? 8 : 4); ++c) {
        const int comb = (int)ra[c] * 127 + (int)rb[c];
        const double cd = __hiloint2double(0x43300000, comb ^ (int)0x80000000) - 4503601774854144.0;
        run[8 * rd + c] = fma(1.0000001, cd, run[8 * rd + c]);
      }
    }
  }
  const long long t1 = clock64();
  double acc = 0.0;
#pragma unroll
  for (int c = 0; c < 36; ++c) acc += run[c];
  if (acc == 1.2345) sink[tid] = acc;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

static bool run_check(int n, int k, int a_signed, int mode) {
  std::vector<uint8_t> ha((size_t)M * k);
  std::vector<int8_t> hb((size_t)n * k);
  srand(4321 + n + k + a_signed + mode);
  for (auto &v : ha) v = a_signed ? (uint8_t)(int8_t)(rand() % 253 - 126) : (uint8_t)(rand() % 256);
  for (auto &v : hb) v = (int8_t)(rand() % 127 - 63);
  uint8_t *da;
  int8_t *db;
  int32_t *dout;
  CHECK(cudaMalloc(&da, ha.size()));
  CHECK(cudaMalloc(&db, hb.size()));
  CHECK(cudaMalloc(&dout, sizeof(int32_t) * M * n));
  CHECK(cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice));
  CHECK(cudaMemset(dout, 0xff, sizeof(int32_t) * M * n));
  const size_t smem = (size_t)(M + n) * k;
  CHECK(cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_check<<<1, 128, smem>>>(da, db, n, k, a_signed, mode, dout);
  CHECK(cudaGetLastError());
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("check N=%3d K=%3d mode %d: CUDA error %s\n", n, k, mode, cudaGetErrorString(e));
    exit(3);
  }
  std::vector<int32_t> hout((size_t)M * n);
  CHECK(cudaMemcpy(hout.data(), dout, sizeof(int32_t) * M * n, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < n; ++j) {
      long long ref = 0;
      for (int kk = 0; kk < k; ++kk) {
        const int av = a_signed ? (int)(int8_t)ha[(size_t)i * k + kk] : (int)ha[(size_t)i * k + kk];
        ref += (long long)av * hb[(size_t)j * k + kk];
      }
      if (ref != hout[(size_t)i * n + j]) {
        if (bad < 4)
          printf("  mismatch (%d,%d): got %d want %lld\n", i, j, hout[(size_t)i * n + j], ref);
        ++bad;
      }
    }
  printf("check N=%3d K=%3d A %s in TMEM via %s: %s (%lld mismatches)\n", n, k,
         a_signed ? "s8" : "u8", mode ? "tcgen05.st" : "tcgen05.cp.128x128b", bad ? "FAIL" : "ok",
         bad);
  cudaFree(da);
  cudaFree(db);
  cudaFree(dout);
  return bad == 0;
}

static void run_rate(int what, int n, int nchain, int iters, int sms) {
  long long *dcyc;
  int32_t *dsink;
  CHECK(cudaMalloc(&dcyc, sizeof(long long) * sms));
  CHECK(cudaMalloc(&dsink, sizeof(int32_t) * 128));
  const size_t smem = (size_t)(6 * M + n) * 160;
  CHECK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_rate<<<sms, 128, smem>>>(what, n, nchain, 4, dcyc, dsink);
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaMemset(dcyc, 0, sizeof(long long) * sms));
  k_rate<<<sms, 128, smem>>>(what, n, nchain, iters, dcyc, dsink);
  CHECK(cudaGetLastError());
  CHECK(cudaDeviceSynchronize());
  std::vector<long long> h(sms);
  CHECK(cudaMemcpy(h.data(), dcyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (auto c : h) mx = c > mx ? c : mx;
  const char *names[] = {"TS (A in TMEM)", "SS (A in smem)", "54 x cp.128x128b only", "54 x cp + TS MMAs",
                         "TS, alternating accs", "TS, two issuing warps"};
  printf("rate  %-22s N=%3d  %3d MMAs/iter x %4d : %8.1f cycles/iter  %6.1f cycles/MMA (tensor floor %5.1f)\n",
         names[what], n, what == 2 ? 0 : nchain, iters, (double)mx / iters,
         what == 2 ? 0.0 : (double)mx / iters / (what == 5 ? 2 * nchain : nchain), 128.0 * n / 256.0);
  cudaFree(dcyc);
  cudaFree(dsink);
}

int main() {
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s, %d SMs\n", prop.name, prop.multiProcessorCount);
  bool ok = true;
  for (int mode = 1; mode >= 0; --mode) {
    ok &= run_check(64, 32, 1, mode);
    ok &= run_check(80, 160, 0, mode);
    ok &= run_check(144, 160, 0, mode);
    ok &= run_check(144, 160, 1, mode);
  }
  if (!ok) printf("CORRECTNESS FAILED somewhere - see above\n");
  const int sms = prop.multiProcessorCount;
  for (int what : {0, 1}) {
    for (int n : {32, 48, 64, 80, 144}) run_rate(what, n, 105, 500, sms);
  }
  for (int n : {48, 96, 144}) run_rate(4, n, 105, 500, sms);
  for (int n : {32, 48, 64}) run_rate(5, n, 105, 500, sms);
  run_rate(0, 96, 105, 500, sms);
  run_rate(0, 48, 315, 500, sms);
  run_rate(0, 80, 210, 500, sms);
  run_rate(2, 80, 105, 500, sms);
  run_rate(3, 80, 210, 500, sms);
  run_rate(3, 144, 105, 500, sms);
  {
    long long *dcyc;
    double *dsink;
    CHECK(cudaMalloc(&dcyc, sizeof(long long) * sms));
    CHECK(cudaMalloc(&dsink, sizeof(double) * 512));
    for (int nw : {8, 16}) {
      k_ldtm<<<sms, nw * 32>>>(600, dcyc, dsink);
      CHECK(cudaGetLastError());
      CHECK(cudaDeviceSynchronize());
      long long c0;
      CHECK(cudaMemcpy(&c0, dcyc, sizeof(long long), cudaMemcpyDeviceToHost));
      printf("drain %2d warps: %7.1f cycles per step of 2 x 36 columns per warp "
             "(%.1f bytes of TMEM per cycle per SM)\n",
             nw, (double)c0 / 600, (double)nw * 32 * 72 * 4 / ((double)c0 / 600));
    }
  }
  return ok ? 0 : 2;
}
