// Microbenchmark: tcgen05.mma kind::i8 (INT8 x INT8 -> INT32 in TMEM) on sm_100a.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o i8_umma i8_umma.cu
//   ./i8_umma
//
// (1) correctness: one CTA tile  D[128 x N] = A[128 x K] . B[N x K]^T, both operands written by
//     ordinary threads into shared memory in the canonical K-major NO-SWIZZLE layout
//     (8-row x 16-byte core matrices; LBO = byte distance between the two 16-byte K columns of
//     one MMA, SBO = byte distance between 8-row groups), accumulators read back with
//     tcgen05.ld and compared with an int32 CPU product.  Mixed signedness (A unsigned, B signed)
//     is checked too: the sliced contraction feeds biased (unsigned) digits on the D side.
// (2) rate: every SM issues a long chain of MMAs on resident operands; cycles per MMA for
//     N = 64 ... 256 (is the issue rate 128*N/256 cycles, and where do small N become
//     shared-memory-bandwidth bound?), and chip-wide TOP/s.  (kind::i8 needs N = 8 or N % 16 == 0.)
//
// This is the gate VERDICT r1 item 8 asks for before an emulated-FP64 contraction is wired in.
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
// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
  return d;                 // base_offset = 0, lbo_mode = 0, layout_type = 0 (SWIZZLE_NONE)
}
// instruction descriptor for kind::i8 (cute::UMMA::InstrDescriptor)
__host__ __device__ inline uint32_t make_idesc(int n, int a_signed, int b_signed) {
  return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
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

// canonical K-major no-swizzle placement of element (row, k) of an operand with `rows` rows:
// K column-major over 16-byte columns, 8-row groups inside a column
__host__ __device__ inline int canon_offset(int rows, int row, int k) {
  return ((k >> 4) * (rows >> 3) + (row >> 3)) * 128 + (row & 7) * 16 + (k & 15);
}

// out: [128][N] int32.  a: [128][K] bytes, b: [N][K] bytes (row-major in global memory).
__global__ void __launch_bounds__(128, 1)
k_check(const uint8_t *__restrict__ a, const int8_t *__restrict__ b, int n, int k, int a_signed,
        int swap, int32_t *__restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t *sa = smem;                 // 128 x k
  uint8_t *sb = smem + M * k;         // n x k
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < M * k; i += blockDim.x) sa[canon_offset(M, i / k, i % k)] = a[i];
  for (int i = tid; i < n * k; i += blockDim.x)
    sb[canon_offset(n, i / k, i % k)] = (uint8_t)b[i];
  if (tid == 0) mbar_init(smem_u32(&bar), 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(
        smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // generic-proxy writes -> visible to the async proxy (tensor core reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(n, a_signed, 1);
    for (int ks = 0; ks < k / 32; ++ks) {
      // K step ks = 16-byte columns 2ks, 2ks+1
      // swap = 1 exchanges the roles of the two offset fields (diagnostic for the field meaning)
      const uint32_t alk = (M / 8) * 128, blk = (n / 8) * 128;
      const uint64_t ad = make_desc(smem_u32(sa) + ks * 2 * alk, swap ? 128 : alk, swap ? alk : 128);
      const uint64_t bd = make_desc(smem_u32(sb) + ks * 2 * blk, swap ? 128 : blk, swap ? blk : 128);
      umma_i8(tmem, ad, bd, idesc, ks > 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warp w reads lanes 32w .. 32w+31; thread = one accumulator row
  const int row = warp * 32 + (tid & 31);
  for (int c0 = 0; c0 < n; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

// rate: one thread per CTA issues `iters` chains of nchain MMAs (K = 32 each) over resident
// operand tiles [128 x kbytes] / [n x kbytes]; cycles between the first issue and the last commit
__global__ void __launch_bounds__(128, 1)
k_rate(int n, int kbytes, int nchain, int iters, long long *__restrict__ cycles,
       int32_t *__restrict__ sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[3];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t *sa = smem, *sb = smem + M * kbytes;
  for (int i = tid; i < (M + n) * kbytes; i += blockDim.x) smem[i] = (uint8_t)((i * 37 + 11) & 0x7f);
  if (tid == 0)
    for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&bar[i]), 1);
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
  if (tid == 0) {
    const uint32_t idesc = make_idesc(n, 0, 1);
    const int ksteps = kbytes / 32;
    const long long t0 = clock64();
    // chain `it` accumulates into TMEM slot it % 3 and commits to bar[it % 3]; before a slot is
    // reused the chain that used it three iterations ago must have completed
    // descriptors of the K steps are loop-invariant: precompute them so that the single issuing
    // thread spends a handful of instructions per MMA (a dependent ALU chain of ~40 instructions
    // per MMA was measured at 175 cycles/MMA, i.e. issue-bound far below the tensor rate)
    uint64_t adv[8], bdv[8];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const int kk = ks < ksteps ? ks : 0;
      adv[ks] = make_desc(smem_u32(sa) + kk * 2 * (M / 8) * 128, (M / 8) * 128, 128);
      bdv[ks] = make_desc(smem_u32(sb) + kk * 2 * (n / 8) * 128, (n / 8) * 128, 128);
    }
    for (int it = 0; it < iters; ++it) {
      const int slot = it % 3;
      if (it >= 3) mbar_wait(smem_u32(&bar[slot]), (uint32_t)((it / 3 - 1) & 1));
      const uint32_t acc = tmem + (uint32_t)(slot * (n <= 160 ? n : 0));
      // nchain = 7 groups of 5 K steps (35 MMAs), like one diagonal of the sliced contraction
      for (int c = 0; c < nchain; c += 5) {
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) umma_i8(acc, adv[ks], bdv[ks], idesc, (c + ks) > 0);
      }
      umma_commit(smem_u32(&bar[slot]));
    }
    for (int it = (iters > 3 ? iters - 3 : 0); it < iters; ++it)
      mbar_wait(smem_u32(&bar[it % 3]), (uint32_t)((it / 3) & 1));
    cycles[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    uint32_t r0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];"
                 : "=r"(r0)
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (r0 == 0x12345678u) sink[tid] = (int32_t)r0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

static bool run_check(int n, int k, int a_signed, int swap = 0) {
  std::vector<uint8_t> ha((size_t)M * k);
  std::vector<int8_t> hb((size_t)n * k);
  srand(1234 + n + k + a_signed);
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
  k_check<<<1, 128, smem>>>(da, db, n, k, a_signed, swap, dout);
  CHECK(cudaGetLastError());
  CHECK(cudaDeviceSynchronize());
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
        if (bad < 5)
          printf("  mismatch (%d,%d): got %d want %lld\n", i, j, hout[(size_t)i * n + j], ref);
        ++bad;
      }
    }
  printf("check N=%3d K=%3d A %s %s: %s (%lld mismatches)\n", n, k, a_signed ? "s8" : "u8",
         swap ? "[LBO/SBO swapped] " : "", bad ? "FAIL" : "ok", bad);
  cudaFree(da);
  cudaFree(db);
  cudaFree(dout);
  return bad == 0;
}

static void run_rate(int n, int kbytes, int nchain, int iters, int sms, double clock_ghz) {
  long long *dcyc;
  int32_t *dsink;
  CHECK(cudaMalloc(&dcyc, sizeof(long long) * sms));
  CHECK(cudaMalloc(&dsink, sizeof(int32_t) * 128));
  const size_t smem = (size_t)(M + n) * kbytes;
  CHECK(cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_rate<<<sms, 128, smem>>>(n, kbytes, nchain, 4, dcyc, dsink);   // warm-up
  CHECK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  k_rate<<<sms, 128, smem>>>(n, kbytes, nchain, iters, dcyc, dsink);
  cudaEventRecord(e1);
  CHECK(cudaGetLastError());
  CHECK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(sms);
  CHECK(cudaMemcpy(h.data(), dcyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (auto c : h) mx = c > mx ? c : mx;
  const double mmas = (double)nchain * iters;
  const double ops = 2.0 * M * n * 32 * mmas * sms;
  printf("rate  N=%3d  chain=%3d x %5d : %7.1f cycles/MMA (ideal %5.1f)  %.3f ms  %.1f TOP/s "
         "(event-timed, %d SMs)  smem operand bytes/cycle = %.0f\n",
         n, nchain, iters, (double)mx / mmas, 128.0 * n / 256.0, ms, ops / (ms * 1e-3) / 1e12, sms,
         (double)(M + n) * 32 / ((double)mx / mmas));
  cudaFree(dcyc);
  cudaFree(dsink);
}

int main(int argc, char **argv) {
  (void)argc;
  (void)argv;
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s, %d SMs, sm_%d%d\n", prop.name, prop.multiProcessorCount, prop.major,
         prop.minor);
  bool ok = true;
  ok &= run_check(16, 32, 1);
  ok &= run_check(144, 32, 1);
  ok &= run_check(144, 160, 1);
  ok &= run_check(144, 160, 0);
  ok &= run_check(80, 160, 0);
  ok &= run_check(256, 64, 1);
  if (!ok) printf("CORRECTNESS FAILED - rates below are meaningless\n");
  const int sms = prop.multiProcessorCount;
  for (int n : {64, 80, 128, 144, 160, 256}) run_rate(n, 160, 35, 2000, sms, 1.9);
  run_rate(144, 160, 35, 2000, 1, 1.9);
  return ok ? 0 : 2;
}
