// Shared internals of libfqe_b200.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "fqe_b200.h"

namespace fqeb {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define FQEB_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t err__ = (call);                                                \
    if (err__ != cudaSuccess) {                                                \
      fqeb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,       \
                      cudaGetErrorString(err__));                              \
      return FQEB_ERR_CUDA;                                                    \
    }                                                                          \
  } while (0)

#define FQEB_CHECK_LAUNCH()                                                    \
  do {                                                                         \
    cudaError_t err__ = cudaGetLastError();                                    \
    if (err__ != cudaSuccess) {                                                \
      fqeb::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__,   \
                      cudaGetErrorString(err__));                              \
      return FQEB_ERR_CUDA;                                                    \
    }                                                                          \
    fqeb::count_launch();                                                      \
  } while (0)

#define FQEB_REQUIRE(cond, ...)                                                \
  do {                                                                         \
    if (!(cond)) {                                                             \
      fqeb::set_error(__VA_ARGS__);                                            \
      return FQEB_ERR_INVALID;                                                 \
    }                                                                          \
  } while (0)

int require_device();  // FQEB_OK or FQEB_ERR_NODEVICE (with message)
int sm_count();
// stream-ordered allocation + upload on the internal non-blocking stream; upload_finish()
// blocks the host until every upload issued so far has landed
int upload_alloc(void **d_ptr, const void *h_src, size_t bytes);
int upload_finish();

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// Kernel attributes (dynamic shared-memory limits) belong to the device: call sites keep one
// of these as a function-local static and (re)apply their attributes when the requested size
// exceeds what the CURRENT device has been given so far.
struct PerDeviceSize {
  std::atomic<size_t> have[64];
  PerDeviceSize() {
    for (auto &h : have) h.store(0);
  }
  // true if `want` exceeds the size recorded for the current device (the caller then sets the
  // attributes and calls record())
  bool needs(size_t want, int *dev_out) {
    int dev = 0;
    cudaGetDevice(&dev);
    *dev_out = dev & 63;
    return want > have[dev & 63].load();
  }
  void record(int dev, size_t want) { have[dev].store(want); }
};

// scratch set of `stream` for graph g (allocated on first use, thread-safe)
struct GraphScratch {
  double *small;
  double *sterm[2];
};
int graph_scratch(const fqeb_graph *g, cudaStream_t stream, GraphScratch *out);
// RAII lock of the graph's mutex (lazy tables such as the occupancy lists)
struct GraphLock {
  explicit GraphLock(const fqeb_graph *g);
  ~GraphLock();
  void *mu;
};

constexpr int kMaxOrb = 63;  // as the reference C path (settings.py c_string_max_norb)

}  // namespace fqeb

// ---- handle definitions -----------------------------------------------------
// Knowles-Handy address of a string: sum_k Z[k, o_k] over its occupied orbitals o_0 < o_1 < ...
// (reference fci_graph.py:401-419, lib/fci_graph.c:123-134)
__device__ __forceinline__ int fqeb_string_address(uint64_t s, const int32_t *__restrict__ z,
                                               int norb) {
  int addr = 0, k = 0;
  while (s) {
    const int bit = __ffsll((long long)s) - 1;
    s &= s - 1;
    addr += z[k * norb + bit];
    ++k;
  }
  return addr;
}


struct fqeb_graph {
  int norb, nele[2];
  int64_t len[2];
  int device;
  bool shared_spin;        // nalpha == nbeta: beta tables alias alpha tables
  int32_t *h_Z[2];         // host [nele][norb]
  int32_t *d_Z[2];
  uint64_t *d_str[2];      // [len]
  // adjoint map: amap[ij][x] = sign*(y+1) with a^+_j a_i |x> = sign |y>  (ij = i*norb+j)
  int32_t *d_amap[2];      // [norb*norb][len]
  int32_t *d_amapT[2];     // [len][norb*norb]
  // merged map of the compressed pair space c = i(i+1)/2 + j (i >= j): the entry of ij or
  // of ji, whichever is non-zero (for i != j at most one is: i occupied and j empty, or the
  // reverse), so that D_c = D[ij] + D[ji] is a single-source gather
  int32_t *d_smap[2];      // [norb(norb+1)/2][len]
  int32_t *d_smapT[2];     // [len][norb(norb+1)/2]
  // compact lists of the lk = nele*(norb-nele+1) non-vanishing adjoint-map entries of
  // every string, as int2 (ij, sign*(y+1)), ascending ij:
  int lk[2];
  int2 *d_clistT[2];       // [len][lk]   by string  (warp-uniform access per row)
  int2 *d_clist[2];        // [lk][len]   by slot    (coalesced access per column)
  // Scratch of the diagonal / column / n-body entry points: one set PER CUDA STREAM (graphs are
  // shared process-wide per sector and device, and two sectors' calls may be in flight on
  // different streams or host threads).  d_small / d_sterm are the first set; the others hang
  // off `sync` and are handed out by fqeb::graph_scratch().
  double *d_small;         // small operator uploads (diag, v, ...)
  size_t small_bytes;
  double *d_sterm[2];      // per-string terms / by-target tables, 16 bytes per string
  void *sync;              // fqeb::GraphSync*: mutex + per-stream scratch sets + lazy-table guard
  // ordered lists of the strings with orbital icol occupied / empty, built on first use by
  // the column-rotation kernels (rotate.cu): [norb][C(norb-1,nele-1)] / [norb][C(norb-1,nele)]
  int32_t *d_occ[2], *d_unocc[2];
  // by-string source tables of the sliced contraction (ozaki.cu k_source_table), zero-padded to a
  // multiple of 16 columns, [compressed pair space?][spin], built on first use
  int32_t *d_ozmapT[2][2];
  int32_t *d_pairs_id;     // identity pair list [norb^2][2] = (ij, -1)
  int32_t *d_rowmap_id;    // identity row map [norb^2]
};

struct fqeb_op {
  int norb;
  int kind;           // FQEB_OP_*
  bool has_h2;
  int device;
  bool sym;           // h2'[ij,kl] symmetric under i<->j and k<->l: pairs compressed to i>=j
  int np;             // size of the pair space the contraction runs over:
                      //   norb^2, or norb(norb+1)/2 when sym
  int32_t *d_pairs;   // [np][2] excitation pairs summed into D row c (second -1 if none)
  int32_t *d_rowmap;  // [norb^2] E row holding pair kl
  // GEMM operand: real row-major [Mp][Kp]; layout depends on kind (see dgemm.cu)
  double *d_A;
  int Mp, Kp;         // padded real dims of the FULL operator (whole pair space)
  double *d_h1;       // complex [norb*norb] (h1' / z), interleaved
  double zr, zi;      // global factor: 1 (real/complex) or i (imag)
  // fused gather+contraction (dgemm.cu k_sigma_fused): usable when one CTA covers the
  // whole row space and the one-body term can be folded into the operand
  bool fuse_ok;
  bool absorb_ok;          // one-body term can be folded into the contraction operand
  double *h_h1p, *h_h2p;   // host copies (complex, interleaved) for per-sector operands
  void *fused_cache;       // std::map<int, double*>*: n_elec -> device operand with h1 absorbed
};
