// FciGraph on the device: Knowles-Handy string tables and signed E_ij maps.
//
// Replaces FciGraph.__init__ (reference src/fqe/fci_graph.py:108-154) and the C
// helpers under it (lib/fci_graph.c:27-147, lib/bitstring.c:29-46).  Everything
// integer here must be BIT-EXACT with the reference:
//   * strings are enumerated in ascending-integer (Gosper / colex) order and then
//     stored at their Knowles-Handy address sum_k Z[k, o_k]  (fci_graph.py:301-333)
//   * a^+_i a_j |s> = (-1)^{popcount(s between i and j)} |t>   (fci_graph.py:229-237)
// The reference loops serially over strings; here one thread unranks one string
// (combinatorial number system) and one thread evaluates one (pair, string) map
// entry, so the tables for norb=16 (2 x 13 MB) are built in well under a ms.
#include "fqeb_common.cuh"

#include <map>
#include <mutex>

#include <stdlib.h>
#include <string.h>
#include <vector>

namespace fqeb {

// mutex + per-stream scratch sets of a graph (see fqeb_graph::sync)
struct GraphSync {
  std::mutex mu;
  std::map<cudaStream_t, GraphScratch> sets;
};

int graph_scratch(const fqeb_graph *g, cudaStream_t stream, GraphScratch *out) {
  GraphSync *sy = static_cast<GraphSync *>(g->sync);
  std::lock_guard<std::mutex> lock(sy->mu);
  auto it = sy->sets.find(stream);
  if (it != sy->sets.end()) {
    *out = it->second;
    return FQEB_OK;
  }
  GraphScratch sc{nullptr, {nullptr, nullptr}};
  if (sy->sets.empty()) {
    // the first stream that shows up gets the set allocated with the graph
    sc.small = g->d_small;
    sc.sterm[0] = g->d_sterm[0];
    sc.sterm[1] = g->d_sterm[1];
  } else {
    if (cudaMalloc(&sc.small, g->small_bytes) != cudaSuccess ||
        cudaMalloc(&sc.sterm[0], sizeof(double) * 2 * (g->len[0] > 0 ? g->len[0] : 1)) != cudaSuccess ||
        cudaMalloc(&sc.sterm[1], sizeof(double) * 2 * (g->len[1] > 0 ? g->len[1] : 1)) != cudaSuccess) {
      if (sc.small) cudaFree(sc.small);
      if (sc.sterm[0]) cudaFree(sc.sterm[0]);
      set_error("graph_scratch: cannot allocate a scratch set for a new stream");
      return FQEB_ERR_NOMEM;
    }
  }
  sy->sets[stream] = sc;
  *out = sc;
  return FQEB_OK;
}

GraphLock::GraphLock(const fqeb_graph *g) : mu(&static_cast<GraphSync *>(g->sync)->mu) {
  static_cast<std::mutex *>(mu)->lock();
}
GraphLock::~GraphLock() { static_cast<std::mutex *>(mu)->unlock(); }


static void host_binom(uint64_t *b /*[65*65]*/) {
  for (int n = 0; n < 65; ++n) {
    for (int k = 0; k < 65; ++k) {
      uint64_t v;
      if (k > n) v = 0;
      else if (k == 0 || k == n) v = 1;
      else v = b[(n - 1) * 65 + k - 1] + b[(n - 1) * 65 + k];
      b[n * 65 + k] = v;
    }
  }
}

// Z[k-1, l-1] = sum_{m=norb-l+1}^{norb-k} [C(m, n-k) - C(m-1, n-k-1)]   (k < n)
// Z[n-1, l-1] = l - n                                   (fci_graph.py:88-95)
static void host_z_matrix(const uint64_t *binom, int norb, int nele, int32_t *z) {
  memset(z, 0, sizeof(int32_t) * (size_t)nele * norb);
  if (nele == 0) return;
  for (int k = 1; k < nele; ++k) {
    for (int l = k; l <= norb - nele + k; ++l) {
      int64_t acc = 0;
      for (int m = norb - l + 1; m <= norb - k; ++m) {
        acc += (int64_t)binom[m * 65 + (nele - k)];
        if (m >= 1 && nele - k - 1 >= 0) acc -= (int64_t)binom[(m - 1) * 65 + (nele - k - 1)];
      }
      z[(k - 1) * norb + (l - 1)] = (int32_t)acc;
    }
  }
  for (int l = nele; l <= norb; ++l) z[(nele - 1) * norb + (l - 1)] = l - nele;
}

// one thread per rank r in ascending-integer order
__global__ void k_build_strings(int nele, int norb, int64_t len,
                                const uint64_t *__restrict__ binom,
                                const int32_t *__restrict__ z,
                                uint64_t *__restrict__ out) {
  const int64_t r0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r0 >= len) return;
  uint64_t r = (uint64_t)r0;
  uint64_t s = 0;
  int c = norb;  // candidates are < c
  for (int t = nele; t >= 1; --t) {
    // largest c' < c with C(c', t) <= r
    int cc = c - 1;
    while (cc > 0 && binom[cc * 65 + t] > r) --cc;
    s |= (1ull << cc);
    r -= binom[cc * 65 + t];
    c = cc;
  }
  out[fqeb_string_address(s, z, norb)] = s;
}

// adjoint-map entry for pair index p = i*norb + j and string x:
//   a^+_j a_i |x> = sign |y>   ->  sign*(y+1), else 0
__global__ void k_build_maps(int norb, int64_t len,
                             const uint64_t *__restrict__ str,
                             const int32_t *__restrict__ z,
                             int32_t *__restrict__ amap,
                             int32_t *__restrict__ amapT) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if (x >= len) return;
  const int i = p / norb, j = p % norb;
  const uint64_t s = str[x];
  const uint64_t bi = 1ull << i, bj = 1ull << j;
  int32_t val = 0;
  if (s & bi) {
    if (i == j) {
      val = (int32_t)(x + 1);
    } else if (!(s & bj)) {
      const int lo = i < j ? i : j, hi = i < j ? j : i;
      const uint64_t between = ((1ull << hi) - 1) & ~((2ull << lo) - 1);
      const int par = __popcll(s & between) & 1;
      const uint64_t t = (s & ~bi) | bj;
      const int y = fqeb_string_address(t, z, norb);
      val = par ? -(y + 1) : (y + 1);
    }
  }
  amap[(int64_t)p * len + x] = val;
  amapT[x * (int64_t)(norb * norb) + p] = val;
}

// merged map of the compressed pair space (see fqeb_graph::d_smap)
__global__ void k_build_symmaps(int norb, int64_t len, const int32_t *__restrict__ amapT,
                                int32_t *__restrict__ smap, int32_t *__restrict__ smapT) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= len) return;
  const int npc = norb * (norb + 1) / 2;
  const int32_t *row = amapT + x * (int64_t)(norb * norb);
  int c = 0;
  for (int i = 0; i < norb; ++i)
    for (int j = 0; j <= i; ++j, ++c) {
      int v = row[i * norb + j];
      if (v == 0 && i != j) v = row[j * norb + i];
      smap[(int64_t)c * len + x] = v;
      smapT[x * (int64_t)npc + c] = v;
    }
}

// compact the non-zero adjoint-map entries of each string (exactly lk of them)
__global__ void k_build_clists(int npair, int64_t len, int lk, const int32_t *__restrict__ amapT,
                               int2 *__restrict__ clistT, int2 *__restrict__ clist) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= len) return;
  const int32_t *row = amapT + x * (int64_t)npair;
  int slot = 0;
  for (int p = 0; p < npair && slot < lk; ++p) {
    const int v = row[p];
    if (v != 0) {
      const int2 e = make_int2(p, v);
      clistT[x * (int64_t)lk + slot] = e;
      clist[(int64_t)slot * len + x] = e;
      ++slot;
    }
  }
}

static int build_spin(fqeb_graph *g, int spin, const uint64_t *d_binom) {
  const int norb = g->norb, nele = g->nele[spin];
  const int64_t len = g->len[spin];
  const int npair = norb * norb;
  const size_t zbytes = sizeof(int32_t) * (size_t)(nele > 0 ? nele : 1) * norb;
  FQEB_CUDA(cudaMalloc(&g->d_Z[spin], zbytes));
  FQEB_CUDA(cudaMemcpy(g->d_Z[spin], g->h_Z[spin], sizeof(int32_t) * (size_t)nele * norb,
                       cudaMemcpyHostToDevice));
  FQEB_CUDA(cudaMalloc(&g->d_str[spin], sizeof(uint64_t) * len));
  FQEB_CUDA(cudaMalloc(&g->d_amap[spin], sizeof(int32_t) * (size_t)npair * len));
  FQEB_CUDA(cudaMalloc(&g->d_amapT[spin], sizeof(int32_t) * (size_t)npair * len));
  const int threads = 256;
  const unsigned blocks = (unsigned)((len + threads - 1) / threads);
  k_build_strings<<<blocks, threads>>>(nele, norb, len, d_binom, g->d_Z[spin], g->d_str[spin]);
  FQEB_CHECK_LAUNCH();
  if (npair > 0) {
    dim3 grid(blocks, npair);
    k_build_maps<<<grid, threads>>>(norb, len, g->d_str[spin], g->d_Z[spin], g->d_amap[spin],
                                    g->d_amapT[spin]);
    FQEB_CHECK_LAUNCH();
    const size_t sbytes = sizeof(int32_t) * (size_t)(norb * (norb + 1) / 2) * len;
    FQEB_CUDA(cudaMalloc(&g->d_smap[spin], sbytes));
    FQEB_CUDA(cudaMalloc(&g->d_smapT[spin], sbytes));
    k_build_symmaps<<<blocks, threads>>>(norb, len, g->d_amapT[spin], g->d_smap[spin],
                                         g->d_smapT[spin]);
    FQEB_CHECK_LAUNCH();
  }
  const int lk = nele * (norb - nele + 1);
  g->lk[spin] = lk;
  const size_t cbytes = sizeof(int2) * (size_t)(lk > 0 ? lk : 1) * len;
  FQEB_CUDA(cudaMalloc(&g->d_clistT[spin], cbytes));
  FQEB_CUDA(cudaMalloc(&g->d_clist[spin], cbytes));
  if (lk > 0) {
    k_build_clists<<<blocks, threads>>>(npair, len, lk, g->d_amapT[spin], g->d_clistT[spin],
                                        g->d_clist[spin]);
    FQEB_CHECK_LAUNCH();
  }
  return FQEB_OK;
}

}  // namespace fqeb

using namespace fqeb;

extern "C" int fqeb_graph_create(int norb, int nalpha, int nbeta, fqeb_graph **out) {
  FQEB_REQUIRE(out != nullptr, "fqeb_graph_create: out is NULL");
  *out = nullptr;
  FQEB_REQUIRE(norb >= 0 && norb <= kMaxOrb, "fqeb_graph_create: norb=%d outside [0,%d]", norb,
               kMaxOrb);
  FQEB_REQUIRE(nalpha >= 0 && nalpha <= norb, "fqeb_graph_create: nalpha=%d invalid for norb=%d",
               nalpha, norb);
  FQEB_REQUIRE(nbeta >= 0 && nbeta <= norb, "fqeb_graph_create: nbeta=%d invalid for norb=%d",
               nbeta, norb);
  int rc = require_device();
  if (rc != FQEB_OK) return rc;

  std::vector<uint64_t> binom(65 * 65);
  host_binom(binom.data());
  const uint64_t la = binom[norb * 65 + nalpha], lb = binom[norb * 65 + nbeta];
  // the reference indexes strings with C int (lib/fci_graph.c:139); keep that bound and
  // require the adjoint map (norb^2 * len entries) to be addressable.
  FQEB_REQUIRE(la < (1ull << 31) && lb < (1ull << 31),
               "fqeb_graph_create: string space too large (%llu x %llu)",
               (unsigned long long)la, (unsigned long long)lb);

  fqeb_graph *g = (fqeb_graph *)calloc(1, sizeof(fqeb_graph));
  if (!g) {
    set_error("fqeb_graph_create: host allocation failed");
    return FQEB_ERR_NOMEM;
  }
  g->norb = norb;
  g->nele[0] = nalpha;
  g->nele[1] = nbeta;
  g->len[0] = (int64_t)la;
  g->len[1] = (int64_t)lb;
  g->shared_spin = (nalpha == nbeta);
  cudaGetDevice(&g->device);

  uint64_t *d_binom = nullptr;
  auto fail = [&](int code) {
    if (d_binom) cudaFree(d_binom);
    fqeb_graph_destroy(g);
    return code;
  };
  if (cudaMalloc(&d_binom, sizeof(uint64_t) * 65 * 65) != cudaSuccess ||
      cudaMemcpy(d_binom, binom.data(), sizeof(uint64_t) * 65 * 65, cudaMemcpyHostToDevice) !=
          cudaSuccess) {
    set_error("fqeb_graph_create: cannot upload binomial table: %s",
              cudaGetErrorString(cudaGetLastError()));
    return fail(FQEB_ERR_CUDA);
  }
  for (int spin = 0; spin < 2; ++spin) {
    const int nele = g->nele[spin];
    g->h_Z[spin] = (int32_t *)calloc((size_t)(nele > 0 ? nele : 1) * (norb > 0 ? norb : 1),
                                     sizeof(int32_t));
    host_z_matrix(binom.data(), norb, nele, g->h_Z[spin]);
    if (spin == 1 && g->shared_spin) {
      g->d_Z[1] = g->d_Z[0];
      g->d_str[1] = g->d_str[0];
      g->d_amap[1] = g->d_amap[0];
      g->d_amapT[1] = g->d_amapT[0];
      g->d_smap[1] = g->d_smap[0];
      g->d_smapT[1] = g->d_smapT[0];
      g->lk[1] = g->lk[0];
      g->d_clistT[1] = g->d_clistT[0];
      g->d_clist[1] = g->d_clist[0];
      continue;
    }
    rc = build_spin(g, spin, d_binom);
    if (rc != FQEB_OK) return fail(rc);
  }
  g->sync = new GraphSync();
  g->small_bytes = sizeof(double) * 2 * (size_t)(4 * 64 * 64 + 4 * 64);
  if (cudaMalloc(&g->d_small, g->small_bytes) != cudaSuccess ||
      cudaMalloc(&g->d_sterm[0], sizeof(double) * 2 * g->len[0]) != cudaSuccess ||
      cudaMalloc(&g->d_sterm[1], sizeof(double) * 2 * g->len[1]) != cudaSuccess) {
    set_error("fqeb_graph_create: scratch allocation failed");
    return fail(FQEB_ERR_CUDA);
  }
  {
    const int npair = norb * norb;
    std::vector<int32_t> ids(3 * (size_t)(npair > 0 ? npair : 1));
    for (int p = 0; p < npair; ++p) {
      ids[2 * p] = p;
      ids[2 * p + 1] = -1;
      ids[2 * npair + p] = p;
    }
    if (cudaMalloc(&g->d_pairs_id, sizeof(int32_t) * ids.size()) != cudaSuccess ||
        cudaMemcpy(g->d_pairs_id, ids.data(), sizeof(int32_t) * ids.size(),
                   cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("fqeb_graph_create: pair table allocation failed");
      return fail(FQEB_ERR_CUDA);
    }
    g->d_rowmap_id = g->d_pairs_id + 2 * npair;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) {
    set_error("fqeb_graph_create: table kernels failed: %s",
              cudaGetErrorString(cudaGetLastError()));
    return fail(FQEB_ERR_CUDA);
  }
  cudaFree(d_binom);
  *out = g;
  return FQEB_OK;
}

extern "C" int fqeb_graph_destroy(fqeb_graph *g) {
  if (!g) return FQEB_OK;
  const int nspin = g->shared_spin ? 1 : 2;
  for (int s = 0; s < nspin; ++s) {
    if (g->d_Z[s]) cudaFree(g->d_Z[s]);
    if (g->d_str[s]) cudaFree(g->d_str[s]);
    if (g->d_amap[s]) cudaFree(g->d_amap[s]);
    if (g->d_amapT[s]) cudaFree(g->d_amapT[s]);
    if (g->d_smap[s]) cudaFree(g->d_smap[s]);
    if (g->d_smapT[s]) cudaFree(g->d_smapT[s]);
    if (g->d_occ[s]) cudaFree(g->d_occ[s]);
    if (g->d_unocc[s]) cudaFree(g->d_unocc[s]);
    if (g->d_clistT[s]) cudaFree(g->d_clistT[s]);
    if (g->d_clist[s]) cudaFree(g->d_clist[s]);
  }
  for (int s = 0; s < 2; ++s)   // always one table per spin (the row lengths differ)
    for (int y = 0; y < 2; ++y)
      if (g->d_ozmapT[y][s]) cudaFree(g->d_ozmapT[y][s]);
  for (int s = 0; s < 2; ++s) free(g->h_Z[s]);
  if (g->sync) {
    GraphSync *sy = static_cast<GraphSync *>(g->sync);
    for (auto &kv : sy->sets) {
      if (kv.second.small == g->d_small) continue;   // the graph's own set, freed below
      cudaFree(kv.second.small);
      cudaFree(kv.second.sterm[0]);
      cudaFree(kv.second.sterm[1]);
    }
    delete sy;
  }
  if (g->d_small) cudaFree(g->d_small);
  if (g->d_pairs_id) cudaFree(g->d_pairs_id);
  for (int s = 0; s < 2; ++s)
    if (g->d_sterm[s]) cudaFree(g->d_sterm[s]);
  free(g);
  return FQEB_OK;
}

extern "C" int fqeb_graph_dims(const fqeb_graph *g, int *norb, int *nalpha, int *nbeta,
                               int64_t *lena, int64_t *lenb) {
  FQEB_REQUIRE(g != nullptr, "fqeb_graph_dims: NULL graph");
  if (norb) *norb = g->norb;
  if (nalpha) *nalpha = g->nele[0];
  if (nbeta) *nbeta = g->nele[1];
  if (lena) *lena = g->len[0];
  if (lenb) *lenb = g->len[1];
  return FQEB_OK;
}

extern "C" int fqeb_graph_get_Z(const fqeb_graph *g, int spin, int32_t *h_out) {
  FQEB_REQUIRE(g && h_out && (spin == 0 || spin == 1), "fqeb_graph_get_Z: bad argument");
  memcpy(h_out, g->h_Z[spin], sizeof(int32_t) * (size_t)g->nele[spin] * g->norb);
  return FQEB_OK;
}

extern "C" int fqeb_graph_get_strings(const fqeb_graph *g, int spin, uint64_t *h_out) {
  FQEB_REQUIRE(g && h_out && (spin == 0 || spin == 1), "fqeb_graph_get_strings: bad argument");
  FQEB_CUDA(cudaMemcpy(h_out, g->d_str[spin], sizeof(uint64_t) * g->len[spin],
                       cudaMemcpyDeviceToHost));
  return FQEB_OK;
}

extern "C" int fqeb_graph_get_map(const fqeb_graph *g, int spin, int32_t *h_out) {
  FQEB_REQUIRE(g && h_out && (spin == 0 || spin == 1), "fqeb_graph_get_map: bad argument");
  const int norb = g->norb;
  const int64_t len = g->len[spin];
  // device holds the adjoint map (pair (i,j) <-> a^+_j a_i); export the forward one.
  for (int i = 0; i < norb; ++i) {
    for (int j = 0; j < norb; ++j) {
      FQEB_CUDA(cudaMemcpy(h_out + (int64_t)(i * norb + j) * len,
                           g->d_amap[spin] + (int64_t)(j * norb + i) * len,
                           sizeof(int32_t) * len, cudaMemcpyDeviceToHost));
    }
  }
  return FQEB_OK;
}

extern "C" int fqeb_graph_device_tables(const fqeb_graph *g, int spin,
                                        const uint64_t **d_strings,
                                        const int32_t **d_map_by_pair,
                                        const int32_t **d_map_by_string) {
  FQEB_REQUIRE(g && (spin == 0 || spin == 1), "fqeb_graph_device_tables: bad argument");
  if (d_strings) *d_strings = g->d_str[spin];
  if (d_map_by_pair) *d_map_by_pair = g->d_amap[spin];
  if (d_map_by_string) *d_map_by_string = g->d_amapT[spin];
  return FQEB_OK;
}
