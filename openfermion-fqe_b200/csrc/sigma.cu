// sigma = H C for a dense restricted (h1', h2') operator: the FCI sigma vector.
//
// Replaces FqeData._apply_array_spatial12 (reference src/fqe/fqe_data.py:582-608,
// 644-710).  Algorithm (Knowles-Handy, the one BASELINE.json's north_star names and
// the reference's Python branch spells out at fqe_data.py:653-657):
//
//     D[ij]  = E_ij C                      gather        (dvec.cu,  HBM-bound)
//     E[kl]  = sum_ij h2'[kl,ij] D[ij]     DMMA GEMM     (dgemm.cu, FP64-bound)
//     sigma += sum_kl E_kl^T E[kl]         scatter       (dvec.cu,  HBM-bound)
// with the first two steps fused into one kernel where the operator allows it (D tiles are
// gathered straight into the GEMM's shared-memory ring and never reach HBM).
//
// The one-body term sum_ij h1'[ij] D[ij] costs nothing: on a sector with n_elec electrons
// sum_k D[kk] = n_elec C, so h1'[kl]/n_elec is added to the operand's diagonal-pair columns
// (absorbed_operand, dgemm.cu).  Operators whose h1' would leave the class of h2' (e.g. a
// complex h1' next to a real h2') get a separate accumulation pass instead.
//
// D and E are norb^2 times larger than C (678 GB at norb=16), so the determinant
// index is streamed in chunks of alpha rows through a caller-provided workspace;
// all three kernels of a chunk are enqueued back-to-back on one stream.  A rank of
// a multi-GPU job passes its shard as [row0,row1) (determinant rows) and/or
// [ij0,ij1) (pair slice, the partition north_star mandates) and obtains a partial
// sigma that one allreduce completes.
#include "fqeb_common.cuh"

#include <math.h>
#include <string.h>
#include <map>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

namespace fqeb {
int launch_make_dvec(const fqeb_graph *g, const double *d_coeff, double *d_dvec, int64_t ldd,
                     int64_t row0, int64_t nrows, int ij0, int ij1, const int32_t *d_pairs,
                     int np_eff, const double *d_h1, double *d_sig, cudaStream_t st);
int launch_make_coeff(const fqeb_graph *g, const double *d_evec, int64_t lde, int64_t pitch,
                      int64_t row0, int64_t nrows, const int32_t *d_rowmap, double zr, double zi,
                      double *d_out, cudaStream_t st, int64_t x0 = 0, int64_t x1 = -1);
int launch_one_body(const fqeb_graph *g, const double *d_coeff, const double *d_h1, int64_t row0,
                    int64_t nrows, double *d_out, cudaStream_t st);
int launch_fused(const fqeb_graph *g, const fqeb_op *op, const double *d_A, const double *d_coeff,
                 int64_t row0, int64_t nrows, int pitch, double *d_evec, int64_t lde, int ij0,
                 int ij1, cudaStream_t st);
int launch_gather(const fqeb_graph *g, bool sym, const double *d_coeff, double *d_dvec,
                  int64_t ldd, int64_t row0, int64_t nrows, int c0, int c1, cudaStream_t st);
int absorbed_operand(const fqeb_op *op, int n_elec, const double **d_A);
int launch_contract(const fqeb_op *op, const double *d_A, const double *d_dvec, int64_t ldd,
                    double *d_evec, int64_t lde, int64_t ncols, int ij0, int ij1,
                    cudaStream_t st);
int dvec_rows_padded(const fqeb_op *op, int nij);
int dvec_rows_zeroed(const fqeb_op *op, int nij);
// INT8-sliced tensor-core contraction (ozaki.cu)
bool ozaki_shape_ok(const fqeb_op *op);
size_t ozaki_workspace_bytes(const fqeb_graph *g);
double *ozaki_stats_ptr(const fqeb_graph *g, void *d_oz);
int ozaki_stats(const fqeb_graph *g, const double *d_coeff, double *d_stats, double *h_absmax,
                double *h_sumsq, double *h_nonzero, cudaStream_t st);
int ozaki_slice(const fqeb_graph *g, const double *d_coeff, const double *d_stats, void *d_planes,
                cudaStream_t st);
double ozaki_error_estimate(double absmax, double sumsq, double nonzero);
int launch_ozaki(const fqeb_graph *g, const fqeb_op *op, const void *d_planes,
                 const double *d_stats, int64_t row0, int64_t nrows, int pitch, double *d_evec,
                 int64_t lde, cudaStream_t st);

// ---- optional per-phase event timing ------------------------------------------------
struct PhaseEvent {
  int phase;
  cudaEvent_t start, stop;
};
static bool g_profile = false;
static std::mutex g_profile_mu;
static std::vector<PhaseEvent> g_phase_events;

struct PhaseTimer {
  bool on;
  PhaseEvent ev;
  cudaStream_t st;
  PhaseTimer(int phase, cudaStream_t s) : on(g_profile), st(s) {
    if (!on) return;
    ev.phase = phase;
    if (cudaEventCreate(&ev.start) != cudaSuccess || cudaEventCreate(&ev.stop) != cudaSuccess) {
      on = false;
      return;
    }
    cudaEventRecord(ev.start, st);
  }
  ~PhaseTimer() {
    if (!on) return;
    cudaEventRecord(ev.stop, st);
    std::lock_guard<std::mutex> lock(g_profile_mu);
    g_phase_events.push_back(ev);
  }
};

static std::atomic<int> g_last_path{0};   // FQEB_PATH_* of the most recent two-body sigma build
// Factor on FQEB_OZAKI_TOL for the builds issued by this thread: fqeb_taylor loosens the
// quantisation bound for terms whose weight in the propagated state is small (see there).
static thread_local double g_ozaki_tol_scale = 1.0;

struct ChunkLayout {
  bool fused;         // D never materialised (k_sigma_fused / k_sigma_ozaki)
  bool ozaki;         // the INT8-sliced contraction may be used: room for the digit planes
  size_t oz_bytes;    // digit planes + statistics block, at the start of the workspace
  int64_t pitch;      // complex elements between consecutive alpha rows inside a D/E row
  int64_t ldd;        // complex elements per D / E row
  int64_t d_rows;     // rows of D (pair slice, padded); 0 when fused
  int64_t e_rows;     // rows of E
  size_t d_bytes, e_bytes;
};

// The fused gather+contraction kernel (D never written to HBM, dgemm.cu k_sigma_fused) is the
// default wherever it applies: one row block (pair space <= 144), real / imaginary operator
// class, absorbable one-body term, at least one electron.  Measured at norb=16: 393 ms fused
// against 79 ms gather + 365 ms contraction.  FQEB_FUSION=0 selects the three-kernel path.
static bool use_fused(const fqeb_graph *g, const fqeb_op *op) {
  const char *env = getenv("FQEB_FUSION");
  const bool enabled = !(env && env[0] == '0');
  return enabled && op->fuse_ok && (g->nele[0] + g->nele[1]) > 0;
}

// The sliced tensor-core contraction (ozaki.cu) replaces the DMMA stream of the fused kernel for
// whole-pair-space builds of real / imaginary operators with a pair space <= 144 (FQEB_OZAKI=0
// disables it); per call it is additionally gated on the state (see fqeb_sigma_restricted).
static bool use_ozaki(const fqeb_graph *g, const fqeb_op *op, int ij0, int ij1) {
  const char *env = getenv("FQEB_OZAKI");
  const bool enabled = !(env && env[0] == '0');
  return enabled && use_fused(g, op) && ozaki_shape_ok(op) && ij0 == 0 && ij1 == op->np;
}

static ChunkLayout layout_for(const fqeb_graph *g, const fqeb_op *op, int64_t rows, int ij0,
                              int ij1) {
  ChunkLayout L;
  L.fused = use_fused(g, op);
  L.ozaki = use_ozaki(g, op, ij0, ij1);
  L.oz_bytes = L.ozaki ? ozaki_workspace_bytes(g) : 0;
  L.e_rows = round_up(op->np, 8);
  if (L.fused) {
    L.pitch = round_up(g->len[1], 64);
    L.ldd = rows * L.pitch;
    L.d_rows = 0;
    L.d_bytes = 0;
  } else {
    L.pitch = g->len[1];
    L.ldd = round_up(rows * g->len[1], fqeb_gemm_col_align());
    L.d_rows = dvec_rows_padded(op, ij1 - ij0);
    L.d_bytes = (size_t)round_up(sizeof(double) * 2 * (size_t)L.d_rows * L.ldd, 256);
  }
  L.e_bytes = (size_t)round_up(sizeof(double) * 2 * (size_t)L.e_rows * L.ldd, 256);
  return L;
}

}  // namespace fqeb

using namespace fqeb;

static int check_shard(const fqeb_graph *g, const fqeb_op *op, int ij0, int ij1) {
  FQEB_REQUIRE(g && op, "sigma: NULL handle");
  FQEB_REQUIRE(op->norb == g->norb, "sigma: operator has %d orbitals, graph has %d", op->norb,
               g->norb);
  FQEB_REQUIRE(ij0 >= 0 && ij0 <= ij1 && ij1 <= op->np,
               "sigma: pair slice [%d,%d) outside the operator's pair space [0,%d)", ij0, ij1,
               op->np);
  return FQEB_OK;
}

extern "C" size_t fqeb_sigma_workspace_bytes(const fqeb_graph *g, const fqeb_op *op,
                                             int64_t rows_per_chunk, int ij0, int ij1) {
  if (check_shard(g, op, ij0, ij1) != FQEB_OK || rows_per_chunk <= 0) return 0;
  if (!op->has_h2 || ij0 == ij1) return 0;
  const ChunkLayout L = layout_for(g, op, rows_per_chunk, ij0, ij1);
  return L.oz_bytes + L.d_bytes + L.e_bytes;
}

extern "C" int64_t fqeb_sigma_rows_for_workspace(const fqeb_graph *g, const fqeb_op *op,
                                                 size_t bytes, int ij0, int ij1) {
  if (check_shard(g, op, ij0, ij1) != FQEB_OK) return -1;
  if (!op->has_h2 || ij0 == ij1) return g->len[0];
  int64_t lo = 0, hi = g->len[0];  // largest rows with workspace_bytes(rows) <= bytes
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) / 2;
    if (fqeb_sigma_workspace_bytes(g, op, mid, ij0, ij1) <= bytes) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

// The sigma build.  With `pending` the scatter of the LAST chunk is not launched but described
// there (fqeb_sigma_restricted_deferred): the caller completes sigma by target-row slices
// (fqeb_scatter_rows) and can start reducing finished slices across ranks meanwhile.
static int sigma_build(const fqeb_graph *g, const fqeb_op *op, const double *d_coeff,
                       double *d_sigma, void *d_workspace, size_t workspace_bytes, int64_t row0,
                       int64_t row1, int ij0, int ij1, void *stream,
                       fqeb_pending_scatter *pending) {
  if (pending) memset(pending, 0, sizeof(*pending));
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  rc = check_shard(g, op, ij0, ij1);
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(d_coeff && d_sigma && d_coeff != d_sigma, "sigma: coeff/sigma NULL or aliased");
  const int64_t lena = g->len[0], lenb = g->len[1];
  FQEB_REQUIRE(row0 >= 0 && row0 <= row1 && row1 <= lena, "sigma: row shard [%lld,%lld) invalid",
               (long long)row0, (long long)row1);
  cudaStream_t st = (cudaStream_t)stream;
  FQEB_CUDA(cudaMemsetAsync(d_sigma, 0, sizeof(double) * 2 * (size_t)lena * lenb, st));
  if (row0 == row1 || ij0 == ij1) return FQEB_OK;

  if (!op->has_h2) {
    // one-body operator: sigma = sum_ij h1[ij] D[ij], D never materialised.  The whole pair
    // range goes through the compact-list kernel; a pair slice (ij-sharded run) through the
    // gather kernel with stores disabled.
    if (ij0 == 0 && ij1 == op->np)
      return launch_one_body(g, d_coeff, op->d_h1, row0, row1 - row0, d_sigma, st);
    return launch_make_dvec(g, d_coeff, nullptr, 0, row0, row1 - row0, ij0, ij1, op->d_pairs,
                            op->np, op->d_h1, d_sigma, st);
  }
  FQEB_REQUIRE((ij0 & 1) == 0, "sigma: pair slice must start at an even index");
  const int64_t rows_max = fqeb_sigma_rows_for_workspace(g, op, workspace_bytes, ij0, ij1);
  if (rows_max < 1 || d_workspace == nullptr) {
    set_error("sigma: workspace of %zu bytes cannot hold one alpha row (need %zu)",
              workspace_bytes, fqeb_sigma_workspace_bytes(g, op, 1, ij0, ij1));
    return FQEB_ERR_NOMEM;
  }
  int64_t rows_chunk = rows_max < (row1 - row0) ? rows_max : (row1 - row0);
  // balance the chunks so the last one is not a sliver
  const int64_t nchunk = (row1 - row0 + rows_chunk - 1) / rows_chunk;
  rows_chunk = (row1 - row0 + nchunk - 1) / nchunk;
  const ChunkLayout L = layout_for(g, op, rows_chunk, ij0, ij1);
  double *d_dvec = (double *)((char *)d_workspace + L.oz_bytes);
  double *d_evec = (double *)((char *)d_workspace + L.oz_bytes + L.d_bytes);
  const int nij = ij1 - ij0;
  if (L.fused) {
    // INT8-sliced tensor-core contraction, unless the state is too strongly peaked for a
    // global fixed-point scale (estimated quantisation error above FQEB_OZAKI_TOL, default
    // 5e-12 relative): then the FP64 DMMA kernel builds the same E.
    bool sliced = L.ozaki;
    double *d_stats = nullptr;
    if (sliced) {
      static const double tol = getenv("FQEB_OZAKI_TOL") ? atof(getenv("FQEB_OZAKI_TOL")) : 5e-12;
      d_stats = ozaki_stats_ptr(g, d_workspace);
      double absmax = 0.0, sumsq = 0.0, nonzero = 0.0;
      PhaseTimer t(0, st);
      rc = ozaki_stats(g, d_coeff, d_stats, &absmax, &sumsq, &nonzero, st);
      if (rc != FQEB_OK) return rc;
      if (sumsq == 0.0) return FQEB_OK;   // zero vector: sigma is already zero
      // NaN / Inf in the state: no fixed-point image exists; the FP64 kernel propagates them
      sliced = isfinite(sumsq) && isfinite(absmax) &&
               ozaki_error_estimate(absmax, sumsq, nonzero) <= tol * g_ozaki_tol_scale;
      if (sliced) {
        rc = ozaki_slice(g, d_coeff, d_stats, d_workspace, st);
        if (rc != FQEB_OK) return rc;
      }
    }
    g_last_path.store(sliced ? FQEB_PATH_SLICED : FQEB_PATH_FUSED);
    const double *d_A = nullptr;
    if (!sliced) {
      rc = absorbed_operand(op, g->nele[0] + g->nele[1], &d_A);
      if (rc != FQEB_OK) return rc;
    }
    double *d_e = d_evec;
    for (int64_t a0 = row0; a0 < row1; a0 += rows_chunk) {
      const int64_t nr = (row1 - a0) < rows_chunk ? (row1 - a0) : rows_chunk;
      {
        PhaseTimer t(1, st);
        rc = sliced ? launch_ozaki(g, op, d_workspace, d_stats, a0, nr, (int)L.pitch, d_e, L.ldd, st)
                    : launch_fused(g, op, d_A, d_coeff, a0, nr, (int)L.pitch, d_e, L.ldd, ij0,
                                   ij1, st);
      }
      if (rc != FQEB_OK) return rc;
      if (pending && a0 + rows_chunk >= row1) {
        *pending = fqeb_pending_scatter{d_e, L.ldd, L.pitch, a0, nr, op->d_rowmap, op->zr, op->zi};
        break;
      }
      {
        PhaseTimer t(2, st);
        rc = launch_make_coeff(g, d_e, L.ldd, L.pitch, a0, nr, op->d_rowmap, op->zr, op->zi,
                               d_sigma, st);
      }
      if (rc != FQEB_OK) return rc;
    }
    return FQEB_OK;
  }
  g_last_path.store(FQEB_PATH_THREE_KERNEL);
  const int zrows = dvec_rows_zeroed(op, nij);
  if (zrows > nij) {
    // the k-padding rows of D that a partial k4 step reads must be exact zeros
    // (they meet finite operator entries)
    FQEB_CUDA(cudaMemsetAsync(d_dvec + 2 * (size_t)nij * L.ldd, 0,
                              sizeof(double) * 2 * (size_t)(zrows - nij) * L.ldd, st));
  }
  const int n_elec = g->nele[0] + g->nele[1];
  const bool absorb = op->absorb_ok && n_elec > 0;
  const double *d_A = nullptr;
  if (absorb) {
    rc = absorbed_operand(op, n_elec, &d_A);
    if (rc != FQEB_OK) return rc;
  }
  for (int64_t a0 = row0; a0 < row1; a0 += rows_chunk) {
    const int64_t nr = (row1 - a0) < rows_chunk ? (row1 - a0) : rows_chunk;
    {
      PhaseTimer t(0, st);
      rc = launch_gather(g, op->sym, d_coeff, d_dvec, L.ldd, a0, nr, ij0, ij1, st);
      if (rc == FQEB_OK && !absorb)
        rc = launch_make_dvec(g, d_coeff, nullptr, 0, a0, nr, ij0, ij1, op->d_pairs, op->np,
                              op->d_h1, d_sigma, st);
    }
    if (rc != FQEB_OK) return rc;
    {
      PhaseTimer t(1, st);
      rc = launch_contract(op, d_A, d_dvec, L.ldd, d_evec, L.ldd, nr * lenb, ij0, ij1, st);
    }
    if (rc != FQEB_OK) return rc;
    if (pending && a0 + rows_chunk >= row1) {
      *pending = fqeb_pending_scatter{d_evec, L.ldd, L.pitch, a0, nr, op->d_rowmap, op->zr, op->zi};
      break;
    }
    {
      PhaseTimer t(2, st);
      rc = launch_make_coeff(g, d_evec, L.ldd, L.pitch, a0, nr, op->d_rowmap, op->zr, op->zi,
                             d_sigma, st);
    }
    if (rc != FQEB_OK) return rc;
  }
  return FQEB_OK;
}

extern "C" int fqeb_sigma_restricted(const fqeb_graph *g, const fqeb_op *op,
                                     const double *d_coeff, double *d_sigma, void *d_workspace,
                                     size_t workspace_bytes, int64_t row0, int64_t row1, int ij0,
                                     int ij1, void *stream) {
  return sigma_build(g, op, d_coeff, d_sigma, d_workspace, workspace_bytes, row0, row1, ij0, ij1,
                     stream, nullptr);
}

extern "C" int fqeb_sigma_restricted_deferred(const fqeb_graph *g, const fqeb_op *op,
                                              const double *d_coeff, double *d_sigma,
                                              void *d_workspace, size_t workspace_bytes,
                                              int64_t row0, int64_t row1, int ij0, int ij1,
                                              fqeb_pending_scatter *pending, void *stream) {
  FQEB_REQUIRE(pending != nullptr, "sigma_deferred: NULL pending descriptor");
  return sigma_build(g, op, d_coeff, d_sigma, d_workspace, workspace_bytes, row0, row1, ij0, ij1,
                     stream, pending);
}

extern "C" int fqeb_scatter_rows(const fqeb_graph *g, const fqeb_pending_scatter *pending,
                                 int64_t x0, int64_t x1, double *d_sigma, void *stream) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(g && pending && d_sigma, "scatter_rows: NULL argument");
  if (pending->nrows == 0 || pending->d_evec == nullptr) return FQEB_OK;   // nothing was deferred
  PhaseTimer t(2, (cudaStream_t)stream);
  return launch_make_coeff(g, pending->d_evec, pending->lde, pending->pitch, pending->row0,
                           pending->nrows, pending->d_rowmap, pending->zr, pending->zi, d_sigma,
                           (cudaStream_t)stream, x0, x1);
}

// Taylor propagator, whole recurrence on the device (reference wavefunction.py:548-567):
//   evol = sum_k op^k / k! |C>,  op prepared from the tensors of -i*t*H (Hamiltonian.iht).
// The loop stops when ||op^k C|| / k! < accuracy; running into max_terms is an error, as the
// reference raises RuntimeError("maximum taylor expansion limit reached").
extern "C" int fqeb_taylor(const fqeb_graph *g, const fqeb_op *op, double *d_evol, double *d_work,
                           double *d_next, void *d_workspace, size_t workspace_bytes,
                           void *d_scratch, double accuracy, int max_terms, int *nterms,
                           void *stream) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(g && op && d_evol && d_work && d_next && d_scratch && nterms,
               "fqeb_taylor: NULL argument");
  FQEB_REQUIRE(d_evol != d_work && d_evol != d_next && d_work != d_next,
               "fqeb_taylor: evol / work / next buffers must be distinct");
  const int64_t n = g->len[0] * g->len[1];
  cudaStream_t st = (cudaStream_t)stream;
  FQEB_CUDA(cudaMemcpyAsync(d_work, d_evol, sizeof(double) * 2 * (size_t)n,
                            cudaMemcpyDeviceToDevice, st));
  // Error budget of the sliced contraction inside the series: a relative error delta_k of the
  // sigma build that produces term k enters the propagated state with the weight of that term
  // (and of the later ones grown from it), ~ delta_k ||term_k||.  Holding every build to the
  // same relative bound would send the late, tiny terms - whose vectors grow heavy tails and
  // fail the global-scale estimate - to the slower FP64 kernel for no gain in the result, so the
  // bound is scaled by ||C|| / (4 ||term_{k-1}||), at least 1, at most 1e4: the absolute error
  // added per term stays below a quarter of what the first term may add.
  struct TolScaleGuard {
    ~TolScaleGuard() { g_ozaki_tol_scale = 1.0; }
  } tol_guard;
  double norm0_sq = 0.0;
  rc = fqeb_znorm2(n, d_evol, d_scratch, &norm0_sq, stream);
  if (rc != FQEB_OK) return rc;
  const double norm0 = sqrt(norm0_sq);
  double factorial = 1.0;
  for (int order = 1; order < max_terms; ++order) {
    rc = fqeb_sigma_restricted(g, op, d_work, d_next, d_workspace, workspace_bytes, 0, g->len[0],
                               0, op->np, stream);
    if (rc != FQEB_OK) return rc;
    double *t = d_work;
    d_work = d_next;
    d_next = t;
    factorial *= (double)order;
    const double coeff = 1.0 / factorial;
    double norm2 = 0.0;
    rc = fqeb_axpy_norm2(n, coeff, 0.0, d_work, d_evol, d_scratch, &norm2, stream);
    if (rc != FQEB_OK) return rc;
    const double term_norm = sqrt(norm2) * coeff;
    if (term_norm < accuracy) {
      *nterms = order;
      return FQEB_OK;
    }
    double scale = term_norm > 0.0 ? 0.25 * norm0 / term_norm : 1.0;
    g_ozaki_tol_scale = scale < 1.0 ? 1.0 : (scale > 1.0e4 ? 1.0e4 : scale);
  }
  *nterms = max_terms;
  set_error("maximum taylor expansion limit reached (%d terms)", max_terms);
  return FQEB_ERR_CONVERGE;
}

extern "C" int fqeb_sigma_last_path(void) { return g_last_path.load(); }

extern "C" int fqeb_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(g_profile_mu);
  g_profile = on != 0;
  return FQEB_OK;
}

extern "C" int fqeb_profile_collect(double *h_ms, int64_t *h_launches) {
  FQEB_REQUIRE(h_ms && h_launches, "fqeb_profile_collect: NULL argument");
  for (int k = 0; k < 3; ++k) {
    h_ms[k] = 0.0;
    h_launches[k] = 0;
  }
  std::lock_guard<std::mutex> lock(g_profile_mu);
  for (auto &ev : g_phase_events) {
    float ms = 0.f;
    if (cudaEventSynchronize(ev.stop) == cudaSuccess &&
        cudaEventElapsedTime(&ms, ev.start, ev.stop) == cudaSuccess) {
      h_ms[ev.phase] += ms;
      h_launches[ev.phase] += 1;
    }
    cudaEventDestroy(ev.start);
    cudaEventDestroy(ev.stop);
  }
  g_phase_events.clear();
  return FQEB_OK;
}

// ---- host-buffer entry: the reference-facing call -----------------------------------
namespace {
std::mutex g_cache_mu;
std::map<std::tuple<int, int, int, int>, fqeb_graph *> g_graph_cache;

fqeb_graph *cached_graph(int norb, int na, int nb, int *rc) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_cache_mu);
  auto key = std::make_tuple(norb, na, nb, dev);
  auto it = g_graph_cache.find(key);
  if (it != g_graph_cache.end()) {
    *rc = FQEB_OK;
    return it->second;
  }
  fqeb_graph *g = nullptr;
  *rc = fqeb_graph_create(norb, na, nb, &g);
  if (*rc == FQEB_OK) g_graph_cache[key] = g;
  return g;
}
}  // namespace

namespace {
// Per-device context of the host-buffer entry point: device copies of C and sigma, the D/E
// workspace and a pinned double buffer for the transfers are allocated once and reused, so
// that a caller looping over fqeb_sigma_restricted_host (a Taylor series driven from the
// reference's Python) pays for allocation and page pinning once.
struct HostCtx {
  double *d_c = nullptr, *d_s = nullptr;
  size_t cbytes = 0;
  void *d_ws = nullptr;
  size_t ws_bytes = 0;
  char *h_stage[2] = {nullptr, nullptr};
  size_t stage_bytes = 0;
  cudaStream_t st = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
};
std::map<int, HostCtx> g_host_ctx;
std::mutex g_host_mu;   // the entry point is serialised per process (one workspace per device)

// Host-side copies between the caller's pageable arrays and the pinned staging buffers are what
// bounds this entry point (2 x 2.65 GB per build at norb = 16): a small persistent pool of copy
// threads, created on first use, splits every block.
class CopyPool {
 public:
  explicit CopyPool(unsigned n) : nthreads_(n) {
    for (unsigned t = 0; t < n; ++t) workers_.emplace_back([this, t]() { run(t); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lock(mu_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto &w : workers_) w.join();
  }
  void copy(char *dst, const char *src, size_t n) {
    std::unique_lock<std::mutex> lock(mu_);
    dst_ = dst;
    src_ = src;
    n_ = n;
    pending_ = nthreads_;
    ++epoch_;
    cv_.notify_all();
    done_.wait(lock, [this]() { return pending_ == 0; });
  }
  unsigned size() const { return nthreads_; }

 private:
  void run(unsigned t) {
    uint64_t seen = 0;
    for (;;) {
      char *dst;
      const char *src;
      size_t n;
      {
        std::unique_lock<std::mutex> lock(mu_);
        cv_.wait(lock, [&]() { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
        dst = dst_;
        src = src_;
        n = n_;
      }
      const size_t per = ((n + nthreads_ - 1) / nthreads_ + 4095) & ~(size_t)4095;
      const size_t lo = (size_t)t * per, hi = lo + per < n ? lo + per : n;
      if (lo < hi) memcpy(dst + lo, src + lo, hi - lo);
      {
        std::lock_guard<std::mutex> lock(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  unsigned nthreads_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  char *dst_ = nullptr;
  const char *src_ = nullptr;
  size_t n_ = 0;
  unsigned pending_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

void parallel_memcpy(char *dst, const char *src, size_t n) {
  if (n < (8u << 20)) {
    memcpy(dst, src, n);
    return;
  }
  // FQEB_COPY_THREADS overrides; default 4 (the copies are bound by host memory bandwidth: more
  // threads were not faster on the boxes measured, scripts/cabi_host_time.py)
  static CopyPool *pool = []() {
    unsigned nt = 4;
    if (const char *env = getenv("FQEB_COPY_THREADS")) nt = (unsigned)atoi(env);
    if (nt < 1) nt = 1;
    if (nt > 16) nt = 16;
    return new CopyPool(nt);   // lives until process exit (workers are blocked on the condition)
  }();
  pool->copy(dst, src, n);
}

// pageable host -> device through the pinned double buffer: the host-side copy of block k+1
// overlaps the DMA of block k
int staged_upload(HostCtx &c, void *d_dst, const void *h_src, size_t bytes) {
  bool used[2] = {false, false};
  int k = 0;
  for (size_t off = 0; off < bytes; off += c.stage_bytes, ++k) {
    const int slot = k & 1;
    const size_t n = bytes - off < c.stage_bytes ? bytes - off : c.stage_bytes;
    if (used[slot]) FQEB_CUDA(cudaEventSynchronize(c.ev[slot]));
    parallel_memcpy(c.h_stage[slot], (const char *)h_src + off, n);
    FQEB_CUDA(cudaMemcpyAsync((char *)d_dst + off, c.h_stage[slot], n, cudaMemcpyHostToDevice, c.st));
    FQEB_CUDA(cudaEventRecord(c.ev[slot], c.st));
    used[slot] = true;
  }
  return FQEB_OK;
}

int staged_download(HostCtx &c, void *h_dst, const void *d_src, size_t bytes) {
  size_t pend_off[2] = {0, 0}, pend_n[2] = {0, 0};
  bool used[2] = {false, false};
  int k = 0;
  for (size_t off = 0; off < bytes; off += c.stage_bytes, ++k) {
    const int slot = k & 1;
    const size_t n = bytes - off < c.stage_bytes ? bytes - off : c.stage_bytes;
    if (used[slot]) {   // drain the block that still sits in this slot
      FQEB_CUDA(cudaEventSynchronize(c.ev[slot]));
      parallel_memcpy((char *)h_dst + pend_off[slot], c.h_stage[slot], pend_n[slot]);
    }
    FQEB_CUDA(cudaMemcpyAsync(c.h_stage[slot], (const char *)d_src + off, n, cudaMemcpyDeviceToHost, c.st));
    FQEB_CUDA(cudaEventRecord(c.ev[slot], c.st));
    pend_off[slot] = off;
    pend_n[slot] = n;
    used[slot] = true;
  }
  for (int j = 0; j < 2; ++j) {   // oldest first
    const int slot = (k + j) & 1;
    if (used[slot]) {
      FQEB_CUDA(cudaEventSynchronize(c.ev[slot]));
      parallel_memcpy((char *)h_dst + pend_off[slot], c.h_stage[slot], pend_n[slot]);
      used[slot] = false;
    }
  }
  return FQEB_OK;
}

int host_ctx_prepare(HostCtx &c, const fqeb_graph *g, const fqeb_op *op) {
  const size_t cbytes = sizeof(double) * 2 * (size_t)g->len[0] * g->len[1];
  if (!c.st) {
    FQEB_CUDA(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
    for (int j = 0; j < 2; ++j) FQEB_CUDA(cudaEventCreateWithFlags(&c.ev[j], cudaEventDisableTiming));
    c.stage_bytes = 64u << 20;
    for (int j = 0; j < 2; ++j) FQEB_CUDA(cudaMallocHost((void **)&c.h_stage[j], c.stage_bytes));
  }
  if (cbytes > c.cbytes) {
    if (c.d_c) cudaFree(c.d_c);
    if (c.d_s) cudaFree(c.d_s);
    c.d_c = c.d_s = nullptr;
    c.cbytes = 0;
    if (cudaMalloc(&c.d_c, cbytes) != cudaSuccess || cudaMalloc(&c.d_s, cbytes) != cudaSuccess) {
      set_error("sigma_host: cannot allocate coefficient buffers (%zu bytes each)", cbytes);
      return FQEB_ERR_NOMEM;
    }
    c.cbytes = cbytes;
  }
  if (op->has_h2) {
    const size_t want = fqeb_sigma_workspace_bytes(g, op, g->len[0], 0, op->np);
    if (want > c.ws_bytes) {
      if (c.d_ws) cudaFree(c.d_ws);
      c.d_ws = nullptr;
      c.ws_bytes = 0;
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      const size_t budget = (size_t)(0.85 * (double)free_b);
      const size_t take = want < budget ? want : budget;
      if (cudaMalloc(&c.d_ws, take) != cudaSuccess) {
        set_error("sigma_host: cannot allocate %zu-byte workspace", take);
        return FQEB_ERR_NOMEM;
      }
      c.ws_bytes = take;
    }
  }
  return FQEB_OK;
}
}  // namespace

// release everything fqeb_sigma_restricted_host keeps between calls (all devices)
extern "C" int fqeb_host_release(void) {
  std::lock_guard<std::mutex> lock(g_host_mu);
  for (auto &kv : g_host_ctx) {
    HostCtx &c = kv.second;
    cudaSetDevice(kv.first);
    if (c.st) cudaStreamSynchronize(c.st);
    if (c.d_c) cudaFree(c.d_c);
    if (c.d_s) cudaFree(c.d_s);
    if (c.d_ws) cudaFree(c.d_ws);
    for (int j = 0; j < 2; ++j) {
      if (c.h_stage[j]) cudaFreeHost(c.h_stage[j]);
      if (c.ev[j]) cudaEventDestroy(c.ev[j]);
    }
    if (c.st) cudaStreamDestroy(c.st);
  }
  g_host_ctx.clear();
  return FQEB_OK;
}

extern "C" int fqeb_sigma_restricted_host(int norb, int nalpha, int nbeta, const double *h_h1p,
                                          const double *h_h2p, const double *h_coeff,
                                          double *h_sigma) {
  int rc = require_device();
  if (rc != FQEB_OK) return rc;
  FQEB_REQUIRE(h_h1p && h_coeff && h_sigma, "sigma_host: NULL argument");
  fqeb_graph *g = cached_graph(norb, nalpha, nbeta, &rc);
  if (rc != FQEB_OK) return rc;
  fqeb_op *op = nullptr;
  rc = fqeb_op_create(norb, h_h1p, h_h2p, &op);
  if (rc != FQEB_OK) return rc;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_host_mu);
  HostCtx &c = g_host_ctx[dev];
  rc = host_ctx_prepare(c, g, op);
  const size_t cbytes = sizeof(double) * 2 * (size_t)g->len[0] * g->len[1];
  if (rc == FQEB_OK) rc = staged_upload(c, c.d_c, h_coeff, cbytes);
  if (rc == FQEB_OK)
    rc = fqeb_sigma_restricted(g, op, c.d_c, c.d_s, c.d_ws, c.ws_bytes, 0, g->len[0], 0, op->np,
                               c.st);
  if (rc == FQEB_OK) rc = staged_download(c, h_sigma, c.d_s, cbytes);
  if (rc == FQEB_OK && cudaStreamSynchronize(c.st) != cudaSuccess) {
    set_error("sigma_host: %s", cudaGetErrorString(cudaGetLastError()));
    rc = FQEB_ERR_CUDA;
  }
  fqeb_op_destroy_async(op, c.st);
  return rc;
}
