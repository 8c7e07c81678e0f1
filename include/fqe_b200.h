/*
 * fqe_b200.h -- C ABI of libfqe_b200.so: the B200 (sm_100a) implementation of
 * OpenFermion-FQE's Hamiltonian-application hot path.
 *
 * This is the drop-in boundary.  The reference reaches its native kernels
 * through ctypes (src/fqe/lib/__init__.py:4-14 loads libfqe.so; per-call
 * wrappers live in src/fqe/lib/fci_graph.py and src/fqe/lib/_fqe_data.pyx).
 * Each entry point below names the reference interface it replaces.  Paths are
 * relative to /root/reference/src/fqe.
 *
 * Conventions
 *   - plain C: pointers, sizes, opaque handles; no C++/torch types.
 *   - every function returns an int status (FQEB_OK == 0); nothing ever calls
 *     exit() (the reference aborts the interpreter on OOM, lib/macros.c:25-48).
 *     fqeb_last_error() returns a thread-local description of the last failure.
 *   - pointers named d_* are DEVICE pointers on the handle's device (the Python
 *     host passes torch.Tensor.data_ptr()); pointers named h_* are HOST memory.
 *   - complex numbers are interleaved (re, im) doubles, i.e. numpy complex128 /
 *     C99 `double complex` / torch.complex128.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Device
 *     entry points are asynchronous with respect to the host unless noted.
 *   - the caller owns every buffer, exactly as in the reference
 *     (lib/fqe_data.c:104-108); workspaces are sized by the *_workspace_bytes
 *     queries and allocated by the caller (torch) so that memory stays under
 *     one allocator.
 */
#ifndef FQE_B200_H_
#define FQE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  FQEB_OK = 0,
  FQEB_ERR_INVALID = 1,   /* bad argument / inconsistent sizes          */
  FQEB_ERR_CUDA = 2,      /* a CUDA runtime call failed                 */
  FQEB_ERR_NOMEM = 3,     /* workspace too small / allocation failed    */
  FQEB_ERR_NODEVICE = 4,  /* no CUDA device: there is NO CPU fallback   */
  FQEB_ERR_CONVERGE = 5   /* polynomial expansion limit reached         */
};

enum { FQEB_ALPHA = 0, FQEB_BETA = 1 };

/* operator coefficient classes detected by fqeb_op_create */
enum {
  FQEB_OP_REAL = 0,    /* h1', h2' purely real           -> real x complex GEMM  */
  FQEB_OP_IMAG = 1,    /* purely imaginary (-i t H, Taylor) -> same, times i      */
  FQEB_OP_COMPLEX = 2  /* general complex                -> 4-real-GEMM complex   */
};

typedef struct fqeb_graph fqeb_graph; /* FciGraph: strings + excitation tables   */
typedef struct fqeb_op fqeb_op;       /* prepared (h1', h2') dense operator      */

const char *fqeb_last_error(void);
int fqeb_version(void);
/* number of CUDA devices visible; 0 means every compute entry point fails with
 * FQEB_ERR_NODEVICE. */
int fqeb_device_count(void);
/* make `device` current for this thread inside the library's CUDA runtime (the
 * Python host calls it with torch.cuda.current_device()). */
int fqeb_set_device(int device);
/* kernels launched by this library since load (bench.py's "gpu_launches"). */
uint64_t fqeb_launch_count(void);

/* ------------------------------------------------------------------------
 * a1-a3  FciGraph: string tables, Z matrix, E_ij maps        (integer, bit-exact)
 * replaces: calculate_Z_matrix, lexicographic_bitstring_generator,
 *           calculate_string_address, build_mapping_strings, map_deexc
 *           (lib/fci_graph.h:22-47, lib/fci_graph.c:27-147, lib/bitstring.c:29-46)
 *           as driven by FciGraph.__init__ (fci_graph.py:108-154).
 * All tables are built ON THE DEVICE, once per (norb, nalpha, nbeta), and stay
 * resident.  The exports copy them back for bit-exact comparison and for the
 * host-side FciGraph mirror (alpha_map/beta_map/_dexca/_dexcb).
 * ------------------------------------------------------------------------ */
int fqeb_graph_create(int norb, int nalpha, int nbeta, fqeb_graph **out);
int fqeb_graph_destroy(fqeb_graph *g);
int fqeb_graph_dims(const fqeb_graph *g, int *norb, int *nalpha, int *nbeta,
                    int64_t *lena, int64_t *lenb);
/* h_out: int32[nele * norb] (row k = electron index), as _get_Z_matrix. */
int fqeb_graph_get_Z(const fqeb_graph *g, int spin, int32_t *h_out);
/* h_out: uint64[len]: string stored at each Knowles-Handy address. */
int fqeb_graph_get_strings(const fqeb_graph *g, int spin, uint64_t *h_out);
/* h_out: int32[norb*norb * len].  Entry [(i*norb+j)*len + s] describes
 * a^+_i a_j |s> = sign |t>:  sign*(t+1), or 0 when the string is annihilated.
 * This is the dense form of FciGraph._alpha_map/_beta_map (fci_graph.py:204-237). */
int fqeb_graph_get_map(const fqeb_graph *g, int spin, int32_t *h_out);
/* device pointers for callers that launch their own kernels (may be NULL-checked) */
int fqeb_graph_device_tables(const fqeb_graph *g, int spin,
                             const uint64_t **d_strings,
                             const int32_t **d_map_by_pair,   /* [n*n][len] adjoint map */
                             const int32_t **d_map_by_string  /* [len][n*n] adjoint map */);

/* ------------------------------------------------------------------------
 * a5/a8/a16  dense restricted operator
 * replaces the tensor preparation inside FqeData._apply_array_spatial12_lm /
 * _halffilling (fqe_data.py:647-652, 691-693): the caller passes
 *   h1p[i,j]      = h1[i,j] - sum_k h2p[i,k,k,j]
 *   h2p[i,j,k,l]  = -h2[i,k,j,l]          (i.e. -moveaxis(h2,1,2))
 * as HOST complex128 arrays; the library classifies them (real / imaginary /
 * complex), expands h2p to the real GEMM operand and uploads both.
 * h2p may be NULL for a one-body operator (FqeData._apply_array_spatial1,
 * fqe_data.py:477-530).
 * ------------------------------------------------------------------------ */
int fqeb_op_create(int norb, const double *h_h1p, const double *h_h2p,
                   fqeb_op **out);
/* Same with flags: FQEB_OP_FLAG_FULL_PAIR_SPACE keeps the norb^2 pair space even when the
 * tensor is pair-symmetric (for callers that hand-build D / E in the reference layout, e.g.
 * the 3-body apply, fqe_data.py:1199-1208). */
enum { FQEB_OP_FLAG_FULL_PAIR_SPACE = 1 };
int fqeb_op_create_ex(int norb, const double *h_h1p, const double *h_h2p, int flags,
                      fqeb_op **out);
int fqeb_op_destroy(fqeb_op *op);
/* Non-blocking destroy: device buffers are released in stream order on `stream`, i.e. after the
 * kernels already enqueued there that may still read the operator; fqeb_op_destroy waits for
 * the whole device instead.  Creating an operator never waits for running kernels either
 * (uploads use an internal stream). */
int fqeb_op_destroy_async(fqeb_op *op, void *stream);
int fqeb_op_kind(const fqeb_op *op, int *kind);
/* Size of the pair space the contraction runs over and whether it is compressed:
 * norb^2, or norb(norb+1)/2 when h2p[ij,kl] == h2p[ji,kl] == h2p[ij,lk] exactly
 * (real-orbital integrals; the compressed route of fqe_data.py:659-681).  Every
 * [ij0, ij1) pair range of the entry points below that take an `op` refers to this
 * space.  Set FQEB_NO_SYMMETRY=1 in the environment to disable the compression. */
int fqeb_op_pair_space(const fqeb_op *op, int *npairs, int *symmetric);

/* ------------------------------------------------------------------------
 * a7  D[i,j,a,b] = sum_I <J|a^+_i a_j|I> C_I     (gather)
 * replaces zdvec_make (lib/fqe_data.h:83-91, lib/fqe_data.c:350-377) as called by
 * FqeData._calculate_dvec_spatial_with_coeff (fqe_data.py:2209-2224).
 * Builds rows [row0, row0+nrows) of the alpha index for pairs [ij0, ij1):
 *   d_dvec[(ij-ij0)*ldd + (a-row0)*lenb + b]      complex128, ldd >= nrows*lenb
 * With row0=0, nrows=lena, ij0=0, ij1=norb^2, ldd=lena*lenb this is exactly the
 * reference's dvec[norb,norb,lena,lenb].
 * ------------------------------------------------------------------------ */
int fqeb_make_dvec(const fqeb_graph *g, const double *d_coeff, double *d_dvec,
                   int64_t ldd, int64_t row0, int64_t nrows, int ij0, int ij1,
                   void *stream);

/* a9  out[a,b] += sum_ij <I|a^+_i a_j|J> E_ij^J       (scatter, by-target gather)
 * replaces zcoeff_make (lib/fqe_data.h:93-101, lib/fqe_data.c:379-406) as called
 * by FqeData._calculate_coeff_spatial_with_dvec (fqe_data.py:2309-2334).
 * d_evec holds alpha rows [row0,row0+nrows) for ALL norb^2 pairs, leading
 * dimension lde (complex elements).  Accumulates zr+i*zi times the result into
 * d_out (caller zero-fills, as in the reference). */
int fqeb_make_coeff(const fqeb_graph *g, const double *d_evec, int64_t lde,
                    int64_t row0, int64_t nrows, double zr, double zi,
                    double *d_out, void *stream);

/* a8  E[kl, det] = sum_ij h2p[kl, ij] D[ij, det]     (FP64 tensor-core GEMM)
 * replaces numpy.einsum("ijkl,klmn->ijmn", h2e, dvec) (fqe_data.py:656).
 * D holds pairs [ij0, ij1) (the K slice); E gets all norb^2 rows.  ncols is the
 * number of determinants in the chunk; ldd/lde are leading dimensions in complex
 * elements and must be multiples of fqeb_gemm_col_align(). */
int fqeb_contract(const fqeb_op *op, const double *d_dvec, int64_t ldd,
                  double *d_evec, int64_t lde, int64_t ncols, int ij0, int ij1,
                  void *stream);
int fqeb_gemm_col_align(void);
/* rows the D buffer must have for a slice of nij pairs (k-padding; the rows beyond
 * nij must be zero-filled by the caller). */
int fqeb_contract_dvec_rows(const fqeb_op *op, int nij);

/* f4  G[m, n] += sum_c conj(bra[m, c]) * ket[n, c]     (reduction over determinants, FP64 DMMA)
 * replaces the contractions of FqeData.rdm12 / rdm1 (fqe_data.py:1726-1838, 1668-1724:
 * numpy.tensordot / einsum of dvec.conj() with dvec over the determinant indices), where bra and
 * ket are D = E_ij C blocks from fqeb_make_dvec.  d_bra: [M][ldb], d_ket: [N][ldk] complex128,
 * d_G: row-major [M][N] complex128, accumulated into (caller zero-fills).  When d_ket_last is not
 * NULL, row N-1 of ket is taken from there (ncols contiguous elements): passing the coefficient
 * block itself yields <D[ij] | C>, the one-particle part, in the same pass.  Split over column
 * slabs with a fixed-order second pass: bitwise reproducible, no atomics. */
int fqeb_gram_accumulate(int M, int N, int64_t ncols, const double *d_bra, int64_t ldb,
                         const double *d_ket, int64_t ldk, const double *d_ket_last,
                         double *d_G, void *stream);

/* ------------------------------------------------------------------------
 * a5/a6  sigma = (h1', h2') applied to C        FqeData.apply_inplace((h1,h2))
 * replaces FqeData._apply_array_spatial12 (fqe_data.py:582-608, 644-710) and the
 * three C kernels under it (lm_apply_array12_same_spin_opt x2,
 * lm_apply_array12_diff_spin_opt; lib/fqe_data.h:164-186).
 * d_sigma is OVERWRITTEN with the result (it must not alias d_coeff).
 * The determinant index is processed in chunks of alpha rows sized to the
 * workspace; [row0,row1) x [ij0,ij1) select this rank's shard (full range for a
 * single GPU): the result is then a PARTIAL sigma to be summed over ranks.
 * ------------------------------------------------------------------------ */
size_t fqeb_sigma_workspace_bytes(const fqeb_graph *g, const fqeb_op *op,
                                  int64_t rows_per_chunk, int ij0, int ij1);
int64_t fqeb_sigma_rows_for_workspace(const fqeb_graph *g, const fqeb_op *op,
                                      size_t bytes, int ij0, int ij1);
int fqeb_sigma_restricted(const fqeb_graph *g, const fqeb_op *op,
                          const double *d_coeff, double *d_sigma,
                          void *d_workspace, size_t workspace_bytes,
                          int64_t row0, int64_t row1, int ij0, int ij1,
                          void *stream);

/* The same build with the scatter of the LAST chunk of alpha rows left to the caller, for
 * multi-GPU runs: fqeb_scatter_rows then completes sigma by slices of TARGET rows [x0, x1), so
 * that the all-reduce of a finished slice (NCCL, another stream) overlaps the scatter of the
 * next one.  `pending` describes the deferred launch (the E chunk lives in the workspace, which
 * must stay untouched until the last slice has been issued); nrows == 0 means nothing was
 * deferred (one-body operator, empty shard) and fqeb_scatter_rows is a no-op. */
typedef struct fqeb_pending_scatter {
  const double *d_evec;      /* E chunk, rows = pair space, in the workspace           */
  int64_t lde, pitch;        /* complex elements per E row / per alpha row inside it   */
  int64_t row0, nrows;       /* alpha rows the chunk covers                            */
  const int32_t *d_rowmap;   /* E row of pair kl                                       */
  double zr, zi;             /* global factor applied by the scatter                   */
} fqeb_pending_scatter;
int fqeb_sigma_restricted_deferred(const fqeb_graph *g, const fqeb_op *op,
                                   const double *d_coeff, double *d_sigma, void *d_workspace,
                                   size_t workspace_bytes, int64_t row0, int64_t row1, int ij0,
                                   int ij1, fqeb_pending_scatter *pending, void *stream);
int fqeb_scatter_rows(const fqeb_graph *g, const fqeb_pending_scatter *pending, int64_t x0,
                      int64_t x1, double *d_sigma, void *stream);
/* Same, HOST buffers in and out (the reference-facing call: numpy coeff in,
 * numpy sigma out; what a maintainer binds at src/fqe/fqe_data.py:685 in place of the three
 * lm_apply_array12_* calls).  Pageable host memory is fine: transfers are staged through a
 * pinned double buffer (host copy of block k+1 overlaps the DMA of block k).  Device copies of
 * C and sigma, the workspace and the staging buffers are kept per device between calls
 * (fqeb_host_release frees them); calls are serialised per process.  Synchronous.           */
int fqeb_sigma_restricted_host(int norb, int nalpha, int nbeta,
                               const double *h_h1p, const double *h_h2p,
                               const double *h_coeff, double *h_sigma);
int fqeb_host_release(void);

/* Which contraction path the most recent two-body fqeb_sigma_restricted call took (a test /
 * benchmark aid; the result is the same sigma to the tolerance of the path):
 *   FQEB_PATH_THREE_KERNEL  gather -> FP64 DMMA GEMM -> scatter, D and E through HBM
 *   FQEB_PATH_FUSED         gather fused into the FP64 DMMA contraction (D stays on chip)
 *   FQEB_PATH_SLICED        gather fused into the INT8-sliced tcgen05 contraction (csrc/ozaki.cu):
 *                           default for real / imaginary operators with a pair space <= 144 when
 *                           the state's estimated quantisation error is below FQEB_OZAKI_TOL
 *                           (environment, default 5e-12); FQEB_OZAKI=0 disables it.            */
enum { FQEB_PATH_NONE = 0, FQEB_PATH_THREE_KERNEL = 1, FQEB_PATH_FUSED = 2, FQEB_PATH_SLICED = 3 };
int fqeb_sigma_last_path(void);
/* Dense INT8 tensor-core throughput of the current device in tera-operations per second
 * (tcgen05.mma kind::i8 on every SM, CUDA-event timed, synchronous): the roofline denominator of
 * the sliced contraction, measured rather than assumed.                                        */
int fqeb_i8_tensor_peak(double *tops);
/* Diagnostic cycle counters of the sliced contraction kernel (enabled by FQEB_OZAKI_PROF=1 in the
 * environment; see csrc/ozaki.cu): h_out[8], summed over CTAs since the previous call.          */
int fqeb_ozaki_profile(uint64_t *h_out);

/* Per-kernel device timing of the sigma build (bench.py's roofline leg).  When
 * enabled, fqeb_sigma_restricted brackets every gather / contraction / scatter
 * launch with CUDA events on the launching stream.  fqeb_profile_collect
 * synchronises, returns the summed milliseconds and launch counts per phase
 * (index 0 gather, 1 contraction, 2 scatter) since the last collect, and resets. */
int fqeb_profile_enable(int on);
int fqeb_profile_collect(double *h_ms /* [3] */, int64_t *h_launches /* [3] */);

/* ------------------------------------------------------------------------
 * a10/a11  diagonal Coulomb
 * replaces zdiagonal_coulomb_apply / zdiagonal_coulomb
 * (lib/fqe_data.h:103-123, lib/fqe_data.c:455-602) called from
 * FqeData.apply_diagonal_coulomb / evolve_diagonal_coulomb (fqe_data.py:263-402).
 * h_diag: complex128[norb], h_array: complex128[norb,norb] (host); d_coeff is
 * updated in place.  The apply / evolve alpha-beta conventions differ exactly as
 * in the reference.
 * ------------------------------------------------------------------------ */
int fqeb_dc_apply(const fqeb_graph *g, const double *h_diag,
                  const double *h_array, double *d_coeff, void *stream);
int fqeb_dc_evolve(const fqeb_graph *g, const double *h_diag,
                   const double *h_array, double *d_coeff, void *stream);

/* ------------------------------------------------------------------------
 * 8f-1  orbital rotation by column operators, one-body diagonal apply / evolve
 * fqeb_apply_columns replaces lm_apply_array1_column_alpha
 * (lib/fqe_data.h:71-82, lib/fqe_data.c:305-347) as looped over icol by
 * FqeData._apply_columns_recursive_alpha (fqe_data.py:1476-1504): for
 * icol = 0..norb-1 in order, C <- (1 + sum_i mat[i,icol] a^+_i a_icol) C on one
 * spin (0 = alpha rows, 1 = beta columns; the reference transposes C for beta,
 * fqe_data.py:1506-1535), in place.  h_mat: complex128[norb,norb] row-major
 * (host), the `process_matrix` output of Wavefunction.transform
 * (wavefunction.py:889-909).
 * fqeb_apply_diagonal / fqeb_evolve_diagonal replace apply_diagonal_inplace /
 * evolve_diagonal_inplace (lib/fqe_data.c:1320-1383; fqe_data.py:153-261):
 * C[a,b] *= A[a] + B[b]  resp.  C[a,b] *= exp(A[a]) * exp(B[b]) with
 * A[a] = sum_{i in a} aarray[i], B[b] = sum_{i in b} barray[i];
 * h_aarray, h_barray: complex128[norb] (host).
 * ------------------------------------------------------------------------ */
int fqeb_apply_columns(const fqeb_graph *g, int spin, const double *h_mat,
                       double *d_coeff, void *stream);
int fqeb_apply_diagonal(const fqeb_graph *g, const double *h_aarray,
                        const double *h_barray, double *d_coeff, void *stream);
int fqeb_evolve_diagonal(const fqeb_graph *g, const double *h_aarray,
                         const double *h_barray, double *d_coeff, void *stream);

/* ------------------------------------------------------------------------
 * 8f-2  individual n-body operators
 * fqeb_nbody_accumulate replaces make_mapping_each (lib/fci_graph.h:64-70,
 * lib/fci_graph.c:223-264) + apply_individual_nbody1_accumulate
 * (lib/fqe_data.h:150-158, lib/fqe_data.c:1157-1182) as called by
 * FqeData.apply_individual_nbody_accumulate (fqe_data.py:1590-1653):
 *   out[ta,tb] += z * pa * pb * in[sa,sb]
 * for the spin-conserving operator  prod a+_{daga} prod a_{undaga} (alpha, na
 * of each) x prod a+_{dagb} prod a_{undagb} (beta, nb of each); operators act
 * right to left in list order.  d_in and d_out must differ.
 * fqeb_sparse_scale replaces evaluate_map_each + sparse_scale
 * (lib/fqe_data.c:1417-1440, 1265-1278; FqeData.apply_cos_inplace and
 * evolve_inplace_individual_nbody_trivial, fqe_data.py:2385-2433, 2497-2580):
 * C[a,b] *= f for the alpha strings with every orbital of bit mask a_occ occupied
 * and every orbital of a_emp empty, and likewise for beta.
 * ------------------------------------------------------------------------ */
int fqeb_nbody_accumulate(const fqeb_graph *g, double zr, double zi,
                          const int *daga, const int *undaga, int na,
                          const int *dagb, const int *undagb, int nb,
                          const double *d_in, double *d_out, void *stream);
int fqeb_sparse_scale(const fqeb_graph *g, uint64_t a_occ, uint64_t a_emp,
                      uint64_t b_occ, uint64_t b_emp, double fr, double fi,
                      double *d_coeff, void *stream);

/* ------------------------------------------------------------------------
 * a15  BLAS-1 on coefficient vectors (n complex elements)
 * replaces FqeData.ax_plus_y / scale / norm and util.vdot
 * (fqe_data.py:2620-2632, 2745-2751, 2701-2707; util.py:506-530).
 * Reductions are deterministic (two-pass); h_out is written after the stream is
 * synchronised. d_scratch: at least fqeb_reduce_scratch_bytes() bytes.
 * ------------------------------------------------------------------------ */
size_t fqeb_reduce_scratch_bytes(void);
int fqeb_zaxpy(int64_t n, double ar, double ai, const double *d_x, double *d_y,
               void *stream);
int fqeb_zscal(int64_t n, double ar, double ai, double *d_x, void *stream);
int fqeb_zaxpby(int64_t n, double ar, double ai, const double *d_x, double br,
                double bi, double *d_y, void *stream); /* y = a x + b y */
int fqeb_znorm2(int64_t n, const double *d_x, void *d_scratch, double *h_out,
                void *stream);
int fqeb_zdotc(int64_t n, const double *d_x, const double *d_y, void *d_scratch,
               double *h_out /* [2] */, void *stream);
/* evol += c*work and ||work||^2 in one pass (the Taylor inner loop,
 * wavefunction.py:563-566). */
int fqeb_axpy_norm2(int64_t n, double cr, double ci, const double *d_work,
                    double *d_evol, void *d_scratch, double *h_out,
                    void *stream);

/* ------------------------------------------------------------------------
 * a13  Taylor propagator, whole recurrence in one call
 * replaces the loop of Wavefunction.apply_generated_unitary(algo='taylor')
 * (wavefunction.py:548-567) for one sector: on entry d_evol holds C, on exit
 * sum_k op^k C / k!, with `op` created from the tensors of -i*t*H
 * (Hamiltonian.iht).  d_work, d_next: scratch of the size of C; d_workspace as
 * for fqeb_sigma_restricted; d_scratch: fqeb_reduce_scratch_bytes().  Stops when
 * ||op^k C|| / k! < accuracy and stores k in *nterms; FQEB_ERR_CONVERGE when
 * max_terms is reached (the reference raises RuntimeError there).  The e_0
 * phase is the caller's (wavefunction.py:600-601).
 * ------------------------------------------------------------------------ */
int fqeb_taylor(const fqeb_graph *g, const fqeb_op *op, double *d_evol,
                double *d_work, double *d_next, void *d_workspace,
                size_t workspace_bytes, void *d_scratch, double accuracy,
                int max_terms, int *nterms, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FQE_B200_H_ */
