/* Plain-C client of libfqe_b200.so: the drop-in boundary has no Python, torch or C++ types.
 *
 *   gcc -std=c99 -Iinclude examples/c_client.c -o c_client \
 *       -Lopenfermion-fqe_b200/fqe_b200/lib -lfqe_b200 \
 *       -Wl,-rpath,$PWD/openfermion-fqe_b200/fqe_b200/lib -lm
 *
 * Builds sigma = H C for a tiny RestrictedHamiltonian (norb = 4, 2 alpha + 2 beta electrons)
 * through the host-buffer entry point, the call a ctypes / cgo / JNI binder of the reference
 * would make in place of lm_apply_array12_*_opt (reference src/fqe/lib/fqe_data.h:83-101).
 * Without a CUDA device the library reports FQEB_ERR_NODEVICE (it has no CPU fallback) and
 * this program says so and exits 0. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "fqe_b200.h"

int main(void) {
  const int norb = 4, nalpha = 2, nbeta = 2;
  const int npair = norb * norb;
  const int ndet = 6 * 6; /* C(4,2)^2 */
  double *h1p = calloc(2 * npair, sizeof(double));
  double *h2p = calloc(2 * (size_t)npair * npair, sizeof(double));
  double *c = calloc(2 * ndet, sizeof(double));
  double *sigma = calloc(2 * ndet, sizeof(double));
  if (!h1p || !h2p || !c || !sigma) return 2;
  /* a real symmetric one-body part and a pair-symmetric two-body part, already folded */
  for (int i = 0; i < norb; ++i)
    for (int j = 0; j < norb; ++j) h1p[2 * (i * norb + j)] = 0.1 * (i + j + 1);
  for (int p = 0; p < npair; ++p)
    for (int q = 0; q < npair; ++q) {
      const int i = p / norb, j = p % norb, k = q / norb, l = q % norb;
      h2p[2 * ((size_t)p * npair + q)] = 0.01 * ((i + 1) * (j + 1) + (k + 1) * (l + 1));
    }
  for (int d = 0; d < ndet; ++d) c[2 * d] = 1.0 / sqrt((double)ndet);

  printf("libfqe_b200 version %d, %d CUDA device(s)\n", fqeb_version(), fqeb_device_count());
  const int rc = fqeb_sigma_restricted_host(norb, nalpha, nbeta, h1p, h2p, c, sigma);
  if (rc == FQEB_ERR_NODEVICE) {
    printf("no device: %s\n", fqeb_last_error());
    return 0;
  }
  if (rc != FQEB_OK) {
    fprintf(stderr, "fqeb_sigma_restricted_host failed (%d): %s\n", rc, fqeb_last_error());
    return 1;
  }
  double norm2 = 0.0;
  for (int d = 0; d < 2 * ndet; ++d) norm2 += sigma[d] * sigma[d];
  printf("|sigma| = %.15f\n", sqrt(norm2));
  free(h1p);
  free(h2p);
  free(c);
  free(sigma);
  return 0;
}
