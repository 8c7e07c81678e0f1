#!/usr/bin/env python
"""Numerical prototype of the INT8-sliced (Ozaki-style) two-electron contraction (CPU, numpy).

The FP64 contraction E = A . D (A = h2' operand, real [P x P]; D = gathered +-C values,
real view [P x 2*ndet]) is replaced by exact integer products of signed radix-R digits:

    C / S  = sum_i  c_i R^-(i+1),   |c_i| <= (R-1)/2           one GLOBAL scale S per state
    A / T  = sum_j  a_j R^-(j+1),   |a_j| <= (R-1)/2           one scale per operator
    D digits = +-c_i(alpha source) +- c_i(beta source)           |.| <= R-1 < 128: fits int8
    E = S T sum_{i+j<=dmax} R^-(i+j+2) (a_j . d_i)               int8 x int8 -> int32, exact

This script measures the relative 2-norm error of sigma against the FP64 oracle as a function of
(R, number of slices, dmax) for uniform-random and for strongly non-uniform states, to fix the
parameters of the tcgen05 kernel (target: <= 1e-11, north_star tolerance 1e-10).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
from oracle import fqe_oracle as O  # noqa: E402


def digits(x, radix, nslice):
    """balanced radix-`radix` digits of x (|x| < 0.5): x ~ sum_i d_i radix^-(i+1)"""
    h = (radix - 1) // 2
    n = np.rint(x * float(radix) ** nslice).astype(np.int64)
    out = []
    for _ in range(nslice):
        d = (n + h) % radix - h
        out.append(d)
        n = (n - d) // radix
    assert np.all(n == 0), "value out of range for the digit expansion"
    return out[::-1]      # most significant first


def sliced_contract(a, dre, radix, ns_a, ns_d, dmax):
    """a: real [M, K]; dre: real [K, N] whose entries are sums of two values with |.| < S/2.
    returns the emulated product"""
    t = 2.0 * np.abs(a).max() * (1 + 2.0 ** -40)
    ad = digits(a / t, radix, ns_a)
    return ad, t


def run(norb, kind_state, radix, ns, dmax, seed=1):
    from fqe_b200 import synth
    na = nb = norb // 2
    g = O.graph(na, nb, norb)
    h1, h2 = synth.integrals(norb, "real8")
    rng = np.random.default_rng(seed)
    c = synth.state(g.lena, g.lenb, seed=seed)
    if kind_state == "spiky":      # magnitudes spread over 12 decades
        mag = 10.0 ** rng.uniform(-12, 0, c.shape)
        c = c * mag
        c /= np.linalg.norm(c)
    elif kind_state == "hf":       # one dominant determinant + small tail
        c = 1e-6 * c
        c[0, 0] = 1.0
        c /= np.linalg.norm(c)
    ref = O.sigma_restricted(g, c, h1, h2)
    # reference pieces in the oracle's own conventions
    h1p, h2p = O.fold_restricted(h1, h2)
    n_elec = na + nb
    npair = norb * norb
    amat = h2p.reshape(npair, npair).copy()
    for k in range(norb):               # absorb the one-body term (sum_k D[kk] = n_elec C)
        amat[:, k * norb + k] += h1p.reshape(-1) / n_elec
    assert np.abs(amat.imag).max() == 0
    amat = amat.real
    # digits of C (global scale) and of the gathered D
    s = 2.0 * max(np.abs(c.real).max(), np.abs(c.imag).max()) * (1 + 2.0 ** -40)
    cre = digits(c.real / s, radix, ns)
    cim = digits(c.imag / s, radix, ns)
    t = 2.0 * np.abs(amat).max() * (1 + 2.0 ** -40)
    ad = digits(amat / t, radix, ns)
    acc = np.zeros((npair, g.lena, g.lenb), dtype=np.complex128)
    worst_digit = 0
    for i in range(ns):
        # D digit plane i: dvec_spatial is linear with +-1 coefficients, so applying it to the
        # (integer) digit plane gives exactly the signed digit sums the kernel forms
        dplane = O.dvec_spatial(g, (cre[i] + 1j * cim[i]).astype(np.complex128))
        dplane = dplane.reshape(npair, g.lena, g.lenb)
        worst_digit = max(worst_digit, int(np.abs(dplane.real).max()), int(np.abs(dplane.imag).max()))
        for j in range(ns):
            if i + j > dmax:
                continue
            prod = np.tensordot(ad[j].astype(np.float64), dplane, axes=([1], [0]))  # exact ints
            assert np.abs(prod.real).max() < 2 ** 31
            acc += prod * float(radix) ** -(i + j + 2)
    evec = (acc * (s * t)).reshape(norb, norb, g.lena, g.lenb)
    out = O.coeff_from_dvec(g, evec)
    return O.rel_err(out, ref), worst_digit


if __name__ == "__main__":
    norb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    for state in ("uniform", "spiky", "hf"):
        for radix, ns, dmax in [(127, 5, 4), (127, 6, 5), (127, 6, 6), (127, 7, 6), (127, 7, 7),
                                (127, 8, 7)]:
            err, wd = run(norb, state, radix, ns, dmax)
            nprod = sum(1 for i in range(ns) for j in range(ns) if i + j <= dmax)
            print(f"norb={norb} state={state:8s} radix={radix} slices={ns} dmax={dmax} "
                  f"products={nprod:2d} max|D digit|={wd:3d}  rel err = {err:.2e}", flush=True)
