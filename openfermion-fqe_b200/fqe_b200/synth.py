"""Seeded synthetic inputs for benchmarks and parity tests (SURVEY 8d recipe).

``rng = numpy.random.default_rng(20260000 + 100*norb + cfg)``; the state has real
and imaginary parts ~ U(-0.5, 0.5) and is normalised; h1 is real symmetric; h2 is
built from an 8-fold symmetric chemist-notation tensor, ``h2 = einsum('ijkl->ikjl',
v) * scale`` and handed over as complex128, so that both reference branches are
valid for it.  ``kind='herm'`` gives a general complex-Hermitian
operator that exercises the complex GEMM; ``kind='general'`` has no symmetry.
"""
from math import comb
from typing import Tuple

import numpy


def seed_for(norb: int, cfg: int = 0) -> int:
    return 20260000 + 100 * norb + cfg


def integrals(norb: int, kind: str = "real8", seed: int = None,
              scale: float = 0.1) -> Tuple[numpy.ndarray, numpy.ndarray]:
    rng = numpy.random.default_rng(seed_for(norb) if seed is None else seed)
    a = rng.standard_normal((norb, norb))
    if kind == "real8":
        h1 = 0.5 * (a + a.T)
        v = rng.standard_normal((norb,) * 4)
        v = v + v.transpose(1, 0, 2, 3)
        v = v + v.transpose(0, 1, 3, 2)
        v = v + v.transpose(2, 3, 0, 1)
        h2 = numpy.einsum("ijkl->ikjl", v) * scale
        return h1.astype(numpy.complex128), numpy.ascontiguousarray(h2).astype(numpy.complex128)
    b = rng.standard_normal((norb, norb))
    w = rng.standard_normal((norb,) * 4) + 1j * rng.standard_normal((norb,) * 4)
    if kind == "herm":
        h1 = 0.5 * ((a + 1j * b) + (a + 1j * b).conj().T)
        h2 = 0.5 * scale * (w + w.conj().transpose(3, 2, 1, 0))
        return h1, numpy.ascontiguousarray(h2)
    if kind == "general":
        return a + 1j * b, scale * w
    raise ValueError(f"unknown integral kind {kind!r}")


def state(lena: int, lenb: int, seed: int) -> numpy.ndarray:
    rng = numpy.random.default_rng(seed)
    c = rng.uniform(-0.5, 0.5, (lena, lenb)) + 1j * rng.uniform(-0.5, 0.5, (lena, lenb))
    return c / numpy.linalg.norm(c)


def sector_dims(nele: int, m_s: int, norb: int) -> Tuple[int, int, int, int]:
    nalpha = (nele + m_s) // 2
    nbeta = nele - nalpha
    return nalpha, nbeta, comb(norb, nalpha), comb(norb, nbeta)


def diagonal_coulomb_matrix(norb: int, seed: int, symmetric: bool = True) -> numpy.ndarray:
    rng = numpy.random.default_rng(seed)
    v = rng.uniform(0.0, 1.0, (norb, norb)) * 8.0 / (norb * norb)
    return 0.5 * (v + v.T) if symmetric else v
