"""String tables, Z matrices and E_ij maps built on the GPU are BIT-EXACT with the
reference (golden fixtures recorded from the reference's FciGraph, the oracle
restatement and the compiled reference C functions)."""
import os

import numpy as np
import pytest

from oracle import fqe_oracle as O
from oracle import ref_harness as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def graphs(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_graphs.npz"))


def _graph(na, nb, norb):
    from fqe_b200.fci_graph import FciGraph
    return FciGraph(na, nb, norb)


@pytest.mark.parametrize("cfg", [(2, 1, 4), (2, 3, 6), (4, 4, 8), (3, 5, 8), (0, 2, 5),
                                 (5, 5, 10), (1, 1, 1), (3, 3, 3), (2, 2, 7)])
def test_tables_match_reference_goldens(graphs, cfg):
    na, nb, norb = cfg
    k = f"{na}_{nb}_{norb}"
    g = _graph(na, nb, norb)
    assert np.array_equal(g.string_alpha_all(), graphs[k + "_astr"])
    assert np.array_equal(g.string_beta_all(), graphs[k + "_bstr"])
    for i in range(norb):
        for j in range(norb):
            assert np.array_equal(g.alpha_map(i, j), graphs[f"{k}_amap_{i}_{j}"]), (i, j)
            assert np.array_equal(g.beta_map(i, j), graphs[f"{k}_bmap_{i}_{j}"]), (i, j)
    assert np.array_equal(g._dexca, graphs[k + "_dexca"])
    assert np.array_equal(g._dexcb, graphs[k + "_dexcb"])


def test_literal_known_answers():
    # reference tests/fci_graph_test.py:28-207 and SURVEY F4
    g = _graph(3, 3, 6)
    assert g.string_alpha_all()[:6].tolist() == [7, 11, 19, 35, 13, 21]
    g = _graph(2, 1, 4)
    assert g.string_alpha_all().tolist() == [3, 5, 9, 6, 10, 12]
    assert g.string_beta_all().tolist() == [1, 2, 4, 8]
    assert g.alpha_map(0, 2).tolist() == [[3, 0, -1], [5, 2, 1]]
    assert g.index_alpha(9) == 2 and g.string_beta(3) == 8
    g = _graph(4, 4, 8)
    assert g.index_alpha((1 << 1) | (1 << 2) | (1 << 3) | (1 << 7)) == 38


@pytest.mark.parametrize("cfg", [(6, 6, 12), (7, 7, 14), (8, 8, 16), (9, 7, 16), (1, 15, 16),
                                 (16, 0, 16), (3, 2, 20), (2, 2, 40), (1, 1, 63)])
def test_large_tables_match_oracle(cfg):
    na, nb, norb = cfg
    g = _graph(na, nb, norb)
    for spin, nele in ((0, na), (1, nb)):
        assert np.array_equal(g.z_matrix(spin), O.z_matrix(norb, nele))
        strings = O.build_strings(nele, norb)
        assert np.array_equal(g._strings(spin), strings)
        # dense signed map against an independent vectorised evaluation
        dense = g._dense_map(spin)
        rng = np.random.default_rng(norb * 100 + nele)
        pairs = [(i, j) for i in range(norb) for j in range(norb)]
        for idx in rng.choice(len(pairs), size=min(40, len(pairs)), replace=False):
            i, j = pairs[idx]
            row = dense[i * norb + j]
            bi, bj = np.uint64(1 << i), np.uint64(1 << j)
            if i == j:
                sel = (strings & bj) != 0
                exp = np.where(sel, np.arange(len(strings)) + 1, 0)
            else:
                sel = ((strings & bj) != 0) & ((strings & bi) == 0)
                tgt = O.string_addresses((strings[sel] | bi) & ~bj, norb, nele)
                lo, hi = min(i, j), max(i, j)
                mask = np.uint64(((1 << hi) - 1) & ~((1 << (lo + 1)) - 1))
                par = np.array([bin(int(x)).count("1") & 1 for x in (strings[sel] & mask)])
                exp = np.zeros(len(strings), dtype=np.int64)
                exp[sel] = (tgt + 1) * (1 - 2 * par)
            assert np.array_equal(row, exp.astype(np.int32)), (i, j)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("cfg", [(6, 6, 12), (7, 7, 14), (8, 8, 16), (5, 8, 13)])
def test_tables_match_compiled_reference(cfg):
    na, nb, norb = cfg
    g = _graph(na, nb, norb)
    rg = R.graph(na, nb, norb)
    assert np.array_equal(g.string_alpha_all(), rg.astr)
    assert np.array_equal(g.string_beta_all(), rg.bstr)
    assert np.array_equal(g._dexca, rg.dexca)
    assert np.array_equal(g._dexcb, rg.dexcb)
    for (i, j) in [(0, 0), (0, norb - 1), (norb - 1, 0), (3, 5), (norb // 2, 1)]:
        assert np.array_equal(g.alpha_map(i, j), rg.alpha_map[(i, j)])
        assert np.array_equal(g.beta_map(i, j), rg.beta_map[(i, j)])


def test_argument_errors():
    from fqe_b200.fci_graph import FciGraph
    for bad in [(-1, 0, 4), (0, -1, 4), (5, 0, 4), (0, 5, 4), (1, 1, -2)]:
        with pytest.raises(ValueError):
            FciGraph(*bad)
