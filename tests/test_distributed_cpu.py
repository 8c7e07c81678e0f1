"""world_size-2 and -3 runs of the sharding logic on CPU with the gloo backend.

The CUDA kernels cannot run here, so each rank evaluates its shard of the sigma
build with the CPU oracle (restricted to the shard's alpha rows / pair slice, which is
exactly what fqeb_sigma_restricted does with [row0,row1) x [ij0,ij1)) and the test
checks that fqe_b200.distributed's shard plan + allreduce reproduce the full sigma."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _partial_sigma(g, c, h1, h2, rows, pairs):
    """Oracle evaluation of one shard: D built for `pairs` x `rows`, contracted with
    h2'[:, pairs], scattered into a full-size partial sigma (+ the shard's h1 term)."""
    from oracle import fqe_oracle as O
    n = g.norb
    npair = n * n
    h1p, h2p = O.fold_restricted(h1, h2)
    d = O.dvec_spatial(g, c).reshape(npair, g.lena, g.lenb)
    r0, r1 = rows
    p0, p1 = pairs
    dsh = np.zeros_like(d)
    dsh[p0:p1, r0:r1] = d[p0:p1, r0:r1]
    out = np.einsum("p,pab->ab", h1p.reshape(-1)[p0:p1], dsh[p0:p1])
    e = (h2p.reshape(npair, npair) @ dsh.reshape(npair, -1)).reshape(n, n, g.lena, g.lenb)
    return out + O.coeff_from_dvec(g, e)


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "openfermion-fqe_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fqe_b200 import synth
        from fqe_b200.distributed import allreduce_sigma, shard_plan
        from oracle import fqe_oracle as O
        na, nb, norb = 3, 2, 6
        g = O.graph(na, nb, norb)
        h1, h2 = synth.integrals(norb, "herm")
        c = synth.state(g.lena, g.lenb, seed=11)
        rows, pairs = shard_plan(mode, rank, world, g.lena, norb * norb)
        part = torch.from_numpy(np.ascontiguousarray(_partial_sigma(g, c, h1, h2, rows, pairs)))
        total = allreduce_sigma(part)
        ref = O.sigma_restricted(g, c, h1, h2)
        q.put((rank, O.rel_err(total.numpy(), ref), rows, pairs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "det"), (2, "pair"), (3, "pair"), (3, "det")])
def test_sharded_sigma_sums_to_full(world, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered_rows, covered_pairs = set(), set()
    for rank, err, rows, pairs in results:
        assert err < 1e-12, (rank, err)
        covered_rows.add(rows)
        covered_pairs.add(pairs)
    if mode == "det":
        assert len(covered_rows) == world and len(covered_pairs) == 1
    else:
        assert len(covered_pairs) == world and len(covered_rows) == 1
