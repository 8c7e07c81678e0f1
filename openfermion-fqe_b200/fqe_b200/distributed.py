"""Multi-GPU sigma: one process per GPU, coefficient vector replicated, work sharded,
one allreduce of sigma.

Two partitions of the sigma build are supported; both leave a PARTIAL sigma of full
size on every rank and need exactly one ``all_reduce(SUM)`` of 2*lena*lenb doubles:

``pair``  the partition BASELINE.json's north_star mandates: the dvec pair index
          ij in [0, norb^2) is split into contiguous slices; rank r gathers D for its
          slice, contracts it against h2'[:, slice] (K = norb^2 / world) and scatters
          all norb^2 rows of its partial E.  Gather and GEMM shrink with the world
          size, the scatter does not.
``det``   the alpha-row (determinant) index is split; rank r runs the complete
          gather -> GEMM -> scatter pipeline on its rows.  All three phases shrink
          with the world size; this is the default.

The sharding arithmetic below is pure Python so it can be exercised on CPU with the
gloo backend (tests/test_distributed_cpu.py).
"""
from typing import List, Tuple

import torch
import torch.distributed as dist


def split_even(total: int, world: int, align: int = 1) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) slices of range(total) for each rank; every boundary except
    the last is a multiple of ``align``; slices differ by at most one aligned unit."""
    if world < 1 or total < 0 or align < 1:
        raise ValueError("bad arguments to split_even")
    units = (total + align - 1) // align
    base, extra = divmod(units, world)
    out, lo = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        hi = min(total, lo + n * align)
        out.append((lo, hi))
        lo = hi
    return out


def shard_plan(mode: str, rank: int, world: int, lena: int, npair: int):
    """(row_range, pair_range) of one rank; ``npair`` is the size of the operator's pair
    space (``DenseOperator.npair``: norb^2, or norb(norb+1)/2 when compressed)."""
    if mode == "det":
        return split_even(lena, world)[rank], (0, npair)
    if mode == "pair":
        # slices start at multiples of 16 pairs so that no k-padding is needed
        return (0, lena), split_even(npair, world, align=16 if npair >= 16 * world else 2)[rank]
    raise ValueError(f"unknown shard mode {mode!r}")


def allreduce_sigma(sigma: torch.Tensor) -> torch.Tensor:
    """Sum partial sigma vectors over all ranks, in place (NCCL over NVLink on GPU,
    gloo in the CPU tests).  complex128 is reduced as float64[..., 2]."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(torch.view_as_real(sigma), op=dist.ReduceOp.SUM)
    return sigma


def sharded_apply(sector, op, mode: str = "det") -> torch.Tensor:
    """sigma = H C with the work of ``sector.apply_operator`` spread over the default
    process group.  Every rank must hold the same coefficients."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sector.apply_operator(op)
    rows, pairs = shard_plan(mode, dist.get_rank(), dist.get_world_size(), sector.lena(),
                             op.npair)
    part = sector.apply_operator(op, row_range=rows, pair_range=pairs)
    return allreduce_sigma(part)


def sharded_apply_host(sector, op, host_coeff: torch.Tensor, host_sigma: torch.Tensor,
                       mode: str = "det") -> None:
    """End-to-end sigma with HOST buffers on every rank.

    Each rank uploads only its 1/world row slice of ``host_coeff`` (pinned), the slices
    are exchanged over NVLink (one broadcast per owner) so that every GPU holds the full
    coefficient matrix, the sharded sigma build and its all-reduce run on the devices,
    and each rank downloads its row slice of the result into ``host_sigma``.  Host <->
    device traffic per rank is 2 * 16 * L^2 / world bytes instead of 2 * 16 * L^2."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        sector.coeff.copy_(host_coeff, non_blocking=True)
        sigma = sector.apply_operator(op)
        host_sigma.copy_(sigma, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return
    rank = dist.get_rank()
    slices = split_even(sector.lena(), world)
    r0, r1 = slices[rank]
    sector.coeff[r0:r1].copy_(host_coeff[r0:r1], non_blocking=True)
    for src, (s0, s1) in enumerate(slices):
        if s1 > s0:
            dist.broadcast(torch.view_as_real(sector.coeff[s0:s1]), src=src)
    sigma = sharded_apply(sector, op, mode)
    host_sigma[r0:r1].copy_(sigma[r0:r1], non_blocking=True)
    torch.cuda.current_stream().synchronize()


class HostApplyStream:
    """A stream of independent end-to-end sigma builds with HOST buffers, software-pipelined:
    the upload of build k+1 and the download of build k-1 run on their own CUDA streams while
    build k computes (double-buffered device coefficients).  Every build still moves its own
    input host -> device and its own result device -> host; only the waiting is overlapped.

        pipe = HostApplyStream(sector, mode="det")
        for c_host, s_host in batches:          # pinned complex128 [lena, lenb] tensors
            pipe.submit(op, c_host, s_host)
        pipe.drain()                            # all results have landed in their s_host

    With more than one rank each rank uploads / downloads only its row slice (as
    ``sharded_apply_host``); the exchange and the all-reduce stay on the compute stream."""

    def __init__(self, sector, mode: str = "det", depth: int = 2):
        from fqe_b200.fqe_data import FqeData
        self.mode = mode
        self.world = dist.get_world_size() if (dist.is_available() and
                                               dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.slices = split_even(sector.lena(), self.world)
        self.s_in = torch.cuda.Stream()
        self.s_out = torch.cuda.Stream()
        self.slots = []
        for _ in range(depth):
            data = FqeData(sector.nalpha(), sector.nbeta(), sector.norb(), sector.get_fcigraph())
            self.slots.append({"data": data, "sigma": None, "h2d": torch.cuda.Event(),
                               "done": torch.cuda.Event(), "d2h": torch.cuda.Event(),
                               "used": False})
        self.count = 0

    def submit(self, op, host_coeff: torch.Tensor, host_sigma: torch.Tensor) -> None:
        slot = self.slots[self.count % len(self.slots)]
        self.count += 1
        cur = torch.cuda.current_stream()
        r0, r1 = self.slices[self.rank]
        data = slot["data"]
        if slot["used"]:
            slot["d2h"].synchronize()          # the result that lived in this slot is home
            self.s_in.wait_event(slot["done"])  # and its coefficients are no longer read
        with torch.cuda.stream(self.s_in):
            data.coeff[r0:r1].copy_(host_coeff[r0:r1], non_blocking=True)
            slot["h2d"].record(self.s_in)
        cur.wait_event(slot["h2d"])
        if self.world > 1:
            for src, (s0, s1) in enumerate(self.slices):
                if s1 > s0:
                    dist.broadcast(torch.view_as_real(data.coeff[s0:s1]), src=src)
        sigma = sharded_apply(data, op, self.mode)
        slot["done"].record(cur)
        sigma.record_stream(self.s_out)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot["done"])
            host_sigma[r0:r1].copy_(sigma[r0:r1], non_blocking=True)
            slot["d2h"].record(self.s_out)
        slot["sigma"] = sigma
        slot["used"] = True

    def drain(self) -> None:
        for slot in self.slots:
            if slot["used"]:
                slot["d2h"].synchronize()
                slot["sigma"] = None
        torch.cuda.current_stream().synchronize()
