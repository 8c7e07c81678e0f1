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

from fqe_b200 import settings


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
    nslices = int(settings.allreduce_slices)
    if nslices <= 1 or not sector.coeff.is_cuda:
        part = sector.apply_operator(op, row_range=rows, pair_range=pairs)
        return allreduce_sigma(part)
    # The scatter of the last chunk is issued by slices of target rows; the all-reduce of a
    # finished slice runs on a side stream while the next slice is scattered, so that of the
    # 2.65 GB reduction (norb = 16) only the last slice's share is exposed.
    part, pending = sector.apply_operator(op, row_range=rows, pair_range=pairs,
                                          defer_last_scatter=True)
    main = torch.cuda.current_stream()
    side = _side_stream(part.device)
    group = _reduce_group()
    flat = torch.view_as_real(part)
    for x0, x1 in split_even(sector.lena(), nslices):
        if x1 == x0:
            continue
        sector.finish_scatter(pending, x0, x1, part)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            dist.all_reduce(flat[x0:x1], op=dist.ReduceOp.SUM, group=group)
    main.wait_stream(side)
    return part


_SIDE_STREAMS = {}
_REDUCE_GROUP = []


def _side_stream(device) -> "torch.cuda.Stream":
    if device not in _SIDE_STREAMS:
        _SIDE_STREAMS[device] = torch.cuda.Stream(device=device, priority=-1)
    return _SIDE_STREAMS[device]


def _reduce_group():
    """Process group of the sliced sigma reduction.  The scatter launches thousands of CTAs per
    SM; a collective kernel on an ordinary stream only gets SMs when that grid has drained, i.e.
    nothing overlaps.  On NCCL the reduction therefore runs in its own communicator whose
    stream has HIGH priority, so that its few CTAs are placed ahead of the scatter's pending
    ones.  (Collective call: every rank reaches it in its first ``sharded_apply``.)"""
    if not _REDUCE_GROUP:
        group = dist.group.WORLD
        if dist.get_backend() == "nccl":
            try:
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
                group = dist.new_group(backend="nccl", pg_options=opts)
            except (AttributeError, RuntimeError, TypeError):
                group = dist.group.WORLD
        _REDUCE_GROUP.append(group)
    return _REDUCE_GROUP[0]


def block_slices(total: int, world: int) -> List[Tuple[int, int]]:
    """Equal blocks of ceil(total / world) rows (the last ones shorter or empty): the layout of
    one all-gather / reduce-scatter over a buffer padded to world * ceil(total / world) rows."""
    per = (total + world - 1) // world
    return [(min(total, r * per), min(total, (r + 1) * per)) for r in range(world)]


class ExchangeBuffers:
    """Device storage of one sector for the host-buffer path on ``world`` ranks: coefficient and
    sigma matrices padded to ``world * ceil(lena / world)`` rows, so that the coefficient
    exchange is ONE in-place ``all_gather_into_tensor`` and the reduction of the partial sigmas
    ONE ``reduce_scatter_tensor`` (each rank only needs the rows it sends home)."""

    def __init__(self, sector, world: int, rank: int):
        from fqe_b200.fqe_data import FqeData
        self.world, self.rank = world, rank
        lena, lenb = sector.lena(), sector.lenb()
        self.per = (lena + world - 1) // world
        self.slices = block_slices(lena, world)
        dev = sector.coeff.device
        self.cstore = torch.zeros((world * self.per, lenb), dtype=torch.complex128, device=dev)
        self.sstore = torch.empty((world * self.per, lenb), dtype=torch.complex128, device=dev)
        self.mine = torch.empty((self.per, lenb), dtype=torch.complex128, device=dev)
        self.data = FqeData(sector.nalpha(), sector.nbeta(), sector.norb(), sector.get_fcigraph())
        self.data.coeff = self.cstore[:lena]          # a contiguous prefix: no copy

    def gather_coeff(self) -> None:
        """every rank has written its block of ``cstore``; afterwards all ranks hold all blocks"""
        if self.world == 1:
            return
        r = self.rank
        dist.all_gather_into_tensor(torch.view_as_real(self.cstore),
                                    torch.view_as_real(self.cstore[r * self.per:(r + 1) * self.per]))

    def build(self, op, mode: str) -> torch.Tensor:
        """partial sigma of this rank's shard into ``sstore``; returns this rank's block of the
        SUM over ranks (rows ``slices[rank]``)"""
        lena = self.data.lena()
        rows, pairs = shard_plan(mode, self.rank, self.world, lena, op.npair)
        self.data.apply_operator(op, row_range=rows, pair_range=pairs, out=self.sstore[:lena])
        if self.world == 1:
            return self.sstore[:lena]
        if self.world * self.per > lena:
            self.sstore[lena:].zero_()
        dist.reduce_scatter_tensor(torch.view_as_real(self.mine), torch.view_as_real(self.sstore))
        r0, r1 = self.slices[self.rank]
        return self.mine[:r1 - r0]


def sharded_apply_host(sector, op, host_coeff: torch.Tensor, host_sigma: torch.Tensor,
                       mode: str = "det", buffers: "ExchangeBuffers" = None) -> None:
    """End-to-end sigma with HOST buffers.

    Every rank uploads its block of rows of ``host_coeff`` (pinned), ONE all-gather over NVLink
    completes the coefficient matrix on every GPU, the sharded sigma build runs, ONE
    reduce-scatter leaves on every rank the rows it is responsible for, and each rank
    downloads those rows into ``host_sigma``.  Host <-> device traffic per rank is
    2 * 16 * L^2 / world bytes.  With ``SharedHostBuffer`` arrays for ``host_coeff`` /
    ``host_sigma`` the caller's process (rank 0) ends up holding the complete host sigma while
    the PCIe traffic is spread over all GPUs."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    if buffers is None:
        buffers = ExchangeBuffers(sector, world, rank)
    r0, r1 = buffers.slices[rank]
    if r1 > r0:
        buffers.cstore[r0:r1].copy_(host_coeff[r0:r1], non_blocking=True)
    buffers.gather_coeff()
    nslices = int(settings.allreduce_slices)
    if world == 1 and nslices > 1:
        # one GPU: the scatter of the last chunk is issued by slices of target rows and every
        # finished slice starts its way home on a side stream while the next one is scattered
        lena = sector.lena()
        sig, pending = buffers.data.apply_operator(op, out=buffers.sstore[:lena],
                                                   defer_last_scatter=True)
        main, side = torch.cuda.current_stream(), _side_stream(sig.device)
        for x0, x1 in split_even(lena, nslices):
            if x1 == x0:
                continue
            buffers.data.finish_scatter(pending, x0, x1, sig)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                host_sigma[x0:x1].copy_(sig[x0:x1], non_blocking=True)
        main.wait_stream(side)
        main.synchronize()
        return
    mine = buffers.build(op, mode)
    if r1 > r0:
        host_sigma[r0:r1].copy_(mine, non_blocking=True)
    torch.cuda.current_stream().synchronize()


class SharedHostBuffer:
    """A pinned host array in POSIX shared memory that every rank of the node maps: rank 0 owns
    the data (``numpy in, numpy out`` for the calling process), the other ranks read / write
    their row blocks directly, so host <-> device copies run on all PCIe links at once."""

    def __init__(self, name: str, shape, dtype=torch.complex128, create: bool = False):
        import numpy
        import os
        self.path = os.path.join("/dev/shm", name)
        nbytes = int(numpy.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        if create:
            with open(self.path, "wb") as fh:
                fh.truncate(nbytes)
        self._map = numpy.memmap(self.path, dtype=numpy.uint8, mode="r+", shape=(nbytes,))
        self.tensor = torch.from_numpy(self._map).view(dtype).reshape(shape)
        self._registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0)
            self._registered = int(rc) == 0
        self._owner = create

    def close(self) -> None:
        import os
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self._registered = False
        if self._owner and os.path.exists(self.path):
            os.unlink(self.path)


class HostApplyStream:
    """A stream of independent end-to-end sigma builds with HOST buffers, software-pipelined:
    the upload of build k+1 and the download of build k-1 run on their own CUDA streams while
    build k computes (double-buffered device coefficients).  Every build still moves its own
    input host -> device and its own result device -> host; only the waiting is overlapped.

        pipe = HostApplyStream(sector, mode="det")
        for c_host, s_host in batches:          # pinned complex128 [lena, lenb] tensors
            pipe.submit(op, c_host, s_host)
        pipe.drain()                            # all results have landed in their s_host

    With more than one rank each rank uploads / downloads only its block of rows (as
    ``sharded_apply_host``); the all-gather and the reduce-scatter stay on the compute stream."""

    def __init__(self, sector, mode: str = "det", depth: int = 2):
        self.mode = mode
        self.world = dist.get_world_size() if (dist.is_available() and
                                               dist.is_initialized()) else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.s_in = torch.cuda.Stream()
        self.s_out = torch.cuda.Stream()
        self.slots = []
        for _ in range(depth):
            self.slots.append({"buf": ExchangeBuffers(sector, self.world, self.rank),
                               "h2d": torch.cuda.Event(), "done": torch.cuda.Event(),
                               "d2h": torch.cuda.Event(), "used": False})
        # the slots were initialised on the current stream: order the first uploads after that
        self.s_in.wait_stream(torch.cuda.current_stream())
        self.s_out.wait_stream(torch.cuda.current_stream())
        self.count = 0

    def submit(self, op, host_coeff: torch.Tensor, host_sigma: torch.Tensor) -> None:
        slot = self.slots[self.count % len(self.slots)]
        self.count += 1
        cur = torch.cuda.current_stream()
        buf = slot["buf"]
        r0, r1 = buf.slices[self.rank]
        if slot["used"]:
            slot["d2h"].synchronize()          # the result that lived in this slot is home
            self.s_in.wait_event(slot["done"])  # and its coefficients are no longer read
        with torch.cuda.stream(self.s_in):
            if r1 > r0:
                buf.cstore[r0:r1].copy_(host_coeff[r0:r1], non_blocking=True)
            slot["h2d"].record(self.s_in)
        cur.wait_event(slot["h2d"])
        buf.gather_coeff()
        mine = buf.build(op, self.mode)
        slot["done"].record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot["done"])
            if r1 > r0:
                host_sigma[r0:r1].copy_(mine, non_blocking=True)
            slot["d2h"].record(self.s_out)
        slot["used"] = True

    def drain(self) -> None:
        for slot in self.slots:
            if slot["used"]:
                slot["d2h"].synchronize()
        torch.cuda.current_stream().synchronize()
