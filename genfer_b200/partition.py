"""Output-axis partitioning of a large truncated product across the GPUs of one box.

The general product (multivariate_taylor.rs:984-1012) recurses over the leading axis first:
``Z[k0, ...] = sum_{j0 <= k0} X[j0, ...] (*) Y[k0 - j0, ...]`` (:1001-1010), so leading-axis output rows
are independent given both operands.  Row k0 costs (k0 + 1) sub-products, so rows are dealt to ranks
with a *folded cyclic* map -- rank r owns the rows k0 with ``k0 mod 2W in {r, 2W-1-r}`` -- which gives
every rank the same MAC count whenever 2W divides the row count (16 rows on 2, 4 or 8 GPUs).  One
operand lives block-sharded along axis 0 and is replicated with ONE all-gather (NCCL over NVLink on the
GPU box, gloo in the CPU tests); the result stays sharded by the same row map.

This module is host logic only: the arithmetic is done by ``row_kernel`` -- on the GPU that is
``Context.mul_rowlist_raw`` (gtp_mul_rowlist_raw); the CPU tests inject the oracle's row product to check
the row map, the gather and the re-assembly under a world_size-2 gloo group.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

PARTITION_THRESHOLD = 10_000_000  # output coefficients; smaller products stay on one GPU (north_star)


def rows_for_rank(n_rows: int, world: int, rank: int) -> List[int]:
    """Folded-cyclic leading-axis rows of `rank` (ascending)."""
    period = 2 * world
    return [k for k in range(n_rows) if k % period in (rank, period - 1 - rank)]


def row_work(k0: int, xlen: int, ylen: int) -> int:
    """Number of (j0, k0-j0) sub-products of row k0: the trip count of :1002-1004."""
    lo = max(0, k0 + 1 - ylen)
    hi = min(k0 + 1, xlen)
    return max(0, hi - lo)


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Contiguous block [lo, hi) of `n` leading-axis slices held by `rank`; also the padded block length."""
    block = (n + world - 1) // world
    lo = min(n, rank * block)
    hi = min(n, lo + block)
    return lo, hi, block


def should_partition(rshape: Sequence[int], world: int) -> bool:
    n = 1
    for d in rshape:
        n *= int(d)
    return world > 1 and n >= PARTITION_THRESHOLD and len(rshape) >= 2 and rshape[0] >= 2 * world


RowKernel = Callable[[Sequence[int], torch.Tensor, Sequence[int], torch.Tensor, Sequence[int], List[int], torch.Tensor], None]


class PartitionedProduct:
    """Z = X (*) Y truncated to `rshape`, output rows sharded over the ranks of `group`.

    `row_kernel(xshape, x, yshape, y, rshape, rows, out_rows)` must fill out_rows[i] with output row rows[i].
    """

    def __init__(self, xshape: Sequence[int], yshape: Sequence[int], rshape: Sequence[int], row_kernel: RowKernel,
                 group: Optional[dist.ProcessGroup] = None):
        self.xshape, self.yshape, self.rshape = tuple(xshape), tuple(yshape), tuple(rshape)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rows = rows_for_rank(self.rshape[0], self.world, self.rank)
        self.row_kernel = row_kernel
        self._gathered: Optional[torch.Tensor] = None
        self._out: Optional[torch.Tensor] = None

    # -- operand replication -----------------------------------------------------------------
    def shard_of(self, full: torch.Tensor) -> torch.Tensor:
        """This rank's zero-padded block of a full operand (used to set up the sharded state)."""
        lo, hi, block = shard_bounds(full.shape[0], self.world, self.rank)
        s = torch.zeros((block,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
        s[: hi - lo] = full[lo:hi]
        return s

    def gather_operand(self, shard: torch.Tensor, n_slices: int) -> torch.Tensor:
        """All-gather the block shards along axis 0; returns the full operand (a prefix view)."""
        if self.world == 1:
            return shard[:n_slices]
        block = shard.shape[0]
        if self._gathered is None or self._gathered.shape != (self.world * block,) + tuple(shard.shape[1:]):
            self._gathered = torch.empty((self.world * block,) + tuple(shard.shape[1:]), dtype=shard.dtype,
                                         device=shard.device)
        dist.all_gather_into_tensor(self._gathered, shard.contiguous(), group=self.group)
        return self._gathered[:n_slices]

    # -- the sharded product --------------------------------------------------------------------
    def __call__(self, x_full: torch.Tensor, y_shard: torch.Tensor) -> torch.Tensor:
        y_full = self.gather_operand(y_shard, self.yshape[0])
        want = (len(self.rows),) + self.rshape[1:]
        if self._out is None or tuple(self._out.shape) != want:
            self._out = torch.empty(want, dtype=x_full.dtype, device=x_full.device)
        if self.rows:
            self.row_kernel(self.xshape, x_full, self.yshape, y_full, self.rshape, self.rows, self._out)
        return self._out

    def assemble(self, out_rows: torch.Tensor) -> torch.Tensor:
        """Full result on every rank (tests / final read-back): all-gather the row shards and un-permute."""
        if self.world == 1:
            return out_rows
        per = max(len(rows_for_rank(self.rshape[0], self.world, r)) for r in range(self.world))
        pad = torch.zeros((per,) + self.rshape[1:], dtype=out_rows.dtype, device=out_rows.device)
        pad[: out_rows.shape[0]] = out_rows
        allr = torch.empty((self.world * per,) + self.rshape[1:], dtype=out_rows.dtype, device=out_rows.device)
        dist.all_gather_into_tensor(allr, pad, group=self.group)
        full = torch.empty(self.rshape, dtype=out_rows.dtype, device=out_rows.device)
        for r in range(self.world):
            for i, k in enumerate(rows_for_rank(self.rshape[0], self.world, r)):
                full[k] = allr[r * per + i]
        return full


def gpu_row_kernel(ctx) -> RowKernel:
    """Row kernel backed by gtp_mul_rowlist_raw on `ctx` (tensors must live on ctx's device).

    Stream contract: the operands (x, the all-gathered y) and `out` are produced / consumed on torch's CURRENT stream,
    the product is launched on the Context's stream.  When the two are the same stream (create the Context with
    ``stream=torch.cuda.current_stream().cuda_stream``, as bench.py does) everything is stream-ordered and nothing
    waits.  When they differ, the kernel is fenced on both sides -- torch's stream is drained before the launch and the
    Context's stream after it -- so the product can neither read y before the all-gather has finished nor be overtaken by
    a later all-gather of `out` (correct for any caller; the same-stream set-up is the fast one)."""

    def run(xshape, x, yshape, y, rshape, rows, out):
        assert x.is_cuda and y.is_cuda and out.is_cuda, "the f64 Taylor path has no CPU fallback"
        assert x.is_contiguous() and y.is_contiguous() and out.is_contiguous()
        same = ctx.stream == torch.cuda.current_stream(x.device).cuda_stream
        if not same:
            torch.cuda.current_stream(x.device).synchronize()
        ctx.mul_rowlist_raw(xshape, x.data_ptr(), yshape, y.data_ptr(), rshape, rows, out.data_ptr())
        if not same:
            ctx.synchronize()

    return run
