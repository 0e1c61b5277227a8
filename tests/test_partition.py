"""Host logic of the output-axis partitioned product (genfer_b200/partition.py).

CPU part: world_size-2 gloo group; the arithmetic is injected from the oracle (`O.mul_rows`, reference loop
order), so what is under test is the folded-cyclic row map, the all-gather of the block-sharded operand and
the re-assembly.  GPU part (-m gpu): the same class over gtp_mul_rowlist_raw on one device.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from genfer_b200.partition import (PartitionedProduct, row_work, rows_for_rank, shard_bounds, should_partition)
from genfer_b200.synth import synth_uniform


@pytest.mark.parametrize("n_rows,world", [(16, 1), (16, 2), (16, 4), (16, 8), (24, 4), (12, 8), (5, 2), (33, 8)])
def test_row_map_is_a_partition(n_rows, world):
    seen = sorted(k for r in range(world) for k in rows_for_rank(n_rows, world, r))
    assert seen == list(range(n_rows))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_map_balances_triangular_work(world):
    """Row k0 of a dense product costs k0+1 sub-products (:1002-1004); 16 rows fold into equal shares."""
    loads = [sum(row_work(k, 16, 16) for k in rows_for_rank(16, world, r)) for r in range(world)]
    assert len(set(loads)) == 1 and sum(loads) == 136


def test_shard_bounds_cover():
    for n, w in ((16, 8), (10, 4), (3, 8)):
        got = []
        for r in range(w):
            lo, hi, block = shard_bounds(n, w, r)
            assert hi - lo <= block
            got += list(range(lo, hi))
        assert got == list(range(n))


def test_partition_threshold():
    assert should_partition((16,) * 6, 8) and not should_partition((16,) * 5, 8) and not should_partition((16,) * 6, 1)


def _oracle_row_kernel(xshape, x, yshape, y, rshape, rows, out):
    from oracle import oracle as O
    r, _ = O.mul_rows(x.numpy(), y.numpy(), rshape, rows)
    out.copy_(torch.from_numpy(r[rows]))


def _worker(rank, world, port, shape, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.from_numpy(synth_uniform(shape, 11))
        y = torch.from_numpy(synth_uniform(shape, 12))
        pp = PartitionedProduct(shape, shape, shape, _oracle_row_kernel)
        y_shard = pp.shard_of(y)                      # this rank only keeps its block of Y ...
        rows = pp(x, y_shard)                         # ... and gets the rest through the all-gather
        full = pp.assemble(rows)
        if rank == 0:
            q.put(full.numpy())
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("shape", [(8, 6, 5), (5, 4, 3)])
def test_world2_gloo_partitioned_product_matches_oracle(shape):
    from oracle import oracle as O
    O.build()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, shape, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = O.mul_raw(synth_uniform(shape, 11), synth_uniform(shape, 12), shape)
    assert np.array_equal(full.view(np.uint64), ref.view(np.uint64))


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,world", [(4, 16, 8), (4, 16, 2), (3, 16, 4), (4, 12, 4), (3, 9, 2)])
def test_gpu_rowlist_shards_tile_the_product(n, d, world):
    """gtp_mul_rowlist_raw on each rank's folded-cyclic rows reproduces the single-launch product bit for bit
    (per-row summation order does not depend on which rows a launch holds ... for the ordered kernel), and to
    1e-12 for the tiled kernel whose split-K chunking depends on the launch."""
    import genfer_b200
    from genfer_b200.partition import gpu_row_kernel
    ctx = genfer_b200.Context(0)
    try:
        shape = (d,) * n
        x = torch.from_numpy(synth_uniform(shape, 1)).cuda()
        y = torch.from_numpy(synth_uniform(shape, 2)).cuda()
        full = torch.empty(shape, dtype=torch.float64, device="cuda")
        ctx.mul_rows_raw(shape, x.data_ptr(), shape, y.data_ptr(), shape, 0, 1, d, full.data_ptr())
        out = torch.full(shape, float("nan"), dtype=torch.float64, device="cuda")
        rk = gpu_row_kernel(ctx)
        for r in range(world):
            rows = rows_for_rank(d, world, r)
            part = torch.empty((len(rows),) + shape[1:], dtype=torch.float64, device="cuda")
            rk(shape, x, shape, y, shape, rows, part)
            ctx.synchronize()
            out[rows] = part
        ctx.synchronize()
        err = torch.max(torch.abs(out - full) / full).item()
        assert err <= 1e-12, err
    finally:
        ctx.close()
