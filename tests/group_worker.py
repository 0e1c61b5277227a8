"""Worker of tests/test_group.py::test_two_rank_group_under_torchrun: real peers, NCCL inside the library."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist


def main():
    import genfer_b200
    from oracle import oracle as O
    from helpers import synth_uniform
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                       # only to ship the NCCL id; the data path is the library's NCCL
    ids = [genfer_b200.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx = genfer_b200.Context.create_group(local, rank, world, ids[0])
    ctx.set_partition_threshold(1000)
    shape = (8, 6, 16, 16)
    x, y = synth_uniform(shape, 5), synth_uniform(shape, 6)
    lo, hi, block = genfer_b200.partition_block(shape[0], world, rank)
    xb = np.ascontiguousarray(x[lo:hi])
    X = genfer_b200.TaylorPoly.from_host_block_ptr(xb.ctypes.data, shape, shape, ctx)
    Y = genfer_b200.TaylorPoly.new(y, shape, ctx)
    Z = X * Y
    assert Z.is_distributed()
    rows = Z.local_rows()
    assert rows == genfer_b200.partition_rows(shape[0], world, rank)
    ref = (O.taylor(x) * O.taylor(y)).array()
    local_rows = np.empty((len(rows),) + shape[1:])
    Z.to_host_local_ptr(local_rows.ctypes.data)
    np.testing.assert_allclose(local_rows, ref[rows], rtol=1e-12)
    W = Z * Y                                            # the chain: Z replicated by grouped broadcasts, W sharded
    np.testing.assert_allclose(W.array(), (O.taylor(ref) * O.taylor(y)).array(), rtol=1e-12)
    s = (Z + Y).array()                                  # any other consumer sees the replicated tensor
    np.testing.assert_allclose(s, ref + y, rtol=1e-12)
    ctx.close()
    dist.barrier()
    if rank == 0:
        print("group worker ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
