"""Image-stripe sharding on CPU: world_size-2 and -3 gloo process groups gather synthetic stripes into the
frame on rank 0; the row ownership must equal the device's Stripes::global_row mapping."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from risltc_b200 import multi


def test_row_ownership_partitions_the_image():
    for height, stripe, world in ((1080, 8, 8), (2160, 8, 8), (123, 8, 3), (36, 16, 2), (7, 8, 4)):
        seen = np.concatenate([multi.owned_rows(height, stripe, r, world) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(height))
        # same mapping as Stripes::global_row (csrc/common.cuh): local row l of rank r
        for r in range(world):
            rows = multi.owned_rows(height, stripe, r, world)
            for l, g in enumerate(rows):
                assert g == ((l // stripe) * world + r) * stripe + l % stripe


def _worker(rank, world, port, width, height, stripe, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = multi.StripeGather(width, height, stripe, rank, world, device="cpu")
    rows = torch.from_numpy(g.rows)
    # every pixel encodes (global row, column, rank): what a device would have accumulated into its slab
    slab = torch.zeros_like(g.slab)
    slab[: len(rows), :, 0] = rows[:, None].float()
    slab[: len(rows), :, 1] = torch.arange(width)[None, :].float()
    slab[: len(rows), :, 2] = float(rank)
    slab[: len(rows), :, 3] = 1.0
    g.slab.copy_(slab)
    full = g.gather()
    if rank == 0:
        torch.save(full.clone(), out)
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height,stripe", [(2, 36, 8), (3, 50, 4)])
def test_gather_over_gloo(tmp_path, world, height, stripe):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "full.pt")
    width = 20
    mp.spawn(_worker, args=(world, port, width, height, stripe, out), nprocs=world, join=True)
    full = torch.load(out)
    assert full.shape == (height, width, 4)
    assert torch.equal(full[:, :, 0], torch.arange(height)[:, None].float().expand(height, width))
    assert torch.equal(full[:, :, 1], torch.arange(width)[None, :].float().expand(height, width))
    owner = (torch.arange(height) // stripe) % world
    assert torch.equal(full[:, 0, 2], owner.float()) and torch.all(full[:, :, 3] == 1.0)
