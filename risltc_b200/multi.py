"""Image-stripe sharding of one frame over the GPUs of a box.

Every pixel's random stream depends only on (pixel index, frame word) (noise_utility.glsl:81-82)
and the shading pass never reads a neighbour, so any partition of the image gives the same image
as the single-GPU run, bit for bit. Rank r of N renders the rows y with
(y // stripe_height) % N == r (interleaved stripes balance background / emitter / cheap pixels),
keeps its accumulation stripes resident across samples per pixel, and the only communication is ONE
gather of the RGBA32F stripes to rank 0 per output frame (NCCL over NVLink; gloo in the CPU tests).

One process per GPU (torch.distributed); scene, BVH, lights and LTC tables are replicated.
"""
import numpy as np
import torch
import torch.distributed as dist


def owned_rows(height, stripe_height, rank, world):
    """Global row indices rendered by `rank`, in the order the device stores them
    (Stripes::global_row in csrc/common.cuh)."""
    rows = np.arange(height, dtype=np.int64)
    return rows[(rows // stripe_height) % world == rank]


def max_owned_rows(height, stripe_height, world):
    return max(len(owned_rows(height, stripe_height, r, world)) for r in range(world))


class StripeGather:
    """Gathers per-rank stripe slabs [owned_rows, W, 4] into the full frame [H, W, 4] on rank 0.

    Slabs are padded to the same number of rows so that a single equal-sized gather suffices (heights
    such as 1080 are not a multiple of stripe_height * world); the de-interleave is one index_copy on
    rank 0. The slab tensor is also the device's accumulation target (risltc_cuda_set_accum_buffer),
    so nothing is copied before the collective."""

    def __init__(self, width, height, stripe_height, rank=None, world=None, device="cpu", group=None):
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.group = group
        self.width, self.height, self.stripe_height = width, height, stripe_height
        self.rows = owned_rows(height, stripe_height, self.rank, self.world)
        self.pad_rows = max_owned_rows(height, stripe_height, self.world)
        self.device = torch.device(device)
        # the accumulation slab of this rank (zero-initialised: accum_num = 0 overwrites it anyway)
        self.slab = torch.zeros((self.pad_rows, width, 4), dtype=torch.float32, device=self.device)
        self.full = None
        self.parts = None
        if self.rank == 0:
            self.full = torch.zeros((height, width, 4), dtype=torch.float32, device=self.device)
            self.parts = [torch.empty_like(self.slab) for _ in range(self.world)]
            self.indices = [torch.from_numpy(owned_rows(height, stripe_height, r, self.world)).to(self.device) for r in range(self.world)]

    def gather(self):
        """One collective: every rank sends its slab, rank 0 assembles and returns the frame (else None)."""
        if self.world == 1:
            self.full.index_copy_(0, self.indices[0], self.slab[: len(self.rows)])
            return self.full
        dist.gather(self.slab, self.parts if self.rank == 0 else None, dst=0, group=self.group)
        if self.rank != 0:
            return None
        for r in range(self.world):
            self.full.index_copy_(0, self.indices[r], self.parts[r][: len(self.indices[r])])
        return self.full


def attach(dev, gatherer):
    """Make a risltc_device_t (api.Device) accumulate straight into the gatherer's slab."""
    assert dev.owned_rows == len(gatherer.rows), "device stripes and gather layout disagree"
    dev.set_accum_buffer(gatherer.slab.data_ptr())
