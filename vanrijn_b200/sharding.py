"""Sample-index sharding across the GPUs of one box.

Samples are pure functions of (seed, pixel, sample index), so GPU g of G renders sample indices
g, g+G, g+2G, ... of every pixel into its own accumulation buffer; the per-GPU (sum XYZ, sum weight)
arrays are then summed into rank 0's AccumulationBuffer with ONE reduce (NCCL on the GPU box, gloo in
the CPU tests).  That is the reference's own decomposition: src/main.rs:199-209 farms whole-frame sample
passes to workers and merges them (accumulation_buffer.rs:62-85).  There is no other data-path exchange.
"""


def shard_samples(rank, world, step, spp):
    """(sample_offset, sample_stride) of `rank` for step `step` when every rank renders `spp` samples per
    step: the step covers global sample indices [step*spp*world, (step+1)*spp*world)."""
    if not (0 <= rank < world) or spp < 0 or step < 0:
        raise ValueError("bad shard request")
    return rank + world * step * spp, world


def shard_sample_indices(rank, world, step, spp):
    off, stride = shard_samples(rank, world, step, spp)
    return [off + j * stride for j in range(spp)]


def reduce_accumulation(colour_sum, weight, dst=0):
    """Sum the per-rank (colour_sum, weight) tensors into rank `dst` (in place) -- the AccumulationBuffer
    merge expressed on sums.  Works for CUDA tensors over NCCL and CPU tensors over gloo."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(colour_sum, dst=dst, op=dist.ReduceOp.SUM)
        dist.reduce(weight, dst=dst, op=dist.ReduceOp.SUM)
    return colour_sum, weight


def finalize_colour(colour_sum, weight):
    """colour = colour_sum * (1 / weight)  (accumulation_buffer.rs:59)."""
    inv = 1.0 / weight
    return colour_sum.reshape(-1, 3) * inv.reshape(-1, 1)
