"""Frame-wise sharding of a batch over the GPUs of one node.

Frames (and frame pairs) are independent units: /root/reference/src/ORBextractor.cc:620-678 reads no state from
a previous frame.  Rank g of G owns the contiguous range [g*F/G, (g+1)*F/G); there is no collective in the
extraction loop.  The only exchange is one final gather of the fixed-capacity result slabs
([frames][capacity] keypoints and descriptors plus a count per frame) to rank 0 -- NCCL over NVLink on the GPU
box, gloo in the CPU tests.  The gathered result is byte-identical to a single-rank run by construction.
"""
import torch
import torch.distributed as dist


def frame_range(rank, world, nframes):
    """Contiguous frames [lo, hi) of rank `rank` out of `world`."""
    if not (0 <= rank < world) or nframes < 0:
        raise ValueError("bad shard request")
    return rank * nframes // world, (rank + 1) * nframes // world


def max_shard(world, nframes):
    return max(frame_range(r, world, nframes)[1] - frame_range(r, world, nframes)[0] for r in range(world))


def gather_slabs(slabs, nframes, group=None, dst=0):
    """slabs: list of tensors whose first dimension is this rank's frame count (keypoints [n,cap,7] f32,
    descriptors [n,cap,32] u8, counts [n] i32 ...).  Returns, on rank `dst`, the list of [nframes, ...] tensors in
    global frame order; None elsewhere.  Shards are padded to the largest shard so one all_gather per slab suffices."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = frame_range(rank, world, nframes)
    for t in slabs:
        if t.shape[0] != hi - lo:
            raise ValueError("slab has %d frames, shard [%d,%d) expects %d" % (t.shape[0], lo, hi, hi - lo))
    if world == 1:
        return list(slabs)
    pad = max_shard(world, nframes)
    out = []
    for t in slabs:
        buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        buf[:hi - lo] = t
        allb = torch.empty((world * pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(allb, buf, group=group)
        if rank == dst:
            parts = []
            for r in range(world):
                a, b = frame_range(r, world, nframes)
                parts.append(allb[r * pad:r * pad + (b - a)])
            out.append(torch.cat(parts, 0))
    return out if rank == dst else None
