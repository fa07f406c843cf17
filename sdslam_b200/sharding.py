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
    global frame order; None elsewhere.  A gather TO ONE RANK: every other rank sends its slab, rank `dst` receives each one
    straight into its place in the result (point-to-point, one batch of isend / irecv) -- no padding, no copy, and nobody but
    `dst` receives anything."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = frame_range(rank, world, nframes)
    for t in slabs:
        if t.shape[0] != hi - lo:
            raise ValueError("slab has %d frames, shard [%d,%d) expects %d" % (t.shape[0], lo, hi, hi - lo))
    if world == 1:
        return list(slabs)
    ops, out = [], None
    if rank == dst:
        out = [torch.empty((nframes,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for t in slabs]
        for o, t in zip(out, slabs):
            o[lo:hi].copy_(t)
            for r in range(world):
                a, b = frame_range(r, world, nframes)
                if r != dst and b > a:
                    ops.append(dist.P2POp(dist.irecv, o[a:b], r, group))
    elif hi > lo:
        for t in slabs:
            ops.append(dist.P2POp(dist.isend, t.contiguous(), dst, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out
