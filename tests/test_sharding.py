"""Multi-GPU host logic on CPU: world_size-2 gloo run of the frame sharding + final gather.
Each rank extracts its contiguous frame range (with the CPU oracle standing in for the device, this being a test
of the sharding plumbing) and rank 0 must end up with exactly the single-rank result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sdslam_b200 import sharding, synth

PARAMS = (120, 1.2, 3, 20)
W, H, NF = 128, 96, 5


def test_frame_range_partitions():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 5, 8, 4096, 4099):
            r = [sharding.frame_range(k, world, n) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == sharding.max_shard(world, n)
    with pytest.raises(ValueError):
        sharding.frame_range(2, 2, 10)


def test_c_abi_shard_range_equals_python():
    """sdorb_shard_range (the split sdorb_extract_batch_multi uses) == sharding.frame_range; needs no GPU."""
    from sdslam_b200 import api
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 5, 8, 4096, 4099):
            for r in range(world):
                assert api.shard_range(r, world, n) == sharding.frame_range(r, world, n)
    with pytest.raises(api.SdorbError):
        api.shard_range(2, 2, 10)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _extract(imgs):
    from oracle import binding as orc
    k, d, c = orc.Extractor(*PARAMS).extract_many(imgs, nthreads=1)
    return (torch.from_numpy(k.view(np.float32).reshape(len(imgs), -1, 7).copy()), torch.from_numpy(d.copy()),
            torch.from_numpy(c.copy()))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        imgs = synth.frames(NF, W, H)
        lo, hi = sharding.frame_range(rank, world, NF)
        got = sharding.gather_slabs(list(_extract(imgs[lo:hi])), NF)
        if rank == 0:
            q.put([g.numpy() for g in got])
        else:
            assert got is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_equals_single_rank():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = [t.numpy() for t in _extract(synth.frames(NF, W, H))]
    assert len(got) == 3
    for g, r in zip(got, ref):
        assert g.dtype == r.dtype and g.shape == r.shape and g.tobytes() == r.tobytes()
    assert int(ref[2].sum()) > 0


def test_gather_single_process_passthrough():
    a = torch.arange(12).reshape(4, 3)
    out = sharding.gather_slabs([a], 4)
    assert out[0] is a
    with pytest.raises(ValueError):
        sharding.gather_slabs([a], 5)


# ------------------------------------------------------------------ frame PAIRS (matching / guided search) shard the same way
NPAIRS, NDESC = 5, 40


def _pair_inputs():
    rng = np.random.default_rng(12)
    A = rng.integers(0, 256, (NPAIRS, NDESC, 32), dtype=np.uint8)
    B = rng.integers(0, 256, (NPAIRS, NDESC, 32), dtype=np.uint8)
    B[:, ::3] = A[:, ::3] ^ 1  # every third row one bit away from its partner
    return A, B


def _match(A, B):
    from oracle import binding as orc
    out = np.zeros((len(A), NDESC, 4), np.int32)
    for p in range(len(A)):
        out[p] = orc.match_best2(A[p], B[p]).view(np.int32).reshape(NDESC, 4)
    return torch.from_numpy(out)


def _pair_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, B = _pair_inputs()
        lo, hi = sharding.frame_range(rank, world, NPAIRS)
        got = sharding.gather_slabs([_match(A[lo:hi], B[lo:hi])], NPAIRS)
        if rank == 0:
            q.put(got[0].numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_pair_sharding_equals_single_rank():
    """Frame pairs (ORBmatcher::DescriptorDistance best / second-best per pair) over two ranks with an odd pair count: rank 0
    ends up with the single-rank result."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pair_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = _match(*_pair_inputs()).numpy()
    assert got.shape == ref.shape and got.tobytes() == ref.tobytes()
    assert int(ref[..., 3].sum()) > 0  # some matches accepted
