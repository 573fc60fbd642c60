"""CPU, world_size 2 over gloo: the clip-sharding host logic and the final gather (no kernels involved)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "face-diffusion-model_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fdm_b200.parallel import sample_sharded
    from oracle.philox_ref import philox_normal

    def run_local(first, count):
        # stand-in for the per-rank sampling job: noise keyed by the GLOBAL clip index
        return torch.stack([torch.from_numpy(philox_normal(7, first + i, 3, 64)).view(4, 16) for i in range(count)])

    out = sample_sharded(run_local, n_clips)
    q.put((rank, out.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sampling_is_world_size_invariant():
    from oracle.philox_ref import philox_normal
    n_clips, world = 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = np.stack([philox_normal(7, c, 3, 64).reshape(4, 16) for c in range(n_clips)])
    for r in range(world):
        assert got[r].shape == (n_clips, 4, 16)
        assert np.array_equal(got[r], ref)  # every rank holds every clip, identical to the unsharded run


def test_shard_range():
    from fdm_b200.parallel import shard_range
    assert [shard_range(64, r, 8) for r in range(8)] == [(8 * r, 8) for r in range(8)]
    assert shard_range(5, 0, 1) == (0, 5)
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)
