"""world_size-2 gloo test (CPU) of the multi-GPU path's host logic: frames shard round-robin, every rank
runs its own frames with no collective, rank 0 gathers the per-frame streams in frame order.  The per-frame
transform here is the CPU oracle (the GPU tests cover the kernels); the result must equal the serial encode."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_FRAMES = 7


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import t3oracle as T
    from ternary_image_codec_b200 import sharding
    oracle = T.Oracle()
    cfg = T.make_cfg(profile=T.P3, uep=2)
    sizes = [1000 + 37 * f for f in range(N_FRAMES)]  # frames of different sizes -> streams of different lengths

    def process(f):
        return torch.from_numpy(oracle.encode_rgb(cfg, T.synth_rgb(5 + f, sizes[f]), 1).reshape(-1).copy())

    local = {f: process(f) for f in sharding.frames_for_rank(N_FRAMES, rank, world)}
    out = sharding.gather_in_frame_order(local, N_FRAMES, rank, world, dst=0)
    if rank == 0:
        ok = all(np.array_equal(out[f].numpy(), oracle.encode_rgb(cfg, T.synth_rgb(5 + f, sizes[f]), 1).reshape(-1)) for f in range(N_FRAMES))
        ret.put(bool(ok))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_assignment():
    from ternary_image_codec_b200 import sharding
    for world in (1, 2, 4, 8):
        seen = sorted(f for r in range(world) for f in sharding.frames_for_rank(240, r, world))
        assert seen == list(range(240))
        assert all(sharding.owner_of(f, world) == r for r in range(world) for f in sharding.frames_for_rank(240, r, world))
        assert max(len(sharding.frames_for_rank(240, r, world)) for r in range(world)) == 240 // world


def test_two_rank_gloo_gather_matches_serial():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert ret.get(timeout=10) is True
