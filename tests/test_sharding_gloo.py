"""N>1 path on CPU: two gloo ranks each accumulate their shard of the sample indices (the oracle
stands in for the device renderer -- this test is about the host-side sharding + reduce logic), the
reduce puts the sum on rank 0, and the result equals the single-process accumulation."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, SPP, STEPS, DEPTH, SEED = 24, 14, 2, 2, 4, 3


def _render_shard(rank, world):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oraclelib as O
    from vanrijn_b200 import scenes, sharding
    orc = O.OracleScene(scenes.scene_main(subdivisions=1, obj=False))
    s = np.zeros(W * H * 3)
    w = np.zeros(W * H)
    for step in range(STEPS):
        for idx in sharding.shard_sample_indices(rank, world, step, SPP):
            r = orc.render((0, W, 0, H), H, W, spp=1, max_depth=DEPTH, seed=SEED, sample_offset=idx, threads=1)
            s += r["colour_sum"]
            w += r["weight"]
    return s, w


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from vanrijn_b200 import sharding
    s, w = _render_shard(rank, world)
    ts, tw = torch.from_numpy(s), torch.from_numpy(w)
    sharding.reduce_accumulation(ts, tw, dst=0)
    if rank == 0:
        np.savez(out_path, s=ts.numpy(), w=tw.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sample_sharding_equals_single_process(tmp_path):
    out = str(tmp_path / "reduced.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    s1, w1 = _single_total()
    assert np.array_equal(got["w"], w1) and np.all(w1 == 2 * SPP * STEPS)
    np.testing.assert_allclose(got["s"], s1, rtol=1e-12, atol=1e-25)


def _single_total():
    """All 2*SPP*STEPS sample indices in one process."""
    sys.path.insert(0, ROOT)
    import oraclelib as O
    from vanrijn_b200 import scenes
    orc = O.OracleScene(scenes.scene_main(subdivisions=1, obj=False))
    s = np.zeros(W * H * 3)
    w = np.zeros(W * H)
    for idx in range(2 * SPP * STEPS):
        r = orc.render((0, W, 0, H), H, W, spp=1, max_depth=DEPTH, seed=SEED, sample_offset=idx, threads=1)
        s += r["colour_sum"]
        w += r["weight"]
    return s, w
