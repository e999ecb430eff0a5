"""world_size-2 gloo tests of the multi-GPU host logic (weight broadcast + tile sharding), on CPU."""

import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from robosat_b200 import synth
    from robosat_b200.dist import broadcast_state_dict, shard_range, unet_state_template

    ref = synth.make_state_dict(2, seed=0)
    sd = ref if rank == 0 else None
    got = broadcast_state_dict(sd, unet_state_template(2), device="cpu")
    same = list(got.keys()) == list(ref.keys()) and all(torch.equal(got[k], ref[k]) and got[k].dtype == ref[k].dtype for k in ref)
    q.put((rank, same, shard_range(101, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_state_dict_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res)
    assert res[0][2] == (0, 51) and res[1][2] == (51, 101)


def test_shard_range_covers_everything():
    from robosat_b200.dist import shard_range

    for n in (0, 1, 7, 100000):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _ar_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from robosat_b200.dist import allreduce_sum_

    flat = torch.arange(10, dtype=torch.float32) * (rank + 1) / world  # gradient of the loss pre-scaled by 1/world
    allreduce_sum_(flat, world)
    q.put((rank, flat.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_is_the_global_mean_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_ar_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    expect = [(i * 1 + i * 2) / 2 for i in range(10)]  # mean over ranks of rank-local gradients i*(rank+1)
    assert res[0][1] == expect and res[1][1] == expect
