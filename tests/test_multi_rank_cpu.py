"""world_size-2 gloo test of the N>1 path: scenes shard by rank with no data-path collective;
only the timing scalars are reduced.  Runs the same sharding/reduction helpers bench.py uses,
with the CPU oracle standing in for the per-rank step (no GPU here)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hope_b200.batched_env import generate_scenes
    from oracle import parking_oracle as po
    import bench
    n = 64
    sc = generate_scenes(2 * n, "mix", bench.scene_seed(rank))       # rank-private scene pool
    env = po.OracleEnv(*[sc[k][:n] for k in ("start", "dest", "bounds", "obs", "nverts")], nthreads=1)
    env.reset_step()
    rng = np.random.default_rng(rank)
    steps = 0
    for _ in range(3):
        env.step(rng.uniform(-1, 1, size=(n, 2)))
        steps += n
    ms_local = 10.0 + rank                                           # pretend device time
    ms = bench.reduce_scalar(ms_local, "max", world, torch.device("cpu"))
    total = bench.reduce_scalar(float(steps), "sum", world, torch.device("cpu"))
    digest = float(np.abs(env.pose).sum())
    q.put((rank, ms, total, digest, float(sc["start"][0, 0])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_scenes_and_reduce_timings():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ms0, tot0, dig0, s0), (r1, ms1, tot1, dig1, s1) = res
    assert ms0 == ms1 == 11.0            # max over ranks
    assert tot0 == tot1 == 2 * 3 * 64    # whole-job env-steps
    assert s0 != s1 and dig0 != dig1     # each rank owns different scenes


def _grad_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hope_b200 import learner
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))
    red = learner.FlatGradAllReduce([net], world)
    x = torch.full((4, 6), float(rank + 1))
    net(x).sum().backward()
    local = torch.cat([p.grad.reshape(-1).clone() for p in net.parameters()])
    red.launch(); red.wait()
    avg = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    q.put((rank, local.numpy(), avg.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_averages_over_ranks():
    """BASELINE cfg 5's only data exchange: one all-reduce of the flattened gradients (gloo stands in for NCCL)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (res[0][1] + res[1][1]) / 2
    assert np.abs(res[0][1] - res[1][1]).max() > 1e-3       # the local gradients really differ
    assert np.allclose(res[0][2], want, atol=1e-6) and np.allclose(res[1][2], want, atol=1e-6)
