"""The N>1 path on CPU: two gloo ranks shard the environments (no data-path collective), time their own work and
agree on max-over-ranks timing / whole-job throughput, exactly the plumbing bench.py uses under torchrun."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from isaacgymenv_b200 import distributed as D

    info = D.rank_info()
    lo, hi = D.env_range(64 * world, info.rank, info.world_size)
    seed = D.shard_seed(42, info.rank)
    g = torch.Generator().manual_seed(seed)
    actions = 2 * torch.rand(hi - lo, 12, generator=g) - 1          # each rank draws only its own shard's actions
    elapsed = 0.010 * (rank + 1)                                     # pretend rank 1 is twice as slow
    t_max = D.reduce_max([elapsed])[0]
    value = D.aggregate_env_steps_per_sec(hi - lo, 100, t_max, world)
    # shards are disjoint and cover everything; no rank saw another rank's environments
    counts = torch.tensor([float(hi - lo)])
    dist.all_reduce(counts)
    out.put((rank, lo, hi, seed, t_max, value, float(counts[0]), float(actions.sum())))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, s0, t0, v0, c0, a0), (r1, lo1, hi1, s1, t1, v1, c1, a1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 64, 64, 128) and (s0, s1) == (42, 43)
    assert t0 == t1 == 0.020                          # max over ranks
    assert v0 == v1 == 2 * 64 * 100 / 0.020           # whole-job aggregate
    assert c0 == c1 == 128.0 and a0 != a1


def _ppo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from isaacgymenv_b200.learning.ppo import PPO, PPOConfig
    from tests.test_ppo import PointEnv

    env = PointEnv(n=64)
    env._g.manual_seed(100 + rank)                     # each rank sees different experience
    env.pos = torch.randn(64, 2, generator=env._g)
    cfg = PPOConfig(horizon_length=10, minibatch_size=320, mini_epochs=2, units=(16, 16, 8), learning_rate=1e-3)
    ppo = PPO(env, cfg, multi_gpu=True, seed=7 + rank)
    ppo.train(max_epochs=3, log_every=1)
    flat = torch.cat([p.detach().reshape(-1) for p in ppo.model.parameters()])
    out.put((rank, flat.double().sum().item(), flat.abs().double().sum().item(), ppo.obs_rms.mean.clone().numpy().tolist(),
             ppo.obs_rms.var.clone().numpy().tolist(), float(ppo.val_rms.mean), ppo.lr))
    dist.destroy_process_group()


def test_two_rank_ppo_keeps_ranks_identical():
    """The learner's only exchange step: gradients are all-reduced (averaged) once per minibatch, the observation / value normalisers
    are pooled after every rollout, the adaptive learning rate follows the all-reduced KL -- so two ranks that see different experience
    end with identical weights, normalisers and learning rate (reference: rl_games multi_gpu under torchrun, README.md:165-172)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ppo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a, b = res
    assert a[1] == b[1] and a[2] == b[2], "weights diverged between ranks"
    assert a[3] == b[3] and a[4] == b[4] and a[5] == b[5], "normalisers diverged between ranks"
    assert a[6] == b[6]
