"""Multi-GPU plumbing for the env step: one process per GPU, environments sharded by rank, NO collective in the step
(environments are independent -- SURVEY.md 8(e); reference: ``utils/rlgames_utils.py:89-107``, ``utils/utils.py:89-94``).
``torch.distributed`` is used only to agree on timings / throughput (and, in a learner, for the gradient all-reduce).
"""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass
class RankInfo:
    rank: int
    local_rank: int
    world_size: int

    @property
    def device(self) -> str:
        return f"cuda:{self.local_rank}"


def rank_info() -> RankInfo:
    """RANK / LOCAL_RANK / WORLD_SIZE as set by torchrun (defaults: single process)."""
    return RankInfo(int(os.getenv("RANK", "0")), int(os.getenv("LOCAL_RANK", "0")), int(os.getenv("WORLD_SIZE", "1")))


def shard_seed(seed: int, rank: int) -> int:
    """Per-rank seed offset, as the reference's ``set_seed(seed, rank=...)`` does (``utils/utils.py:89-94``)."""
    return int(seed) + int(rank)


def env_range(total_envs: int, rank: int, world_size: int):
    """Contiguous block of global environment ids owned by ``rank`` (weak scaling uses total = per_gpu * world)."""
    base, rem = divmod(int(total_envs), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def reduce_max(values, device=None):
    """MAX over ranks of a list of floats (timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def aggregate_env_steps_per_sec(envs_per_rank: int, steps: int, elapsed_s_max: float, world_size: int) -> float:
    """Whole-job throughput: every rank's env-steps divided by the slowest rank's time."""
    return float(world_size) * envs_per_rank * steps / elapsed_s_max


def shutdown(ppo=None, grace_s: float = 30.0):
    """Leave a multi-rank run: release captured graphs, agree that everybody is done, destroy the process group -- and if the teardown
    does not return within ``grace_s`` (seen with graphs that captured NCCL work), exit the process: all results are out by now."""
    import os
    import sys
    import threading

    import torch.distributed as dist

    if ppo is not None and hasattr(ppo, "release_graphs"):
        ppo.release_graphs()
    if not (dist.is_available() and dist.is_initialized()):
        return
    dist.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    t = threading.Timer(grace_s, lambda: os._exit(0))
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()
