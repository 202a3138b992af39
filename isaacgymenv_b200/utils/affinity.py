"""CPU placement of a rank next to its GPU.

`b2g_task_step_host` completes a step by polling a page-locked word the GPU writes (no stream sync), so the calling thread
spins on a host core for the length of the step.  With one process per GPU (reference: utils/rlgames_utils.py:89-107, one
rank per device under torchrun) the ranks must not share cores, and each should sit on the NUMA node its GPU's PCIe root
hangs off.  `pin_to_gpu_numa` gives every local rank a disjoint slice of the CPUs NVML reports as local to its GPU."""
from __future__ import annotations

import os


def _gpu_cpus(index: int):
    """CPUs local to GPU `index` (NVML's ideal affinity), intersected with what this process may use."""
    allowed = sorted(os.sched_getaffinity(0))
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (max(allowed) // 64) + 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        cpus = [c for c in cpus if c in set(allowed)]
        if cpus:
            return cpus, "nvml"
    except Exception:
        pass
    return allowed, "sched_getaffinity"


def pin_to_gpu_numa(local_rank: int, local_world: int, physical_index: int | None = None):
    """Restrict this process to its share of the CPUs local to its GPU.  Ranks whose GPUs share a CPU set split it evenly in
    local-rank order.  Returns a small description (goes into bench.py's JSON line); never raises."""
    try:
        idx = physical_index
        if idx is None:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip() != ""]
            idx = int(ids[local_rank]) if ids and ids[local_rank].isdigit() else local_rank
        cpus, src = _gpu_cpus(idx)
        lw = max(int(local_world), 1)
        per = max(len(cpus) // lw, 1)
        mine = cpus[(local_rank % lw) * per:(local_rank % lw) * per + per] or cpus
        os.sched_setaffinity(0, mine)
        return {"cpus": f"{mine[0]}-{mine[-1]}" if mine == list(range(mine[0], mine[-1] + 1)) else ",".join(map(str, mine)), "count": len(mine),
                "source": src}
    except Exception as exc:      # placement is an optimisation, not a requirement
        return {"error": f"{type(exc).__name__}: {exc}"[:120]}
