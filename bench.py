#!/usr/bin/env python
"""Benchmark of the hot path: one `VecTask.step` of the flat Anymal task at 4096 envs per GPU
(BASELINE.json configs[1]), synthetic random actions.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = env-steps/s with inputs resident in HBM (device actions, one fused
kernel launch per step, timed with CUDA events, L2 flushed between timed steps); `e2e` = the same metric through
the C-ABI host call (pinned host actions in, obs/reward/reset/time-outs out, copies inside the timed region).
`--impl reference` times the CPU restatement of the same step (oracle port; Isaac Gym/PhysX is a closed binary
that is not installed) on the box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 4096
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"
ALGO_BYTES_PER_ENV_STEP = 788      # boundary traffic of one Anymal env-step
WORKLOAD = "Anymal flat-terrain, 4096 envs/GPU, implicit PD position drive (Kp 85, Kd 2), dt 0.02 s x 2 sub-steps, random actions"
# other hot-path configs (--task): algorithmic bytes per env-step, env count, kernels per step, workload text
TASKS = {
    "Anymal": (788, 4096, "k_anymal_step", None, WORKLOAD),
    "Hound": (836, 4096, "k_anymal_step", None, "Hound flat-terrain, 4096 envs/GPU, implicit PD position drive, dt 0.02 s x 2 sub-steps, random actions"),
    "Cartpole": (88, 512, "k_cartpole_step", None, "Cartpole, 512 envs, effort control, dt 0.0166 s x 2 sub-steps, random actions"),
    "AnymalTerrain": (2256, 4096, "k_terrain_phys", {"env": {"terrain": {"terrainType": "trimesh"}}},
                      "AnymalTerrain rough heightfield (10 levels x 20 types, 1200x2000 int16) + 140-point height scan, 4096 envs/GPU, explicit PD "
                      "decimation 4 (+1 sim step), dt 0.005 s, observation noise on, random actions"),
    "HoundTerrain": (2256, 4096, "k_terrain_phys", None, "HoundTerrain (plane), 4096 envs/GPU, explicit PD decimation 4 (+1), dt 0.005 s, random actions"),
    "UsefulHound": (2696, 4096, "k_terrain_phys", None, "UsefulHound hound + 6-DOF arm (OSC), 4096 envs/GPU, 18 actions, 204 obs, explicit PD decimation 4 (+1), "
                    "dt 0.005 s, random actions"),
    "Houndarm": (356, 8192, "k_houndarm_step", None, "Houndarm fixed-base 6-DOF arm reach (OSC: two 6x6 inversions per step), 8192 envs/GPU, dt 0.01667 s x 2 sub-steps, "
                 "random actions"),
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def oracle_setup(n_envs, threads):
    import numpy as np
    from isaacgymenv_b200 import _abi
    from oracle.cpu_baseline import CpuAnymalStep
    from tests import kernel_checks as kc

    art = kc.load_robot("anymal")
    sp = kc.flat_params()
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    c = kc.anymal_cfg(art)
    return CpuAnymalStep(_abi.pack_model(art), sp, props, kc.cfg_dict(c, art.num_dofs), n_envs, threads=threads, dtype=np.float32)


def time_cpu(n_envs, steps, warmup, threads):
    import numpy as np

    cpu = oracle_setup(n_envs, threads)
    rng = np.random.default_rng(42)
    acts = [(2 * rng.random((n_envs, 12), dtype=np.float32) - 1) for _ in range(8)]
    for i in range(warmup):
        cpu.step(acts[i % 8])
    t0 = time.perf_counter()
    for i in range(steps):
        cpu.step(acts[i % 8])
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_envs = 1024     # bounded sample of the 4096-env batch (envs are independent: throughput is per env-step)
    value, per_step = time_cpu(n_envs, args.steps, args.warmup, threads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{n_envs} envs per step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n_envs} of 4096 envs x {args.steps} steps; CPU restatement of the step (oracle), NOT PhysX: Isaac Gym is a closed binary that is not installed"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    import isaacgymenv_b200
    from isaacgymenv_b200 import _lib

    lib = _lib.load()
    algo_bytes, envs_per_gpu, kernel_name, overrides, workload = TASKS[args.task]
    if args.num_envs > 0 and args.num_envs != envs_per_gpu:
        envs_per_gpu = args.num_envs
        workload += f" [scaling study: {envs_per_gpu} envs per GPU instead of the config's count]"
    env = isaacgymenv_b200.make(seed=42 + rank, task=args.task, num_envs=envs_per_gpu, sim_device=dev, rl_device=dev, headless=True, overrides=overrides)
    n, na = env.num_envs, env.num_actions
    is_terrain = hasattr(env, "common_step_counter")
    g = torch.Generator(device=dev).manual_seed(42 + rank)
    pool = [2.0 * torch.rand(n, na, device=dev, generator=g) - 1.0 for _ in range(16)]
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 192 MiB > 126 MB L2

    def tick():
        if is_terrain:      # the task's common_step_counter drives pushes and the noise stream
            env.common_step_counter += 1
            lib.b2g_task_terrain_set_step(env.sim.handle, int(env.common_step_counter))

    def step_dev(i):
        tick()
        _lib.check(lib.b2g_task_step(env.sim.handle, C.c_void_p(pool[i % 16].data_ptr()), sptr), "step")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step_dev(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- device-resident timing: per-step CUDA events, L2 flushed (untimed) between timed steps ----
    launches0 = lib.b2g_sim_launch_count(env.sim.handle)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(float(i))
        starts[i].record(stream)
        step_dev(i)
        stops[i].record(stream)
    barrier()
    launches = lib.b2g_sim_launch_count(env.sim.handle) - launches0
    cold_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    # ---- same K steps back to back, warm L2 (the steady state a learner sees) ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step_dev(i)
    e1.record(stream)
    barrier()
    warm_ms = e0.elapsed_time(e1)
    # ---- end to end: pinned host actions in, obs/rew/reset/time-outs out, through the C-ABI host call ----
    h_act = [p.cpu().pin_memory() for p in pool]
    # result buffers carved from ONE pinned allocation in the layout the library asks for
    offs, tot = (C.c_int64 * 4)(), C.c_int64()
    _lib.check(lib.b2g_task_host_layout(env.sim.handle, offs, C.byref(tot)), "host_layout")
    h_arena = torch.empty(tot.value, dtype=torch.uint8).pin_memory()
    h_obs = h_arena[offs[0]:offs[0] + n * env.num_obs * 4].view(torch.float32).view(n, env.num_obs)
    h_rew = h_arena[offs[1]:offs[1] + n * 4].view(torch.float32)
    h_reset = h_arena[offs[2]:offs[2] + n * 8].view(torch.int64)
    h_to = h_arena[offs[3]:offs[3] + n * 8].view(torch.int64)

    def step_host(i):
        tick()
        _lib.check(lib.b2g_task_step_host(env.sim.handle, C.c_void_p(h_act[i % 16].data_ptr()), C.c_void_p(h_obs.data_ptr()),
                                                 C.c_void_p(h_rew.data_ptr()), C.c_void_p(h_reset.data_ptr()), C.c_void_p(h_to.data_ptr()), sptr), "step_host")

    for i in range(3):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        step_host(i)
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()
    assert torch.isfinite(h_obs).all() and torch.isfinite(h_rew).all()

    times = torch.tensor([cold_ms, warm_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    cold_ms, warm_ms, e2e_ms = times.tolist()
    if rank == 0:
        total = world * n * args.steps
        value = total / (cold_ms * 1e-3)
        peak, peak_src = measured_peak()
        kernel_s = cold_ms * 1e-3 / args.steps
        achieved = algo_bytes * n / kernel_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tp):      # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu captures
            try:
                tj = json.load(open(tp))
                traffic = tj.get("k_anymal_step_dram_bytes_per_launch") if args.task == "Anymal" else tj.get("per_task", {}).get(args.task, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu = None
        if args.task == "Anymal":
            threads = os.cpu_count() or 1
            cpu_steps = 6
            cv, _ = time_cpu(1024, cpu_steps, 1, threads)
            cpu = {"value": cv, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"1024 of 4096 envs x {cpu_steps} steps; CPU restatement of the step (oracle), NOT PhysX (Isaac Gym not installed)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": cold_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "task": args.task, "envs_per_gpu": n, "l2": "flushed between timed steps (192 MiB fill, untimed)",
                           "timing": "per-step CUDA events on the launch stream, summed; max over ranks"},
                "value_warm_l2": total / (warm_ms * 1e-3), "ms_per_step_warm_l2": warm_ms / args.steps,
                "e2e": {"value": total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n * na * 4,
                        "d2h_bytes_per_step": n * env.num_obs * 4 + n * 4 + n * 8 + n * 8, "ms_per_step": e2e_ms / args.steps,
                        "path": "b2g_task_step_host (C ABI), one blocking call per step: pinned host actions (read in place by the kernel over PCIe) -> "
                                "obs/rew/reset/time_outs stored by the SMs into the caller's pinned buffer (b2g_task_host_layout; tail of the fused "
                                "step kernel for the flat tasks / of k_terrain_post for the rough-terrain tasks, k_mirror_host otherwise), completion by a published sequence word the host polls"
                                + (" [B2G_HOST_MIRROR=0: copy-engine D2H + stream sync]" if os.environ.get("B2G_HOST_MIRROR", "1")[:1] == "0" else "")},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                             "kernel": kernel_name, "peak_source": peak_src,
                             "note": "latency/issue-bound by construction: the per-launch working set (a few MB) is L2-resident; see profiles/ and DESIGN.md"},
                "cpu_baseline": cpu, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--num-envs", type=int, default=0, help="environments per GPU (default: the config's own count, 4096 for Anymal); "
                    "other values are a scaling study, not the headline config")
    ap.add_argument("--task", default="Anymal", choices=sorted(TASKS), help="hot-path config to time (default: the headline Anymal config)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 50:
            args.steps = 50      # CPU arm: bounded so the run ends within minutes
        args.warmup = min(args.warmup, 3)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
