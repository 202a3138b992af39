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
import gc
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
    "Manipulator": (376, 8192, "k_houndarm_step", None, "Manipulator fixed-base 7-DOF Franka arm reach (OSC: a 7x7 and a 6x6 inversion per step), 8192 envs/GPU, "
                    "dt 0.01667 s x 2 sub-steps, random actions"),
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle-reason samples during the timed region: NVML in a thread of this process (one light query per GPU every
    20 ms -- a per-rank `nvidia-smi -lms` subprocess takes driver locks the ranks' launches then wait on), nvidia-smi as the
    fallback.  Rank 0 samples every GPU of the job."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, indices):
        self.indices = list(indices) if isinstance(indices, (list, tuple, range)) else [indices]
        self.rows, self.proc, self.stop_flag, self.thread, self.mode = [], None, False, None, None
        self.sm, self.mx, self.reasons = [], [], set()

    def _nvml_loop(self, nv, handles):
        while not self.stop_flag:
            for h in handles:
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                    get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                    bits = int(get(h))
                    for bit, name in self.REASONS:
                        if bits & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.02)

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip().isdigit()]
            handles = [nv.nvmlDeviceGetHandleByIndex(int(vis[i]) if i < len(vis) else i) for i in self.indices]
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handles), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.mode = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(map(str, self.indices)), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.thread:
            time.sleep(0.06)
            self.stop_flag = True
            self.thread.join(timeout=1.0)
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = list(self.sm), list(self.mx), set(self.reasons)
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.mode}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm), "source": self.mode,
                "gpus": len(self.indices)}


def oracle_setup(n_envs, threads):
    from isaacgymenv_b200 import _abi
    from oracle.cpu_baseline import CpuAnymalStepNative
    from tests import kernel_checks as kc

    art = kc.load_robot("anymal")
    sp = kc.flat_params()
    props = _abi.default_dof_props(art, _abi.DOF_MODE_POS, 85.0, 2.0)
    return CpuAnymalStepNative(_abi.pack_model(art), sp, props, kc.anymal_cfg(art), n_envs, threads=threads)


def time_cpu(n_envs, steps, warmup, threads):
    """The CPU restatement of the whole flat-task step (C, OpenMP over the environments, -O3 -march=native built on this
    machine) on `threads` host threads: returns (env-steps/s, s per step, compiler flags)."""
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
    return n_envs * steps / dt, dt / steps, cpu.flags


def bench_config(task, n):
    """`config` of the JSON line: identical for the two arms (the driver compares them)."""
    return {"workload": TASKS[task][4], "task": task, "envs_per_gpu": n}


CPU_PREROLL = 10      # untimed CPU steps so the robots have landed (contact steady state), on top of --warmup


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0)) or 1
    n_envs = ENVS_PER_GPU      # the whole 4096-env batch of the headline config
    value, per_step, flags = time_cpu(n_envs, args.steps, args.warmup + CPU_PREROLL, threads)
    sample = (f"all {n_envs} envs x {args.steps} steps after {args.warmup} warm-up + {CPU_PREROLL} pre-roll steps; CPU restatement of the whole step "
              f"(oracle port in C, OpenMP over envs, gcc {flags}), NOT PhysX: Isaac Gym is a closed binary that is not installed")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config("Anymal", n_envs),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def measure_task(task, num_envs, steps, warmup, preroll, dev, rank, world, dist, sample_clocks=True):
    """One hot-path config on this rank's GPU: pre-roll (untimed, so the robots have landed and the contact steady state is what
    gets timed), warm-up, K timed steps three ways (L2 flushed / warm L2 / end to end through host buffers).  Times are maxima
    over ranks."""
    import torch

    import isaacgymenv_b200
    from isaacgymenv_b200 import _lib

    lib = _lib.load()
    algo_bytes, envs_per_gpu, kernel_name, overrides, workload = TASKS[task]
    if num_envs > 0 and num_envs != envs_per_gpu:
        envs_per_gpu = num_envs
        workload += f" [scaling study: {envs_per_gpu} envs per GPU instead of the config's count]"
    local_rank = int(dev.split(":")[1])
    env = isaacgymenv_b200.make(seed=42 + rank, task=task, num_envs=envs_per_gpu, sim_device=dev, rl_device=dev, headless=True, overrides=overrides)
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

    # clocks are sampled from the pre-roll to the end of the e2e loop (the same kernels under the same load as the timed steps)
    sampler = ClockSampler(list(range(world)) if world > 1 else local_rank) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    for i in range(preroll):
        step_dev(i)
    for i in range(max(warmup, 3)):
        step_dev(i)
    barrier()
    # ---- device-resident timing: per-step CUDA events, L2 flushed (untimed) between timed steps ----
    gc.collect()
    gc.disable()      # a collector pause between an event record and the launch behind it would be timed (re-enabled after the e2e loop)
    launches0 = lib.b2g_sim_launch_count(env.sim.handle)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    barrier()
    for i in range(steps):
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
    for i in range(steps):
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
    # the interpreter's cycle collector can stop this thread for milliseconds (torch's module graph is large); a 20-step region is 1.3 ms
    gc.collect()
    gc.disable()
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(steps):
        step_host(i)
    # b2g_task_step_host blocks until the results are in the caller's host buffers: the K steps end HERE on this rank.  The closing
    # barrier stays outside the interval (an NCCL barrier costs a noticeable share of a 20-step, 1.2 ms region); max over ranks below
    wall_ms = (time.perf_counter() - t0) * 1e3
    e1.record(stream)
    gc.enable()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), wall_ms)
    clocks = sampler.stop() if sampler else None
    assert torch.isfinite(h_obs).all() and torch.isfinite(h_rew).all()
    stats = None
    if hasattr(lib, "b2g_sim_contact_stats"):
        st = (C.c_int64 * 4)()
        if lib.b2g_sim_contact_stats(env.sim.handle, st, 0) == 0:
            stats = {"active_contacts": int(st[0]), "dropped_candidates": int(st[1]), "env_substeps_with_drop": int(st[2]), "env_substeps": int(st[3])}

    times = torch.tensor([cold_ms, warm_ms, e2e_ms], dtype=torch.float64, device=dev)
    e2e_ranks = [e2e_ms]
    if world > 1:
        allr = [torch.zeros_like(times) for _ in range(world)]
        dist.all_gather(allr, times)
        e2e_ranks = [float(t[2]) for t in allr]
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    cold_ms, warm_ms, e2e_ms = times.tolist()
    res = dict(task=task, n=n, na=na, num_obs=env.num_obs, algo_bytes=algo_bytes, kernel=kernel_name, workload=workload, cold_ms=cold_ms, warm_ms=warm_ms,
               e2e_ms=e2e_ms, e2e_ranks_ms=e2e_ranks, launches=int(launches), clocks=clocks, contact_stats=stats)
    del env, flush, pool, h_arena
    torch.cuda.empty_cache()
    return res


def measure_ppo(dev, rank, world, dist, num_envs=8192, epochs=10, warm=2, fused_update=True):
    """BASELINE.json config 5: Anymal, 8192 envs per GPU, PPO per cfg/train/AnymalPPO.yaml (horizon 24, minibatch 32768, 5 mini-epochs
    -> 6 x 5 = 30 gradient all-reduces per iteration over NCCL), one process per GPU, seeds 42 + rank (reference README.md:165-172,
    utils/utils.py:89-94).  Rollout (tcgen05 policy kernel + fused env step) and minibatch update (with its all-reduce) replay from
    CUDA graphs.  Returns env-steps/s INCLUDING the learner, and the stand-alone cost of one iteration's all-reduces."""
    import torch

    import isaacgymenv_b200
    from isaacgymenv_b200.learning.ppo import PPO
    from isaacgymenv_b200.train import load_train_config, ppo_config_from_train_cfg

    env = isaacgymenv_b200.make(seed=42 + rank, task="Anymal", num_envs=num_envs, sim_device=dev, rl_device=dev, headless=True)
    cfg = ppo_config_from_train_cfg(load_train_config("AnymalPPO"))
    cfg.tf32 = True
    ppo = PPO(env, cfg, multi_gpu=world > 1, seed=42 + rank, fused_rollout=True, cuda_graphs=True, fused_update=fused_update)
    ppo.train(max_epochs=warm, log_every=10 ** 9)          # graph capture + warm-up epochs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    log = ppo.train(max_epochs=epochs, log_every=10 ** 9)
    e1.record()
    barrier()
    ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    T, N = cfg.horizon_length, env.num_envs
    n_mb = max((T * N) // min(cfg.minibatch_size, T * N), 1)
    n_ar = cfg.mini_epochs * n_mb
    nparam = sum(p.numel() for p in ppo.model.parameters())
    nccl_ms = 0.0
    if world > 1:
        buf = torch.zeros(nparam, device=dev)
        for _ in range(5):
            dist.all_reduce(buf)
        barrier()
        e0.record()
        for _ in range(n_ar * 10):
            dist.all_reduce(buf)
        e1.record()
        barrier()
        nccl_ms = e0.elapsed_time(e1) / 10
    t = torch.tensor([ms, nccl_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, nccl_ms = t.tolist()
    per_iter = ms / epochs
    out = {"workload": f"Anymal PPO (cfg/train/AnymalPPO.yaml), {N} envs/GPU, horizon {T}, minibatch {cfg.minibatch_size}, {cfg.mini_epochs} mini-epochs",
           "envs_per_gpu": N, "n_gpus": world, "iterations": epochs, "ms_per_iteration": per_iter,
           "env_steps_per_sec_incl_learner": world * N * T * epochs / (ms * 1e-3), "unit": UNIT,
           "allreduces_per_iteration": n_ar if world > 1 else 0, "allreduce_bytes": nparam * 4,
           "nccl_ms_per_iteration_standalone": nccl_ms, "nccl_share_of_iteration": (nccl_ms / per_iter) if per_iter > 0 else None,
           "fused_update_kernels": bool(fused_update), "update_in_cuda_graph": ppo._g_update is not None, "rollout_in_cuda_graph": ppo._g_rollout is not None,
           "update_capture_error": ppo.update_capture_error, "mean_episode_reward_last": (log.mean_episode_reward[-1] if log.mean_episode_reward else None)}
    del ppo, env
    torch.cuda.empty_cache()
    return out


def traffic_for(task):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the committed `ncu --set full` captures
    (profiles/traffic.json; the capture is cold-cache like the flushed timing)."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        tj = json.load(open(tp))
        return tj.get("k_anymal_step_dram_bytes_per_launch") if task == "Anymal" else tj.get("per_task", {}).get(task, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def roofline_of(res, steps):
    peak, peak_src = measured_peak()
    kernel_s = res["cold_ms"] * 1e-3 / steps
    achieved = res["algo_bytes"] * res["n"] / kernel_s / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic_for(res["task"]),
            "kernel": res["kernel"], "peak_source": peak_src, "algo_bytes_per_env_step": res["algo_bytes"],
            "note": "latency/issue-bound by construction: the per-launch working set (a few MB) is L2-resident; see profiles/ and DESIGN.md"}


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
    from isaacgymenv_b200.utils import affinity

    pin = affinity.pin_to_gpu_numa(local_rank, world)      # the e2e path polls host memory: keep this rank on its GPU's NUMA node
    res = measure_task(args.task, args.num_envs, args.steps, args.warmup, args.preroll, dev, rank, world, dist)
    n, na = res["n"], res["na"]
    # the other BASELINE configs ride in the same line (shorter runs, one GPU only: they are parity-test configs, not the headline)
    others = {}
    if args.other_configs and world == 1 and args.task == "Anymal" and args.num_envs <= 0:
        for t in ("AnymalTerrain", "UsefulHound", "Cartpole", "Manipulator"):
            try:
                k = min(args.steps, 200)
                r = measure_task(t, 0, k, min(args.warmup, 20), args.preroll, dev, rank, world, dist, sample_clocks=False)
                tot = r["n"] * k
                others[t] = {"workload": r["workload"], "envs": r["n"], "steps": k, "ms_per_step": r["cold_ms"] / k, "value": tot / (r["cold_ms"] * 1e-3),
                             "value_warm_l2": tot / (r["warm_ms"] * 1e-3), "e2e_value": tot / (r["e2e_ms"] * 1e-3), "unit": UNIT,
                             "gpu_launches": r["launches"], "roofline": roofline_of(r, k), "contact_stats": r["contact_stats"]}
            except Exception as exc:      # never lose the headline line to a side config
                others[t] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    line = None
    if rank == 0:
        total = world * n * args.steps
        cold_ms, warm_ms, e2e_ms = res["cold_ms"], res["warm_ms"], res["e2e_ms"]
        value = total / (cold_ms * 1e-3)
        cpu = None
        if args.task == "Anymal" and world == 1:      # CPU baseline beside it: rank 0 at N = 1 only
            threads = len(os.sched_getaffinity(0)) or 1
            cpu_steps = 40
            cv, _, flags = time_cpu(ENVS_PER_GPU, cpu_steps, CPU_PREROLL, threads)
            cpu = {"value": cv, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"all {ENVS_PER_GPU} envs x {cpu_steps} steps after {CPU_PREROLL} pre-roll steps; CPU restatement of the whole step (oracle port in C, "
                             f"OpenMP over envs, gcc {flags}), NOT PhysX (Isaac Gym not installed)"}
        cfg = bench_config(args.task, n)
        if res["workload"] != cfg["workload"]:
            cfg["workload"] = res["workload"]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": cold_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg, "preroll_steps": args.preroll,
                "l2": "flushed between timed steps (192 MiB fill, untimed)", "timing": "per-step CUDA events on the launch stream, summed; max over ranks",
                "value_warm_l2": total / (warm_ms * 1e-3), "ms_per_step_warm_l2": warm_ms / args.steps,
                "e2e": {"value": total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n * na * 4,
                        "d2h_bytes_per_step": n * res["num_obs"] * 4 + n * 4 + n * 8 + n * 8, "ms_per_step": e2e_ms / args.steps,
                        "ms_per_step_by_rank": [round(t / args.steps, 5) for t in res.get("e2e_ranks_ms", [e2e_ms])],
                        "path": "b2g_task_step_host (C ABI), one blocking call per step: pinned host actions (read in place by the kernel over PCIe) -> "
                                "obs/rew/reset/time_outs stored by the SMs into the caller's pinned buffer (b2g_task_host_layout; tail of the fused "
                                "step kernel for the flat tasks / of k_terrain_post for the rough-terrain tasks, k_mirror_host otherwise), completion by a published sequence word the host polls"
                                + (" [B2G_HOST_MIRROR=0: copy-engine D2H + stream sync]" if os.environ.get("B2G_HOST_MIRROR", "1")[:1] == "0" else ""),
                        "cpu_affinity": pin},
                "gpu_launches": res["launches"],
                "roofline": roofline_of(res, args.steps),
                "cpu_baseline": cpu, "clocks": res["clocks"], "contact_stats": res["contact_stats"]}
        if others:
            line["other_configs"] = others
    # ---- config 5 rides in the same line at every N.  The headline line above is complete before it starts, and a watchdog prints that
    # line and leaves if the learner leg does not come back (a stuck collective must never cost the measurement) ----
    if args.ppo and args.task == "Anymal" and args.num_envs <= 0:
        finished = threading.Event()

        def watchdog():
            if not finished.wait(timeout=float(args.ppo_timeout)):
                if rank == 0:
                    line["ppo_config5"] = {"error": f"learner leg did not finish within {args.ppo_timeout} s"}
                    print(json.dumps(line), flush=True)
                os._exit(0)

        threading.Thread(target=watchdog, daemon=True).start()
        try:
            ppo = measure_ppo(dev, rank, world, dist)
        except Exception as exc:
            ppo = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        finished.set()
        if rank == 0:
            line["ppo_config5"] = ppo
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # no destroy_process_group(): tearing a communicator down while CUDA graphs that captured its all-reduces are still alive has
        # been seen to hang (2 x B200, torch 2.11 / NCCL 2.28); the ranks agree that everybody is done and leave
        import gc

        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--num-envs", type=int, default=0, help="environments per GPU (default: the config's own count, 4096 for Anymal); "
                    "other values are a scaling study, not the headline config")
    ap.add_argument("--preroll", type=int, default=300, help="untimed steps before the warm-up (robots start above the ground and land around step 6; "
                    "the timed region must be the contact steady state whatever --steps is)")
    ap.add_argument("--other-configs", type=int, default=1, help="1: also time AnymalTerrain (trimesh), UsefulHound and Cartpole (N=1 only) and report them "
                    "under other_configs")
    ap.add_argument("--ppo", type=int, default=1, help="1: also run BASELINE config 5 (Anymal PPO, 8192 envs/GPU, NCCL gradient all-reduce when N > 1) "
                    "for a few iterations and report it under ppo_config5")
    ap.add_argument("--ppo-timeout", type=int, default=240, help="seconds the learner leg may take before the line is printed without it")
    ap.add_argument("--task", default="Anymal", choices=sorted(TASKS), help="hot-path config to time (default: the headline Anymal config)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 200:
            args.steps = 200      # CPU arm: bounded so the run ends within minutes (13 ms per 4096-env step on 16 cores)
        args.warmup = min(args.warmup, 50)
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
