#!/usr/bin/env python
"""Step time of the REFERENCE's own task classes (unmodified files from a checkout, B2G_REFERENCE_ROOT) running on libb200gym through the
isaacgym shim -- the reference's hook structure: torch / TorchScript task code, one k_simulate launch per gym.simulate -- next to this
package's fused task of the same name on the same GPU.  Random actions, 4096 envs, CUDA-event timing after a pre-roll.
    B2G_REFERENCE_ROOT=/path/to/IsaacgymEnv python tools/time_reference_classes.py > gpurun_out/reference_classes_step_time.json"""
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import isaacgymenv_b200 as b2g  # noqa: E402

REF = os.environ["B2G_REFERENCE_ROOT"]
b2g.install_isaacgym_shim(REF)
vt = importlib.import_module("isaacgymenvs.tasks.base.vec_task")


def time_env(env, nact, n, steps=200, pre=100):
    g = torch.Generator(device="cuda:0").manual_seed(1)
    acts = [2 * torch.rand(n, nact, device="cuda:0", generator=g) - 1 for _ in range(16)]
    for k in range(pre):
        env.step(acts[k % 16])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        env.step(acts[k % 16])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3


out = {}
n = 4096
for module, cls, task, nact in (("tasks.anymal", "Anymal", "Anymal", 12), ("tasks.anymal_terrain", "AnymalTerrain", "AnymalTerrain", 12),
                                ("tasks.useful_hound", "UsefulHound", "UsefulHound", 18), ("tasks.manipulator", "Manipulator", "Manipulator", 6)):
    cfg = b2g.load_task_config(task, None)
    cfg["env"]["numEnvs"] = n
    cfg["sim"]["use_gpu_pipeline"] = True
    cfg["sim"].setdefault("physx", {})["use_gpu"] = True
    vt.EXISTING_SIM = None
    torch.manual_seed(0)
    ref = getattr(importlib.import_module("isaacgymenvs." + module), cls)(cfg=cfg, rl_device="cuda:0", sim_device="cuda:0", graphics_device_id=-1, headless=True,
                                                                         virtual_screen_capture=False, force_render=False)
    us_ref = time_env(ref, nact, n)
    del ref
    vt.EXISTING_SIM = None
    ours = b2g.make(seed=0, task=task, num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True)
    us_fused = time_env(ours, nact, n)
    del ours
    out[task] = {"envs": n, "reference_class_on_shim_us_per_step": round(us_ref, 1), "fused_task_us_per_step": round(us_fused, 1),
                 "reference_class_env_steps_per_s": round(n / us_ref * 1e6), "fused_env_steps_per_s": round(n / us_fused * 1e6)}
    print(task, out[task], file=sys.stderr, flush=True)
print(json.dumps(out, indent=1))
