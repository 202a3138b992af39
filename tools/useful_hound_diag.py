#!/usr/bin/env python
"""Why do UsefulHound robots fly?  Random leg actions, zero arm actions, 300 steps: distribution of the root height for a few settings."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import isaacgymenv_b200

def run(task, ov, label, arm_zero=True, n=4096, steps=300):
    env = isaacgymenv_b200.make(seed=7, task=task, num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True, overrides=ov)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(3)
    na = env.num_acts
    zmax = torch.zeros(n, device="cuda")
    first = {}
    for k in range(steps):
        a = 2 * torch.rand(n, na, device="cuda", generator=g) - 1
        if arm_zero and na > 12:
            a[:, 12:] = 0
        env.step(a)
        z = env.root_states[:, 2]
        zmax = torch.maximum(zmax, z)
    st = env.sim.contact_stats() if hasattr(env.sim, "contact_stats") else None
    out = {"label": label, "max_z": float(zmax.max()), "envs_above_2m": int((zmax > 2).sum()), "envs_above_1m": int((zmax > 1).sum()), "q99_zmax": float(zmax.quantile(0.99)),
           "median_zmax": float(zmax.median()), "stats": st}
    print(json.dumps(out), flush=True)
    return out

res = []
res.append(run("HoundTerrain", {}, "HoundTerrain default"))
res.append(run("UsefulHound", {}, "UsefulHound default (6 slots)"))
res.append(run("UsefulHound", {"sim": {"physx": {"max_depenetration_velocity": 5.0}}}, "UsefulHound max_depen 5"))
res.append(run("UsefulHound", {"sim": {"physx": {"max_contacts_per_chain": 8}}}, "UsefulHound 8 slots"))
res.append(run("UsefulHound", {}, "UsefulHound random arm", arm_zero=False))
json.dump(res, open("gpurun_out/r02x_useful_hound_diag.json", "w"), indent=1)
