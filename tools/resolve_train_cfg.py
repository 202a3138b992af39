#!/usr/bin/env python
"""Resolve the reference's cfg/train/<Task>PPO.yaml files (Hydra/OmegaConf interpolations -> the defaults of cfg/config.yaml)
into plain yaml under isaacgymenv_b200/cfg/train/.  Run in the build container (needs /root/reference); outputs are committed.
Interpolations handled: ${...seed}, ${...checkpoint}, ${....multi_gpu}, ${....experiment}, ${....max_iterations},
${....task.env.numEnvs}, ${resolve_default:X,...} -> X, ${if:${...checkpoint},True,False} -> False, ${.name}."""
import os
import re
import sys

import yaml

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import isaacgymenv_b200  # noqa: E402

REF = os.environ.get("B2G_REFERENCE_ROOT", "/root/reference") + "/isaacgymenvs/cfg/train"
OUT = os.path.join(os.path.dirname(os.path.abspath(isaacgymenv_b200.__file__)), "cfg", "train")
TASKS = ["Anymal", "Hound", "Cartpole", "AnymalTerrain", "HoundTerrain", "UsefulHound", "Houndarm"]


def resolve(text, task):
    num_envs = isaacgymenv_b200.load_task_config(task)["env"]["numEnvs"]
    text = re.sub(r"\$\{resolve_default:([^,}]+),\$\{[^}]*\}\}", r"\1", text)
    text = re.sub(r"\$\{if:\$\{[^}]*\},True,False\}", "False", text)
    text = re.sub(r"\$\{\.+seed\}", "42", text)
    text = re.sub(r"\$\{\.+checkpoint\}", "''", text)
    text = re.sub(r"\$\{\.+multi_gpu\}", "False", text)
    text = re.sub(r"\$\{\.+task\.env\.numEnvs\}", str(num_envs), text)
    return text


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for task in TASKS:
        src = os.path.join(REF, f"{task}PPO.yaml")
        if not os.path.isfile(src):
            print("missing", src)
            continue
        cfg = yaml.safe_load(resolve(open(src).read(), task))
        c = cfg["params"]["config"]
        if isinstance(c.get("full_experiment_name"), str) and "${" in c["full_experiment_name"]:
            c["full_experiment_name"] = c["name"]
        left = [m for m in re.findall(r"\$\{[^}]*\}", yaml.safe_dump(cfg))]
        assert not left, (task, left)
        with open(os.path.join(OUT, f"{task}PPO.yaml"), "w") as fh:
            fh.write(f"# cfg/train/{task}PPO.yaml of the reference with the Hydra interpolations resolved to the defaults of cfg/config.yaml\n")
            yaml.safe_dump(cfg, fh, sort_keys=False, default_flow_style=None, width=110)
        print(task, c["horizon_length"], c["minibatch_size"], c["mini_epochs"], cfg["params"]["network"]["mlp"]["units"], c["max_epochs"])
