"""Hydra-style train CLI (isaacgymenv_b200/train.py, reference isaacgymenvs/train.py + cfg/train/*.yaml): override parsing and
the mapping of the reference's train yaml onto the in-repo learner; on the GPU a short train -> checkpoint -> test round trip."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_override_parsing():
    from isaacgymenv_b200.train import parse_overrides

    top, task_over, train_over = parse_overrides(["task=AnymalTerrain", "num_envs=512", "headless=True", "test=False", "seed=7",
                                                  "task.env.learn.pushInterval_s=8", "train.params.config.horizon_length=16", "checkpoint=runs/x.pth"])
    assert top["task"] == "AnymalTerrain" and top["num_envs"] == 512 and top["seed"] == 7 and top["test"] is False and top["checkpoint"] == "runs/x.pth"
    assert task_over == {"env": {"learn": {"pushInterval_s": 8}}} and train_over == {"params": {"config": {"horizon_length": 16}}}
    with pytest.raises(SystemExit):
        parse_overrides(["bogus=1"])
    with pytest.raises(SystemExit):
        parse_overrides(["task"])


@pytest.mark.parametrize("task,units,horizon,mb,epochs", [("Anymal", (256, 128, 64), 24, 32768, 1000), ("Cartpole", (32, 32), 16, 8192, 100),
                                                           ("AnymalTerrain", (512, 256, 128), 24, 16384, 1500), ("UsefulHound", (512, 256, 128), 24, 16384, 15000),
                                                           ("Houndarm", (256, 128, 64), 32, 16384, 10000)])
def test_train_yaml_maps_onto_ppo_config(task, units, horizon, mb, epochs):
    """Values of the reference's cfg/train/<Task>PPO.yaml (resolved copies under cfg/train/)."""
    from isaacgymenv_b200.train import load_train_config, ppo_config_from_train_cfg

    tc = load_train_config(f"{task}PPO")
    assert tc["params"]["algo"]["name"] == "a2c_continuous" and tc["params"]["config"]["normalize_input"] is True
    c = ppo_config_from_train_cfg(tc)
    assert c.units == units and c.horizon_length == horizon and c.minibatch_size == mb and c.max_epochs == epochs
    assert c.gamma == 0.99 and c.tau == 0.95 and c.e_clip == 0.2 and c.mini_epochs in (5, 8)
    assert ppo_config_from_train_cfg(tc, max_iterations=3).max_epochs == 3
    tc2 = load_train_config(f"{task}PPO", {"params": {"config": {"horizon_length": 8}}})
    assert ppo_config_from_train_cfg(tc2).horizon_length == 8


@pytest.mark.gpu
def test_train_then_test_roundtrip(tmp_path):
    ck = str(tmp_path / "nn" / "Cartpole.pth")
    run = [sys.executable, "-m", "isaacgymenv_b200.train", "task=Cartpole", "num_envs=256", "max_iterations=12", f"output={ck}"]
    out = subprocess.run(run, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["mode"] == "train" and res["epochs"] == 12 and os.path.isfile(ck)
    out = subprocess.run([sys.executable, "-m", "isaacgymenv_b200.train", "task=Cartpole", "num_envs=64", "test=True", f"checkpoint={ck}"],
                         cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["mode"] == "test" and res["mean_episode_length"] > 5
