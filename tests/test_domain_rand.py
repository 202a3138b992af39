"""Tensorised domain randomisation (isaacgymenv_b200/utils/domain_rand.py) against the reference's samplers
(utils/dr_utils.py, golden file tests/golden/dr_utils.npz made by gen_golden.py --dr-only): bucketing value for value,
sampler statistics per schedule step; the masking / frequency logic on a stand-in task; and on the GPU the Anymal task with
task.randomize=True (kernels reading per-env scales)."""
import json
import os

import numpy as np
import pytest
import torch

from isaacgymenv_b200.utils import domain_rand as DR

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "dr_utils.npz"))
BLOCKS = json.loads(str(G["blocks_json"]))


@pytest.mark.parametrize("name", ["friction", "gauss_buckets"])
def test_bucketing_matches_reference(name):
    v = torch.from_numpy(G[f"bucket_in_{name}"])
    out = DR.bucketed(v, BLOCKS[name]).numpy()
    ref = G[f"bucket_out_{name}"]
    # values that sit exactly on a bucket edge may fall either side in floating point: allow one bucket there
    lo, hi = (BLOCKS[name]["range"] if BLOCKS[name]["distribution"] == "uniform"
              else (BLOCKS[name]["range"][0] - 2 * np.sqrt(BLOCKS[name]["range"][1]), BLOCKS[name]["range"][0] + 2 * np.sqrt(BLOCKS[name]["range"][1])))
    width = (hi - lo) / BLOCKS[name]["num_buckets"]
    diff = np.abs(out - ref)
    assert (diff < 1e-9).mean() > 0.9
    assert diff.max() <= width * 1.0001


@pytest.mark.parametrize("name", ["mass", "friction", "gravity", "gauss_scaling", "loguniform"])
@pytest.mark.parametrize("step", [0, 500, 1500, 3000, 10000])
def test_sampler_statistics_match_reference(name, step):
    g = torch.Generator().manual_seed(99)
    s = DR.sample(BLOCKS[name], (200000,), step, "cpu", g).double().numpy()
    mean, std, lo, hi = G[f"stat_{name}_{step}"]
    tol = 4 * max(std, 1e-12) / np.sqrt(200000) * 2 + 1e-9
    assert abs(s.mean() - mean) < tol, (s.mean(), mean)
    assert abs(s.std() - std) < 0.01 * max(std, 1e-9) + 1e-9, (s.std(), std)
    if BLOCKS[name]["distribution"] != "gaussian":
        span = max(hi - lo, 1e-9)
        assert abs(s.min() - lo) < 1e-3 * span + 1e-6 and abs(s.max() - hi) < 1e-3 * span + 1e-6


class _FakeGym:
    def __init__(self, n):
        self.frame = 0
        ls = torch.zeros(n, 13, DR._abi.LINK_SCALE_COLS)
        ls[:, :, :3] = 1.0
        self.t = {DR._abi.T_ENV_SCALE: torch.ones(n, 4), DR._abi.T_FRICTION: torch.ones(n), DR._abi.T_LINK_SCALE: ls}
        self.gravity_sets = []

        class V:
            x, y, z = 0.0, 0.0, -9.81

        class P:
            gravity = V()

        self.params = P()

    def get_frame_count(self, sim):
        return self.frame

    def _tensor(self, sim, kind):
        return self.t[kind]

    def get_sim_params(self, sim):
        return self.params

    def set_sim_params(self, sim, p):
        self.gravity_sets.append((p.gravity.x, p.gravity.y, p.gravity.z))


class _FakeTask:
    def __init__(self, n=64):
        self.num_envs, self.device, self.seed = n, "cpu", 7
        self.gym, self.sim = _FakeGym(n), None
        self.reset_buf = torch.zeros(n, dtype=torch.long)
        self.randomize_buf = torch.zeros(n, dtype=torch.long)
        self.dr_randomizations = {}


def _params():
    return {
        "frequency": 10,
        "observations": {"range": [0, 0.002], "operation": "additive", "distribution": "gaussian"},
        "actions": {"range": [0.0, 0.02], "operation": "additive", "distribution": "gaussian"},
        "sim_params": {"gravity": {"range": [0, 0.4], "operation": "additive", "distribution": "gaussian"}},
        "actor_params": {"anymal": {
            "color": True,
            "rigid_body_properties": {"mass": {"range": [0.5, 1.5], "operation": "scaling", "distribution": "uniform", "setup_only": True}},
            "rigid_shape_properties": {"friction": {"num_buckets": 50, "range": [0.7, 1.3], "operation": "scaling", "distribution": "uniform"},
                                       "restitution": {"range": [0.0, 0.7], "operation": "scaling", "distribution": "uniform"}},
            "dof_properties": {"damping": {"range": [0.5, 1.5], "operation": "scaling", "distribution": "uniform"},
                               "stiffness": {"range": [0.5, 1.5], "operation": "scaling", "distribution": "loguniform"},
                               "lower": {"range": [0, 0.01], "operation": "additive", "distribution": "gaussian"}}}},
    }


def test_randomizer_masks_frequency_and_setup_only():
    task = _FakeTask()
    dr = DR.DomainRandomizer(task, count_steps=True)
    p = _params()
    dr.apply(p)                                   # first pass: every env, every parameter, noise closures, gravity
    sc, fr = task.gym.t[DR._abi.T_LINK_SCALE], task.gym.t[DR._abi.T_FRICTION]
    assert (task.gym.t[DR._abi.T_ENV_SCALE] == 1).all()          # the per-env factors stay at the caller's disposal
    assert ((sc[:, :, 0] >= 0.5) & (sc[:, :, 0] <= 1.5)).all() and ((sc[:, 1:, 1:3] >= 0.5) & (sc[:, 1:, 1:3] <= 1.5)).all()
    # one draw per body / per DOF (utils/dr_utils.py:135-238), not one per environment
    assert sc[:, :, 0].std(1).min() > 0.05 and sc[:, 1:, 1].std(1).min() > 0.05 and sc[:, 1:, 2].std(1).min() > 0.05
    assert (sc[:, 0, 1:3] == 1).all() and (sc[:, :, 4:] == 0).all()          # the root has no DOF; `upper` is not in this block
    assert sc[:, 1:, 3].abs().max() < 0.06 and sc[:, 1:, 3].std() > 1e-3          # additive gaussian limit offsets
    assert ((fr >= 0.7) & (fr < 1.3)).all() and len(torch.unique(fr)) <= 50
    assert len(task.gym.gravity_sets) == 1 and set(task.dr_randomizations) == {"observations", "actions"}
    assert sorted(dr.skipped) == ["anymal.color", "anymal.rigid_shape_properties.restitution"]
    x = torch.zeros(64, 12)
    y = task.dr_randomizations["actions"]["noise_lambda"](x)
    assert 0.01 < y.std() < 0.03 and y.shape == x.shape
    # before `frequency` frames: nothing moves, even for resetting envs
    sc0, fr0 = sc.clone(), fr.clone()
    task.gym.frame = 5
    task.randomize_buf += 5
    task.reset_buf[:8] = 1
    dr.apply(p)
    assert torch.equal(sc, sc0) and torch.equal(fr, fr0) and len(task.gym.gravity_sets) == 1
    # after it: gravity + noise are redrawn; only the resetting envs get new physical parameters; mass is setup_only
    task.gym.frame = 12
    task.randomize_buf += 7
    dr.apply(p)
    assert len(task.gym.gravity_sets) == 2
    assert torch.equal(sc[:, :, 0], sc0[:, :, 0]), "setup_only mass must not change after the first pass"
    assert not torch.equal(sc[:8, 1:, 1:3], sc0[:8, 1:, 1:3]) and torch.equal(sc[8:], sc0[8:])
    assert not torch.equal(fr[:8], fr0[:8]) and torch.equal(fr[8:], fr0[8:])
    assert (task.randomize_buf[:8] == 0).all() and (task.randomize_buf[8:] == 12).all()
    assert dr.num_applied() == 64 + 8


def test_reference_quirk_no_step_counting_means_no_rerandomisation():
    task = _FakeTask()
    dr = DR.DomainRandomizer(task, count_steps=False)
    p = _params()
    dr.apply(p)
    sc0 = task.gym.t[DR._abi.T_LINK_SCALE].clone()
    task.gym.frame = 100
    task.reset_buf[:] = 1
    dr.apply(p)                                   # randomize_buf never advanced (vec_task.py:322,632-635): no env qualifies
    assert torch.equal(task.gym.t[DR._abi.T_LINK_SCALE], sc0) and len(task.gym.gravity_sets) == 2


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_anymal_with_domain_randomisation_on_gpu(fused):
    import isaacgymenv_b200

    n = 256
    rp = isaacgymenv_b200.load_task_config("Anymal")["task"]["randomization_params"]
    rp = json.loads(json.dumps(rp))
    rp["frequency"] = 20
    rp["count_steps"] = True
    for blk in (rp["sim_params"]["gravity"], rp["actor_params"]["anymal"]["rigid_body_properties"]["mass"],
                rp["actor_params"]["anymal"]["rigid_shape_properties"]["friction"], rp["actor_params"]["anymal"]["dof_properties"]["damping"],
                rp["actor_params"]["anymal"]["dof_properties"]["stiffness"]):
        blk["schedule_steps"] = 1          # full-strength randomisation from the first frames
    rp["actor_params"]["anymal"]["rigid_body_properties"]["mass"]["setup_only"] = False     # also at construction (schedule at step 0 is 0)
    over = {"task": {"randomize": True, "randomization_params": rp}, "env": {"fusedStep": fused}}
    env = isaacgymenv_b200.make(seed=3, task="Anymal", num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True, overrides=over)
    plain = isaacgymenv_b200.make(seed=3, task="Anymal", num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True,
                                  overrides={"env": {"fusedStep": fused}})
    g = torch.Generator(device="cuda").manual_seed(1)
    for i in range(80):
        a = 2 * torch.rand(n, 12, device="cuda", generator=g) - 1
        o, r, d, _ = env.step(a)
        op, rp_, dp, _ = plain.step(a)
        assert torch.isfinite(o["obs"]).all() and torch.isfinite(r).all()
    sc = env._dr.link_scale
    assert ((sc[:, :, :3] > 0.45) & (sc[:, :, :3] < 1.55)).all()
    # the construction-time pass happens at frame 0, where the linear schedule still gives scale 1; the environments that were
    # reset after `frequency` steps carry fresh draws -- one per DOF
    redrawn = (sc[:, 1:, 1:3] != 1.0).flatten(1).any(dim=1)
    assert int(redrawn.sum()) >= 8 and sc[redrawn][:, 1:, 1].std(1).min() > 0.05, int(redrawn.sum())
    assert env._dr.num_applied() >= n + int(redrawn.sum()), "environments reset after `frequency` steps are re-randomised (possibly more than once)"
    assert "observations" in env.dr_randomizations and "actions" in env.dr_randomizations
    g = env.gym.get_sim_params(env.sim).gravity
    assert abs(g.z + 9.81) < 3.0 and (abs(g.x) > 1e-4 or abs(g.y) > 1e-4)
    assert (o["obs"] - op["obs"]).abs().max() > 1e-3, "randomised dynamics must differ from the nominal twin"
    assert float(env.root_states[:, 2].abs().max()) < 5.0
