"""Host-side logic that needs no GPU: configs, the make() factory, sim-parameter parsing, spaces, the gym shim's
asset queries (what tasks/anymal.py:184-224 reads), rank/shard helpers."""
import numpy as np
import pytest

import isaacgymenv_b200 as b2g
from isaacgymenv_b200 import distributed, gymapi, spaces
from isaacgymenv_b200.tasks.base.vec_task import VecTask


def test_task_configs_load_and_carry_the_reference_constants():
    a = b2g.load_task_config("Anymal")
    assert a["env"]["numEnvs"] == 4096 and a["env"]["clipObservations"] == 5.0 and a["env"]["clipActions"] == 1.0
    assert a["sim"]["dt"] == 0.02 and a["sim"]["substeps"] == 2
    assert a["env"]["control"] == {"stiffness": 85.0, "damping": 2.0, "actionScale": 0.5, "controlFrequencyInv": 1}
    assert a["sim"]["physx"]["num_position_iterations"] == 4 and a["sim"]["physx"]["num_velocity_iterations"] == 1
    assert a["env"]["learn"]["torqueRewardScale"] == -0.000025 and a["env"]["learn"]["episodeLength_s"] == 50
    h = b2g.load_task_config("Hound")
    assert h["env"]["urdfAsset"]["collapseFixedJoints"] is False and h["env"]["defaultJointAngles"]["FL_knee_joint"] == -1.5708
    c = b2g.load_task_config("Cartpole", {"env": {"numEnvs": 64}})
    assert c["env"]["numEnvs"] == 64 and c["env"]["maxEffort"] == 400.0 and c["sim"]["physx"]["num_velocity_iterations"] == 0
    with pytest.raises(ValueError):
        b2g.load_task_config("NoSuchTask")


def test_task_map_has_the_hot_path_tasks():
    m = b2g.task_map()
    assert {"Anymal", "Hound"} <= set(m)
    for cls in m.values():
        assert issubclass(cls, VecTask)


def test_parse_sim_params():
    cfg = b2g.load_task_config("Anymal")
    sp = VecTask._parse_sim_params(None, "physx", cfg["sim"])
    assert sp.dt == 0.02 and sp.substeps == 2 and sp.up_axis == gymapi.UP_AXIS_Z
    assert (sp.gravity.x, sp.gravity.y, sp.gravity.z) == (0.0, 0.0, -9.81)
    assert sp.physx.num_position_iterations == 4 and sp.physx.contact_offset == 0.02 and sp.physx.max_depenetration_velocity == 100.0
    bad = dict(cfg["sim"], up_axis="x")
    with pytest.raises(ValueError):
        VecTask._parse_sim_params(None, "physx", bad)


def test_spaces_box():
    b = spaces.Box(np.full(48, -np.inf), np.full(48, np.inf))
    assert b.shape == (48,) and b.contains(np.zeros(48, np.float32))
    a = spaces.Box(np.full(12, -1.0), np.full(12, 1.0))
    assert a.contains(a.sample()) and not a.contains(np.full(12, 2.0))


def test_gym_asset_queries_without_a_sim():
    gym = gymapi.acquire_gym()
    assert gym is gymapi.acquire_gym()
    o = gymapi.AssetOptions()
    o.collapse_fixed_joints, o.replace_cylinder_with_capsule, o.density = True, True, 0.001
    asset = gym.load_asset(None, "/nonexistent/assets", "urdf/anymal_c/urdf/anymal.urdf", o)     # falls back to the compiled model
    assert gym.get_asset_dof_count(asset) == 12 and gym.get_asset_rigid_body_count(asset) == 13
    names = gym.get_asset_rigid_body_names(asset)
    assert [n for n in names if "THIGH" in n] == ["LF_THIGH", "LH_THIGH", "RF_THIGH", "RH_THIGH"]
    props = gym.get_asset_dof_properties(asset)
    assert props.dtype.names[:4] == ("hasLimits", "lower", "upper", "driveMode") and props["effort"][0] == 80.0
    props["driveMode"][:] = gymapi.DOF_MODE_POS
    props["stiffness"][:] = 85.0
    assert gym.get_asset_rigid_shape_properties(asset)[0].friction == 1.0
    with pytest.raises(FileNotFoundError):
        gym.load_asset(None, "/nonexistent", "urdf/unknown_robot.urdf", o)


def test_rank_helpers(monkeypatch):
    monkeypatch.setenv("RANK", "3")
    monkeypatch.setenv("LOCAL_RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "8")
    r = distributed.rank_info()
    assert (r.rank, r.local_rank, r.world_size, r.device) == (3, 1, 8, "cuda:1")
    assert distributed.shard_seed(42, 3) == 45
    ranges = [distributed.env_range(4096 * 8, k, 8) for k in range(8)]
    assert ranges[0] == (0, 4096) and ranges[7] == (7 * 4096, 8 * 4096)
    assert sum(b - a for a, b in [distributed.env_range(10, k, 4) for k in range(4)]) == 10
    assert distributed.aggregate_env_steps_per_sec(4096, 100, 0.5, 8) == 8 * 4096 * 100 / 0.5
