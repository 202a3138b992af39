"""Drop-in proof against the reference's OWN task files (north_star: "`isaacgymenvs.make(...)` ... stay drop-in").

These tests need a checkout of the reference (``B2G_REFERENCE_ROOT``, default /root/reference) and skip without one -- the
GPU box has none, so the GPU half runs only where a maintainer provides it.  Nothing here is copied from the reference: its
files are parsed (ast) or imported in place.

* every ``self.gym.*`` method and every ``gymapi.* / gymtorch.*`` name the hot-path files use exists in this package's shim;
* the reference's task modules import UNMODIFIED under ``install_isaacgym_shim`` and their constructors run through the
  reference's own ``VecTask.__init__`` down to ``gym.create_sim`` = ``b2g_sim_create`` of the C ABI;
* on a GPU: the reference's ``Anymal`` / ``Cartpole`` classes construct and step on libb200gym, and the reference Anymal's
  trajectory equals this package's own generic (un-fused) task bit for bit under the same seed and actions.
"""
import ast
import importlib
import os

import pytest

REF = os.environ.get("B2G_REFERENCE_ROOT", "/root/reference")
HAVE_REF = os.path.isdir(os.path.join(REF, "isaacgymenvs", "tasks"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="no reference checkout (B2G_REFERENCE_ROOT)")

HOT_PATH_FILES = ["tasks/base/vec_task.py", "tasks/anymal.py", "tasks/hound.py", "tasks/cartpole.py", "tasks/anymal_terrain.py", "tasks/Hound_terrain.py",
                  "tasks/useful_hound.py", "tasks/hound_arm.py", "tasks/manipulator.py", "utils/dr_utils.py"]
# names only reached with a viewer / camera sensors / an external params generator (headless=True training never calls them;
# SURVEY 8(b) "viewer / DR (not needed headless)")
VIEWER_ONLY = {"create_viewer", "subscribe_viewer_keyboard_event", "query_viewer_has_closed", "query_viewer_action_events", "poll_viewer_events",
               "step_graphics", "draw_viewer", "sync_frame_time", "destroy_viewer", "viewer_camera_look_at", "render_all_camera_sensors",
               "clear_lines", "add_lines", "start_access_image_tensors", "end_access_image_tensors", "create_camera_sensor", "set_camera_location",
               "get_camera_image_gpu_tensor", "write_viewer_image_to_file", "get_viewer_camera_handle", "set_light_parameters"}


def _used_names():
    gym_methods, api_names, torch_names = {}, {}, {}
    for rel in HOT_PATH_FILES:
        path = os.path.join(REF, "isaacgymenvs", rel)
        tree = ast.parse(open(path).read(), filename=path)
        for node in ast.walk(tree):
            if not isinstance(node, ast.Attribute):
                continue
            v = node.value
            if isinstance(v, ast.Attribute) and v.attr == "gym" and isinstance(v.value, ast.Name) and v.value.id == "self":
                gym_methods.setdefault(node.attr, f"{rel}:{node.lineno}")
            elif isinstance(v, ast.Name) and v.id == "gym" and rel != "tasks/base/vec_task.py":
                gym_methods.setdefault(node.attr, f"{rel}:{node.lineno}")
            elif isinstance(v, ast.Name) and v.id == "gymapi":
                api_names.setdefault(node.attr, f"{rel}:{node.lineno}")
            elif isinstance(v, ast.Name) and v.id == "gymtorch":
                torch_names.setdefault(node.attr, f"{rel}:{node.lineno}")
    return gym_methods, api_names, torch_names


@needs_ref
def test_shim_covers_every_gym_name_the_hot_path_files_use():
    from isaacgymenv_b200 import gymapi, gymtorch

    gym_methods, api_names, torch_names = _used_names()
    assert len(gym_methods) > 40 and "simulate" in gym_methods and "acquire_dof_state_tensor" in gym_methods
    missing = {n: where for n, where in gym_methods.items() if n not in VIEWER_ONLY and not hasattr(gymapi.Gym, n)}
    assert not missing, f"gym methods used by the reference's hot-path files but absent from the shim: {missing}"
    missing = {n: where for n, where in api_names.items() if not hasattr(gymapi, n)}
    assert not missing, f"gymapi names used by the reference but absent from the shim: {missing}"
    missing = {n: where for n, where in torch_names.items() if not hasattr(gymtorch, n)}
    assert not missing, f"gymtorch names absent from the shim: {missing}"
    viewer_used = sorted(n for n in gym_methods if n in VIEWER_ONLY)
    # the viewer names must still exist as callables so that headless=False fails with a clear message, not an AttributeError
    for n in viewer_used:
        assert hasattr(gymapi.Gym, n), n


@needs_ref
@pytest.mark.parametrize("module,cls", [("tasks.anymal", "Anymal"), ("tasks.hound", "Hound"), ("tasks.cartpole", "Cartpole"), ("tasks.anymal_terrain", "AnymalTerrain"),
                                        ("tasks.Hound_terrain", "HoundTerrain"), ("tasks.useful_hound", "UsefulHound"), ("tasks.hound_arm", "Houndarm"), ("tasks.manipulator", "Manipulator")])
def test_reference_task_modules_import_unmodified_under_the_shim(module, cls):
    import isaacgymenv_b200 as b2g

    b2g.install_isaacgym_shim(REF)
    mod = importlib.import_module("isaacgymenvs." + module)
    klass = getattr(mod, cls)
    assert klass.__module__ == "isaacgymenvs." + module
    assert os.path.realpath(mod.__file__).startswith(os.path.realpath(REF)), "must be the reference's own file"
    base = importlib.import_module("isaacgymenvs.tasks.base.vec_task")
    assert issubclass(klass, base.VecTask)
    import isaacgym

    assert isaacgym.gymapi is b2g.gymapi and isaacgym.gymtorch is b2g.gymtorch


def _task_cfg(task, n):
    import isaacgymenv_b200 as b2g

    cfg = b2g.load_task_config(task, None)
    cfg["env"]["numEnvs"] = n
    cfg["sim"]["use_gpu_pipeline"] = True
    cfg["sim"].setdefault("physx", {})["use_gpu"] = True
    return cfg


@needs_ref
def test_reference_constructor_reaches_the_c_abi_without_a_gpu():
    """No CUDA device here: the reference's Anymal.__init__ -> VecTask.__init__ -> create_sim must arrive at b2g_sim_create and fail
    THERE, loudly (no CPU fallback) -- everything above it (Env.__init__, spaces, sim-params parsing, acquire_gym) is the reference's code
    running on the shim."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    import isaacgymenv_b200 as b2g
    from isaacgymenv_b200._lib import B2GError

    b2g.install_isaacgym_shim(REF)
    mod = importlib.import_module("isaacgymenvs.tasks.anymal")
    vt = importlib.import_module("isaacgymenvs.tasks.base.vec_task")
    vt.EXISTING_SIM = None
    with pytest.raises(B2GError, match="no usable CUDA device"):
        mod.Anymal(cfg=_task_cfg("Anymal", 16), rl_device="cuda:0", sim_device="cuda:0", graphics_device_id=-1, headless=True,
                   virtual_screen_capture=False, force_render=False)


@needs_ref
@pytest.mark.gpu
def test_reference_anymal_and_cartpole_step_on_libb200gym():
    import torch

    import isaacgymenv_b200 as b2g

    b2g.install_isaacgym_shim(REF)
    vt = importlib.import_module("isaacgymenvs.tasks.base.vec_task")
    n = 64
    torch.manual_seed(42)
    vt.EXISTING_SIM = None           # the reference keeps ONE sim per process (vec_task.py:55-64)
    ref = importlib.import_module("isaacgymenvs.tasks.anymal").Anymal(cfg=_task_cfg("Anymal", n), rl_device="cuda:0", sim_device="cuda:0",
                                                                       graphics_device_id=-1, headless=True, virtual_screen_capture=False, force_render=False)
    torch.manual_seed(42)
    ours = b2g.make(seed=42, task="Anymal", num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True, overrides={"env": {"fusedStep": False}})
    assert ref.num_envs == n and ref.obs_buf.shape == (n, 48) and ref.reset_buf.dtype == torch.int64
    g = torch.Generator(device="cuda:0").manual_seed(3)
    resets = 0
    worst = 0.0
    for k in range(40):
        # same starting point for both: the reference task's tensors are the truth, this package's generic task follows them
        ours.root_states.copy_(ref.root_states); ours.dof_state.copy_(ref.dof_state); ours.commands.copy_(ref.commands)
        ours.progress_buf.copy_(ref.progress_buf); ours.reset_buf.copy_(ref.reset_buf)
        a = 2 * torch.rand(n, 12, device="cuda:0", generator=g) - 1
        st = torch.cuda.get_rng_state()
        o_r, r_r, d_r, e_r = ref.step(a)
        torch.cuda.set_rng_state(st)          # reset_idx draws from torch's CUDA generator: replay the same draws
        o_o, r_o, d_o, e_o = ours.step(a)
        assert torch.isfinite(o_r["obs"]).all() and torch.isfinite(r_r).all()
        # same kernels underneath, the reference's TorchScript arithmetic on top: agreement to rounding (TorchScript may contract a * b + c)
        worst = max(worst, float((o_r["obs"] - o_o["obs"]).abs().max()), float((r_r - r_o).abs().max()))
        assert torch.allclose(o_r["obs"], o_o["obs"], rtol=1e-5, atol=1e-5), f"step {k}: {float((o_r['obs'] - o_o['obs']).abs().max())}"
        assert torch.allclose(r_r, r_o, rtol=1e-5, atol=1e-7) and torch.equal(d_r, d_o) and torch.equal(e_r["time_outs"], e_o["time_outs"]), f"step {k}"
        assert torch.equal(ref.root_states, ours.root_states) or torch.allclose(ref.root_states, ours.root_states, rtol=1e-6, atol=1e-6)
        resets += int(d_r.sum())
    assert resets > 0
    vt.EXISTING_SIM = None
    cart = importlib.import_module("isaacgymenvs.tasks.cartpole").Cartpole(cfg=_task_cfg("Cartpole", 32), rl_device="cuda:0", sim_device="cuda:0",
                                                                            graphics_device_id=-1, headless=True, virtual_screen_capture=False, force_render=False)
    for k in range(30):
        o, r, d, e = cart.step(2 * torch.rand(32, 1, device="cuda:0", generator=g) - 1)
    assert o["obs"].shape == (32, 4) and torch.isfinite(o["obs"]).all() and torch.isfinite(r).all()


@needs_ref
@pytest.mark.parametrize("curriculum", [True, False])
def test_terrain_grid_equals_the_references_terrain_class(curriculum):
    """The reference's own ``Terrain`` class (tasks/anymal_terrain.py:543-673), executed unmodified with ``isaacgym.terrain_utils`` =
    this package's sub-terrain builders, against this package's ``Terrain``: same tile layout, curriculum mapping (column -> type,
    row -> difficulty), border, height samples and environment origins, value for value, from the same numpy random stream.  (The
    sub-terrain builders themselves restate Isaac Gym's published ``terrain_utils`` -- that module is not in the reference tree.)"""
    import numpy as np

    import isaacgymenv_b200 as b2g
    from isaacgymenv_b200.terrain import Terrain

    b2g.install_isaacgym_shim(REF)
    ref_mod = importlib.import_module("isaacgymenvs.tasks.anymal_terrain")
    cfg = dict(terrainType="trimesh", curriculum=curriculum, mapLength=8.0, mapWidth=8.0, numLevels=4, numTerrains=10,
               terrainProportions=[0.1, 0.1, 0.35, 0.25, 0.2], slopeTreshold=0.5)
    np.random.seed(7)
    ref = ref_mod.Terrain(cfg, 256)
    np.random.seed(7)
    ours = Terrain(cfg, 256, seed=None)
    assert ref.height_field_raw.shape == ours.height_field_raw.shape == (4 * 80 + 400, 10 * 80 + 400)
    assert np.array_equal(ref.height_field_raw, ours.height_field_raw)
    assert np.array_equal(ref.env_origins, ours.env_origins)
    assert (ref.tot_rows, ref.tot_cols, ref.border, ref.env_rows, ref.env_cols) == (ours.tot_rows, ours.tot_cols, ours.border, ours.env_rows, ours.env_cols)
    assert np.abs(ours.height_field_raw).max() > 20 and len(np.unique(ours.env_origins[:, :, 2])) > 3


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("module,cls,task,nd", [("tasks.manipulator", "Manipulator", "Manipulator", 7), ("tasks.hound_arm", "Houndarm", "Houndarm", 6)])
def test_reference_arm_tasks_step_on_libb200gym(module, cls, task, nd):
    """The reference's own arm-reach classes, unmodified, on the shim: load_asset resolves their URDF (or its compiled model when the asset
    tree is absent), acquire_jacobian_tensor / acquire_mass_matrix_tensor feed their torch OSC law, set_dof_actuation_force_tensor +
    simulate move the arm.  Closed loop: their controller drives the end effector towards the command; this package's generic task,
    glued to the same state each step, returns the same observations and reward."""
    import torch

    import isaacgymenv_b200 as b2g

    b2g.install_isaacgym_shim(REF)
    vt = importlib.import_module("isaacgymenvs.tasks.base.vec_task")
    vt.EXISTING_SIM = None
    n = 64
    torch.manual_seed(11)
    ref = getattr(importlib.import_module("isaacgymenvs." + module), cls)(cfg=_task_cfg(task, n), rl_device="cuda:0", sim_device="cuda:0", graphics_device_id=-1,
                                                                         headless=True, virtual_screen_capture=False, force_render=False)
    assert ref.num_envs == n and ref.num_dofs == nd and ref.obs_buf.shape == (n, 10) and ref._mm.shape == (n, nd, nd) and ref._j_eef.shape == (n, 6, nd)
    torch.manual_seed(11)
    ours = b2g.make(seed=11, task=task, num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True, overrides={"env": {"fusedStep": False}})
    ref._refresh()
    ref.commands[:] = ref.states["eef_pos"] + torch.tensor([0.05, -0.04, 0.03], device="cuda:0")
    d0 = float((ref.states["eef_pos"] - ref.commands).norm(dim=-1).mean())
    for k in range(60):
        ours._dof_state.copy_(ref._dof_state); ours.commands.copy_(ref.commands)
        ours.progress_buf.copy_(ref.progress_buf); ours.reset_buf.copy_(ref.reset_buf)
        ours._refresh()
        ref._refresh()
        err = ref.commands - ref.states["eef_pos"]
        a = torch.cat([torch.clamp(err / 0.1, -1, 1), torch.zeros(n, 3, device="cuda:0")], dim=1)
        o_r, r_r, d_r, e_r = ref.step(a)
        o_o, r_o, d_o, e_o = ours.step(a)
        assert torch.isfinite(o_r["obs"]).all() and torch.isfinite(r_r).all() and not d_r.any()
        # the same kernels under both; the OSC law is float32 torch.inverse in both (ill-conditioned on these arms): agreement to ~1e-3
        assert torch.allclose(o_r["obs"], o_o["obs"], rtol=2e-3, atol=2e-3), (k, float((o_r["obs"] - o_o["obs"]).abs().max()))
        assert torch.allclose(r_r, r_o, rtol=2e-3, atol=1e-4) and torch.equal(d_r, d_o)
    ref._refresh()
    d1 = float((ref.states["eef_pos"] - ref.commands).norm(dim=-1).mean())
    assert d1 < 0.35 * d0, (d0, d1)
    vt.EXISTING_SIM = None


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("module,cls,task,nact,nobs", [("tasks.anymal_terrain", "AnymalTerrain", "AnymalTerrain", 12, 188), ("tasks.Hound_terrain", "HoundTerrain", "HoundTerrain", 12, 188),
                                                       ("tasks.useful_hound", "UsefulHound", "UsefulHound", 18, 204)])
@pytest.mark.parametrize("terrain", ["plane", "trimesh"])
def test_reference_rough_terrain_tasks_step_on_libb200gym(module, cls, task, nact, nobs, terrain):
    """BASELINE configs 3 and 4 through the reference's own task classes, unmodified, on the shim: their ``Terrain`` class builds the
    height field on this package's ``terrain_utils``, ``add_triangle_mesh`` / ``add_ground`` hand it to the library, the decimation loop
    (``set_dof_actuation_force_tensor`` + ``simulate`` + ``refresh_dof_state_tensor`` four times per step), reward, height scan, curriculum,
    pushes and resets are the reference's code.  Contract checks over a rollout: shapes, finiteness, robots stand under small actions until
    the episode limit resets them, fall and get reset under random actions, the height scan sees the terrain."""
    import torch

    import isaacgymenv_b200 as b2g

    b2g.install_isaacgym_shim(REF)
    vt = importlib.import_module("isaacgymenvs.tasks.base.vec_task")
    vt.EXISTING_SIM = None
    n = 128
    cfg = _task_cfg(task, n)
    cfg["env"]["terrain"]["terrainType"] = terrain
    if terrain == "trimesh":
        cfg["env"]["terrain"].update(numLevels=3, numTerrains=4, mapLength=8.0, mapWidth=8.0)
    cfg["env"]["learn"]["episodeLength_s"] = 1.0      # 50 policy steps: time-outs inside the window
    torch.manual_seed(5)
    ref = getattr(importlib.import_module("isaacgymenvs." + module), cls)(cfg=cfg, rl_device="cuda:0", sim_device="cuda:0", graphics_device_id=-1, headless=True,
                                                                         virtual_screen_capture=False, force_render=False)
    assert ref.num_envs == n and ref.num_actions == nact and ref.obs_buf.shape == (n, nobs)
    g = torch.Generator(device="cuda:0").manual_seed(3)
    resets = early = at_limit = 0
    for k in range(120):
        a = 2 * torch.rand(n, nact, device="cuda:0", generator=g) - 1
        if k < 56:
            a = 0.05 * a      # first a quiet stretch: the robots keep standing until the 50-step episode limit; then random actions throw them over
        o, r, d, e = ref.step(a)
        assert o["obs"].shape == (n, nobs) and torch.isfinite(o["obs"]).all() and torch.isfinite(r).all() and torch.isfinite(ref.root_states).all(), k
        assert e["time_outs"].shape == (n,)      # (always False here: the reference's post_physics_step has reset progress_buf before VecTask.step looks)
        resets += int(d.sum())
        early += int(d.sum()) if k < 45 else 0
        at_limit += int(d.sum()) if 47 <= k <= 51 else 0
    if cls == "AnymalTerrain":      # ANYmal stands under small actions until the episode limit; the hound's reset poses (default angles x 0.5 .. 1.5 under
        # its explicit PD gains) put a thigh on the ground for part of the robots, which the Hound variant of check_termination ends at once
        assert early < n // 4, f"{early} robots fell while standing still"
        assert at_limit >= n // 2, f"only {at_limit} episodes ended at the 50-step limit"
        assert resets > at_limit + n // 2, (resets, at_limit)
    else:
        assert resets > n, resets
    assert float(ref.root_states[:, 2].max()) < 25.0
    if terrain == "trimesh":
        assert float(ref.measured_heights.std()) > 1e-3        # the scan reads a non-flat field
    vt.EXISTING_SIM = None


@needs_ref
@pytest.mark.gpu
def test_reference_hound_steps_on_libb200gym():
    """The reference's flat-terrain ``Hound`` class (tasks/hound.py), unmodified, on the shim: contract, finiteness, resets under random actions."""
    import torch

    import isaacgymenv_b200 as b2g

    b2g.install_isaacgym_shim(REF)
    vt = importlib.import_module("isaacgymenvs.tasks.base.vec_task")
    vt.EXISTING_SIM = None
    n = 64
    torch.manual_seed(2)
    ref = importlib.import_module("isaacgymenvs.tasks.hound").Hound(cfg=_task_cfg("Hound", n), rl_device="cuda:0", sim_device="cuda:0", graphics_device_id=-1,
                                                                    headless=True, virtual_screen_capture=False, force_render=False)
    assert ref.obs_buf.shape == (n, 48) and ref.num_actions == 12
    g = torch.Generator(device="cuda:0").manual_seed(3)
    resets = 0
    for k in range(80):
        o, r, d, e = ref.step(2 * torch.rand(n, 12, device="cuda:0", generator=g) - 1)
        assert torch.isfinite(o["obs"]).all() and torch.isfinite(r).all()
        resets += int(d.sum())
    assert resets > 0
    vt.EXISTING_SIM = None


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("module,cls,task,nact", [("tasks.anymal_terrain", "AnymalTerrain", "AnymalTerrain", 12), ("tasks.Hound_terrain", "HoundTerrain", "HoundTerrain", 12)])
def test_fused_terrain_step_equals_the_reference_class_on_the_same_physics(module, cls, task, nact):
    """The whole policy step of BASELINE config 3, two ways on the same state and actions: (a) the REFERENCE's class, unmodified -- its own
    decimation loop (explicit PD torques, set_dof_actuation_force_tensor, gym.simulate x 4 + the base class's fifth), its post_physics_step
    in torch / TorchScript -- on the library's generic kernels; (b) this package's fused task (k_terrain_phys + k_terrain_post).  Noise and
    pushes off (they draw from different generators), environments that reset in the step left out (so do the reset draws).  What must
    agree: observations incl. the height scan, reward, the reset decision -- to the rounding two instantiations of the same sub-step code
    leave after five sim steps."""
    import torch

    import isaacgymenv_b200 as b2g

    b2g.install_isaacgym_shim(REF)
    vt = importlib.import_module("isaacgymenvs.tasks.base.vec_task")
    n = 64

    def cfg_of():
        cfg = _task_cfg(task, n)
        cfg["env"]["learn"]["addNoise"] = False
        cfg["env"]["learn"]["pushInterval_s"] = 1.0e6
        return cfg

    vt.EXISTING_SIM = None
    ref_cls = getattr(importlib.import_module("isaacgymenvs." + module), cls)       # first import draws from torch's generator: do it before seeding
    torch.manual_seed(9)
    ref = ref_cls(cfg=cfg_of(), rl_device="cuda:0", sim_device="cuda:0", graphics_device_id=-1, headless=True, virtual_screen_capture=False, force_render=False)
    torch.manual_seed(9)
    ov = {"env": {"learn": {"addNoise": False, "pushInterval_s": 1.0e6}}}
    ours = b2g.make(seed=9, task=task, num_envs=n, sim_device="cuda:0", rl_device="cuda:0", headless=True, overrides=ov)
    # same physics needs the same per-environment friction: both classes draw the 100 buckets first thing after the seed (anymal_terrain.py:239)
    assert max(abs(x - y) for x, y in zip(ref.sim.env_friction, ours.sim.env_friction)) < 1e-6
    g = torch.Generator(device="cuda:0").manual_seed(4)
    compared, dobs, dscan, drew = 0, [], [], []
    for k in range(30):
        for name in ("root_states", "dof_state", "commands", "last_actions", "last_dof_vel", "feet_air_time", "progress_buf"):
            getattr(ours, name).copy_(getattr(ref, name).view_as(getattr(ours, name)))
        ours.reset_buf.copy_(ref.reset_buf.to(ours.reset_buf.dtype))
        a = 0.6 * (2 * torch.rand(n, nact, device="cuda:0", generator=g) - 1)
        o_r, r_r, d_r, _ = ref.step(a)
        o_o, r_o, d_o, _ = ours.step(a)
        d_r, d_o = d_r != 0, d_o != 0
        keep = ~d_r & ~d_o
        # a contact force within rounding of the 1 N termination threshold may fall either side of it in the two paths
        assert int((d_r != d_o).sum()) <= 1, (k, int((d_r != d_o).sum()))
        if int(keep.sum()) == 0:
            continue
        compared += int(keep.sum())
        dd = (o_r["obs"][keep] - o_o["obs"][keep]).abs()
        dobs.append(torch.cat([dd[:, :48], dd[:, 188:]], dim=1).flatten())      # base velocities, gravity, commands, joint state, actions
        dscan.append(dd[:, 48:188].flatten())                                    # the 140-point height scan
        drew.append((r_r[keep] - r_o[keep]).abs())
    assert compared > 10 * n
    dobs, dscan, drew = torch.cat(dobs), torch.cat(dscan), torch.cat(drew)
    # height scan on the default plane: 5 x (root z - 0.5), so this is the root height's agreement
    scan_off = float((dscan > 1e-3).float().mean())
    assert scan_off < 0.01, scan_off
    # one step = five contact-rich sim steps through two differently compiled instantiations of the same code: nearly every entry agrees to
    # float32 rounding, a foot that makes or breaks contact a sub-step earlier in one of them shows up in that environment's velocities
    q = lambda t, p: float(t.float().kthvalue(max(1, int(p * t.numel()))).values)
    stats = dict(obs_median=q(dobs, 0.5), obs_p99=q(dobs, 0.99), obs_p999=q(dobs, 0.999), obs_max=float(dobs.max()), rew_p99=q(drew, 0.99), rew_max=float(drew.max()))
    print(f"{task}: {compared} env-steps compared, scan points off {scan_off:.4f}", {k_: f"{v:.2e}" for k_, v in stats.items()})
    # measured on a B200 (profiles/r02_dropin_fused_vs_reference_class.log): median 0, p99 5e-7, p99.9 2e-6 (Anymal) / 1e-3 (Hound), max 8e-5 / 2.5e-2
    assert stats["obs_median"] < 1e-6 and stats["obs_p99"] < 1e-4 and stats["obs_p999"] < 1e-2 and stats["obs_max"] < 0.2, stats
    assert stats["rew_p99"] < 1e-5 and stats["rew_max"] < 5e-3, stats
    vt.EXISTING_SIM = None
