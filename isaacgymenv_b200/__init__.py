"""isaacgymenv_b200 -- B200-native vectorised environment step behind the IsaacGymEnvs task API.

``make()`` mirrors the reference's ``isaacgymenvs.make`` (``isaacgymenvs/__init__.py:14-55``): it loads the task's
yaml config, applies ``num_envs`` and returns the task object (a ``VecTask``), ready for ``reset()`` / ``step()``.
"""
from __future__ import annotations

import copy
import os
from typing import Any, Dict, Optional

__all__ = ["make", "load_task_config", "task_map", "install_isaacgym_shim"]

_CFG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg", "task")


def task_map():
    """Name -> class, like ``isaacgymenvs.tasks.isaacgym_task_map`` (hot-path tasks only)."""
    from .tasks.anymal import Anymal
    from .tasks.hound import Hound

    m = {"Anymal": Anymal, "Hound": Hound}
    try:
        from .tasks.cartpole import Cartpole

        m["Cartpole"] = Cartpole
    except ImportError:
        pass
    try:
        from .tasks.anymal_terrain import AnymalTerrain
        from .tasks.hound_terrain import HoundTerrain

        m["AnymalTerrain"] = AnymalTerrain
        m["HoundTerrain"] = HoundTerrain
        from .tasks.useful_hound import UsefulHound

        m["UsefulHound"] = UsefulHound
    except ImportError:
        pass
    try:
        from .tasks.hound_arm import Houndarm, Manipulator

        m["Houndarm"] = Houndarm
        m["Manipulator"] = Manipulator
    except ImportError:
        pass
    return m


def _deep_update(d: Dict[str, Any], u: Dict[str, Any]):
    for k, v in u.items():
        if isinstance(v, dict) and isinstance(d.get(k), dict):
            _deep_update(d[k], v)
        else:
            d[k] = v


def load_task_config(task: str, overrides: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
    """cfg/task/<task>.yaml as a plain dict (what the reference hands to the task after ``omegaconf_to_dict``)."""
    import yaml

    path = os.path.join(_CFG_DIR, f"{task}.yaml")
    if not os.path.isfile(path):
        raise ValueError(f"unknown task {task!r}: no {path}")
    with open(path) as fh:
        cfg = yaml.safe_load(fh)
    if overrides:
        _deep_update(cfg, copy.deepcopy(overrides))
    return cfg


def make(seed: int, task: str, num_envs: int, sim_device: str, rl_device: str, graphics_device_id: int = -1, headless: bool = True,
         multi_gpu: bool = False, virtual_screen_capture: bool = False, force_render: bool = False, cfg: Optional[Dict[str, Any]] = None,
         overrides: Optional[Dict[str, Any]] = None):
    """Create a task.  Same positional arguments as the reference's ``isaacgymenvs.make``; ``cfg`` may be a ready task
    config dict, otherwise cfg/task/<task>.yaml is loaded and ``num_envs`` applied.  With ``multi_gpu`` the device
    becomes ``cuda:$LOCAL_RANK`` and the seed is offset by the rank (``utils/rlgames_utils.py:89-107``,
    ``utils/utils.py:89-94``)."""
    if cfg is None:
        cfg_dict = load_task_config(task, overrides)
        if num_envs is not None and num_envs != "":       # '' / None = the task yaml's own count (cfg/config.yaml:8 num_envs: '')
            cfg_dict["env"]["numEnvs"] = int(num_envs)
    else:
        cfg_dict = copy.deepcopy(cfg)
        if overrides:
            _deep_update(cfg_dict, copy.deepcopy(overrides))
    if multi_gpu:
        local_rank = int(os.getenv("LOCAL_RANK", "0"))
        global_rank = int(os.getenv("RANK", "0"))
        sim_device = rl_device = f"cuda:{local_rank}"
        cfg_dict["rank"] = local_rank
        cfg_dict["rl_device"] = rl_device
        seed = seed + global_rank
    cfg_dict["seed"] = int(seed)
    on_gpu = sim_device.startswith(("cuda", "gpu"))
    cfg_dict["sim"]["use_gpu_pipeline"] = bool(cfg_dict["sim"].get("use_gpu_pipeline", True)) and on_gpu
    cfg_dict["sim"].setdefault("physx", {})["use_gpu"] = on_gpu
    cls = task_map()[cfg_dict["name"]]
    return cls(cfg=cfg_dict, rl_device=rl_device, sim_device=sim_device, graphics_device_id=graphics_device_id, headless=headless,
               virtual_screen_capture=virtual_screen_capture, force_render=force_render)


def install_isaacgym_shim(reference_root: Optional[str] = None):
    """Register this package as the ``isaacgym`` the reference's task files import, so that they run UNMODIFIED on libb200gym:

    * ``isaacgym.gymapi`` / ``isaacgym.gymtorch``: this package's gym shim over the C ABI;
    * ``isaacgym.terrain_utils``: the terrain generator (``from isaacgym.terrain_utils import *``, tasks/anymal_terrain.py:542);
    * ``isaacgym.torch_utils``: the tensor helpers; ``isaacgym.gymutil``: viewer-side helpers as no-ops (headless only);
    * ``gym`` / ``gym.spaces`` (tasks/base/vec_task.py:34-35) when OpenAI gym is not installed; ``np.Inf`` (vec_task.py:107, gone in numpy 2).

    With ``reference_root`` (a checkout of the reference) the ``isaacgymenvs`` package is made importable WITHOUT running its
    ``__init__`` (which pulls in hydra / omegaconf / rl_games, ``isaacgymenvs/__init__.py:1-5``), so
    ``from isaacgymenvs.tasks.anymal import Anymal`` gives the reference's own class, ready to be constructed on this engine."""
    import importlib.util
    import sys
    import types

    import numpy as np

    from . import gymapi, gymtorch, spaces, terrain
    from .utils import torch_math

    mod = sys.modules.get("isaacgym")
    if mod is None or not getattr(mod, "_b2g_shim", False):
        mod = types.ModuleType("isaacgym")
        mod._b2g_shim = True
        sys.modules["isaacgym"] = mod
    gymutil = types.ModuleType("isaacgym.gymutil")

    class _Geometry:           # debug-visualisation helpers: constructed by viewer code only
        def __init__(self, *a, **k):
            pass

    gymutil.AxesGeometry = gymutil.WireframeSphereGeometry = gymutil.WireframeBoxGeometry = _Geometry
    gymutil.draw_lines = lambda *a, **k: None
    gymutil.parse_arguments = lambda *a, **k: types.SimpleNamespace()
    for name, m in (("gymapi", gymapi), ("gymtorch", gymtorch), ("terrain_utils", terrain), ("torch_utils", torch_math), ("gymutil", gymutil)):
        setattr(mod, name, m)
        sys.modules["isaacgym." + name] = m
    if "gym" not in sys.modules and importlib.util.find_spec("gym") is None:
        g = types.ModuleType("gym")
        g.spaces = spaces
        g.Space, g.Env = spaces.Space, object
        g._b2g_shim = True
        sys.modules["gym"], sys.modules["gym.spaces"] = g, spaces
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    if reference_root:
        base = os.path.join(reference_root, "isaacgymenvs")
        if not os.path.isdir(base):
            raise FileNotFoundError(f"no isaacgymenvs package under {reference_root}")
        for name, path in (("isaacgymenvs", base), ("isaacgymenvs.tasks", os.path.join(base, "tasks")),
                           ("isaacgymenvs.tasks.base", os.path.join(base, "tasks", "base")), ("isaacgymenvs.utils", os.path.join(base, "utils"))):
            if name not in sys.modules:
                pkg = types.ModuleType(name)
                pkg.__path__ = [path]
                sys.modules[name] = pkg
    return mod
