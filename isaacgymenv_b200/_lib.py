"""Loader + thin ctypes wrapper of ``libb200gym.so`` (the CUDA library behind the gym API shim).

There is no CPU fallback: if the library is missing or no CUDA device is usable the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2G_LIB_PATH: development switch for A/B timing of experimental builds (tools/build_variant.sh)
LIB_PATH = os.environ.get("B2G_LIB_PATH") or os.path.join(_HERE, "lib", "libb200gym.so")
_lib = None


class B2GError(RuntimeError):
    pass


def load():
    """Load the shared library (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise B2GError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.b2g_last_error.restype = C.c_char_p
    lib.b2g_sim_launch_count.restype = C.c_int64
    lib.b2g_sim_launch_count.argtypes = [C.c_void_p]
    lib.b2g_sim_contact_stats.restype = C.c_int
    lib.b2g_sim_contact_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int]
    vp, ip = C.c_void_p, C.c_int
    protos = {
        "b2g_abi_version": [],
        "b2g_sizeof": [ip],
        "b2g_sim_create": [ip, C.POINTER(_abi.SimParams), C.POINTER(vp)],
        "b2g_sim_destroy": [vp],
        "b2g_sim_set_params": [vp, C.POINTER(_abi.SimParams)],
        "b2g_sim_get_params": [vp, C.POINTER(_abi.SimParams)],
        "b2g_sim_add_ground": [vp, C.c_float, C.c_float, C.c_float],
        "b2g_sim_add_heightfield": [vp, C.POINTER(_abi.Heightfield), vp],
        "b2g_sim_add_articulation": [vp, C.POINTER(_abi.Model), C.POINTER(_abi.DofProps), ip, C.POINTER(C.c_float), C.c_float, ip],
        "b2g_sim_prepare": [vp],
        "b2g_sim_set_dof_props": [vp, C.POINTER(_abi.DofProps)],
        "b2g_sim_tensor": [vp, ip, C.POINTER(_abi.TensorDesc)],
        "b2g_sim_simulate": [vp, vp],
        "b2g_sim_refresh": [vp, ip, vp],
        "b2g_sim_set_indexed": [vp, ip, vp, vp, ip, vp],
        "b2g_sim_set_tensor": [vp, ip, vp, vp],
        "b2g_sim_forward_dynamics": [vp, vp, vp, vp],
        "b2g_task_anymal_create": [vp, C.POINTER(_abi.AnymalCfg)],
        "b2g_task_tensor": [vp, ip, C.POINTER(_abi.TensorDesc)],
        "b2g_task_anymal_reset_all": [vp, vp],
        "b2g_task_anymal_step": [vp, vp, vp],
        "b2g_task_anymal_post_only": [vp, vp, vp],
        "b2g_task_set_rand_override": [vp, ip],
        "b2g_task_anymal_step_host": [vp, vp, vp, vp, vp, vp, vp],
        "b2g_task_cartpole_create": [vp, C.POINTER(_abi.CartpoleCfg)],
        "b2g_task_houndarm_create": [vp, C.POINTER(_abi.HoundarmCfg)],
        "b2g_task_terrain_create": [vp, C.POINTER(_abi.TerrainCfg), vp, vp],
        "b2g_task_terrain_set_step": [vp, C.c_int64],
        "b2g_task_terrain_set_init_done": [vp, ip],
        "b2g_task_terrain_device_step": [vp, ip],
        "b2g_task_step": [vp, vp, vp],
        "b2g_task_post_only": [vp, vp, vp],
        "b2g_task_osc_probe": [vp, vp, vp],
        "b2g_task_step_host": [vp, vp, vp, vp, vp, vp, vp],
        "b2g_dlpack_from_desc": [C.POINTER(_abi.TensorDesc), C.POINTER(vp)],
        "b2g_task_host_layout": [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
        "b2g_policy_create": [ip, ip, C.POINTER(C.c_int), ip, C.POINTER(vp)],
        "b2g_policy_set_layer": [vp, ip, vp, vp, vp],
        "b2g_policy_set_obs_norm": [vp, vp, vp, C.c_float, C.c_float, vp],
        "b2g_policy_forward": [vp, vp, ip, vp, vp, vp],
    }
    lib.b2g_policy_destroy.argtypes = [vp]
    lib.b2g_policy_destroy.restype = None
    lib.b2g_policy_launch_count.argtypes = [vp]
    lib.b2g_policy_launch_count.restype = C.c_int64
    for name, args in protos.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    if lib.b2g_abi_version() != _abi.B2G_ABI_VERSION:
        raise B2GError("libb200gym.so ABI version mismatch; rebuild")
    for which, st in enumerate((_abi.Model, _abi.SimParams, _abi.DofProps, _abi.Heightfield, _abi.TensorDesc, _abi.AnymalCfg, _abi.CartpoleCfg, _abi.TerrainCfg, _abi.HoundarmCfg)):
        if lib.b2g_sizeof(which) != C.sizeof(st):
            raise B2GError(f"struct layout mismatch for {st.__name__}: C {lib.b2g_sizeof(which)} vs ctypes {C.sizeof(st)}")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().b2g_last_error().decode("utf-8", "replace")
        raise B2GError(f"{what or 'libb200gym'} failed ({rc}): {msg}")


EXPORTED_SYMBOLS = [
    "b2g_abi_version", "b2g_last_error", "b2g_sim_create", "b2g_sim_destroy", "b2g_sim_set_params", "b2g_sim_get_params",
    "b2g_sim_add_ground", "b2g_sim_add_heightfield", "b2g_sim_add_articulation", "b2g_sim_prepare", "b2g_sim_set_dof_props",
    "b2g_sim_tensor", "b2g_sim_simulate", "b2g_sim_refresh", "b2g_sim_set_indexed", "b2g_sim_set_tensor",
    "b2g_sim_forward_dynamics", "b2g_task_anymal_create", "b2g_task_tensor", "b2g_task_anymal_reset_all", "b2g_task_anymal_step", "b2g_task_set_rand_override", "b2g_task_anymal_post_only", "b2g_task_anymal_step_host", "b2g_task_cartpole_create", "b2g_task_houndarm_create", "b2g_task_terrain_create", "b2g_task_terrain_set_step", "b2g_task_terrain_set_init_done", "b2g_task_terrain_device_step", "b2g_task_step", "b2g_task_post_only", "b2g_task_osc_probe", "b2g_task_step_host", "b2g_sim_launch_count", "b2g_sim_contact_stats", "b2g_sizeof", "b2g_dlpack_from_desc", "b2g_task_host_layout",
    "b2g_ppo_head", "b2g_ppo_head_workspace_floats", "b2g_adam_clip_step", "b2g_running_stat_update", "b2g_stat_workspace_doubles", "b2g_normalize_store", "b2g_rollout_sample", "b2g_rollout_counter_advance", "b2g_rollout_post", "b2g_gae_finish",
    "b2g_mlp_bias_elu", "b2g_mlp_elu_backward", "b2g_mlp_elu_backward_workspace_floats", "b2g_mlp_heads_forward", "b2g_mlp_heads_backward", "b2g_mlp_heads_backward_workspace_floats", "b2g_mlp_heads_backward_scatter", "b2g_gather_rows", "b2g_random_permutation",
    "b2g_policy_create", "b2g_policy_destroy", "b2g_policy_set_layer", "b2g_policy_set_obs_norm", "b2g_policy_forward", "b2g_policy_launch_count",
]


def desc_to_torch(desc: _abi.TensorDesc):
    """``gymtorch.wrap_tensor``: non-owning torch view of a sim-owned device buffer, through DLPack."""
    import torch

    managed = C.c_void_p()
    check(load().b2g_dlpack_from_desc(C.byref(desc), C.byref(managed)), "dlpack export")
    C.pythonapi.PyCapsule_New.restype = C.py_object
    C.pythonapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    capsule = C.pythonapi.PyCapsule_New(managed, b"dltensor", None)
    return torch.from_dlpack(capsule)
