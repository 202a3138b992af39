"""Tensorised domain randomisation -- the B200-side replacement of ``VecTask.apply_randomizations``
(reference ``tasks/base/vec_task.py:610-840``; samplers ``utils/dr_utils.py:71-133``; parameters ``cfg/task/Anymal.yaml:104-170``).

The reference walks over environments in Python and pushes property structs through Isaac Gym setters one actor at a time
("this part is not tensorised yet", ``vec_task.py:752``).  Here every physical parameter that the step kernels read per
environment is one device tensor, and a randomisation pass is a handful of masked tensor writes without host synchronisation:

=====================================  ==========================================================================
reference parameter                    tensor it lands in
=====================================  ==========================================================================
``rigid_body_properties.mass``         ``B2G_T_LINK_SCALE[:, l, 0]``: one factor per rigid body, as the reference draws them (``dr_utils.py:135-238``
                                       walks every entry of the property list); it scales the mass and the inertia of the link the body rides
                                       on (Isaac Gym's setter recomputes inertia, ``dr_utils.py:63``); bodies merged into one link share its
                                       first body's factor
``dof_properties.stiffness``           ``B2G_T_LINK_SCALE[:, 1 + d, 1]``: one factor per DOF
``dof_properties.damping``             ``B2G_T_LINK_SCALE[:, 1 + d, 2]``
``dof_properties.lower`` / ``upper``   ``B2G_T_LINK_SCALE[:, 1 + d, 3 / 4]``: additive offsets of the joint limits
``rigid_shape_properties.friction``    ``B2G_T_FRICTION`` (bucketed like ``dr_utils.get_bucketed_val``)
``sim_params.gravity``                 ``b2g_sim_set_params`` (global, host side)
``observations`` / ``actions``         noise closures applied by ``VecTask.step`` (same maths as ``vec_task.py:648-718``)
=====================================  ==========================================================================

``B2G_T_ENV_SCALE`` (one mass / stiffness / damping factor per environment) stays available to callers that want a single factor;
the kernels multiply the two.  Not modelled (recorded in :attr:`skipped`): ``color``, ``scale``, ``restitution`` (the contact model has
e = 0), tendon properties, external parameter generators.

Faithful quirk: this fork never advances ``randomize_buf`` (``vec_task.py:322,632-635`` are its only uses), so after the first
pass only the non-environment parameters (noise, gravity) are re-drawn every ``frequency`` frames; ``count_steps=True`` restores
upstream's per-step increment so that physical parameters are re-drawn on resets.
"""
from __future__ import annotations

import math
import operator
from typing import Any, Dict, List

import torch

from .. import _abi

LINK_COLUMN = {"mass": 0, "stiffness": 1, "damping": 2, "lower": 3, "upper": 4}


def schedule_scaling(params: Dict[str, Any], step: int) -> float:
    """``dr_utils.py:81-86`` / ``vec_task.py:655-661``."""
    sched = params.get("schedule", None)
    if sched == "linear":
        return 1.0 / params["schedule_steps"] * min(step, params["schedule_steps"])
    if sched == "constant":
        return 0.0 if step < params["schedule_steps"] else 1.0
    return 1.0


def sample(params: Dict[str, Any], shape, step: int, device, generator=None) -> torch.Tensor:
    """``dr_utils.generate_random_samples`` on the device."""
    lo, hi = float(params["range"][0]), float(params["range"][1])
    dist, op = params["distribution"], params["operation"]
    s = schedule_scaling(params, step)
    if op == "additive":
        lo, hi = lo * s, hi * s
    elif op == "scaling":
        if dist == "gaussian":
            lo, hi = lo * s + 1.0 * (1.0 - s), hi * s            # mean interpolates to 1, spread grows with the schedule
        else:
            lo, hi = lo * s + 1.0 * (1.0 - s), hi * s + 1.0 * (1.0 - s)
    else:
        raise ValueError(f"unknown operation {op!r}")
    if dist == "gaussian":
        return lo + hi * torch.randn(shape, device=device, generator=generator)
    if dist == "uniform":
        return lo + (hi - lo) * torch.rand(shape, device=device, generator=generator)
    if dist == "loguniform":
        return torch.exp(math.log(lo) + (math.log(hi) - math.log(lo)) * torch.rand(shape, device=device, generator=generator))
    raise ValueError(f"unknown distribution {dist!r}")


def bucketed(values: torch.Tensor, params: Dict[str, Any]) -> torch.Tensor:
    """``dr_utils.get_bucketed_val``: snap to the lower edge of one of ``num_buckets`` equal bins over the range."""
    nb = int(params.get("num_buckets", 0) or 0)
    if nb <= 0:
        return values
    if params["distribution"] == "uniform":
        lo, hi = float(params["range"][0]), float(params["range"][1])
    else:
        lo = float(params["range"][0]) - 2.0 * math.sqrt(float(params["range"][1]))
        hi = float(params["range"][0]) + 2.0 * math.sqrt(float(params["range"][1]))
    width = (hi - lo) / nb
    # bisect(buckets, v) - 1: values below the first edge wrap to the LAST bucket (python index -1), as in the reference
    idx = torch.floor((values - lo) / width).long()
    idx = torch.where(idx < 0, torch.full_like(idx, nb - 1), torch.clamp(idx, max=nb - 1))
    return lo + width * idx.to(values.dtype)


class DomainRandomizer:
    def __init__(self, task, count_steps: bool = False):
        self.task = task
        self.device = task.device
        self.count_steps = bool(count_steps)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(getattr(task, "seed", 0) or 0) + 0x5EED)
        self.first = True
        self.last_rand_step = 0
        self.skipped: List[str] = []
        self.env_scale = None
        self.link_scale = None
        self.friction = None
        self.friction0 = None
        self.gravity0 = None
        self.applied_envs = 0          # host-visible only when queried (kept as a device tensor)
        self._applied = torch.zeros((), device=self.device, dtype=torch.long)

    # ---- tensors the kernels read ----
    def _tensors(self):
        if self.env_scale is None:
            gym, sim = self.task.gym, self.task.sim
            self.env_scale = gym._tensor(sim, _abi.T_ENV_SCALE)
            self.link_scale = gym._tensor(sim, _abi.T_LINK_SCALE)           # (N, nd + 1, 6)
            self.friction = gym._tensor(sim, _abi.T_FRICTION)
            self.friction0 = self.friction.clone()
            # rigid body b -> the link it rides on; a link keeps the factor of its first body
            art = getattr(getattr(sim, "asset", None), "art", None)
            nl = self.link_scale.shape[1]
            first = torch.full((nl,), -1, dtype=torch.long)
            if art is not None:
                for b, l in enumerate(art.body_link):
                    if first[int(l)] < 0:
                        first[int(l)] = b
                self.n_bodies = len(art.body_link)
            else:
                first = torch.arange(nl)
                self.n_bodies = nl
            self.link_first_body = first.clamp_min(0).to(self.device)
        return self.link_scale, self.friction

    def _skip(self, what: str):
        if what not in self.skipped:
            self.skipped.append(what)

    # ---- one pass ----
    def apply(self, dr_params: Dict[str, Any], reset_mask: torch.Tensor = None):
        """``reset_mask``: environments being reset now (default: ``task.reset_buf != 0``)."""
        task = self.task
        freq = dr_params.get("frequency", 1)
        step = int(task.gym.get_frame_count(task.sim))
        n = task.num_envs
        if self.first:
            do_nonenv = True
            mask = torch.ones(n, dtype=torch.bool, device=self.device)
        else:
            do_nonenv = (step - self.last_rand_step) >= freq
            if reset_mask is None:
                reset_mask = task.reset_buf != 0
            mask = (task.randomize_buf >= freq) & reset_mask.to(self.device).bool()
            task.randomize_buf.masked_fill_(mask, 0)
        if do_nonenv:
            self.last_rand_step = step
        for name in ("observations", "actions"):
            if name in dr_params and do_nonenv:
                task.dr_randomizations[name] = self._noise_closure(name, dr_params[name], step)
        if "sim_params" in dr_params and do_nonenv:
            self._sim_params(dr_params["sim_params"], step)
        for actor, props in (dr_params.get("actor_params", {}) or {}).items():
            self._actor(actor, props, mask, step)
        self._applied += mask.sum()
        self.first = False

    def _noise_closure(self, name, p, step):
        dist, op_type = p["distribution"], p["operation"]
        op = operator.add if op_type == "additive" else operator.mul
        s = schedule_scaling(p, step)
        a, b = float(p["range"][0]), float(p["range"][1])
        ac, bc = (float(x) for x in p.get("range_correlated", [0.0, 0.0]))
        if dist == "gaussian":
            if op_type == "additive":
                a, b, ac, bc = a * s, b * s, ac * s, bc * s
            else:
                b, a = b * s, a * s + 1.0 * (1.0 - s)
                bc, ac = bc * s, ac * s + 1.0 * (1.0 - s)
            state = {"mu": a, "var": b, "mu_corr": ac, "var_corr": bc}

            def noise_lambda(tensor, state=state):
                corr = state.get("corr", None)
                if corr is None:
                    corr = torch.randn(tensor.shape, device=tensor.device, generator=self.gen)
                    state["corr"] = corr
                corr = corr * state["var_corr"] + state["mu_corr"]
                return op(tensor, corr + torch.randn(tensor.shape, device=tensor.device, generator=self.gen) * state["var"] + state["mu"])
        elif dist == "uniform":
            if op_type == "additive":
                a, b, ac, bc = a * s, b * s, ac * s, bc * s
            else:
                a, b = a * s + 1.0 * (1.0 - s), b * s + 1.0 * (1.0 - s)
                ac, bc = ac * s + 1.0 * (1.0 - s), bc * s + 1.0 * (1.0 - s)
            state = {"lo": a, "hi": b, "lo_corr": ac, "hi_corr": bc}

            def noise_lambda(tensor, state=state):
                corr = state.get("corr", None)
                if corr is None:
                    corr = torch.randn(tensor.shape, device=tensor.device, generator=self.gen)
                    state["corr"] = corr
                corr = corr * (state["hi_corr"] - state["lo_corr"]) + state["lo_corr"]
                return op(tensor, corr + torch.rand(tensor.shape, device=tensor.device, generator=self.gen) * (state["hi"] - state["lo"]) + state["lo"])
        else:
            raise ValueError(f"{name}: unknown distribution {dist!r}")
        state["noise_lambda"] = noise_lambda
        return state

    def _sim_params(self, attrs, step):
        gym, sim = self.task.gym, self.task.sim
        prop = gym.get_sim_params(sim)
        if self.gravity0 is None:
            self.gravity0 = (prop.gravity.x, prop.gravity.y, prop.gravity.z)
        for attr, p in attrs.items():
            if attr != "gravity":
                self._skip(f"sim_params.{attr}")
                continue
            smp = sample(p, (3,), step, "cpu").tolist()       # three host floats: the parameter block is host state
            g0 = self.gravity0
            if p["operation"] == "scaling":
                prop.gravity.x, prop.gravity.y, prop.gravity.z = g0[0] * smp[0], g0[1] * smp[1], g0[2] * smp[2]
            else:
                prop.gravity.x, prop.gravity.y, prop.gravity.z = g0[0] + smp[0], g0[1] + smp[1], g0[2] + smp[2]
        gym.set_sim_params(sim, prop)

    def _actor(self, actor, props, mask, step):
        scale, friction = self._tensors()
        n = self.task.num_envs
        sim_started = not self.first
        for prop_name, attrs in props.items():
            if prop_name in ("color", "scale"):
                self._skip(f"{actor}.{prop_name}")
                continue
            if not isinstance(attrs, dict):
                continue
            for attr, p in attrs.items():
                if p.get("setup_only", False) and sim_started:
                    continue
                if prop_name == "rigid_shape_properties" and attr == "friction":
                    smp = sample(p, (n,), step, self.device, self.gen)
                    new = self.friction0 * smp if p["operation"] == "scaling" else self.friction0 + smp
                    friction.copy_(torch.where(mask, bucketed(new, p), friction))
                elif prop_name == "rigid_body_properties" and attr == "mass":
                    if p["operation"] != "scaling":
                        self._skip(f"{actor}.{prop_name}.{attr} (additive)")
                        continue
                    smp = sample(p, (n, self.n_bodies), step, self.device, self.gen)[:, self.link_first_body]      # (n, links)
                    scale[:, :, 0] = torch.where(mask[:, None], smp.to(scale.dtype), scale[:, :, 0])
                elif prop_name == "dof_properties" and attr in ("stiffness", "damping", "lower", "upper"):
                    col = LINK_COLUMN[attr]
                    additive = attr in ("lower", "upper")
                    if (p["operation"] == "scaling") == additive:       # the kernels scale gains and offset limits
                        self._skip(f"{actor}.{prop_name}.{attr} ({p['operation']})")
                        continue
                    smp = sample(p, (n, scale.shape[1] - 1), step, self.device, self.gen)                              # one draw per DOF
                    scale[:, 1:, col] = torch.where(mask[:, None], smp.to(scale.dtype), scale[:, 1:, col])
                else:
                    self._skip(f"{actor}.{prop_name}.{attr}")

    def num_applied(self) -> int:
        return int(self._applied)
