"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes front-end of ``oracle/_build/liboracle_dyn.so``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product path (``isaacgymenv_b200``) never does.

PARITY UNPINNED for dynamics (see ``oracle/dyn/oracle_dyn_impl.h``): this is a CPU restatement of the
published rigid-body algorithms, not PhysX; the reference has no golden vector at this boundary.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_dyn.so")
_lib = None


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, "dyn", f) for f in ("oracle_dyn.c", "oracle_dyn_impl.h", "oracle_task_impl.h")]
    src.append(os.path.join(_HERE, "..", "include", "b200gym.h"))
    stale = (not os.path.isfile(_LIB_PATH)) or any(
        os.path.isfile(s) and os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        try:
            build()
        except Exception:
            if not os.path.isfile(_LIB_PATH):
                raise
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def _suffix_types(dtype):
    if np.dtype(dtype) == np.float64:
        return "_f64", C.c_double
    if np.dtype(dtype) == np.float32:
        return "_f32", C.c_float
    raise TypeError(dtype)


def simulate(model, params, props, root, dof, target, actuation, heightfield=None, hf_samples=None, friction=None):
    """One ``gym.simulate``: updates ``root`` (N,13) and ``dof`` (N,nd,2) in place; returns
    (dof_force (N,nd), contact (N,nb,3)).  dtype of ``root`` selects the f64 or f32 instantiation."""
    suf, ct = _suffix_types(root.dtype)
    n = root.shape[0]
    nd, nb = model.n_dof, model.n_bodies
    dof_force = np.zeros((n, nd), dtype=root.dtype)
    contact = np.zeros((n, nb, 3), dtype=root.dtype)
    fr = None if friction is None else np.ascontiguousarray(friction, dtype=np.float32)
    hs = None if hf_samples is None else np.ascontiguousarray(hf_samples, dtype=np.int16)
    fn = getattr(lib(), "orc_simulate" + suf)
    fn.restype = C.c_int
    rc = fn(C.byref(model), C.byref(params), C.byref(props),
            C.byref(heightfield) if heightfield is not None else None, _ptr(hs, C.c_int16), _ptr(fr, C.c_float),
            C.c_int(n), _ptr(root, ct), _ptr(dof, ct),
            _ptr(np.ascontiguousarray(target, dtype=root.dtype), ct),
            _ptr(np.ascontiguousarray(actuation, dtype=root.dtype), ct), _ptr(dof_force, ct), _ptr(contact, ct))
    if rc != 0:
        raise RuntimeError(f"oracle simulate failed ({rc})")
    return dof_force, contact


def simulate_ref(model, params, props, root, dof, target, actuation, heightfield=None, hf_samples=None, friction=None, hard_limits=False,
                 max_iter=500, tol=1e-9):
    """One ``gym.simulate`` with the CONVERGED reference solver (every candidate a contact, sequential Gauss-Seidel to
    convergence, optional hard joint limits): what the production solver's error is measured against.  Updates ``root`` /
    ``dof`` in place; returns (dof_force, contact, info (N,4) = [contacts, position sweeps, velocity sweeps, hit max_iter])."""
    suf, ct = _suffix_types(root.dtype)
    n = root.shape[0]
    nd, nb = model.n_dof, model.n_bodies
    dof_force = np.zeros((n, nd), dtype=root.dtype)
    contact = np.zeros((n, nb, 3), dtype=root.dtype)
    info = np.zeros((n, 4), dtype=np.int32)
    fr = None if friction is None else np.ascontiguousarray(friction, dtype=np.float32)
    hs = None if hf_samples is None else np.ascontiguousarray(hf_samples, dtype=np.int16)
    fn = getattr(lib(), "orc_simulate_ref" + suf)
    fn.restype = C.c_int
    rc = fn(C.byref(model), C.byref(params), C.byref(props),
            C.byref(heightfield) if heightfield is not None else None, _ptr(hs, C.c_int16), _ptr(fr, C.c_float),
            C.c_int(n), _ptr(root, ct), _ptr(dof, ct),
            _ptr(np.ascontiguousarray(target, dtype=root.dtype), ct),
            _ptr(np.ascontiguousarray(actuation, dtype=root.dtype), ct), _ptr(dof_force, ct), _ptr(contact, ct),
            C.c_int(1 if hard_limits else 0), C.c_int(int(max_iter)), C.c_double(float(tol)), info.ctypes.data_as(C.POINTER(C.c_int)))
    if rc != 0:
        raise RuntimeError(f"oracle reference simulate failed ({rc})")
    return dof_force, contact, info


def forward_dynamics(model, params, root, dof, tau):
    suf, ct = _suffix_types(root.dtype)
    n = root.shape[0]
    qdd = np.zeros((n, model.n_dof), dtype=root.dtype)
    a0 = np.zeros((n, 6), dtype=root.dtype)
    fn = getattr(lib(), "orc_forward_dynamics" + suf)
    fn.restype = C.c_int
    rc = fn(C.byref(model), C.byref(params), C.c_int(n), _ptr(root, ct), _ptr(dof, ct),
            _ptr(np.ascontiguousarray(tau, dtype=root.dtype), ct), _ptr(qdd, ct), _ptr(a0, ct))
    if rc != 0:
        raise RuntimeError(f"oracle forward dynamics failed ({rc})")
    return qdd, a0


def crba_rnea(model, params, root13, dof):
    suf, ct = _suffix_types(root13.dtype)
    n = model.n_dof + (0 if model.fixed_base else 6)
    H = np.zeros((n, n), dtype=root13.dtype)
    Cb = np.zeros(n, dtype=root13.dtype)
    fn = getattr(lib(), "orc_crba_rnea" + suf)
    fn.restype = C.c_int
    rc = fn(C.byref(model), C.byref(params), _ptr(root13, ct), _ptr(dof, ct), _ptr(H, ct), _ptr(Cb, ct))
    if rc != 0:
        raise RuntimeError("oracle crba failed")
    return H, Cb


def energy_momentum(model, params, root13, dof):
    suf, ct = _suffix_types(root13.dtype)
    ke, pe = ct(0), ct(0)
    mom = np.zeros(6, dtype=root13.dtype)
    fn = getattr(lib(), "orc_energy_momentum" + suf)
    fn.restype = C.c_int
    rc = fn(C.byref(model), C.byref(params), _ptr(root13, ct), _ptr(dof, ct), C.byref(ke), C.byref(pe), _ptr(mom, ct))
    if rc != 0:
        raise RuntimeError("oracle energy failed")
    return ke.value, pe.value, mom
