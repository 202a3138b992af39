"""Import the reference's task-math functions in THIS container (fixture generation only).

The reference (``/root/reference``) needs Isaac Gym, gym and hydra at import time.  None is
installed here, and none is needed by the ``@torch.jit.script`` reward/observation functions
themselves, so we register permissive stub modules and bare package objects (so the heavy
``__init__`` files are skipped) and then import the task modules straight from the read-only
reference tree.  Recipe: SURVEY.md section 8(c).

This file is test infrastructure.  It is used ONLY by ``tests/golden/gen_golden.py`` to
produce the committed ``*.npz`` fixtures; nothing that runs on the GPU box may import it
(``/root/reference`` does not exist there).
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("B2G_REFERENCE_ROOT", "/root/reference")


class _Permissive:
    """Object that swallows any attribute access / call (stands in for gymapi symbols)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Permissive()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Permissive()


def _stub_module(name):
    mod = types.ModuleType(name)

    def _getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Permissive()

    mod.__getattr__ = _getattr
    sys.modules[name] = mod
    return mod


def _bare_package(name, path):
    pkg = types.ModuleType(name)
    pkg.__path__ = [path]
    sys.modules[name] = pkg
    return pkg


def install_stubs():
    if "isaacgymenvs" in sys.modules and getattr(sys.modules["isaacgymenvs"], "_b2g_stub", False):
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if not hasattr(np, "Inf"):
        np.Inf = np.inf  # removed in numpy 2; the reference uses it (vec_task.py:107)
    ig = _stub_module("isaacgym")
    for sub in ("gymtorch", "gymapi", "gymutil", "terrain_utils", "torch_utils"):
        setattr(ig, sub, _stub_module("isaacgym." + sub))
    sys.modules["isaacgym.gymapi"].SimParams = object
    gym = _stub_module("gym")
    spaces = _stub_module("gym.spaces")
    gym.spaces = spaces

    class Box:
        def __init__(self, *a, **k):
            pass

    class Space:
        pass

    spaces.Box = Box
    spaces.Space = Space
    base = os.path.join(REFERENCE_ROOT, "isaacgymenvs")
    pkg = _bare_package("isaacgymenvs", base)
    pkg._b2g_stub = True
    _bare_package("isaacgymenvs.tasks", os.path.join(base, "tasks"))
    _bare_package("isaacgymenvs.tasks.base", os.path.join(base, "tasks", "base"))
    _bare_package("isaacgymenvs.utils", os.path.join(base, "utils"))


def load(modname):
    """Return reference module ``isaacgymenvs.<modname>`` (e.g. ``tasks.anymal``)."""
    install_stubs()
    import importlib

    return importlib.import_module("isaacgymenvs." + modname)
