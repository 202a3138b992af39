"""Compiled-articulation files (JSON) -- the in-tree form of the hot-path robots.

The URDFs themselves belong to the reference's asset tree (not copied here).  ``tools/compile_assets.py``
runs the model compiler over them with the asset options each task uses and stores the *result*
(masses, inertias, joint frames, contact spheres) under ``isaacgymenv_b200/assets/compiled``; tasks load
these when the URDF is not on disk (e.g. on the GPU box).
"""
from __future__ import annotations

import dataclasses
import json
import os

import numpy as np

from .urdf import Articulation, AssetOptions

COMPILED_DIR = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets", "compiled"))


def options_key(asset_file: str, opts: AssetOptions) -> str:
    stem = asset_file.replace("\\", "/").replace("/", "__").replace(".urdf", "")
    flags = f"c{int(opts.collapse_fixed_joints)}k{int(opts.replace_cylinder_with_capsule)}f{int(opts.fix_base_link)}"
    return f"{stem}.{flags}"


def save_articulation(art: Articulation, path: str) -> None:
    d = {}
    for f in dataclasses.fields(art):
        v = getattr(art, f.name)
        if isinstance(v, np.ndarray):
            a = np.where(np.isfinite(v), v, np.sign(v) * 3.0e38) if v.dtype.kind == "f" else v
            d[f.name] = {"dtype": str(v.dtype), "shape": list(v.shape), "data": a.reshape(-1).tolist()}
        else:
            d[f.name] = v
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        json.dump(d, fh, indent=None, separators=(",", ":"))
        fh.write("\n")


def load_articulation(path: str) -> Articulation:
    with open(path) as fh:
        d = json.load(fh)
    kw = {}
    for f in dataclasses.fields(Articulation):
        v = d[f.name]
        if isinstance(v, dict) and "dtype" in v:
            a = np.array(v["data"], dtype=v["dtype"]).reshape(v["shape"])
            if a.dtype.kind == "f":
                big = np.abs(a) >= 1.0e38
                a = a.copy()
                a[big & (a > 0)] = np.inf
                a[big & (a < 0)] = -np.inf
            kw[f.name] = a
        else:
            kw[f.name] = v
    return Articulation(**kw)


def find_compiled(asset_file: str, opts: AssetOptions, rootpath: str = ""):
    """Compiled model of ``asset_file`` (relative to the asset root, e.g. ``urdf/anymal_c/urdf/anymal.urdf``).  Callers split the
    path between root and file name differently (the reference's Cartpole passes root ``.../assets/urdf`` and file ``cartpole.urdf``,
    tasks/cartpole.py:82-88), so trailing components of ``rootpath`` are tried in front of the file name as well."""
    parts = [c for c in os.path.normpath(rootpath).replace("\\", "/").split("/") if c not in ("", ".", "..")] if rootpath else []
    for k in range(0, min(len(parts), 4) + 1):
        rel = "/".join(parts[len(parts) - k:] + [asset_file]) if k else asset_file
        p = os.path.join(COMPILED_DIR, options_key(rel, opts) + ".json")
        if os.path.isfile(p):
            return p
    return None
