"""Procedural terrain for the rough-terrain tasks: a stand-in for ``isaacgym.terrain_utils`` (absent from the
reference tree -- un-vendored Isaac Gym module, imported at ``tasks/anymal_terrain.py:542``) plus the ``Terrain`` grid
generator of ``tasks/anymal_terrain.py:543-673``.

The sub-terrain builders keep the names and arguments the reference calls (``SubTerrain``, ``pyramid_sloped_terrain``,
``random_uniform_terrain``, ``pyramid_stairs_terrain``, ``discrete_obstacles_terrain``, ``stepping_stones_terrain``,
``convert_heightfield_to_trimesh``) and produce ``int16`` height fields in units of ``vertical_scale``; their exact
random content is this package's own (PARITY UNPINNED: no reference output exists to compare against -- DESIGN.md).
``Terrain`` lays the tiles out exactly as the reference does (levels along x, types along y, border, env origins at
the tile centres at the height of the central 2 m x 2 m patch).
"""
from __future__ import annotations

import numpy as np


class SubTerrain:
    def __init__(self, terrain_name="terrain", width=256, length=256, vertical_scale=1.0, horizontal_scale=1.0):
        self.terrain_name = terrain_name
        self.vertical_scale = vertical_scale
        self.horizontal_scale = horizontal_scale
        self.width = width
        self.length = length
        self.height_field_raw = np.zeros((self.width, self.length), dtype=np.int16)


def _rng(rng):
    return rng if rng is not None else np.random


def random_uniform_terrain(terrain, min_height, max_height, step=1, downsampled_scale=None, rng=None):
    """Coarse grid of heights drawn from {min_height, min_height+step, ..., max_height}, bilinearly upsampled and
    ADDED to the field."""
    rng = _rng(rng)
    if downsampled_scale is None:
        downsampled_scale = terrain.horizontal_scale
    lo, hi, st = int(min_height / terrain.vertical_scale), int(max_height / terrain.vertical_scale), max(int(step / terrain.vertical_scale), 1)
    levels = np.arange(lo, hi + st, st)
    nw = max(int(terrain.width * terrain.horizontal_scale / downsampled_scale), 2)
    nl = max(int(terrain.length * terrain.horizontal_scale / downsampled_scale), 2)
    coarse = rng.choice(levels, (nw, nl)).astype(np.float64)
    xs = np.linspace(0, nw - 1, terrain.width)
    ys = np.linspace(0, nl - 1, terrain.length)
    x0 = np.clip(np.floor(xs).astype(int), 0, nw - 2)
    y0 = np.clip(np.floor(ys).astype(int), 0, nl - 2)
    fx = (xs - x0)[:, None]
    fy = (ys - y0)[None, :]
    c00 = coarse[x0][:, y0]
    c10 = coarse[x0 + 1][:, y0]
    c01 = coarse[x0][:, y0 + 1]
    c11 = coarse[x0 + 1][:, y0 + 1]
    up = c00 * (1 - fx) * (1 - fy) + c10 * fx * (1 - fy) + c01 * (1 - fx) * fy + c11 * fx * fy
    terrain.height_field_raw += np.rint(up).astype(np.int16)
    return terrain


def pyramid_sloped_terrain(terrain, slope=1, platform_size=1.0):
    """Square pyramid rising (slope > 0) or sinking towards the centre, cut by a flat platform on top."""
    x = np.arange(terrain.width)
    y = np.arange(terrain.length)
    cx, cy = terrain.width / 2, terrain.length / 2
    xx = (cx - np.abs(cx - x)) / cx
    yy = (cy - np.abs(cy - y)) / cy
    max_height = int(slope * (terrain.horizontal_scale / terrain.vertical_scale) * (terrain.width / 2))
    field = (max_height * xx[:, None] * yy[None, :]).astype(np.float64)
    terrain.height_field_raw += field.astype(np.int16)
    half = int(platform_size / terrain.horizontal_scale / 2)
    x1, x2, y1, y2 = terrain.width // 2 - half, terrain.width // 2 + half, terrain.length // 2 - half, terrain.length // 2 + half
    edge = terrain.height_field_raw[x1, y1]
    lo, hi = min(edge, 0), max(edge, 0)
    terrain.height_field_raw = np.clip(terrain.height_field_raw, lo, hi).astype(np.int16)
    return terrain


def pyramid_stairs_terrain(terrain, step_width, step_height, platform_size=1.0):
    """Concentric square steps going up (step_height > 0) or down towards a central platform."""
    sw = max(int(step_width / terrain.horizontal_scale), 1)
    sh = int(step_height / terrain.vertical_scale)
    plat = int(platform_size / terrain.horizontal_scale)
    height = 0
    x0, x1, y0, y1 = 0, terrain.width, 0, terrain.length
    while (x1 - x0) > plat and (y1 - y0) > plat:
        x0 += sw
        x1 -= sw
        y0 += sw
        y1 -= sw
        height += sh
        terrain.height_field_raw[x0:x1, y0:y1] = height
    return terrain


def discrete_obstacles_terrain(terrain, max_height, min_size, max_size, num_rects, platform_size=1.0, rng=None):
    """Random axis-aligned boxes / pits of height in {-h, -h/2, h/2, h}; flat platform in the centre."""
    rng = _rng(rng)
    mh = int(max_height / terrain.vertical_scale)
    smin, smax = int(min_size / terrain.horizontal_scale), int(max_size / terrain.horizontal_scale)
    plat = int(platform_size / terrain.horizontal_scale)
    (i, j) = terrain.height_field_raw.shape
    heights = [-mh, -mh // 2, mh // 2, mh]
    sizes = np.arange(smin, max(smax, smin + 1), 4)
    for _ in range(num_rects):
        w = int(rng.choice(sizes))
        l = int(rng.choice(sizes))
        sx = int(rng.choice(np.arange(0, max(i - w, 1), 4)))
        sy = int(rng.choice(np.arange(0, max(j - l, 1), 4)))
        terrain.height_field_raw[sx:sx + w, sy:sy + l] = int(rng.choice(heights))
    x1, x2, y1, y2 = (i - plat) // 2, (i + plat) // 2, (j - plat) // 2, (j + plat) // 2
    terrain.height_field_raw[x1:x2, y1:y2] = 0
    return terrain


def stepping_stones_terrain(terrain, stone_size, stone_distance, max_height, platform_size=1.0, depth=-10, rng=None):
    """Square stones of random height separated by gaps that drop to `depth`; flat platform in the centre."""
    rng = _rng(rng)
    ss = max(int(stone_size / terrain.horizontal_scale), 1)
    sd = max(int(stone_distance / terrain.horizontal_scale), 1)
    mh = int(max_height / terrain.vertical_scale)
    plat = int(platform_size / terrain.horizontal_scale)
    heights = np.arange(-mh - 1, mh, 1)
    terrain.height_field_raw[:, :] = int(depth / terrain.vertical_scale)
    sy = 0
    while sy < terrain.length:
        ey = min(terrain.length, sy + ss)
        sx = int(rng.integers(0, ss)) if hasattr(rng, "integers") else int(rng.randint(0, ss))
        ex = max(0, sx - sd)
        terrain.height_field_raw[0:ex, sy:ey] = int(rng.choice(heights))
        while sx < terrain.width:
            ex = min(terrain.width, sx + ss)
            terrain.height_field_raw[sx:ex, sy:ey] = int(rng.choice(heights))
            sx += ss + sd
        sy += ss + sd
    x1, x2 = (terrain.width - plat) // 2, (terrain.width + plat) // 2
    y1, y2 = (terrain.length - plat) // 2, (terrain.length + plat) // 2
    terrain.height_field_raw[x1:x2, y1:y2] = 0
    return terrain


def convert_heightfield_to_trimesh(height_field_raw, horizontal_scale, vertical_scale, slope_threshold=None):
    """Vertices (rows*cols, 3) float32 on the regular grid and triangles (2 per cell) uint32.  The slope-threshold
    correction of Isaac Gym's version (which shifts vertices to build vertical walls) is not applied: this engine
    collides against the height grid itself (DESIGN.md deviations)."""
    hf = np.asarray(height_field_raw)
    rows, cols = hf.shape
    x = np.linspace(0, (rows - 1) * horizontal_scale, rows)
    y = np.linspace(0, (cols - 1) * horizontal_scale, cols)
    xx, yy = np.meshgrid(x, y, indexing="ij")
    vertices = np.zeros((rows * cols, 3), dtype=np.float32)
    vertices[:, 0] = xx.reshape(-1)
    vertices[:, 1] = yy.reshape(-1)
    vertices[:, 2] = hf.reshape(-1) * vertical_scale
    tris = np.zeros((2 * (rows - 1) * (cols - 1), 3), dtype=np.uint32)
    idx = np.arange(rows * cols).reshape(rows, cols)
    a, b, c, d = idx[:-1, :-1].reshape(-1), idx[1:, :-1].reshape(-1), idx[1:, 1:].reshape(-1), idx[:-1, 1:].reshape(-1)
    tris[0::2] = np.stack([a, b, c], axis=1)
    tris[1::2] = np.stack([a, c, d], axis=1)
    return vertices, tris


class Terrain:
    """Grid of sub-terrains: ``numLevels`` rows (difficulty) x ``numTerrains`` columns (type), 20 m flat border,
    0.1 m cells, 5 mm height units (reference ``tasks/anymal_terrain.py:543-673``)."""

    def __init__(self, cfg, num_robots, seed=42):
        self.type = cfg["terrainType"]
        if self.type in ("none", "plane"):
            return
        # seed=None draws from numpy's global generator, exactly where the reference's Terrain draws from (tests compare the two classes)
        self.rng = np.random.default_rng(seed) if seed is not None else np.random
        self.horizontal_scale = 0.1
        self.vertical_scale = 0.005
        self.border_size = 20
        self.env_length = cfg["mapLength"]
        self.env_width = cfg["mapWidth"]
        props = cfg["terrainProportions"]
        self.proportions = [float(np.sum(props[:i + 1])) for i in range(len(props))]
        self.env_rows = cfg["numLevels"]
        self.env_cols = cfg["numTerrains"]
        self.num_maps = self.env_rows * self.env_cols
        self.num_per_env = int(num_robots / self.num_maps)
        self.env_origins = np.zeros((self.env_rows, self.env_cols, 3))
        self.width_per_env_pixels = int(self.env_width / self.horizontal_scale)
        self.length_per_env_pixels = int(self.env_length / self.horizontal_scale)
        self.border = int(self.border_size / self.horizontal_scale)
        self.tot_cols = int(self.env_cols * self.width_per_env_pixels) + 2 * self.border
        self.tot_rows = int(self.env_rows * self.length_per_env_pixels) + 2 * self.border
        self.height_field_raw = np.zeros((self.tot_rows, self.tot_cols), dtype=np.int16)
        if cfg["curriculum"]:
            self.curiculum(num_robots, num_terrains=self.env_cols, num_levels=self.env_rows)
        else:
            self.randomized_terrain()
        self.heightsamples = self.height_field_raw
        self.vertices, self.triangles = convert_heightfield_to_trimesh(self.height_field_raw, self.horizontal_scale, self.vertical_scale,
                                                                       cfg.get("slopeTreshold", 0.5))

    def _new_tile(self):
        return SubTerrain("terrain", width=self.width_per_env_pixels, length=self.width_per_env_pixels, vertical_scale=self.vertical_scale,
                          horizontal_scale=self.horizontal_scale)

    def _place(self, terrain, i, j):
        sx, sy = self.border + i * self.length_per_env_pixels, self.border + j * self.width_per_env_pixels
        self.height_field_raw[sx:sx + self.length_per_env_pixels, sy:sy + self.width_per_env_pixels] = terrain.height_field_raw
        x1 = int((self.env_length / 2.0 - 1) / self.horizontal_scale)
        x2 = int((self.env_length / 2.0 + 1) / self.horizontal_scale)
        y1 = int((self.env_width / 2.0 - 1) / self.horizontal_scale)
        y2 = int((self.env_width / 2.0 + 1) / self.horizontal_scale)
        z = np.max(terrain.height_field_raw[x1:x2, y1:y2]) * self.vertical_scale
        self.env_origins[i, j] = [(i + 0.5) * self.env_length, (j + 0.5) * self.env_width, z]

    def randomized_terrain(self):
        rng = self.rng
        for k in range(self.num_maps):
            (i, j) = np.unravel_index(k, (self.env_rows, self.env_cols))
            terrain = self._new_tile()
            choice = rng.uniform(0, 1)
            if choice < 0.1:
                if rng.choice([0, 1]):
                    pyramid_sloped_terrain(terrain, rng.choice([-0.3, -0.2, 0, 0.2, 0.3]))
                    random_uniform_terrain(terrain, min_height=-0.1, max_height=0.1, step=0.05, downsampled_scale=0.2, rng=rng)
                else:
                    pyramid_sloped_terrain(terrain, rng.choice([-0.3, -0.2, 0, 0.2, 0.3]))
            elif choice < 0.6:
                pyramid_stairs_terrain(terrain, step_width=0.31, step_height=rng.choice([-0.15, 0.15]), platform_size=3.0)
            else:
                discrete_obstacles_terrain(terrain, 0.15, 1.0, 2.0, 40, platform_size=3.0, rng=rng)
            self._place(terrain, i, j)

    def curiculum(self, num_robots, num_terrains, num_levels):
        rng = self.rng
        for j in range(num_terrains):
            for i in range(num_levels):
                terrain = self._new_tile()
                difficulty = i / num_levels
                choice = j / num_terrains
                slope = difficulty * 0.4
                step_height = 0.05 + 0.175 * difficulty
                obstacle_height = 0.025 + difficulty * 0.15
                stone_size = 2 - 1.8 * difficulty
                if choice < self.proportions[0]:
                    if choice < 0.05:
                        slope *= -1
                    pyramid_sloped_terrain(terrain, slope=slope, platform_size=3.0)
                elif choice < self.proportions[1]:
                    if choice < 0.15:
                        slope *= -1
                    pyramid_sloped_terrain(terrain, slope=slope, platform_size=3.0)
                    random_uniform_terrain(terrain, min_height=-0.1, max_height=0.1, step=0.025, downsampled_scale=0.2, rng=rng)
                elif choice < self.proportions[3]:
                    if choice < self.proportions[2]:
                        step_height *= -1
                    pyramid_stairs_terrain(terrain, step_width=0.31, step_height=step_height, platform_size=3.0)
                elif choice < self.proportions[4]:
                    discrete_obstacles_terrain(terrain, obstacle_height, 1.0, 2.0, 40, platform_size=3.0, rng=rng)
                else:
                    stepping_stones_terrain(terrain, stone_size=stone_size, stone_distance=0.1, max_height=0.0, platform_size=3.0, rng=rng)
                self._place(terrain, i, j)
