"""Scene definitions: YAML/variant loading and target-density assets (host side).

Restates `PlasticineEnv.load_varaints` (`plb/envs/env.py:62-86`): load the scene
tree on top of the defaults, overlay `VARIANTS[version-1]` (lists merged element
by element with `merge_lists`), and point `ENV.loss.target_path` at the asset of
that version (character [-5] of the path is replaced by the version digit).

Scene trees come either from a YAML file (a path ending in .yml that exists on
disk, e.g. inside a PlasticineLab checkout) or from the bundled `scenes.json`
(written by `tools/import_reference_scenes.py`, values verbatim).
"""
from __future__ import annotations

import json
import os

import numpy as np

from ..config import CfgNode, load_dict, merge_lists
from ..config.cfg import _decode_tree

_HERE = os.path.dirname(os.path.abspath(__file__))
_SCENES = None
_TARGETS = None


def bundled_scenes() -> dict:
    global _SCENES
    if _SCENES is None:
        with open(os.path.join(_HERE, "scenes.json")) as f:
            _SCENES = json.load(f)
    return _SCENES


def scene_tree(cfg_path: str) -> dict:
    """`cfg_path` is 'move.yml' / 'move' (bundled) or a path to a YAML file."""
    if os.path.isfile(cfg_path):
        import yaml
        with open(cfg_path) as f:
            return yaml.safe_load(f)
    key = os.path.basename(cfg_path)
    if key.endswith(".yml"):
        key = key[:-4]
    scenes = bundled_scenes()
    if key not in scenes:
        raise FileNotFoundError(f"no scene '{cfg_path}' (bundled: {sorted(scenes)})")
    return scenes[key]


def load_variants(cfg_path: str, version: int) -> CfgNode:
    assert version >= 1
    tree = scene_tree(cfg_path)
    cfg = load_dict(tree)
    variants = _decode_tree(tree["VARIANTS"][version - 1]) if tree.get("VARIANTS") else {}
    new_cfg = CfgNode(variants)
    if "PRIMITIVES" in new_cfg:
        new_cfg.PRIMITIVES = merge_lists([dict(p) for p in cfg.PRIMITIVES],
                                         [None if p is None else dict(p) for p in new_cfg.PRIMITIVES])
    if "SHAPES" in new_cfg:
        new_cfg.SHAPES = merge_lists([dict(p) for p in cfg.SHAPES],
                                     [None if p is None else dict(p) for p in new_cfg.SHAPES])
    cfg.merge_from_other_cfg(new_cfg)
    name = list(cfg.ENV.loss.target_path)
    if len(name) >= 5:
        name[-5] = str(version)
    cfg.ENV.loss.target_path = "".join(name)
    cfg.VARIANTS = None
    return cfg


def load_target(path_or_key: str) -> np.ndarray:
    """Dense float64 64^3 target-density grid.

    Accepts a real .npy path (used as is) or a reference-style relative path
    such as 'envs/assets/Move3D-v1.npy' / the bare key 'Move3D-v1', resolved
    in the bundled `assets/targets.npz` (bit-exact copies of the reference
    grids, `plb/envs/assets/*.npy`).
    """
    global _TARGETS
    if os.path.isfile(path_or_key):
        return np.load(path_or_key)
    key = os.path.basename(path_or_key)
    if key.endswith(".npy"):
        key = key[:-4]
    if _TARGETS is None:
        _TARGETS = np.load(os.path.join(_HERE, "assets", "targets.npz"))
    if key + ".idx" not in _TARGETS:
        raise FileNotFoundError(f"no target grid '{path_or_key}'")
    out = np.zeros(64 * 64 * 64, dtype=np.float64)
    out[_TARGETS[key + ".idx"]] = _TARGETS[key + ".val"]
    return out.reshape(64, 64, 64)


def resample_target(grid: np.ndarray, n_grid: int, total_mass: float | None = None) -> np.ndarray:
    """Nearest-neighbour up-sample of a cubic target grid to n_grid^3, rescaled
    so that its sum is `total_mass` (N * p_mass, the invariant every stock asset
    satisfies) -- the rule SURVEY.md 8(d) fixes for the 128^3+ benchmark configs,
    where the stock 64^3 assets cannot be loaded (`plb/engine/losses/loss.py:29,52`).
    """
    n0 = grid.shape[0]
    if n_grid != n0:
        assert n_grid % n0 == 0
        r = n_grid // n0
        grid = np.repeat(np.repeat(np.repeat(grid, r, 0), r, 1), r, 2)
    grid = np.ascontiguousarray(grid, dtype=np.float64)
    if total_mass is not None and grid.sum() > 0:
        grid = grid * (total_mass / grid.sum())
    return grid
