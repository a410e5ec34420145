"""Minimal config node + loaders (host side; no device code).

Reference behaviour restated here:
  * defaults: `plb/config/default_config.py:12-78`
  * `load(path, opts)`: `plb/config/utils.py:33-40`
  * yacs leaf decoding: a YAML string leaf is passed through
    `ast.literal_eval`; on failure it stays a string (so "(0.5, 0.2)" becomes
    a tuple while "0.2049/2" and "(127<<16)" stay strings that
    `Shapes` later `eval`s, `plb/engine/shapes/shape_maker.py:23`).
  * variant merging: `plb/envs/utils.py:3-31`.
"""
from __future__ import annotations

import ast
import copy
from typing import Any


class CfgNode(dict):
    """dict with attribute access (the slice of yacs.CfgNode the callers use)."""

    def __init__(self, init=None, new_allowed=True):
        super().__init__()
        if init:
            for k, v in init.items():
                self[k] = _wrap(v)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:  # pragma: no cover - mirrors AttributeError of yacs
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value

    # yacs API used by reference callers; all no-ops or trivial here
    def defrost(self):
        return self

    def freeze(self):
        return self

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_other_cfg(self, other):
        _merge_into(self, other, strict=True)

    def merge_from_dict(self, other):
        _merge_into(self, _wrap(_decode_tree(other)), strict=False)

    def merge_from_list(self, opts):
        assert len(opts) % 2 == 0
        for k, v in zip(opts[0::2], opts[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = _wrap(_decode(v))

    def to_dict(self):
        return _unwrap(self)


def _wrap(v):
    if isinstance(v, CfgNode):
        return v
    if isinstance(v, dict):
        return CfgNode(v)
    if isinstance(v, list):
        return [_wrap(i) for i in v]
    return v


def _unwrap(v):
    if isinstance(v, dict):
        return {k: _unwrap(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)) and any(isinstance(i, dict) for i in v):
        return [_unwrap(i) for i in v]
    return v


def _decode(v: Any) -> Any:
    """yacs `_decode_cfg_value`: literal_eval strings, keep the rest."""
    if not isinstance(v, str):
        return v
    try:
        return ast.literal_eval(v)
    except (ValueError, SyntaxError, TypeError, MemoryError, RecursionError):
        return v


def _decode_tree(v):
    if isinstance(v, dict):
        return {k: _decode_tree(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_decode_tree(x) for x in v]
    return _decode(v)


def _merge_into(a: CfgNode, b: dict, strict: bool):
    for k, v in b.items():
        if isinstance(v, dict) and isinstance(a.get(k), dict):
            _merge_into(a[k], v, strict)
        else:
            a[k] = _wrap(copy.deepcopy(v))


def merge_dict(a, b):
    """`plb/envs/utils.py:3-19`: overlay b's leaves on a; unknown keys are an error."""
    if b is None:
        return a
    a = copy.deepcopy(a)
    for key in a:
        if key in b:
            if not isinstance(b[key], dict):
                a[key] = b[key]
            else:
                assert not isinstance(a[key], list)
                a[key] = merge_dict(a[key], b[key])
    for key in b:
        if key not in a:
            raise ValueError("Key is not in dict A!")
    return a


def merge_lists(a, b):
    """`plb/envs/utils.py:22-31`: element-wise merge_dict over the shorter list."""
    assert isinstance(a, list) and isinstance(b, list)
    outs = []
    for i in range(len(a)):
        x = a[i]
        if i < len(b):
            x = merge_dict(a[i], b[i])
        outs.append(x)
    return outs


def get_cfg_defaults() -> CfgNode:
    """Same tree and values as `plb/config/default_config.py:12-78`."""
    c = CfgNode()
    c.SIMULATOR = CfgNode(dict(
        dim=3, quality=1, yield_stress=50.0, dtype="float64", max_steps=1024,
        n_particles=9000, E=5e3, nu=0.2, ground_friction=1.5, gravity=(0, -1, 0)))
    c.PRIMITIVES = []
    c.SHAPES = []
    c.RENDERER = CfgNode(dict(
        spp=50, max_ray_depth=2, image_res=(512, 512), voxel_res=(168, 168, 168),
        target_res=(64, 64, 64), dx=1.0 / 150, sdf_threshold=0.37 * 0.56, bake_size=6,
        use_roulette=False, light_direction=(2.0, 1.0, 0.7), camera_pos=(0.5, 1.2, 4.0),
        camera_rot=(0.2, 0), use_directional_light=False, max_num_particles=1000000))
    c.ENV = CfgNode(dict(
        loss=dict(soft_contact=False, weight=dict(sdf=10, density=10, contact=1), target_path=""),
        n_observed_particles=200))
    c.VARIANTS = []
    return c


def load_dict(tree: dict | None = None, opts=None) -> CfgNode:
    cfg = get_cfg_defaults()
    if tree is not None:
        cfg.merge_from_dict(tree)
    if opts is not None:
        cfg.merge_from_list(list(opts))
    return cfg


def load(path=None, opts=None) -> CfgNode:
    """`plb/config/utils.py:33-40` (yaml file + option list on top of defaults)."""
    tree = None
    if path is not None:
        import yaml
        with open(path, "r") as f:
            tree = yaml.safe_load(f)
    return load_dict(tree, opts)


def make_cls_config(obj, cfg=None, **kwargs) -> CfgNode:
    """`plb/config/utils.py:4-13`."""
    _cfg = obj.default_config()
    if cfg is not None:
        if isinstance(cfg, str):
            import yaml
            with open(cfg, "r") as f:
                _cfg.merge_from_dict(yaml.safe_load(f))
        else:
            _cfg.merge_from_dict(_unwrap(cfg) if isinstance(cfg, dict) else cfg)
    if kwargs:
        _cfg.merge_from_list(sum(list(kwargs.items()), ()))
    return _cfg
