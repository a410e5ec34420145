"""Configuration tree for the engine.

Mirrors the semantics of the reference's yacs-based config
(`plb/config/default_config.py:12-78`, `plb/config/utils.py:4-40`) without
depending on yacs (not installed in this image): an attribute-access dict
(`CfgNode`), the same defaults, the same "strings that parse as Python
literals become values, others stay strings" rule yacs applies to YAML
leaves, and the same list/dict variant merge (`plb/envs/utils.py:3-31`).
"""
from .cfg import CfgNode, get_cfg_defaults, load, load_dict, merge_lists, merge_dict, make_cls_config

__all__ = ["CfgNode", "get_cfg_defaults", "load", "load_dict", "merge_lists", "merge_dict",
           "make_cls_config"]
