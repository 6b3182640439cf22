"""Flow-relevant configuration: the attribute bag ``get_flow`` reads (flow/flow.py:24-48, flow/mobiusflow.py:7-14,
flow/affineflow.py:5-73).  Defaults follow config.py:119-166 overlaid by settings/base.yml:5-21; the yml files under
``rotationnormflow_b200/settings`` carry the flow keys of the reference's settings/*.yml."""
from __future__ import annotations

import os
import types

import yaml

SETTINGS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "settings")

DEFAULTS = dict(
    dist="mobiusflow", condition=0, layers=24, segments=64, rot="16Trans", lu=0, feature_dim=512, embedding=0,
    embedding_dim=512, last_affine=0, first_affine=1, frequent_permute=0, pretrain_fisher="", number_queries=5000,
)


def _str2type(v):
    """config.py:224-233 maps the strings None/true/false."""
    if isinstance(v, str):
        low = v.lower()
        if low == "none":
            return None
        if low == "true":
            return True
        if low == "false":
            return False
    return v


def load_config(name_or_path: str | None = None, **overrides) -> types.SimpleNamespace:
    """``load_config('symsol', feature_dim=2048)`` -> namespace usable by ``get_flow``."""
    cfg = dict(DEFAULTS)
    base = os.path.join(SETTINGS_DIR, "base.yml")          # always loaded first (config.py:95-96)
    if os.path.exists(base):
        with open(base) as f:
            cfg.update({k: _str2type(v) for k, v in (yaml.safe_load(f) or {}).items()})
    if name_or_path is not None:
        path = name_or_path if os.path.exists(name_or_path) else os.path.join(SETTINGS_DIR, name_or_path + ".yml")
        with open(path) as f:
            y = yaml.safe_load(f) or {}
        cfg.update({k: _str2type(v) for k, v in y.items()})
    cfg.update(overrides)
    return types.SimpleNamespace(**cfg)
