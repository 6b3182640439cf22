"""TEST INFRASTRUCTURE ONLY.  Imports the UNMODIFIED reference from /root/reference.

Usable only in the build container (the GPU box has no /root/reference); it is
what ``oracle/make_golden.py`` uses to mint ``tests/golden/*.npz`` and what
``tests/test_oracle.py`` uses (when the directory exists) to pin the
restatement in ``oracle/rnf_oracle.py`` against the real code.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np
import torch
import yaml

from . import stubs

REF_ROOT = os.environ.get("RNF_REFERENCE_ROOT", "/root/reference")

# flow-relevant defaults: config.py:119-166 overlaid by settings/base.yml:5-21
_DEFAULTS = dict(
    dist="mobiusflow", condition=0, layers=24, segments=64, rot="16Trans", lu=0,
    feature_dim=512, embedding=0, embedding_dim=512, last_affine=0, first_affine=1,
    frequent_permute=0, pretrain_fisher="", number_queries=5000,
)
_FLOW_KEYS = tuple(_DEFAULTS.keys())


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "flow"))


def ref_config(name: str | None = None, **overrides) -> types.SimpleNamespace:
    """Attribute bag equivalent to ``get_config`` for the flow-relevant keys
    (settings/base.yml, then settings/<name>.yml, then overrides)."""
    cfg = dict(_DEFAULTS)
    if name is not None:
        with open(os.path.join(REF_ROOT, "settings", name + ".yml")) as f:
            y = yaml.safe_load(f) or {}
        for k in _FLOW_KEYS:
            if k in y:
                cfg[k] = y[k]
    cfg.update(overrides)
    return types.SimpleNamespace(**cfg)


def _import_reference():
    stubs.install()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import flow.flow as ref_flow  # noqa: E402  (reference module)
    return ref_flow


def build_reference_flow(cfg, seed: int = 0, dtype=torch.float32):
    """``torch.manual_seed(seed); np.random.seed(seed); get_flow(cfg)`` on the real reference."""
    ref_flow = _import_reference()
    torch.manual_seed(seed)
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref_flow.get_flow(cfg)
    m = m.to(dtype)
    m.eval()
    return m


@contextlib.contextmanager
def default_dtype(dtype):
    """flow/mobiusflow.py:80,178 allocate default-dtype outputs: fp64 runs need this."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        yield
    finally:
        torch.set_default_dtype(old)


def run_reference(m, R, feature=None, inverse=False):
    dt = next(m.parameters()).dtype
    with torch.no_grad(), default_dtype(dt):
        R = R.to(dt)
        f = None if feature is None else feature.to(dt)
        out, ldj = m.inverse(R, f) if inverse else m(R, f)
    return out, ldj


def reference_sd():
    stubs.install()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import utils.sd as sd  # noqa: E402
    return sd


def reference_fisher():
    stubs.install()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import utils.fisher as fisher  # noqa: E402
    return fisher
