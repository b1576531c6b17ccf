"""Run the reference's own scripts on top of sug_b200 without touching them.

    import sug_b200.compat; sug_b200.compat.install()      # before `import model...`
    # or:  python -m sug_b200.compat train_dg_single_gpu.py --cfg ...   (from the reference checkout)

``install()`` registers this package's drop-in modules under the reference's import names
(``model.Model``, ``model.model_utils``, ``model.point_utils``, ``model.model_pointnet``,
``model.mmd``) so that ``train_dg_single_gpu.py:23-27`` picks them up, and provides the third-party
names the trainer imports unconditionally but that are irrelevant to the DGCNN / PointNet path
(``chamfer_distance``, ``tensorboardX``, ``MinkowskiEngine``, ``pytorch3d``, ``easydict``,
``turtle``) when they are not installed.  Nothing is patched inside the reference files."""
from __future__ import annotations

import importlib
import runpy
import os
import sys
import types


class _EasyDict(dict):
    """Minimal EasyDict (attribute access, recursive) for utils/config.py."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        elif isinstance(v, list):
            v = [_EasyDict(x) if isinstance(x, dict) and not isinstance(x, _EasyDict) else x for x in v]
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _ensure(name: str, **attrs):
    try:
        return importlib.import_module(name)
    except Exception:
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m


def install(fused_adam=None):
    """Register the drop-in modules under the reference's import names.  ``fused_adam`` (default: env
    ``SUG_B200_FUSED_ADAM=1``): additionally let the trainer's ``torch.optim.Adam(...)`` calls
    (train_dg_single_gpu.py:191-203) build ``sug_b200.optim.FusedAdam`` for CUDA parameters -- same arithmetic,
    one launch per param group; keyword arguments FusedAdam does not know fall back to torch's class."""
    from . import Model, mmd, model_pointnet, model_utils, ops, point_utils
    import torch

    if fused_adam is None:
        fused_adam = os.environ.get("SUG_B200_FUSED_ADAM", "0") == "1"
    if fused_adam and not getattr(torch.optim.Adam, "_sug_wrapped", False):
        from .optim import FusedAdam
        _TorchAdam = torch.optim.Adam

        def _adam(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, **kw):
            params = list(params)
            flat = [p for g in params for p in (g["params"] if isinstance(g, dict) else [g])]
            if kw or not flat or not all(isinstance(p, torch.Tensor) and p.is_cuda for p in flat):
                return _TorchAdam(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, **kw)
            return FusedAdam(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        _adam._sug_wrapped = True
        torch.optim.Adam = _adam

    class ChamferDistance(torch.nn.Module):
        """Call signature of the third-party package: (dist1, dist2, idx1, idx2); indices unused
        by the reference (mmd.py:170)."""

        def forward(self, p1, p2):
            d1, d2 = ops.chamfer(p1, p2)
            return d1, d2, None, None

    class _Writer:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, _):
            return lambda *a, **k: None

    _ensure("easydict", EasyDict=_EasyDict)
    _ensure("chamfer_distance", ChamferDistance=ChamferDistance)
    _ensure("tensorboardX", SummaryWriter=_Writer)
    _ensure("turtle", distance=None)
    for n in ("MinkowskiEngine", "pytorch3d", "pytorch3d.ops", "open3d", "h5py", "matplotlib", "matplotlib.pyplot"):
        _ensure(n)
    pkg = sys.modules.get("model")
    if pkg is None:
        pkg = types.ModuleType("model")
        pkg.__path__ = []  # a namespace: other `model.*` modules of the reference keep importing from disk
        sys.modules["model"] = pkg
    for name, mod in (("Model", Model), ("model_utils", model_utils), ("point_utils", point_utils),
                      ("model_pointnet", model_pointnet), ("mmd", mmd)):
        sys.modules["model." + name] = mod
        setattr(pkg, name, mod)
    return pkg


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m sug_b200.compat <reference script> [args...]")
    install()
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name="__main__")
