"""Run the reference's own scripts on top of sug_b200 without touching them.

    import sug_b200.compat; sug_b200.compat.install()      # before `import model...`
    # or:  python -m sug_b200.compat train_dg_single_gpu.py --cfg ...   (from the reference checkout)

``install()`` registers this package's drop-in modules under the reference's import names
(``model.Model``, ``model.model_utils``, ``model.point_utils``, ``model.model_pointnet``,
``model.mmd``) so that ``train_dg_single_gpu.py:8-27`` picks them up.  Every OTHER ``model.*`` module
(``model.KPConv_model`` imported unconditionally at train_dg_single_gpu.py:27, ``model.pointnet2_utils``,
...) still loads from the reference checkout: the ``model`` package object keeps the checkout's
``model/`` directory on its ``__path__``.  Third-party names the trainer imports unconditionally but
that are irrelevant to the DGCNN / PointNet path (``chamfer_distance``, ``tensorboardX``,
``MinkowskiEngine``, ``pytorch3d``, ``easydict``, ``turtle``, ``h5py``, ``matplotlib``, ``open3d``) are
provided when they are not installed: real re-implementations for ``EasyDict`` and ``ChamferDistance``,
inert stand-ins for the rest.  Nothing is patched inside the reference files.

What this has been exercised on is stated in INTEGRATION.md §1 (imports of the trainer, model / optimizer
construction, ``copy.deepcopy(model).eval()``, the checkpoint round trip, ``--help``); the dataset files the
trainer reads from its hard-coded root are not synthesised here."""
from __future__ import annotations

import importlib
import importlib.machinery
import importlib.util
import os
import runpy
import sys
import types

DROPIN_MODULES = ("Model", "model_utils", "point_utils", "model_pointnet", "mmd")


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


class _Inert:
    """Stand-in object for names of absent third-party packages: importable, subclassable, callable at
    import time -- and loud as soon as the code path that really needs the package runs."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("this third-party package is not installed; sug_b200.compat only provides an import-time "
                           "stand-in for it (it is outside the DGCNN / PointNet hot path)")

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert()


class _StubModule(types.ModuleType):
    """Module whose unknown attributes resolve to inert classes (``from pkg import anything`` works)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (_Inert,), {"__module__": self.__name__})
        setattr(self, name, obj)
        return obj


def _ensure(name: str, **attrs):
    """Import ``name`` if it is installed, else register a stub module under that name."""
    try:
        return importlib.import_module(name)
    except Exception:
        m = _StubModule(name)
        m.__path__ = []  # so that `import name.sub` is attempted (and answered by a stub registered below)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        if "." in name:
            parent, _, leaf = name.rpartition(".")
            if parent in sys.modules:
                setattr(sys.modules[parent], leaf, m)
        return m


def find_reference_root(start=None):
    """Directory of the reference checkout (the one holding ``model/Model.py`` and ``model/mmd.py``):
    the current directory, the script's directory or an entry of ``sys.path``.  None if there is none."""
    cands = [start, os.getcwd(), *(sys.path or [])]
    for c in cands:
        if not c:
            continue
        d = os.path.abspath(c)
        if os.path.isfile(os.path.join(d, "model", "Model.py")) and os.path.isfile(os.path.join(d, "model", "mmd.py")):
            return d
    return None


def install(fused_adam=None, reference_root=None):
    """Register the drop-in modules under the reference's import names.  ``fused_adam`` (default: env
    ``SUG_B200_FUSED_ADAM=1``): additionally let the trainer's ``torch.optim.Adam(...)`` calls
    (train_dg_single_gpu.py:191-203) build ``sug_b200.optim.FusedAdam`` for CUDA parameters -- same arithmetic,
    one launch per param group; keyword arguments FusedAdam does not know fall back to torch's class.
    ``reference_root``: the reference checkout (default: found from the cwd / ``sys.path``)."""
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
    for n in ("MinkowskiEngine", "pytorch3d", "pytorch3d.ops", "open3d", "h5py", "matplotlib", "matplotlib.pyplot",
              "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        _ensure(n)

    root = reference_root or find_reference_root()
    pkg = sys.modules.get("model")
    if pkg is None:
        pkg = types.ModuleType("model")
        # the checkout's model/ directory stays on the package path: every module this package does NOT
        # replace (model.KPConv_model, model.pointnet2_utils, ...) is found on disk as before
        pkg.__path__ = [os.path.join(root, "model")] if root else []
        pkg.__spec__ = importlib.machinery.ModuleSpec("model", None, is_package=True)
        pkg.__spec__.submodule_search_locations = pkg.__path__
        sys.modules["model"] = pkg
    for name, mod in (("Model", Model), ("model_utils", model_utils), ("point_utils", point_utils),
                      ("model_pointnet", model_pointnet), ("mmd", mmd)):
        sys.modules["model." + name] = mod
        setattr(pkg, name, mod)
    return pkg


def main(argv):
    if len(argv) < 2:
        raise SystemExit("usage: python -m sug_b200.compat <reference script> [args...]")
    script = os.path.abspath(argv[1])
    sdir = os.path.dirname(script)
    if sdir not in sys.path:  # what `python script.py` does: the script's directory leads sys.path
        sys.path.insert(0, sdir)
    install(reference_root=find_reference_root(sdir))
    sys.argv = argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main(sys.argv)
