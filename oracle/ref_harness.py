"""Import the UNMODIFIED reference (/root/reference) in the build container.

TEST INFRASTRUCTURE ONLY — used by tests/golden/make_golden.py (to generate the committed
fixtures) and by tests that are skipped when /root/reference is absent (it does not exist on
the GPU box).  Nothing here is copied from the reference; it only makes its modules importable
on a CPU-only box:

  * ``sys.modules`` stubs for packages the reference imports but this image lacks
    (turtle, chamfer_distance, h5py, matplotlib, easydict, tensorboardX, MinkowskiEngine,
    pytorch3d, open3d).  ``chamfer_distance.ChamferDistance`` is a brute-force module with the
    third-party package's call signature (dist1, dist2, idx1, idx2).
  * a device shim: the reference hard-codes ``device='cuda:0'`` / ``.to(device='cuda')``
    (model_utils.py:86,195; mmd.py:61-62,295).  With no GPU those are redirected to the CPU.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("SUG_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


class _BruteChamfer(torch.nn.Module):
    def forward(self, p1, p2):
        d = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
        d1, i1 = d.min(2)
        d2, i2 = d.min(1)
        return d1, d2, i1, i2


def _install_stubs():
    names = ["turtle", "chamfer_distance", "h5py", "matplotlib", "matplotlib.pyplot", "easydict",
             "tensorboardX", "MinkowskiEngine", "pytorch3d", "pytorch3d.ops", "open3d"]
    for n in names:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = types.ModuleType(n)
    if not hasattr(sys.modules["turtle"], "distance"):
        sys.modules["turtle"].distance = None
    if not hasattr(sys.modules["chamfer_distance"], "ChamferDistance"):
        sys.modules["chamfer_distance"].ChamferDistance = _BruteChamfer


def _is_cuda(dev) -> bool:
    return dev is not None and str(dev).startswith("cuda")


@contextlib.contextmanager
def cpu_device_shim():
    """Redirect hard-coded CUDA devices to the CPU when no GPU is present."""
    if torch.cuda.is_available():
        yield
        return
    o_arange, o_to = torch.arange, torch.Tensor.to

    def arange(*a, **kw):
        if _is_cuda(kw.get("device")):
            kw.pop("device")
        return o_arange(*a, **kw)

    def to(self, *a, **kw):
        if _is_cuda(kw.get("device")):
            kw = dict(kw)
            kw["device"] = "cpu"
        a = tuple("cpu" if (isinstance(x, (str, torch.device)) and _is_cuda(x)) else x for x in a)
        return o_to(self, *a, **kw)

    torch.arange, torch.Tensor.to = arange, to
    try:
        yield
    finally:
        torch.arange, torch.Tensor.to = o_arange, o_to


def load():
    """Returns a namespace with the reference modules: Model, model_utils, model_pointnet,
    point_utils, mmd, common_utils."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.dont_write_bytecode = True
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import model.Model as Model
        import model.model_utils as model_utils
        import model.model_pointnet as model_pointnet
        import model.point_utils as point_utils
        import model.mmd as mmd
        import utils.common_utils as common_utils
    return types.SimpleNamespace(Model=Model, model_utils=model_utils, model_pointnet=model_pointnet,
                                 point_utils=point_utils, mmd=mmd, common_utils=common_utils)
