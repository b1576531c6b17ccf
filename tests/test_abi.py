"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol that include/sug_b200.h declares; the host modules mirror the reference's state_dict; and
the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes

import pytest
import torch

from oracle import sug_oracle as O


def test_library_exports_every_header_symbol():
    from sug_b200 import _lib
    lib = _lib.load()
    names = _lib.header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sug_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.sug_version() >= 100
    assert isinstance(lib.sug_last_error(), bytes)


def test_workspace_queries_are_host_only():
    from sug_b200 import _lib
    lib = _lib.load()
    assert lib.sug_edgeconv_ws_bytes(64, 1024, 128, 256, 20) > 64 * 1024 * 256 * 4
    assert lib.sug_mlp_pool_ws_bytes(64, 1024, 512, 512) > 0
    assert lib.sug_mmd_ws_bytes(64, 4106) >= 4 * 64 * 64 * 4
    assert lib.sug_knn_ws_bytes(64, 64, 1024, 20) > 0


def test_bad_arguments_return_status_not_crash():
    from sug_b200 import _lib
    lib = _lib.load()
    # argument validation happens before any CUDA call
    st = lib.sug_knn_f32(None, 1, 3, 16, 4, 48, 1, 16, None, None, 0, None)
    assert st == -1 and b"null" in lib.sug_last_error()
    buf = ctypes.c_void_p(16)
    st = lib.sug_knn_f32(buf, 1, 3, 16, 40, 48, 1, 16, buf, None, 0, None)
    assert st == -1 and b"k=40" in lib.sug_last_error()
    st = lib.sug_edgeconv_fwd(buf, 3, buf, buf, buf, buf, buf, buf, 1, 16, 3, 6, 4, 1e-5, 0.1, 0.01, 1, buf, 8, buf,
                              buf, buf, buf, buf, None, 0, None)
    assert st == -1 and b"Cout" in lib.sug_last_error()


@pytest.mark.parametrize("name,spec", [("DGCNN", "Net_MDA:DGCNN"), ("Pointnet", "Net_MDA:Pointnet")])
def test_state_dict_matches_reference_layout(name, spec):
    from sug_b200 import Model
    net = Model.Net_MDA(name)
    sd = net.state_dict()
    want = O.state_spec(spec)  # checked against the reference modules by tests/golden/make_golden.py
    assert set(sd) == set(want)
    for k, shp in want.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    net.load_state_dict(O.synth_state(spec), strict=True)
    if name == "DGCNN":
        assert sum(p.numel() for p in net.parameters()) == 10908893


def test_dgcnn_cls_state_dict():
    from sug_b200 import model_pointnet
    net = model_pointnet.DGCNN()
    assert set(net.state_dict()) == set(O.state_spec("DGCNN_cls"))


def test_optimizer_groups_follow_the_trainer():
    from sug_b200 import Model, step
    net = Model.Net_MDA("DGCNN")
    od, og, oc = step.make_optimizers(net)
    n_g = len([1 for k, _ in net.g.named_parameters() if "pred_offset" not in k])
    assert len(og.param_groups) == n_g and len(oc.param_groups) == 2 and len(od.param_groups) == 3


def test_no_cpu_fallback():
    from sug_b200 import Model, model_utils, mmd
    x, _ = O.synth_clouds(2, 64, 0)
    with pytest.raises(RuntimeError, match="CUDA"):
        model_utils.knn(x.squeeze(-1), 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        mmd.mix_rbf_mmd2(torch.randn(4, 8), torch.randn(4, 8), mmd.sigma_list)
    net = Model.Net_MDA("DGCNN")
    with pytest.raises(RuntimeError, match="CUDA"):
        net(O.synth_clouds(2, 1024, 0)[0])


def test_unsupported_backbones_are_explicit():
    from sug_b200 import Model
    for n in ("Pointnet2", "PTran", "KPConv"):
        with pytest.raises(NotImplementedError):
            Model.Net_MDA(n)


def test_fused_adam_has_no_cpu_path_and_mirrors_adam_groups():
    """optim.FusedAdam keeps torch.optim.Adam's param_groups / hyper-parameters, refuses CPU parameters
    (no CPU fallback) and exports its chunk size through the C ABI."""
    from sug_b200 import _lib
    from sug_b200.optim import FusedAdam
    assert _lib.load().sug_adam_chunk() >= 1024
    p = torch.nn.Parameter(torch.zeros(8))
    opt = FusedAdam([{"params": [p], "lr": 3e-3}], lr=1e-3, weight_decay=5e-4)
    g = opt.param_groups[0]
    assert g["lr"] == 3e-3 and g["betas"] == (0.9, 0.999) and g["eps"] == 1e-8 and g["weight_decay"] == 5e-4
    opt.step()  # no gradient yet: a no-op like torch.optim.Adam
    p.grad = torch.ones(8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        opt.step()
    with pytest.raises(ValueError):
        FusedAdam([p], lr=-1.0)


def test_shared_trunk_is_opt_in():
    """The common-subexpression switch of DGCNN must be off unless asked for."""
    import os
    from sug_b200 import Model
    assert os.environ.get("SUG_B200_SHARE_TRUNK", "0") != "1"
    net = Model.Net_MDA("DGCNN")
    assert net.g.share_trunk is False and net.g._trunk == {}


def test_join_slices_autograd_contract():
    """ops.join_slices: producers that wrote into consecutive channel slices of one buffer get the matching
    slice of the gradient; anything else is refused.  (Pure autograd plumbing: runs on CPU tensors.)"""
    from sug_b200 import ops
    buf = torch.zeros(2, 5, 12)
    a = torch.randn(2, 5, 4, requires_grad=True)
    b = torch.randn(2, 5, 8, requires_grad=True)

    class _Write(torch.autograd.Function):  # stands in for an op that writes its output into a slot
        @staticmethod
        def forward(ctx, x, slot):
            slot[0].data.copy_(2 * x)  # like the CUDA kernels: a raw write, no autograd version bump
            return slot[0]

        @staticmethod
        def backward(ctx, g):
            return 2 * g, None

    pa, pb = _Write.apply(a, [buf[:, :, 0:4]]), _Write.apply(b, [buf[:, :, 4:12]])
    cat = ops.join_slices(buf, pa, pb)
    assert cat.shape == buf.shape and cat.data_ptr() == buf.data_ptr()
    w = torch.randn(2, 5, 12)
    (cat * w).sum().backward()
    assert torch.allclose(a.grad, 2 * w[:, :, 0:4]) and torch.allclose(b.grad, 2 * w[:, :, 4:12])
    with pytest.raises(RuntimeError, match="consecutive"):
        ops.join_slices(buf, pb, pa)
    with pytest.raises(RuntimeError, match="cover"):
        ops.join_slices(buf, pa)
