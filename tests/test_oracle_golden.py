"""Pin oracle/sug_oracle.py against the fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import sug_oracle as O


def feat_input(B, C, N, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.standard_normal((B, C, N)).astype(np.float32))


def close(a, b, rtol=1e-4, atol=1e-5):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("C,N", [(3, 256), (64, 256), (128, 128)])
def test_knn(golden, C, N):
    x = O.synth_clouds(2, N, 10 + C)[0].squeeze(-1) if C == 3 else feat_input(2, C, N, 10 + C)
    idx = O.knn(x, 20)
    assert idx.dtype == torch.int64
    np.testing.assert_array_equal(idx.numpy(), golden(f"knn_c{C}")["idx"].astype(np.int64))


def test_edgeconv_block(golden):
    g = golden("edgeconv_block")
    B, C, N, Co, k = 2, 8, 128, 16, 20
    x = feat_input(B, C, N, 21).requires_grad_(True)
    rng = np.random.Generator(np.random.PCG64(22))
    W = torch.from_numpy(rng.standard_normal((Co, 2 * C, 1, 1)).astype(np.float32) * 0.3).requires_grad_(True)
    ga = rng.uniform(0.5, 1.5, Co).astype(np.float32)
    ga[::3] *= -1
    sd = {"b.conv.0.weight": W, "b.conv.1.weight": torch.from_numpy(ga).requires_grad_(True),
          "b.conv.1.bias": torch.from_numpy(rng.standard_normal(Co).astype(np.float32) * 0.1).requires_grad_(True),
          "b.conv.1.running_mean": torch.zeros(Co), "b.conv.1.running_var": torch.ones(Co),
          "b.conv.1.num_batches_tracked": torch.zeros((), dtype=torch.int64)}
    Rw = torch.from_numpy(rng.standard_normal((B, Co, N)).astype(np.float32))
    out = O.edgeconv(x, sd, "b", True, k=k)
    (out * Rw).sum().backward()
    close(out, g["out"])
    close(x.grad, g["dx"], 1e-3, 1e-5)
    close(W.grad, g["dW"], 1e-3, 1e-4)
    close(sd["b.conv.1.weight"].grad, g["dgamma"], 1e-3, 1e-4)
    close(sd["b.conv.1.bias"].grad, g["dbeta"], 1e-3, 1e-4)
    close(sd["b.conv.1.running_mean"], g["running_mean"])
    close(sd["b.conv.1.running_var"], g["running_var"])
    assert int(sd["b.conv.1.num_batches_tracked"]) == 1


def test_adapt_layer(golden):
    g = golden("adapt_layer")
    sd = O.synth_state("Net_MDA:DGCNN")
    loc = O.synth_clouds(2, 256, 31)[0].squeeze(-1)
    fea = feat_input(2, 64, 256, 32)
    torch.manual_seed(5)
    o, nf, no = O.adapt_layer_off(fea, loc, sd, "g.node_fea_adapt", True)
    close(o.unsqueeze(3), g["out"])
    close(nf, g["node_fea"])
    close(no, g["node_off"])


def test_dgcnn_g(golden):
    g = golden("dgcnn_g")
    x, _ = O.synth_clouds(2, 1024, 41)
    sd = O.synth_state("Net_MDA:DGCNN")
    torch.manual_seed(7)
    f, n = O.dgcnn_trunk(x, sd, "g.", True, adapt=True)
    close(f, g["feat_train"], 1e-3, 1e-4)
    close(n, g["node_train"], 1e-3, 1e-4)
    torch.manual_seed(8)
    f, n = O.dgcnn_trunk(x, sd, "g.", False, adapt=True)
    close(f, g["feat_eval"], 1e-3, 1e-4)
    close(n, g["node_eval"], 1e-3, 1e-4)
    close(sd["g.conv1.conv.1.running_mean"], g["rm1"])
    close(sd["g.conv4.conv.1.running_var"], g["rv4"], 1e-4, 1e-6)
    close(sd["g.bn5.running_mean"], g["rm5"], 1e-4, 1e-5)
    close(sd["g.bn5.running_var"], g["rv5"], 1e-4, 1e-6)


def test_net_mda(golden):
    g = golden("net_mda_dgcnn")
    x, _ = O.synth_clouds(2, 1024, 41)
    sd = O.synth_state("Net_MDA:DGCNN")
    torch.manual_seed(11)
    y1, y2, s1, s2 = O.net_mda(x, sd, True, semantic_adaption=True)
    torch.manual_seed(12)
    ns = O.net_mda(x, sd, True, node_adaptation_s=True)
    torch.manual_seed(13)
    nt = O.net_mda(x, sd, True, node_adaptation_t=True)
    for a, k in ((y1, "y1"), (y2, "y2"), (s1, "s1"), (s2, "s2"), (ns, "node_s"), (nt, "node_t")):
        close(a, g[k], 1e-3, 1e-4)


def test_dgcnn_cls(golden):
    x, _ = O.synth_clouds(2, 1024, 41)
    sd = O.synth_state("DGCNN_cls", 667)
    close(O.dgcnn_cls(x, sd, False), golden("dgcnn_cls")["logits_eval"], 1e-3, 1e-4)


def test_pointnet_g(golden):
    g = golden("pointnet_g")
    x, _ = O.synth_clouds(2, 1024, 41)
    sd = O.synth_state("Net_MDA:Pointnet", 668)
    torch.manual_seed(17)
    f, n, off = O.pointnet_g(x, sd, "g.", True)
    close(f, g["feat"], 1e-3, 1e-4)
    close(n, g["node_fea"], 1e-3, 1e-4)
    close(off, g["node_off"], 1e-3, 1e-5)
    close(sd["g.conv5.conv.1.running_var"], g["rv5"], 1e-4, 1e-6)


def test_net_mda_pointnet(golden):
    """Net_MDA('Pointnet') (Model.py:485-520): eval-mode logits of both heads, train-mode node branches."""
    g = golden("net_mda_pointnet")
    x, _ = O.synth_clouds(2, 1024, 41)
    sd = O.synth_state("Net_MDA:Pointnet", 668)
    with torch.no_grad():
        torch.manual_seed(21)
        y1, y2 = O.net_mda(x, sd, False, model_name="Pointnet")
    close(y1, g["y1_eval"], 1e-3, 1e-4)
    close(y2, g["y2_eval"], 1e-3, 1e-4)
    torch.manual_seed(22)
    close(O.net_mda(x, sd, True, model_name="Pointnet", node_adaptation_s=True), g["node_s"], 1e-3, 1e-4)
    torch.manual_seed(23)
    close(O.net_mda(x, sd, True, model_name="Pointnet", node_adaptation_t=True), g["node_t"], 1e-3, 1e-4)


def mmd_inputs():
    rng = np.random.Generator(np.random.PCG64(51))
    m = 16
    t = lambda a: torch.from_numpy(np.asarray(a))
    X = t(rng.standard_normal((m, 4096)).astype(np.float32))
    Y = t((rng.standard_normal((m, 4096)) * 1.1 + 0.1).astype(np.float32))
    Xs = t(rng.standard_normal((m, 256)).astype(np.float32) * 0.5)
    Ys = t(rng.standard_normal((m, 256)).astype(np.float32) * 0.6)
    ls = t(rng.integers(0, 10, m).astype(np.int64))
    lt = t(rng.integers(0, 10, m).astype(np.int64))
    w = t(rng.uniform(0, 2, (1, m)).astype(np.float32))
    ps = t(rng.standard_normal((m, 10)).astype(np.float32))
    pt = t(rng.standard_normal((m, 10)).astype(np.float32))
    ds, dt = O.synth_clouds(m, 256, 52)[0], O.synth_clouds(m, 256, 53)[0]
    return X, Y, Xs, Ys, ls, lt, w, ps, pt, ds, dt


def test_mmd(golden):
    g = golden("mmd")
    X, Y, Xs, Ys, ls, lt, w, ps, pt, ds, dt = mmd_inputs()
    Xg, Yg = X.clone().requires_grad_(True), Y.clone().requires_grad_(True)
    close(O.mix_rbf_mmd2(Xg, Yg), g["v_plain"], 1e-5, 1e-6)
    v = O.mix_rbf_mmd2(Xg, Yg, sample_weights=w)
    v.backward()
    close(v, g["v_w"], 1e-5, 1e-6)
    close(Xg.grad, g["dX"], 1e-3, 1e-8)
    close(Yg.grad, g["dY"], 1e-3, 1e-8)
    Xsg, Ysg = Xs.clone().requires_grad_(True), Ys.clone().requires_grad_(True)
    v = O.mix_rbf_mmd2(Xsg, Ysg, sample_weights=w)
    v.backward()
    close(v, g["v_sem"], 1e-5, 1e-6)
    close(Xsg.grad, g["dXs"], 1e-3, 1e-7)
    close(O.geometric_weights(ds, dt), g["geo_w"], 1e-5, 1e-7)
    close(O.prob_weights_soft(ps, pt, ls, lt, 0.5), g["sem_w"], 1e-4, 1e-7)
    close(O.mmd_cal(ls, X, lt, Y, O.SUG_CFG["GEO_MMD"], ds, dt), g["geo"], 1e-5, 1e-6)
    close(O.mmd_cal(ls, Xs, lt, Ys, O.SUG_CFG["SEM_MMD"], ps, pt), g["sem"], 1e-5, 1e-6)
    c1, c2 = O.chamfer(ds.squeeze(-1).transpose(1, 2), dt.squeeze(-1).transpose(1, 2))
    close(c1, g["cd1"], 1e-5, 1e-7)
    close(c2, g["cd2"], 1e-5, 1e-7)


def test_sug_step(golden):
    g = golden("sug_step")
    Bs = 12
    data, label = O.synth_clouds(Bs, 1024, 0)
    data_t, label_t = O.synth_clouds(Bs, 1024, 1)
    sd = O.clone_state(O.synth_state("Net_MDA:DGCNN"), requires_grad=True)
    crit = O.FocalLoss([0.1] * 10, 0.0)
    torch.manual_seed(101)
    r = O.sug_losses(sd, data, label, data_t, label_t, crit, drop_p=0.0)
    r["loss"].backward()
    for k in ("loss", "loss_cls", "loss_geo", "loss_sem"):
        close(r[k], g[k], 1e-4, 1e-6)
    close(r["pred_s1"], g["pred_s1"], 1e-3, 1e-4)
    close(r["pred_t1"], g["pred_t1"], 1e-3, 1e-4)
    n_checked = 0
    for k, v in g.items():
        if k.startswith("gn."):
            close(sd[k[3:]].grad.norm(), v, 2e-3, 1e-7)
            n_checked += 1
        elif k.startswith("gf."):
            gr = sd[k[3:]].grad
            assert float((gr - torch.from_numpy(v)).norm()) <= 2e-3 * float(np.linalg.norm(v)) + 1e-7, k
    assert n_checked >= 50
    # parameters the reference never touches must stay gradient-free here too
    assert sd["g.input_transform_net.fc3.weight"].grad is None
    assert sd["g.node_fea_adapt.trans.conv.0.weight"].grad is None
