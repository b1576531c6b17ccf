"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

Run in the build container only (the reference tree does not travel to the GPU box):

    python tests/golden/make_golden.py

Inputs and weights come from oracle.sug_oracle.synth_* (numpy PCG64 => reproducible on any
box), so the fixtures hold OUTPUTS only.  The reference runs on the CPU in fp32 with the
device shim of oracle/ref_harness.py; every stochastic call is preceded by torch.manual_seed.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness, sug_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
warnings.simplefilter("ignore")


ONLY = None  # set by --only NAME: write just that fixture


def npz(name, **arrs):
    if ONLY is not None and name != ONLY:
        return
    conv = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **conv)
    print(f"{name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in conv.items()))


def feat_input(B, C, N, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.standard_normal((B, C, N)).astype(np.float32))


def load_ref_module(mod, spec_name, seed=666):
    sd = O.synth_state(spec_name, seed)
    ref_sd = mod.state_dict()
    assert set(ref_sd.keys()) == set(sd.keys()), (set(ref_sd) ^ set(sd))
    for k in ref_sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), (k, ref_sd[k].shape, sd[k].shape)
    mod.load_state_dict(sd, strict=True)
    return mod


def main():
    R = ref_harness.load()
    torch.set_num_threads(os.cpu_count())
    with ref_harness.cpu_device_shim():
        # ---- knn (model_utils.py:178-185) ------------------------------------------------
        for C, N in ((3, 256), (64, 256), (128, 128)):
            x = O.synth_clouds(2, N, 10 + C)[0].squeeze(-1) if C == 3 else feat_input(2, C, N, 10 + C)
            idx = R.model_utils.knn(x, 20)
            npz(f"knn_c{C}", idx=idx.to(torch.int16))

        # ---- EdgeConv block: get_graph_feature -> conv_2d(leakyrelu) -> max, fwd + bwd -----
        B, C, N, Co, k = 2, 8, 128, 16, 20
        x = feat_input(B, C, N, 21).requires_grad_(True)
        blk = R.model_utils.conv_2d(2 * C, Co, 1, activation="leakyrelu", bias=False)
        rng = np.random.Generator(np.random.PCG64(22))
        with torch.no_grad():
            blk.conv[0].weight.copy_(torch.from_numpy(rng.standard_normal((Co, 2 * C, 1, 1)).astype(np.float32) * 0.3))
            g = rng.uniform(0.5, 1.5, Co).astype(np.float32)
            g[::3] *= -1
            blk.conv[1].weight.copy_(torch.from_numpy(g))
            blk.conv[1].bias.copy_(torch.from_numpy(rng.standard_normal(Co).astype(np.float32) * 0.1))
        blk.train()
        Rw = torch.from_numpy(rng.standard_normal((B, Co, N)).astype(np.float32))
        out = blk(R.model_utils.get_graph_feature(x, k=k)).max(dim=-1)[0]
        (out * Rw).sum().backward()
        npz("edgeconv_block", out=out, dx=x.grad, dW=blk.conv[0].weight.grad, dgamma=blk.conv[1].weight.grad,
            dbeta=blk.conv[1].bias.grad, running_mean=blk.conv[1].running_mean, running_var=blk.conv[1].running_var)

        # ---- adapt_layer_off (model_utils.py:92-128) ---------------------------------------
        sd = O.synth_state("Net_MDA:DGCNN")
        ad = R.model_utils.adapt_layer_off()
        ad.load_state_dict({k[len("g.node_fea_adapt."):]: v for k, v in sd.items() if k.startswith("g.node_fea_adapt.")})
        ad.train()
        loc = O.synth_clouds(2, 256, 31)[0].squeeze(-1)
        fea = feat_input(2, 64, 256, 32)
        torch.manual_seed(5)
        o, nf, no = ad(fea.unsqueeze(3), loc)
        npz("adapt_layer", out=o, node_fea=nf, node_off=no)

        # ---- Model.DGCNN (Model.py:54-121), train then eval on the updated running stats ----
        x, _ = O.synth_clouds(2, 1024, 41)
        net = load_ref_module(R.Model.Net_MDA("DGCNN"), "Net_MDA:DGCNN")
        net.train()
        torch.manual_seed(7)
        f_tr, n_tr, _ = net.g(x, node=True)
        net.eval()
        torch.manual_seed(8)
        with torch.no_grad():
            f_ev, n_ev, _ = net.g(x, node=True)
        npz("dgcnn_g", feat_train=f_tr, node_train=n_tr, feat_eval=f_ev, node_eval=n_ev,
            rm1=net.g.conv1.conv[1].running_mean, rv4=net.g.conv4.conv[1].running_var,
            rm5=net.g.bn5.running_mean, rv5=net.g.bn5.running_var)

        # ---- Net_MDA('DGCNN') three modes (Model.py:485-520), train mode, seeded ------------
        net = load_ref_module(R.Model.Net_MDA("DGCNN"), "Net_MDA:DGCNN")
        net.train()
        torch.manual_seed(11)
        y1, y2, s1, s2 = net(x, semantic_adaption=True)
        torch.manual_seed(12)
        ns = net(x, node_adaptation_s=True)
        torch.manual_seed(13)
        nt = net(x, node_adaptation_t=True)
        npz("net_mda_dgcnn", y1=y1, y2=y2, s1=s1, s2=s2, node_s=ns, node_t=nt)

        # ---- model_pointnet.DGCNN (config-1 model), eval ------------------------------------
        cls = load_ref_module(R.model_pointnet.DGCNN(), "DGCNN_cls", seed=667)
        cls.eval()
        with torch.no_grad():
            lg = cls(x)
        npz("dgcnn_cls", logits_eval=lg)

        # ---- Pointnet_g (Model.py:235-283) ---------------------------------------------------
        pn = load_ref_module(R.Model.Net_MDA("Pointnet"), "Net_MDA:Pointnet", seed=668)
        pn.train()
        torch.manual_seed(17)
        pf, pnode, poff = pn.g(x, node=True)
        npz("pointnet_g", feat=pf, node_fea=pnode, node_off=poff, rv5=pn.g.conv5.conv[1].running_var)

        # ---- Net_MDA('Pointnet') modes (Model.py:485-520): eval logits, train-mode node branches ---
        pn2 = load_ref_module(R.Model.Net_MDA("Pointnet"), "Net_MDA:Pointnet", seed=668)
        pn2.eval()
        with torch.no_grad():
            torch.manual_seed(21)
            e1, e2 = pn2(x)
        pn2.train()
        torch.manual_seed(22)
        pns = pn2(x, node_adaptation_s=True)
        torch.manual_seed(23)
        pnt = pn2(x, node_adaptation_t=True)
        npz("net_mda_pointnet", y1_eval=e1, y2_eval=e2, node_s=pns, node_t=pnt)

        # ---- MMD (mmd.py) --------------------------------------------------------------------
        rng = np.random.Generator(np.random.PCG64(51))
        m = 16
        X = torch.from_numpy(rng.standard_normal((m, 4096)).astype(np.float32))
        Y = torch.from_numpy((rng.standard_normal((m, 4096)) * 1.1 + 0.1).astype(np.float32))
        Xs = torch.from_numpy(rng.standard_normal((m, 256)).astype(np.float32) * 0.5)
        Ys = torch.from_numpy(rng.standard_normal((m, 256)).astype(np.float32) * 0.6)
        ls = torch.from_numpy(rng.integers(0, 10, m).astype(np.int64))
        lt = torch.from_numpy(rng.integers(0, 10, m).astype(np.int64))
        w = torch.from_numpy(rng.uniform(0, 2, (1, m)).astype(np.float32))
        ps = torch.from_numpy(rng.standard_normal((m, 10)).astype(np.float32))
        pt = torch.from_numpy(rng.standard_normal((m, 10)).astype(np.float32))
        ds, dt = O.synth_clouds(m, 256, 52)[0], O.synth_clouds(m, 256, 53)[0]
        Xg, Yg = X.clone().requires_grad_(True), Y.clone().requires_grad_(True)
        v_plain = R.mmd.mix_rbf_mmd2(Xg, Yg, R.mmd.sigma_list)
        v_w = R.mmd.mix_rbf_mmd2(Xg, Yg, R.mmd.sigma_list, sample_weights=w)
        v_w.backward()
        Xsg, Ysg = Xs.clone().requires_grad_(True), Ys.clone().requires_grad_(True)
        v_sem_plain = R.mmd.mix_rbf_mmd2(Xsg, Ysg, R.mmd.sigma_list, sample_weights=w)
        v_sem_plain.backward()
        geo_w = R.mmd.geometric_weights(ds, dt, weighting="mean2one")
        sem_w = R.mmd.prob_weights_soft(ps, pt, ls, lt, 0.5, "mean2one")
        geo = R.mmd.mmd_cal(ls, X, lt, Y, O.SUG_CFG["GEO_MMD"], data_s=ds, data_t=dt)
        sem = R.mmd.mmd_cal(ls, Xs, lt, Ys, O.SUG_CFG["SEM_MMD"], data_s=ps, data_t=pt)
        cd1, cd2, _, _ = ref_harness._BruteChamfer()(ds.squeeze(-1).transpose(1, 2), dt.squeeze(-1).transpose(1, 2))
        npz("mmd", v_plain=v_plain, v_w=v_w, dX=Xg.grad, dY=Yg.grad, v_sem=v_sem_plain, dXs=Xsg.grad, dYs=Ysg.grad,
            geo_w=geo_w, sem_w=sem_w, geo=geo, sem=sem, cd1=cd1, cd2=cd2)

        # ---- the public helper functions as users call them (reference layouts) -----------------------
        # model_utils.get_graph_feature (188-210), point_utils.{index_points, query_ball_point, upsample_inter,
        # square_distance} (60-165), focal_loss with gamma = 2 and its alpha statefulness (131-176)
        xa = feat_input(2, 8, 64, 71)
        gf = R.model_utils.get_graph_feature(xa, k=5)
        pc = O.synth_clouds(2, 128, 72)[0].squeeze(-1)                      # [2,3,128]
        torch.manual_seed(73)
        fi = R.point_utils.farthest_point_sample(pc, 16)
        ctr = R.point_utils.index_points(pc, fi)                            # [2,3,16]
        qb = R.point_utils.query_ball_point(0.3, 8, pc, ctr)
        qk = R.point_utils.query_ball_point(None, 8, pc, ctr)
        grp = R.point_utils.index_points(pc, qb)                            # [2,3,16,8]
        nodes_f = feat_input(2, 32, 16, 74)
        up = R.point_utils.upsample_inter(pc, ctr, None, nodes_f, 3)
        sqd = R.point_utils.square_distance(pc, ctr)                        # [2,128,16]
        fl = R.model_utils.focal_loss(num_classes=10, gamma=2, alpha=[0.05 * (i + 1) for i in range(10)], size_average=False)
        rngf = np.random.Generator(np.random.PCG64(75))
        lg = torch.from_numpy(rngf.standard_normal((12, 10)).astype(np.float32) * 2)
        lb = torch.from_numpy(rngf.integers(0, 10, 12).astype(np.int64))
        f1 = fl(lg, lb)
        f2 = fl(lg * 0.5, (lb + 3) % 10)  # second call: alpha has been re-gathered by the first (line 168)
        npz("api_funcs", graph_feature=gf, fps=fi, centres=ctr, ball=qb, knn8=qk, grouped=grp, upsample=up, sqdist=sqd,
            focal1=f1, focal2=f2)

        # ---- the MMD modes the SUG config does not use (mmd.py:25-41, 69-77, 178-202, 274-312) ---
        lt_same = ls.clone()
        lt_same[::3] = (lt_same[::3] + 1) % 10  # about two thirds of the pairs share their label
        hard = R.mmd.mmd_cal(ls, Xs, lt_same, Ys, {"NAME": "HARD_MMD"})
        off = R.mmd.mmd_cal(ls, Xs, lt, Ys, {"NAME": "OFF"})
        unb = R.mmd.mix_rbf_mmd2(Xs, Ys, R.mmd.sigma_list, biased=False)
        cdv = R.mmd.cd_distance(ds.squeeze(-1).transpose(1, 2), dt.squeeze(-1).transpose(1, 2), ref_harness._BruteChamfer())
        # ("naive_inverse" / "exp_inverse" build Python lists and crash at mmd.py:202 in the reference)
        w_none = R.mmd.distance2weights(cdv, method="none")
        geo_none = R.mmd.geometric_weights(ds, dt, weighting="none")
        # MAX_HARD_MMD (mmd.py:96-105) and the entropy distance behind entropy_weights (mmd.py:155-166; with a weighting
        # the reference can execute -- its default "exp_inverse" crashes at mmd.py:202)
        max_hard = R.mmd.mmd_cal(ls, Xs, lt, Ys, {"NAME": "MAX_HARD_MMD"})
        mh_s, mh_t = R.common_utils.get_most_overlapped_element(ls, lt)
        prob_s, prob_t = torch.softmax(Xs[:, :10], 1), torch.softmax(Ys[:, :10], 1)
        ent_dis = R.mmd.entropy_dis(prob_s, prob_t)
        ent_w = R.mmd.entropy_weights(prob_s, prob_t, weighting="mean2one")
        npz("mmd_modes", hard=hard, off=off, unbiased=unb, lt_same=lt_same, cd=cdv, w_none=w_none, geo_none=geo_none,
            max_hard=max_hard, mh_s=np.asarray(mh_s), mh_t=np.asarray(mh_t), ent_dis=ent_dis, ent_w=ent_w)

        # ---- whole SUG step (train_dg_single_gpu.py:260-329), dropout off -----------------------------
        # B = 12 (>= 10: the reference focal_loss re-gathers its own alpha, model_utils.py:168) and B = 64 + 64, the
        # batch of BASELINE.json configs[1] that bench.py times (needs ~30 GiB of host memory, a few minutes)
        for name, Bs in (("sug_step", 12), ("sug_step_b64", 64)):
            if ONLY is not None and name != ONLY:
                continue
            if ONLY is None and Bs == 64 and os.environ.get("SUG_GOLDEN_B64", "0") != "1":
                print("sug_step_b64: skipped (set SUG_GOLDEN_B64=1 or use --only sug_step_b64)")
                continue
            ref_step(R, name, Bs)


def ref_step(R, name, Bs):
    data, label = O.synth_clouds(Bs, 1024, 0)
    data_t, label_t = O.synth_clouds(Bs, 1024, 1)
    net = load_ref_module(R.Model.Net_MDA("DGCNN"), "Net_MDA:DGCNN")
    net.train()
    for hd in (net.c1, net.c2):
        hd.dropout1.p = 0.0
        hd.dropout2.p = 0.0
    crit = R.model_utils.focal_loss(num_classes=10, gamma=0.0, alpha=[0.1] * 10)
    cfg = O.SUG_CFG
    # every neighbour list the reference computes, as one 16-bit hash per row of the SORTED list: lets the GPU test
    # count the rows whose neighbour SET differs from the reference's at full size without storing 40 MB of indices
    hashes = []
    ref_knn = R.model_utils.knn

    def knn_rec(x, k):
        idx = ref_knn(x, k)
        hashes.append(O.knn_row_hash(idx))
        return idx
    R.model_utils.knn = knn_rec
    try:
        torch.manual_seed(101)
        ps1, ps2, ss1, ss2 = net(data, semantic_adaption=True)
        pt1, pt2, st1, st2 = net(data_t, semantic_adaption=True)
        loss_s = 0.5 * crit(ps1, label) + 0.5 * crit(ps2, label)
        loss_t = 0.5 * crit(pt1, label) + 0.5 * crit(pt2, label)
        loss_cls = cfg["CLS_WEIGHT"] * (0.5 * loss_s + 0.5 * loss_t)
        fns = net(data, node_adaptation_s=True)
        fnt = net(data_t, node_adaptation_t=True)
    finally:
        R.model_utils.knn = ref_knn
    loss_geo = cfg["MMD_WEIGHT"] * R.mmd.mmd_cal(label, fns, label_t, fnt, cfg["GEO_MMD"], data_s=data, data_t=data_t)
    l1 = R.mmd.mmd_cal(label, ss1, label_t, st1, cfg["SEM_MMD"], data_s=ps1, data_t=pt1)
    l2 = R.mmd.mmd_cal(label, ss2, label_t, st2, cfg["SEM_MMD"], data_s=ps2, data_t=pt2)
    loss_sem = cfg["MMD_WEIGHT"] * (0.5 * l1 + 0.5 * l2)
    loss = loss_cls + loss_geo + loss_sem
    loss.backward()
    grads = {}
    for n, p in net.named_parameters():
        if p.grad is not None:
            grads["gn." + n] = p.grad.norm()
    keep = ["g.conv1.conv.0.weight", "g.conv2.conv.1.weight", "g.conv4.conv.1.bias", "g.bn5.weight",
            "g.node_fea_adapt.pred_offset.0.weight", "g.conv1d.bias", "c1.mlp3.weight", "c1.mlp3.bias", "c2.mlp3.weight",
            "c2.mlp3.bias"]
    full = {"gf." + n: dict(net.named_parameters())[n].grad for n in keep}
    extra = {}
    if Bs > 12:  # the full-size fixture also pins the node / semantic features and the neighbour graphs
        # ... and carries the gradient norms of the SAME loss with the MMD autograd evaluated in fp64 (oracle; with the
        # fp32 MMD the oracle reproduces the reference's norms to 1e-6, tests/test_oracle_golden.py): the reference's own
        # fp32 MMD gradient is cancellation noise for the 4096-d node features, so the parameters
        # upstream of the MMD cannot be judged against gn.* alone -- the oracle cannot be run at this size on the GPU box
        rm_conv4, rv_bn5 = net.g.conv4.conv[1].running_mean.clone(), net.g.bn5.running_var.clone()
        loss = loss.detach().clone()
        del net, ps2, pt2, ss2, st1, st2, fnt
        import gc
        gc.collect()
        sd = O.clone_state(O.synth_state("Net_MDA:DGCNN"), requires_grad=True)
        torch.manual_seed(101)
        out64 = O.sug_losses(sd, data, label, data_t, label_t, O.FocalLoss([0.1] * 10, 0.0), drop_p=0.0, mmd_dtype=torch.float64)
        out64["loss"].backward()
        for n, v in sd.items():
            if v.grad is not None:
                grads["gn64." + n] = v.grad.norm()
        extra = dict(knn_hash=torch.stack(hashes).numpy().astype(np.uint16), feat_node_s=fns[:, :64], sem_s1=ss1[:, :32],
                     rm_conv4=rm_conv4, rv_bn5=rv_bn5)
    npz(name, loss=loss, loss_cls=loss_cls, loss_geo=loss_geo, loss_sem=loss_sem, pred_s1=ps1, pred_t1=pt1,
        **grads, **full, **extra)


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--only":
        ONLY = sys.argv[2]
    main()
