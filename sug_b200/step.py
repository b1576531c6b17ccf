"""One SUG domain-generalisation training step: the body of the reference's
``train_dg_single_gpu.py:246-335`` under
``tools/cfgs/cfgs_sproject/DG_unified_loss_onedataset_shapenet.yaml``, written against the
drop-in modules of this package.  Used by bench.py, the smoke test and the parity tests; the
reference trainer itself runs unchanged through ``sug_b200.compat``.
"""
from __future__ import annotations

import torch

from . import mmd

SUG_CFG = {  # METHODS section of DG_unified_loss_onedataset_shapenet.yaml
    "MMD_WEIGHT": 0.5, "CLS_WEIGHT": 1.0, "TARGET_LOSS": 1.0, "SRC_LOSS_WEIGHT": 1.0, "ADV_WEIGHT": 0.0,
    "PURE_CLS_EPOCH": 0,
    "GEO_MMD": [{"NAME": "SOFT_MMD", "LABEL_SCALE": 50, "GEO_WEIGHTS": "mean2one", "GEO_SCALE": 1}],
    "SEM_MMD": [{"NAME": "SOFT_MMD", "LABEL_SCALE": 5, "SEM_WEIGHTS": "mean2one", "LABEL_WEIGHT": 0.5,
                 "SEM_SCALE": 1}],
}
OPT_CFG = {"LR": 1e-4, "LR_SCALER": 1.0, "WEIGHT_DECAY": 5e-4}  # OPTIMIZATION section


def sug_losses(model, data, label, data_t, label_t, criterion, cfg=SUG_CFG, mmd_fn=mmd.mmd_cal):
    """train_dg_single_gpu.py:260-324: four Net_MDA forwards, class-weighted CE on both heads and
    both sub-domains (target logits are scored against the SOURCE labels, lines 287-288), the
    geometric MMD on the node features and the semantic MMD on both heads."""
    pred_s1, pred_s2, sem_s1, sem_s2 = model(data, semantic_adaption=True)
    pred_t1, pred_t2, sem_t1, sem_t2 = model(data_t, semantic_adaption=True)
    loss_s = 0.5 * criterion(pred_s1, label) + 0.5 * criterion(pred_s2, label)
    if cfg["TARGET_LOSS"] > 0:
        loss_t = 0.5 * criterion(pred_t1, label) + 0.5 * criterion(pred_t2, label)
        loss = 0.5 * loss_s + 0.5 * loss_t
    else:
        loss = cfg["SRC_LOSS_WEIGHT"] * loss_s
    loss_cls = cfg["CLS_WEIGHT"] * loss
    feat_node_s = model(data, node_adaptation_s=True)
    feat_node_t = model(data_t, node_adaptation_t=True)
    geo, sem = cfg["GEO_MMD"][0], cfg["SEM_MMD"][0]
    loss_geo = cfg["MMD_WEIGHT"] * geo["GEO_SCALE"] * mmd_fn(label, feat_node_s, label_t, feat_node_t, geo,
                                                             data_s=data, data_t=data_t)
    l1 = sem["SEM_SCALE"] * mmd_fn(label, sem_s1, label_t, sem_t1, sem, data_s=pred_s1, data_t=pred_t1)
    l2 = sem["SEM_SCALE"] * mmd_fn(label, sem_s2, label_t, sem_t2, sem, data_s=pred_s2, data_t=pred_t2)
    loss_sem = cfg["MMD_WEIGHT"] * (0.5 * l1 + 0.5 * l2)
    return {"loss": loss_cls + loss_geo + loss_sem, "loss_cls": loss_cls, "loss_geo": loss_geo,
            "loss_sem": loss_sem, "pred_s1": pred_s1, "pred_t1": pred_t1}


def make_optimizers(model, opt=OPT_CFG):
    """train_dg_single_gpu.py:191-203: three Adam optimizers; ``g`` is stepped by two of them."""
    lr, wd = opt["LR"], opt["WEIGHT_DECAY"]
    params = [{'params': v} for k, v in model.g.named_parameters() if 'pred_offset' not in k]
    opt_g = torch.optim.Adam(params, lr=lr, weight_decay=wd)
    opt_c = torch.optim.Adam([{'params': model.c1.parameters()}, {'params': model.c2.parameters()}], lr=lr,
                             weight_decay=wd)
    opt_dis = torch.optim.Adam([{'params': model.g.parameters()}, {'params': model.attention_s.parameters()},
                                {'params': model.attention_t.parameters()}], lr=lr * opt["LR_SCALER"],
                               weight_decay=wd)
    return opt_dis, opt_g, opt_c


def train_step(model, optimizers, data, label, data_t, label_t, criterion, cfg=SUG_CFG, grad_hook=None):
    """Forward, backward and the three optimizer steps in the reference's order
    (train_dg_single_gpu.py:329-335).  ``grad_hook(model)`` runs between backward and the steps
    (data-parallel gradient all-reduce)."""
    out = sug_losses(model, data, label, data_t, label_t, criterion, cfg)
    out["loss"].backward()
    if grad_hook is not None:
        grad_hook(model)
    for o in optimizers:
        o.step()
    for o in (optimizers[1], optimizers[2], optimizers[0]):
        o.zero_grad()
    return out
