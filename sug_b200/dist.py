"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU
box, gloo in CPU tests).  The encoder shards over independent clouds with no data-path collective
(plain per-rank BatchNorm, exactly like the reference's DDP intent, train_dg.py:216-217); the only
exchanges are the gradient all-reduce and — with ``mmd_scope='global'`` — the all-gather of the MMD
inputs named by BASELINE.json's north_star.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, 0
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def grads_of(model: torch.nn.Module):
    return [p.grad for p in model.parameters() if p.grad is not None]


def allreduce_grads(model: torch.nn.Module, grads=None):
    """Average the gradients of every parameter that received one, IN PLACE, as ONE coalesced NCCL launch
    (ncclGroupStart / End around one all-reduce per tensor with the AVG reduction: no flatten / unflatten copies, no
    separate division).  Parameters the step never touches (g.input_transform_net.*, g.node_fea_adapt.trans.*;
    806 793 of 10.9 M) keep ``grad is None`` on every rank, as with the reference's ``find_unused_parameters=True``
    (train_dg.py:217), so Adam skips them exactly as on one GPU.  gloo (CPU tests) has neither coalescing nor AVG:
    there the tensors are reduced one by one and divided."""
    w = world_size()
    if w == 1:
        return
    if grads is None:
        grads = grads_of(model)
    if not grads:
        return
    if dist.get_backend() == "nccl":
        with dist._coalescing_manager(device=grads[0].device, async_ops=False):
            for g in grads:
                dist.all_reduce(g, op=dist.ReduceOp.AVG)
    else:
        for g in grads:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            g.div_(w)


class _AllGatherRows(torch.autograd.Function):
    """[m, D] per rank -> [W*m, D]; every rank then evaluates the same global loss, so the backward
    hands each rank its own slice scaled by W (the gradient average divides it back)."""

    @staticmethod
    def forward(ctx, x):
        w = world_size()
        ctx.m = x.shape[0]
        x = x.contiguous()
        out = x.new_empty((w * x.shape[0],) + tuple(x.shape[1:]))
        dist.all_gather_into_tensor(out, x)
        return out

    @staticmethod
    def backward(ctx, g):
        r, w = dist.get_rank(), world_size()
        return g[r * ctx.m:(r + 1) * ctx.m] * w


def all_gather_rows(x: torch.Tensor) -> torch.Tensor:
    if world_size() == 1:
        return x
    if x.requires_grad:
        return _AllGatherRows.apply(x)
    with torch.no_grad():
        return _AllGatherRows.apply(x)


def global_mmd_cal(label_s, feat_s, label_t, feat_t, args, data_s=None, data_t=None, KPC=False):
    """``mmd.mmd_cal`` over the global batch: features, labels and the SDA weight inputs of all
    ranks are gathered first (m = 64 * world).  This changes m, the mean2one scale and hence the
    loss value with respect to the reference's per-rank MMD (train_dg.py:357-368); it is opt-in."""
    from . import mmd
    g = all_gather_rows
    return mmd.mmd_cal(g(label_s), g(feat_s), g(label_t), g(feat_t), args,
                       data_s=None if data_s is None else g(data_s.detach()),
                       data_t=None if data_t is None else g(data_t.detach()), KPC=KPC)
