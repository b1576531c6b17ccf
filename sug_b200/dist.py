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


class OverlappedGradAllReduce:
    """Gradient all-reduce that starts while the backward is still running (what DistributedDataParallel's buckets
    do for the reference, train_dg.py:216-217), written for a step that is captured into ONE CUDA graph.

    The parameters are grouped by WHEN their gradient is complete in the SUG step's backward:
      group 0  attention_s / attention_t  (8.4 M of the 10.1 M trained parameters: used by the node passes only, their
               gradients are final right after the MMD backward, before any encoder backward has run),
      group 1  the classifier heads c1 / c2 (final once both semantic passes have been back-propagated through them),
      group 2  the encoder g (final at the very end).
    A post-accumulate-grad hook counts the finished parameters of each group; the group's coalesced in-place AVG
    all-reduce is issued on a side stream the moment the count is complete (fork / join edges when captured), so
    only the last, small group (4.7 MB) is exposed.  ``finish()`` joins the side stream and must be called after
    ``backward()`` and before the optimisers."""

    def __init__(self, model: torch.nn.Module):
        groups = [[], [], []]
        for name, p in model.named_parameters():
            if name.startswith("attention_"):
                groups[0].append(p)
            elif name.startswith("c1.") or name.startswith("c2."):
                groups[1].append(p)
            else:
                groups[2].append(p)
        self.groups = groups
        self.need = None          # parameters of each group that receive a gradient (known after one backward)
        self.count = [0] * len(groups)
        self.launched = [False] * len(groups)
        self.side = None
        self.enabled = False
        for gi, ps in enumerate(groups):
            for p in ps:
                p.register_post_accumulate_grad_hook(lambda _p, gi=gi: self._ready(gi))

    def calibrate(self):
        """Call once after a backward: records which parameters take part."""
        self.need = [sum(1 for p in ps if p.grad is not None) for ps in self.groups]

    def begin(self):
        """Call before ``backward()``."""
        self.count = [0] * len(self.groups)
        self.launched = [False] * len(self.groups)
        self.enabled = self.need is not None and world_size() > 1
        if self.enabled and self.side is None:
            self.side = torch.cuda.Stream()

    def _launch(self, gi):
        grads = [p.grad for p in self.groups[gi] if p.grad is not None]
        self.launched[gi] = True
        if not grads:
            return
        cur = torch.cuda.current_stream()
        self.side.wait_stream(cur)
        with torch.cuda.stream(self.side):
            # one flat bucket per group: a single NCCL all-reduce instead of one latency-bound operation per tensor
            # (54 tensors, 50 of them under 1 MB: measured +0.6 ms per step at 8 GPUs when reduced one by one)
            flat = torch.cat([g.reshape(-1) for g in grads])
            if dist.get_backend() == "nccl":
                dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                flat.div_(world_size())
            views, off = [], 0
            for g in grads:
                views.append(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
            torch._foreach_copy_(grads, views)

    def _ready(self, gi):
        if not self.enabled:
            return
        self.count[gi] += 1
        if self.count[gi] == self.need[gi] and not self.launched[gi]:
            self._launch(gi)

    def finish(self):
        """Call after ``backward()``: reduces whatever has not been launched and joins the side stream."""
        if not self.enabled:
            return False
        for gi in range(len(self.groups)):
            if not self.launched[gi]:
                self._launch(gi)
        torch.cuda.current_stream().wait_stream(self.side)
        self.enabled = False
        return True


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
    loss value with respect to the reference's per-rank MMD (train_dg.py:357-368); it is opt-in.

    The geometric SDA weights are a function of the per-pair Chamfer distances only (mmd.py:107-131, 198-201), so
    every rank evaluates the Chamfer kernel on ITS cloud pairs and the distances (one float per pair) are gathered --
    not the clouds: the same numbers as gathering the clouds first, without every rank redoing all ranks' pairs."""
    from . import mmd

    def gather_packed(cols):
        """ONE all-gather for several per-sample tensors: [m, d_i] float columns are concatenated, gathered (autograd
        aware) and split again -- an all-gather of a few hundred KB is latency-bound, so 3 instead of ~20 collectives
        per step."""
        cols = [c.reshape(c.shape[0], -1).float() for c in cols]
        widths = [c.shape[1] for c in cols]
        return torch.split(all_gather_rows(torch.cat(cols, dim=1)), widths, dim=1)

    def labels(t):
        return t.round().long().reshape(-1)

    geo, sem = args.get("GEO_WEIGHTS", None), args.get("SEM_WEIGHTS", None)
    if args["NAME"] == "SOFT_MMD" and data_s is not None and geo:
        pc_s, pc_t = data_s.detach(), data_t.detach()
        if pc_s.shape[1] == 3:
            pc_s = pc_s.reshape(pc_s.shape[0], 3, -1).transpose(1, 2)
            pc_t = pc_t.reshape(pc_t.shape[0], 3, -1).transpose(1, 2)
        d_loc = mmd.cd_distance(pc_s, pc_t)
        fs, ft, ls, lt, d_all = gather_packed([feat_s, feat_t, label_s, label_t, d_loc])
        weights = mmd.distance2weights(d_all.reshape(-1), method=geo).reshape(1, -1)
        return mmd.soft_mmd(labels(ls), fs, labels(lt), ft, float(args["LABEL_SCALE"]), sample_weights=weights)
    if args["NAME"] == "SOFT_MMD" and data_s is not None and sem:
        fs, ft, ls, lt, ps, pt = gather_packed([feat_s, feat_t, label_s, label_t, data_s.detach(), data_t.detach()])
        weights = mmd.prob_weights_soft(ps, pt, labels(ls), labels(lt), args["LABEL_WEIGHT"], sem)
        return mmd.soft_mmd(labels(ls), fs, labels(lt), ft, float(args["LABEL_SCALE"]), sample_weights=weights)
    g = all_gather_rows
    return mmd.mmd_cal(g(label_s), g(feat_s), g(label_t), g(feat_t), args,
                       data_s=None if data_s is None else g(data_s.detach()),
                       data_t=None if data_t is None else g(data_t.detach()), KPC=KPC)
