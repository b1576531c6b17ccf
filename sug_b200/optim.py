"""``FusedAdam``: torch.optim.Adam semantics (the optimizer of train_dg_single_gpu.py:191-203) as ONE
multi-tensor CUDA launch per step (``sug_adam_f32``, csrc/optim.cu).

Same update rule as ``torch.optim.Adam`` (L2 ``weight_decay`` folded into the gradient, bias-corrected
moments, ``amsgrad=False``), same skipping of parameters whose ``.grad`` is ``None``, same
``param_groups`` / ``state`` layout (``step`` is one device scalar shared by the tensors of a group).
The step counter and the learning rate are device scalars, so a step can be captured in a CUDA graph
and an LR scheduler that rewrites ``group['lr']`` is honoured on the next (eager or replayed) step via
``sync_lr()``.  No CPU fallback: parameters must live on a CUDA device.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib


class _GroupPlan:
    """Device tables for one param group: rebuilt whenever a pointer (typically a fresh ``.grad``) moves."""

    def __init__(self):
        self.sig = None
        self.tables = None      # device int64 [5, T]
        self.blk = None         # device int32 [2, n_blocks]
        self.n_blocks = 0
        self.n_params = 0
        self.stages = []        # pinned staging buffers [h_tab, h_blk, event, frozen]; a captured graph
                                # re-reads its (frozen) buffer at every replay, so those are never reused
        self.keep = None

    def stage(self, shape_tab, shape_blk, capturing):
        for st in self.stages:
            if st[3] or st[0].shape != shape_tab or st[1].shape != shape_blk:
                continue
            if st[2] is not None and not capturing:
                st[2].synchronize()  # the previous asynchronous upload from this buffer has finished
            st[3] = capturing
            return st
        if capturing:
            raise RuntimeError("FusedAdam: run one eager step before capturing the step into a CUDA graph "
                               "(pinned staging buffers cannot be allocated during capture)")
        st = [torch.empty(shape_tab, dtype=torch.int64).pin_memory(), torch.empty(shape_blk, dtype=torch.int32).pin_memory(),
              None, False]
        self.stages.append(st)
        return st


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("FusedAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._plans = {}
        self._gstate = {}
        self._chunk = int(_lib.load().sug_adam_chunk())

    # -- state ------------------------------------------------------------------------------------
    def _group_state(self, gi, group, device):
        gs = self._gstate.get(gi)
        if gs is None:
            gs = self._gstate[gi] = {"step": torch.zeros((), dtype=torch.float32, device=device),
                                  "lr": torch.full((), float(group["lr"]), dtype=torch.float32, device=device),
                                  "lr_host": float(group["lr"])}
        return gs

    def sync_lr(self):
        """Push ``group['lr']`` (as rewritten by LR schedulers) to the device scalars.  Called by ``step``;
        call it yourself before replaying a CUDA graph that contains the step."""
        for gi, group in enumerate(self.param_groups):
            gs = self._gstate.get(gi)
            if gs is not None and gs["lr_host"] != float(group["lr"]):
                gs["lr"].fill_(float(group["lr"]))
                gs["lr_host"] = float(group["lr"])

    def _plan(self, gi, group, gs):
        ps = [p for p in group["params"] if p.grad is not None]
        for p in ps:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdam runs on contiguous fp32 CUDA parameters only (there is no CPU fallback)")
            if p.grad.is_sparse:
                raise RuntimeError("FusedAdam does not support sparse gradients")
            st = self.state[p]
            if not st:
                st["step"] = gs["step"]
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        plan = self._plans.setdefault(gi, _GroupPlan())
        grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
        # every pointer the device tables hold: a fresh .grad, or moments replaced by load_state_dict, rebuilds them
        sig = tuple((p.data_ptr(), g.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(),
                     p.numel()) for p, g in zip(ps, grads))
        if sig != plan.sig:
            dev = ps[0].device if ps else None
            rows = [[p.data_ptr() for p in ps], [g.data_ptr() for g in grads],
                    [self.state[p]["exp_avg"].data_ptr() for p in ps],
                    [self.state[p]["exp_avg_sq"].data_ptr() for p in ps], [p.numel() for p in ps]]
            bt, bc = [], []
            for t, p in enumerate(ps):
                nchunk = (p.numel() + self._chunk - 1) // self._chunk
                bt.extend([t] * nchunk)
                bc.extend(range(nchunk))
            plan.n_blocks, plan.n_params = len(bt), sum(p.numel() for p in ps)
            if ps:
                capturing = torch.cuda.is_current_stream_capturing()
                t_tab, t_blk = torch.tensor(rows, dtype=torch.int64), torch.tensor([bt, bc], dtype=torch.int32)
                st = plan.stage(t_tab.shape, t_blk.shape, capturing)
                st[0].copy_(t_tab)
                st[1].copy_(t_blk)
                if plan.tables is None or plan.tables.shape != t_tab.shape or plan.blk.shape != t_blk.shape:
                    plan.tables = torch.empty(t_tab.shape, dtype=torch.int64, device=dev)
                    plan.blk = torch.empty(t_blk.shape, dtype=torch.int32, device=dev)
                plan.tables.copy_(st[0], non_blocking=True)
                plan.blk.copy_(st[1], non_blocking=True)
                if not capturing:
                    st[2] = torch.cuda.Event()
                    st[2].record(torch.cuda.current_stream(dev))
            plan.sig = sig
            plan.keep = [g for p, g in zip(ps, grads) if g is not p.grad]  # contiguous copies must outlive the launch
        return plan, ps

    # -- (de)serialisation --------------------------------------------------------------------------
    def _adopt_loaded_state(self):
        """After ``load_state_dict`` / unpickling: the cached device tables point at the OLD moment buffers and the
        group step counters know nothing of the loaded ``step`` values.  Drop the tables, restore one device step
        scalar per group from the loaded state (torch.optim.Adam layout: a ``step`` per parameter) and re-alias every
        ``state[p]['step']`` to it, so that bias correction continues where the checkpoint stopped and a later
        ``state_dict()`` saves the live counter."""
        self._plans = {}
        self._gstate = {}
        for gi, group in enumerate(self.param_groups):
            steps = []
            for p in group["params"]:
                st = self.state.get(p)
                if st and "step" in st:
                    steps.append(float(st["step"]))
            if not steps:
                continue
            if max(steps) != min(steps):
                raise RuntimeError("FusedAdam keeps one step counter per param group; the loaded state has parameters of "
                                   f"group {gi} at different steps ({min(steps)} .. {max(steps)})")
            dev = next(p.device for p in group["params"] if self.state.get(p))
            gs = self._group_state(gi, group, dev)
            gs["step"].fill_(steps[0])
            for p in group["params"]:
                st = self.state.get(p)
                if st:
                    st["step"] = gs["step"]
                    for k in ("exp_avg", "exp_avg_sq"):
                        if not st[k].is_contiguous():
                            st[k] = st[k].contiguous()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._adopt_loaded_state()

    def __setstate__(self, state):
        super().__setstate__(state)
        self._chunk = int(_lib.load().sug_adam_chunk())
        self._adopt_loaded_state()

    # -- step -------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        capturing = None
        for gi, group in enumerate(self.param_groups):
            first = next((p for p in group["params"] if p.grad is not None), None)
            if first is None:
                continue
            if not first.is_cuda:
                raise RuntimeError("FusedAdam runs on contiguous fp32 CUDA parameters only (there is no CPU fallback)")
            if capturing is None:
                capturing = torch.cuda.is_current_stream_capturing()
            gs = self._group_state(gi, group, first.device)
            if not capturing:
                self.sync_lr()
            plan, ps = self._plan(gi, group, gs)
            if not ps:
                continue
            tab, blk = plan.tables, plan.blk
            T = tab.shape[1]
            base, es = tab.data_ptr(), 8 * T
            b1, b2 = group["betas"]
            stream = ctypes.c_void_p(torch.cuda.current_stream(first.device).cuda_stream)
            _lib.check(lib.sug_adam_f32(ctypes.c_void_p(base), ctypes.c_void_p(base + es), ctypes.c_void_p(base + 2 * es),
                                        ctypes.c_void_p(base + 3 * es), ctypes.c_void_p(base + 4 * es),
                                        ctypes.c_void_p(blk.data_ptr()), ctypes.c_void_p(blk.data_ptr() + 4 * plan.n_blocks),
                                        plan.n_blocks, plan.n_params, ctypes.c_void_p(gs["step"].data_ptr()),
                                        ctypes.c_void_p(gs["lr"].data_ptr()), float(b1), float(b2), float(group["eps"]),
                                        float(group["weight_decay"]), stream), "sug_adam_f32")
        return loss
