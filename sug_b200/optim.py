"""``FusedAdam``: torch.optim.Adam semantics (the optimizer of train_dg_single_gpu.py:191-203) as ONE
multi-tensor CUDA launch per step for ALL param groups (``sug_adam_multi_f32``, csrc/optim.cu; the trainer's optimizer_g
has one group per parameter).

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
        self.n_groups = 0
        self.offsets = None
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

    def _plan_all(self, live):
        """One set of device tables for ALL param groups with gradients (``live`` = [(gi, group, gs)])."""
        ps, owner = [], []
        for gi, group, gs in live:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam runs on contiguous fp32 CUDA parameters only (there is no CPU fallback)")
                if p.grad.is_sparse:
                    raise RuntimeError("FusedAdam does not support sparse gradients")
                st = self.state[p]
                if not st:
                    st["step"] = gs["step"]
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                ps.append(p)
                owner.append((group, gs))
        plan = self._plans.setdefault("all", _GroupPlan())
        grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
        # every pointer / hyper-parameter the device tables hold: a fresh .grad, moments replaced by load_state_dict
        # or an edited group rebuilds them
        sig = tuple((p.data_ptr(), g.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(),
                     p.numel(), gs["step"].data_ptr(), gs["lr"].data_ptr(), tuple(group["betas"]), group["eps"],
                     group["weight_decay"]) for p, g, (group, gs) in zip(ps, grads, owner))
        if sig != plan.sig:
            dev = ps[0].device if ps else None
            rows = [[p.data_ptr() for p in ps], [g.data_ptr() for g in grads],
                    [self.state[p]["exp_avg"].data_ptr() for p in ps],
                    [self.state[p]["exp_avg_sq"].data_ptr() for p in ps], [p.numel() for p in ps],
                    [gs["step"].data_ptr() for _, gs in owner], [gs["lr"].data_ptr() for _, gs in owner]]
            steps = sorted({gs["step"].data_ptr() for _, gs in owner})
            hyper = [[float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"])] for g, _ in owner]
            bt, bc = [], []
            for t, p in enumerate(ps):
                nchunk = (p.numel() + self._chunk - 1) // self._chunk
                bt.extend([t] * nchunk)
                bc.extend(range(nchunk))
            plan.n_blocks, plan.n_params, plan.n_groups = len(bt), sum(p.numel() for p in ps), len(steps)
            if ps:
                capturing = torch.cuda.is_current_stream_capturing()
                T = len(ps)
                # one int64 table: 7 rows of T entries, the distinct step-counter addresses, then (as raw bits) the
                # float hyper-parameters and the int32 block map -- a single upload
                t_tab = torch.tensor(rows, dtype=torch.int64).reshape(-1)
                t_steps = torch.tensor(steps, dtype=torch.int64)
                t_hyper = torch.tensor(hyper, dtype=torch.float32).reshape(-1)
                t_blk = torch.tensor([bt, bc], dtype=torch.int32).reshape(-1)
                pad = lambda n: (n + 1) // 2 * 2  # noqa: E731  (int32 / fp32 words -> whole int64 words)
                host = torch.zeros(t_tab.numel() + t_steps.numel() + pad(t_hyper.numel()) // 2 + pad(t_blk.numel()) // 2,
                                   dtype=torch.int64)
                o1 = t_tab.numel()
                o2 = o1 + t_steps.numel()
                o3 = o2 + pad(t_hyper.numel()) // 2
                host[:o1] = t_tab
                host[o1:o2] = t_steps
                host[o2:o3].view(torch.float32)[:t_hyper.numel()] = t_hyper
                host[o3:].view(torch.int32)[:t_blk.numel()] = t_blk
                st = plan.stage(host.shape, (1,), capturing)
                st[0].copy_(host)
                if plan.tables is None or plan.tables.shape != host.shape:
                    plan.tables = torch.empty(host.shape, dtype=torch.int64, device=dev)
                plan.tables.copy_(st[0], non_blocking=True)
                if not capturing:
                    st[2] = torch.cuda.Event()
                    st[2].record(torch.cuda.current_stream(dev))
                plan.offsets = (T, o1, o2, o3)
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
        live = []
        for gi, group in enumerate(self.param_groups):
            first = next((p for p in group["params"] if p.grad is not None), None)
            if first is None:
                continue
            if not first.is_cuda:
                raise RuntimeError("FusedAdam runs on contiguous fp32 CUDA parameters only (there is no CPU fallback)")
            live.append((gi, group, self._group_state(gi, group, first.device)))
        if not live:
            return loss
        dev = next(p for p in live[0][1]["params"] if p.grad is not None).device
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()
        plan, ps = self._plan_all(live)
        if not ps:
            return loss
        T, o1, o2, o3 = plan.offsets
        base = plan.tables.data_ptr()
        row = lambda r: ctypes.c_void_p(base + 8 * T * r)  # noqa: E731
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.sug_adam_multi_f32(row(0), row(1), row(2), row(3), row(4), row(5), row(6), ctypes.c_void_p(base + 8 * o2),
                                          ctypes.c_void_p(base + 8 * o3), ctypes.c_void_p(base + 8 * o3 + 4 * plan.n_blocks),
                                          plan.n_blocks, plan.n_params, ctypes.c_void_p(base + 8 * o1), plan.n_groups, stream),
                   "sug_adam_multi_f32")
        return loss
