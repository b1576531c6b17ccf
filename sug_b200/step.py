"""One SUG domain-generalisation training step: the body of the reference's
``train_dg_single_gpu.py:246-335`` under
``tools/cfgs/cfgs_sproject/DG_unified_loss_onedataset_shapenet.yaml``, written against the
drop-in modules of this package.  Used by bench.py, the smoke test and the parity tests; the
reference trainer itself runs unchanged through ``sug_b200.compat``.
"""
from __future__ import annotations

import os

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


_PAIR_STREAMS = {}


class ConcurrentPasses:
    """Runs the encoder passes of a step (source / target batch, semantic / node variant) CONCURRENTLY on up to four
    CUDA streams, then the deferred BatchNorm updates, and offers the same streams as lanes for the loss terms.

    The passes share nothing but the weights (read-only) and the BatchNorm running buffers; the latter are updated
    through ``ops.BNRecorder`` AFTER the passes, pass by pass in the reference's order
    (train_dg_single_gpu.py:260-264, 309-310), so losses, gradients and buffers are what the sequential execution
    produces.  What the concurrency buys: the step's kernels come in 3.46 waves of CTAs (64 clouds on 148 SMs) and many
    of them are latency-bound (FPS, counting sorts, index kernels, small GEMMs): the other streams fill the tails.  The
    backward follows automatically (autograd runs every node on the stream of its forward).  Captured into the step's
    CUDA graph as fork / join edges.  Off for CPU tensors, eval mode and the opt-in shared trunk; then ``run`` calls
    the function in place, ``lanes`` returns None and ``lane`` / ``merge`` / ``join`` do nothing.

        with ConcurrentPasses(model, device) as cp:
            a = cp.run(0, lambda: model(x)); b = cp.run(1, lambda: model(y))     # concurrent
        lanes = cp.lanes(2)
        with cp.lane(lanes, 0): la = loss_a(a)
        with cp.lane(lanes, 1): lb = loss_b(b)
        cp.merge(lanes, (la, lb)); cp.join()                                     # current stream owns everything again
    """

    def __init__(self, model, device):
        from . import ops
        self.ops = ops
        self.dev = torch.device(device)
        self.on = (self.dev.type == "cuda" and model.training and ENABLE_CONCURRENT_PASSES
                   and not getattr(getattr(model, "g", None), "share_trunk", False))
        if self.on:
            key = (self.dev.index, torch.cuda.current_stream(self.dev).cuda_stream)
            if key not in _PAIR_STREAMS:
                _PAIR_STREAMS[key] = tuple(torch.cuda.Stream(device=self.dev) for _ in range(4))
            self.streams = _PAIR_STREAMS[key][:max(1, min(4, N_CONCURRENT))]
            self.rec = ops.BNRecorder()

    def __enter__(self):
        if self.on:
            self.main = torch.cuda.current_stream(self.dev)
            for st in self.streams:
                st.wait_stream(self.main)
            self.prev = self.ops.BN_RECORDER
            self.ops.BN_RECORDER = self.rec
        return self

    def run(self, i, fn):
        if not self.on:
            return fn()
        self.rec.current = i
        with torch.cuda.stream(self.streams[i % len(self.streams)]):
            out = fn()
        for t in (out if isinstance(out, (tuple, list)) else (out,)):
            if isinstance(t, torch.Tensor):
                t.record_stream(self.main)
        return out

    def __exit__(self, et, ev, tb):
        if self.on:
            self.ops.BN_RECORDER = self.prev
            for st in self.streams:
                self.main.wait_stream(st)
            if et is None:
                # the ~40 small multi-tensor launches of the deferred running-statistics updates go to a side stream:
                # nothing of the step reads the running buffers, so they run next to the losses; join() waits for them
                side = self.streams[-1]
                side.wait_stream(self.main)
                with torch.cuda.stream(side):
                    self.rec.apply()
                self.pending = side
        return False

    def lanes(self, n):
        """Up to ``n`` side streams forked from the current stream (None when the concurrency is off)."""
        if not self.on or len(self.streams) < 2:
            return None
        main = torch.cuda.current_stream(self.dev)
        lanes = [self.streams[i % (len(self.streams) - 1)] for i in range(n)]  # the last stream carries the BN updates
        for st in set(lanes):
            st.wait_stream(main)
        return lanes

    def lane(self, lanes, i):
        import contextlib
        return torch.cuda.stream(lanes[i]) if lanes else contextlib.nullcontext()

    def merge(self, lanes, tensors):
        """Joins the lanes into the current stream; ``tensors`` were produced on them."""
        if lanes:
            main = torch.cuda.current_stream(self.dev)
            for st in set(lanes):
                main.wait_stream(st)
            for t in tensors:
                t.record_stream(main)

    def join(self):
        """Makes the current stream wait for the deferred BatchNorm updates (call once, after the losses)."""
        side, self.pending = getattr(self, "pending", None), None
        if side is not None:
            torch.cuda.current_stream(self.dev).wait_stream(side)


ENABLE_CONCURRENT_PASSES = os.environ.get("SUG_B200_CONCURRENT_PASSES", "1") == "1"
N_CONCURRENT = int(os.environ.get("SUG_B200_N_CONCURRENT", "4"))


def sug_losses(model, data, label, data_t, label_t, criterion, cfg=SUG_CFG, mmd_fn=mmd.mmd_cal):
    """train_dg_single_gpu.py:260-324: four Net_MDA forwards, class-weighted CE on both heads and
    both sub-domains (target logits are scored against the SOURCE labels, lines 287-288), the
    geometric MMD on the node features and the semantic MMD on both heads.  The four passes run concurrently on up to
    four streams (``ConcurrentPasses``: same results, BatchNorm buffers updated afterwards in the reference's order, on a
    side stream), and so do the three MMD terms and the classification losses."""
    # all four forwards are independent of each other (the node passes recompute the encoder): four streams
    with ConcurrentPasses(model, data.device) as cp:
        pred_s1, pred_s2, sem_s1, sem_s2 = cp.run(0, lambda: model(data, semantic_adaption=True))
        pred_t1, pred_t2, sem_t1, sem_t2 = cp.run(1, lambda: model(data_t, semantic_adaption=True))
        feat_node_s = cp.run(2, lambda: model(data, node_adaptation_s=True))
        feat_node_t = cp.run(3, lambda: model(data_t, node_adaptation_t=True))
    geo, sem = cfg["GEO_MMD"][0], cfg["SEM_MMD"][0]
    # the three MMD terms are independent chains of small kernels (weights, Gram matrix, kernel sums): on three streams
    # when the passes ran concurrently (not with a collective inside mmd_fn: one communicator, one issue order)
    lanes = cp.lanes(3) if mmd_fn is mmd.mmd_cal else None
    with cp.lane(lanes, 0):
        loss_geo = cfg["MMD_WEIGHT"] * geo["GEO_SCALE"] * mmd_fn(label, feat_node_s, label_t, feat_node_t, geo,
                                                                 data_s=data, data_t=data_t)
    with cp.lane(lanes, 1):
        l1 = sem["SEM_SCALE"] * mmd_fn(label, sem_s1, label_t, sem_t1, sem, data_s=pred_s1, data_t=pred_t1)
    with cp.lane(lanes, 2):
        l2 = sem["SEM_SCALE"] * mmd_fn(label, sem_s2, label_t, sem_t2, sem, data_s=pred_s2, data_t=pred_t2)
    # the classification losses are issued after the forks, so they run next to the MMD chains
    loss_s = 0.5 * criterion(pred_s1, label) + 0.5 * criterion(pred_s2, label)
    if cfg["TARGET_LOSS"] > 0:
        loss_t = 0.5 * criterion(pred_t1, label) + 0.5 * criterion(pred_t2, label)
        loss = 0.5 * loss_s + 0.5 * loss_t
    else:
        loss = cfg["SRC_LOSS_WEIGHT"] * loss_s
    loss_cls = cfg["CLS_WEIGHT"] * loss
    cp.merge(lanes, (loss_geo, l1, l2))
    loss_sem = cfg["MMD_WEIGHT"] * (0.5 * l1 + 0.5 * l2)
    cp.join()
    return {"loss": loss_cls + loss_geo + loss_sem, "loss_cls": loss_cls, "loss_geo": loss_geo,
            "loss_sem": loss_sem, "pred_s1": pred_s1, "pred_t1": pred_t1}


def make_optimizers(model, opt=OPT_CFG, capturable=False, fused=None):
    """train_dg_single_gpu.py:191-203: three Adam optimizers; ``g`` is stepped by two of them.
    On CUDA the update is ``optim.FusedAdam`` (torch.optim.Adam arithmetic, one multi-tensor launch per
    param group, graph-capturable by construction); ``fused=False`` selects ``torch.optim.Adam``
    (``capturable=True`` keeps its step counters on the device for use inside a CUDA graph)."""
    lr, wd = opt["LR"], opt["WEIGHT_DECAY"]
    if fused is None:
        fused = next(model.parameters()).is_cuda
    if fused:
        from .optim import FusedAdam
        mk = lambda groups, lr_: FusedAdam(groups, lr=lr_, weight_decay=wd)
    else:
        mk = lambda groups, lr_: torch.optim.Adam(groups, lr=lr_, weight_decay=wd, capturable=capturable)
    params = [{'params': v} for k, v in model.g.named_parameters() if 'pred_offset' not in k]
    opt_g = mk(params, lr)
    opt_c = mk([{'params': model.c1.parameters()}, {'params': model.c2.parameters()}], lr)
    opt_dis = mk([{'params': model.g.parameters()}, {'params': model.attention_s.parameters()},
                  {'params': model.attention_t.parameters()}], lr * opt["LR_SCALER"])
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


class GraphedTrainStep:
    """The whole training step (4 forwards, losses, backward, 3 Adam updates) as ONE CUDA graph.

    Eager execution of the step is bound by the ~2 000 Python-side launches, not by the GPU; the
    graph removes that.  Shapes are static (``batch`` clouds of ``points`` points per sub-domain).
    Host-side state the reference keeps per call is preserved:
      * FPS start indices are still drawn with ``torch.randint`` on the CPU generator, four draws per
        step in call order (point_utils.py:17), and copied into a static device buffer;
      * ``focal_loss`` re-gathers its own ``alpha`` on every call (model_utils.py:168); inside the graph
        every step starts from the constructor's alpha, which is identical for the uniform class
        weights used here.
    ``optimizers`` come from ``make_optimizers`` (``FusedAdam``; or torch Adam with ``capturable=True``).
    """

    def __init__(self, model, optimizers, criterion, batch, points, device, cfg=SUG_CFG, mmd_fn=None,
                 grad_hook=None, warmup=3, nccl_in_graph=False, overlap=True):
        from . import point_utils
        self.model, self.opts, self.crit = model, optimizers, criterion
        self.cfg, self.mmd_fn, self.hook = cfg, (mmd_fn or mmd.mmd_cal), grad_hook
        self.N = points
        self.nccl_in_graph = bool(nccl_in_graph)
        # one graph with NCCL inside: the all-reduce is issued group by group from gradient hooks and overlaps the
        # remaining backward (dist.OverlappedGradAllReduce); the first eager warm-up step calibrates it
        self.overlap = None
        if self.nccl_in_graph and grad_hook is not None and overlap:
            from .dist import OverlappedGradAllReduce
            self.overlap = OverlappedGradAllReduce(model)
        dev = torch.device(device)
        self.data = torch.zeros(batch, 3, points, 1, device=dev)
        self.data_t = torch.zeros(batch, 3, points, 1, device=dev)
        self.label = torch.zeros(batch, dtype=torch.long, device=dev)
        self.label_t = torch.zeros(batch, dtype=torch.long, device=dev)
        self.fps_dev = torch.zeros(4, batch, dtype=torch.int32, device=dev)
        # pinned staging ring for the FPS start indices: a slot is rewritten only after the asynchronous upload
        # that last read it has completed (its event), so the host can run ahead of the GPU by several steps
        self.fps_host = [torch.zeros(4, batch, dtype=torch.int32).pin_memory() for _ in range(4)]
        self.fps_event = [None] * len(self.fps_host)
        self._fps_slot = 0
        self._call = 0
        self._alpha0 = criterion.alpha.detach().clone().to(dev) if hasattr(criterion, "alpha") else None

        def feed():
            t = self.fps_dev[self._call % 4]
            self._call += 1
            return t
        self._feed = feed
        self._pu = point_utils
        self.graph = None
        self.out = None

    def _fwd_bwd(self):
        self._call = 0
        if self._alpha0 is not None:
            self.crit.alpha = self._alpha0
        out = sug_losses(self.model, self.data, self.label, self.data_t, self.label_t, self.crit, self.cfg, self.mmd_fn)
        if self.overlap is not None:
            self.overlap.begin()
        out["loss"].backward()
        return {k: v.detach() for k, v in out.items()}

    def _body(self):
        out = self._fwd_bwd()
        if self.overlap is not None and self.overlap.finish():
            pass  # the gradient all-reduce ran group by group behind the backward
        elif self.hook is not None:
            self.hook(self.model)
        if self.overlap is not None and self.overlap.need is None:
            self.overlap.calibrate()
        for o in self.opts:
            o.step()
        return out

    def _draw_fps(self):
        slot = self._fps_slot
        self._fps_slot = (slot + 1) % len(self.fps_host)
        if self.fps_event[slot] is not None:
            self.fps_event[slot].synchronize()
        host = self.fps_host[slot]
        for i in range(4):  # same CPU-RNG consumption as four eager forwards
            host[i] = torch.randint(0, self.N, (host.shape[1],), dtype=torch.long).to(torch.int32)
        self.fps_dev.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.fps_event[slot] = ev

    def warm(self, data, label, data_t, label_t, iters=3):
        """Eager warm-up on a side stream (required before capture; these passes DO train the model)."""
        for dst, src in ((self.data, data), (self.label, label), (self.data_t, data_t), (self.label_t, label_t)):
            dst.copy_(src)
        self._pu.set_fps_start_feed(self._feed)
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(iters):
                    self._draw_fps()
                    self._body()
                    for o in self.opts:
                        o.zero_grad(set_to_none=True)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
        finally:
            self._pu.set_fps_start_feed(None)
        return self

    def capture(self):
        """Record the step into a CUDA graph (call ``warm`` first).  Nothing executes during the
        capture (the model is not updated), but one set of FPS start indices is drawn.  Gradients
        must be None so that the backward's first write of every ``.grad`` is an assignment inside
        the graph's pool."""
        for o in self.opts:
            o.zero_grad(set_to_none=True)
        self._pu.set_fps_start_feed(self._feed)
        try:
            self._draw_fps()
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            if self.hook is None:
                with torch.cuda.graph(self.graph):
                    self.out = self._body()
            elif self.nccl_in_graph:
                # data parallel, one graph: the coalesced gradient all-reduce is captured with the rest (NCCL supports
                # stream capture; the capture must not be disturbed by the process group's watchdog thread, hence
                # the thread-local capture mode)
                with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                    self.out = self._body()
            else:
                # data parallel, NCCL outside the graphs: graph A = forwards + backward, then ONE eager coalesced
                # all-reduce (AVG, in place on the gradients: no flatten / unflatten copies), graph B = the optimisers
                with torch.cuda.graph(self.graph):
                    self.out = self._fwd_bwd()
                    self._grads = [p.grad for p in self.model.parameters() if p.grad is not None]
                self.graph_b = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_b, pool=self.graph.pool()):
                    for o in self.opts:
                        o.step()
        finally:
            self._pu.set_fps_start_feed(None)
        torch.cuda.synchronize()
        return self

    def __call__(self, data, label, data_t, label_t):
        """One training step.  Inputs may live on the host (pinned) or on the device."""
        self.data.copy_(data, non_blocking=True)
        self.label.copy_(label, non_blocking=True)
        self.data_t.copy_(data_t, non_blocking=True)
        self.label_t.copy_(label_t, non_blocking=True)
        self._draw_fps()
        for o in self.opts:  # LR schedulers rewrite group['lr'] on the host: refresh the device scalars
            if hasattr(o, "sync_lr"):
                o.sync_lr()
        self.graph.replay()
        if self.hook is not None and not self.nccl_in_graph:
            self.hook(self.model, self._grads)
            self.graph_b.replay()
        return self.out


class GraphedEval:
    """An eval-mode forward (``model(x)`` under ``torch.no_grad()``) as one CUDA graph for a fixed input shape --
    e.g. ``model_pointnet.DGCNN`` at LiDAR scale (BASELINE.json configs[4]: N = 16 384, k = 20), where the ~30 launches
    of a forward are otherwise paced by the host.  The adapt layer of ``Net_MDA`` draws its FPS start on the CPU; pass
    ``fps_points`` (= N) to feed those indices from a device buffer so that the draw stays outside the graph.

        fwd = GraphedEval(net.eval(), x_example)      # captures after 2 eager warm-up passes
        logits = fwd(x)                               # x: host (pinned) or device tensor of the same shape
    """

    def __init__(self, model, example, fps_points=None, warmup=2):
        from . import point_utils
        assert not model.training, "GraphedEval captures an eval-mode forward"
        self.model = model
        self.x = example.detach().clone()
        self._pu = point_utils
        self.fps_dev = None
        if fps_points is not None:
            self.N = int(fps_points)
            self.fps_dev = torch.zeros(self.x.shape[0], dtype=torch.int32, device=self.x.device)
            self.fps_host = [torch.zeros(self.x.shape[0], dtype=torch.int32).pin_memory() for _ in range(4)]
            self.fps_event = [None] * 4
            self._slot = 0
        feed = (lambda: self.fps_dev) if self.fps_dev is not None else None
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        self._pu.set_fps_start_feed(feed)
        try:
            with torch.cuda.stream(s), torch.no_grad():
                for _ in range(warmup):
                    self._draw()
                    model(self.x)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self._draw()
            self.graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self.out = model(self.x)
        finally:
            self._pu.set_fps_start_feed(None)
        torch.cuda.synchronize()

    def _draw(self):
        if self.fps_dev is None:
            return
        slot = self._slot
        self._slot = (slot + 1) % len(self.fps_host)
        if self.fps_event[slot] is not None:
            self.fps_event[slot].synchronize()
        self.fps_host[slot].copy_(torch.randint(0, self.N, (self.fps_host[slot].shape[0],), dtype=torch.long).to(torch.int32))
        self.fps_dev.copy_(self.fps_host[slot], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.fps_event[slot] = ev

    def __call__(self, x):
        self.x.copy_(x, non_blocking=True)
        self._draw()
        self.graph.replay()
        return self.out
