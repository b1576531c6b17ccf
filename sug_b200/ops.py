"""Host-side operators over libsug_b200.so: torch tensors in, torch tensors out, autograd-aware.

PyTorch is used only for device memory, the current stream and the autograd tape; all arithmetic
of the hot path runs in the CUDA library through the C ABI (include/sug_b200.h).  Feature tensors
are *point-major*: ``[B, N, C]`` with the channel stride 1, i.e. one row per point.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib

_WS = {}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sug_b200 operators run on CUDA tensors only (there is no CPU fallback)")


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


_SIDE = {}


def _side_stream(device):
    """One auxiliary stream per (device, current stream) for latency-bound index kernels that overlap the main stream's work
    (adapt_layer_off.prefetch_indices / forward_pm)."""
    key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)
    st = _SIDE.get(key)
    if st is None:
        st = _SIDE[key] = torch.cuda.Stream(device=device)
    return st


class BNRecorder:
    """Deferred BatchNorm side effects.  While a recorder is installed (``ops.BN_RECORDER``) the fused ops of this module
    do NOT touch running_mean / running_var / num_batches_tracked; they record (buffers, batch statistics) per pass
    instead, and ``apply()`` performs the momentum updates afterwards, pass by pass in the recorded order.  That is what
    lets two encoder passes of one step run CONCURRENTLY on two streams (step.sug_losses): the exponential moving
    average is order dependent (first the source pass, then the target pass, train_dg_single_gpu.py:260-264), so the
    updates cannot race -- the batch statistics, the outputs and the gradients do not depend on the running buffers."""

    def __init__(self):
        self.passes = {}
        self.current = 0

    def _p(self):
        return self.passes.setdefault(self.current, {"bn": [], "nbt": []})

    def add(self, running_mean, running_var, save, count, momentum, eps):
        self._p()["bn"].append((running_mean, running_var, save, float(count), float(momentum), float(eps)))

    def tick(self, nbt):
        self._p()["nbt"].append(nbt)

    @torch.no_grad()
    def apply(self):
        for key in sorted(self.passes):
            ent = self.passes[key]
            if ent["nbt"]:
                torch._foreach_add_(ent["nbt"], 1)
            if not ent["bn"]:
                continue
            rms = [e[0] for e in ent["bn"]]
            rvs = [e[1] for e in ent["bn"]]
            C = [e[0].numel() for e in ent["bn"]]
            means = [e[2][:c] for e, c in zip(ent["bn"], C)]
            invstds = [e[2][c:] for e, c in zip(ent["bn"], C)]
            mom = [e[4] for e in ent["bn"]]
            # biased batch variance back from invstd = 1 / sqrt(var + eps); unbiased for the running estimate
            var = torch._foreach_pow(invstds, -2.0)
            torch._foreach_sub_(var, [e[5] for e in ent["bn"]])
            torch._foreach_clamp_min_(var, 0.0)
            torch._foreach_mul_(var, [m * (e[3] / (e[3] - 1.0) if e[3] > 1.0 else 1.0) for e, m in zip(ent["bn"], mom)])
            torch._foreach_mul_(rms, [1.0 - m for m in mom])
            torch._foreach_mul_(rvs, [1.0 - m for m in mom])
            torch._foreach_add_(rms, torch._foreach_mul(means, mom))
            torch._foreach_add_(rvs, var)
        self.passes = {}


BN_RECORDER = None


def bn_tick(bn, training):
    """num_batches_tracked += 1 like nn.BatchNorm does in training mode -- deferred when a recorder is installed."""
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        if BN_RECORDER is not None:
            BN_RECORDER.tick(bn.num_batches_tracked)
        else:
            bn.num_batches_tracked.add_(1)
    return bn


def _running(running_mean, running_var, save, count, momentum, eps, training):
    """The running buffers a fused op should update itself: none while a recorder defers the update."""
    if training and BN_RECORDER is not None and running_mean is not None:
        BN_RECORDER.add(running_mean, running_var, save, count, momentum, eps)
        return None, None
    return running_mean, running_var


def _rows(x: torch.Tensor, vec: bool = False) -> torch.Tensor:
    """[B,N,C] view whose (b,n) rows have one uniform stride and unit channel stride
    (``vec``: additionally 16-byte aligned rows for float4 access)."""
    bad = x.stride(2) != 1 or x.stride(0) != x.shape[1] * x.stride(1) or x.dtype != torch.float32
    if vec and not bad:
        bad = x.stride(1) % 4 != 0 or x.data_ptr() % 16 != 0
    if bad:
        x = x.float().contiguous()
    return x


# ------------------------------------------------------------------------------------------------
# kNN
# ------------------------------------------------------------------------------------------------
def knn_cm(x: torch.Tensor, k: int) -> torch.Tensor:
    """x [B,C,N] (reference layout, model_utils.py:178) -> int32 [B,N,k], nearest first."""
    _need_cuda(x)
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    B, C, N = x.shape
    if C >= 16:  # feature inputs take the tensor-core path, which wants one row per point
        return knn_pm(x.transpose(1, 2).contiguous(), k)
    if x.stride(2) != 1:
        x = x.contiguous()
    idx = torch.empty(B, N, k, dtype=torch.int32, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.sug_knn_f32(_ptr(x), B, C, N, k, x.stride(0), x.stride(2), x.stride(1), _ptr(idx), None, 0,
                                   _stream()), "sug_knn_f32")
    return idx


def knn_pm(x: torch.Tensor, k: int) -> torch.Tensor:
    """x [B,N,C] point-major -> int32 [B,N,k]."""
    _need_cuda(x)
    x = x.detach()
    if x.dtype != torch.float32:
        x = x.float()
    B, N, C = x.shape
    if x.stride(2) != 1 or x.stride(0) != N * x.stride(1):
        x = x.contiguous()
    idx = torch.empty(B, N, k, dtype=torch.int32, device=x.device)
    lib = _lib.load()
    ws = _workspace(lib.sug_knn_ws_bytes(B, C, N, k), x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.sug_knn_f32(_ptr(x), B, C, N, k, x.stride(0), x.stride(1), 1, _ptr(idx), _ptr(ws), ws.numel(),
                                   _stream()), "sug_knn_f32")
    return idx


def knn_reverse(idx: torch.Tensor):
    B, N, k = idx.shape
    rev_ptr = torch.empty(B, N + 1, dtype=torch.int32, device=idx.device)
    rev_edge = torch.empty(B, N * k, dtype=torch.int32, device=idx.device)
    lib = _lib.load()
    with torch.cuda.device(idx.device):
        _lib.check(lib.sug_knn_reverse(_ptr(idx), B, N, k, _ptr(rev_ptr), _ptr(rev_edge), _stream()),
                   "sug_knn_reverse")
    return rev_ptr, rev_edge


# ------------------------------------------------------------------------------------------------
# EdgeConv
# ------------------------------------------------------------------------------------------------
class _EdgeConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx, weight, gamma, beta, running_mean, running_var, training, eps, momentum, slope, slot):
        x = _rows(x)
        B, N, C = x.shape
        Cout = weight.shape[0]
        k = idx.shape[2]
        dev = x.device
        w2 = weight.detach().reshape(Cout, 2 * C).contiguous()
        P = B * N
        out = _take_slot(slot, (B, N, Cout), dev)
        ab = torch.empty(P, 2 * Cout, dtype=torch.float32, device=dev)
        lib = _lib.load()
        ws = _workspace(lib.sug_edgeconv_ws_bytes(B, N, C, Cout, k), dev)
        if training:
            ext = torch.empty(P, Cout, dtype=torch.float32, device=dev)
            ssum = torch.empty(P, Cout, dtype=torch.float32, device=dev)
            arg = torch.empty(P, Cout, dtype=torch.uint8, device=dev)
            save = torch.empty(2 * Cout, dtype=torch.float32, device=dev)
        else:
            ext = ssum = arg = save = None
        running_mean, running_var = _running(running_mean, running_var, save, P * k, momentum, eps, training)
        with torch.cuda.device(dev):
            _lib.check(lib.sug_edgeconv_fwd(_ptr(x), x.stride(1), _ptr(idx), _ptr(w2), _ptr(gamma.detach()),
                                            _ptr(beta.detach()), _ptr(running_mean), _ptr(running_var), B, N, C, Cout,
                                            k, eps, momentum, slope, int(training), _ptr(out), out.stride(1), _ptr(ab),
                                            _ptr(ext), _ptr(arg), _ptr(ssum), _ptr(save), _ptr(ws), ws.numel(),
                                            _stream()), "sug_edgeconv_fwd")
        if training:
            ctx.save_for_backward(x, idx, w2, gamma, beta, ab, ext, arg, ssum, save)
            ctx.slope = slope
            ctx.wshape = weight.shape
        return out

    @staticmethod
    def backward(ctx, gout):
        x, idx, w2, gamma, beta, ab, ext, arg, ssum, save = ctx.saved_tensors
        B, N, C = x.shape
        Cout = w2.shape[0]
        k = idx.shape[2]
        dev = x.device
        gout = _rows(gout, vec=True)
        lib = _lib.load()
        need_dx = ctx.needs_input_grad[0]
        dx = torch.empty(B, N, C, dtype=torch.float32, device=dev) if need_dx else None
        dw = torch.empty(Cout, 2 * C, dtype=torch.float32, device=dev)
        dgamma = torch.empty(Cout, dtype=torch.float32, device=dev)
        dbeta = torch.empty(Cout, dtype=torch.float32, device=dev)
        dab = torch.empty(B * N, 2 * Cout, dtype=torch.float32, device=dev)
        ws = _workspace(lib.sug_edgeconv_ws_bytes(B, N, C, Cout, k), dev)
        with torch.cuda.device(dev):
            rev_ptr, rev_edge = knn_reverse(idx)
            _lib.check(lib.sug_edgeconv_bwd(_ptr(gout), gout.stride(1), _ptr(x), x.stride(1), _ptr(idx),
                                            _ptr(rev_ptr), _ptr(rev_edge), _ptr(w2), _ptr(gamma.detach()),
                                            _ptr(beta.detach()), _ptr(ab), _ptr(ext), _ptr(arg), _ptr(ssum),
                                            _ptr(save), B, N, C, Cout, k, ctx.slope, _ptr(dx), C, 0, _ptr(dw),
                                            _ptr(dgamma), _ptr(dbeta), _ptr(dab), _ptr(ws), ws.numel(), _stream()),
                       "sug_edgeconv_bwd")
        return dx, None, dw.view(ctx.wshape), dgamma, dbeta, None, None, None, None, None, None, None


def edgeconv(x, idx, weight, gamma, beta, running_mean, running_var, training: bool, eps: float = 1e-5,
             momentum: float = 0.1, slope: float = 0.01, out=None):
    """Fused get_graph_feature -> 1x1 conv -> BatchNorm2d -> LeakyReLU -> max over k.
    x [B,N,C] point-major, idx int32 [B,N,k], weight [Cout,2C,1,1] -> [B,N,Cout].
    ``out``: optional [B,N,Cout] view (unit channel stride) of a wider buffer to write into, see ``join_slices``."""
    _need_cuda(x, idx, weight)
    return _EdgeConvFn.apply(x, idx, weight, gamma, beta, running_mean, running_var, bool(training), float(eps),
                             float(momentum), float(slope), None if out is None else [out])


def _take_slot(slot, shape, dev):
    """Output tensor of an op: a fresh one, or the caller's slice of a wider point-major buffer.  The slice is
    handed over inside a list so that autograd does not see it as an input (the op's output is then an ordinary
    output that happens to live in that buffer).  The kernels write through raw pointers, so no autograd
    version counter moves; a later torch in-place op on the buffer or on another slice would (correctly) make
    autograd refuse the slices that were already returned -- the buffer must only be read afterwards."""
    if slot is None:
        return torch.empty(*shape, dtype=torch.float32, device=dev)
    out = slot[0]
    if (tuple(out.shape) != tuple(shape) or out.dtype != torch.float32 or out.stride(-1) != 1 or out.stride(-2) % 4 != 0
            or out.data_ptr() % 16 != 0 or (out.dim() == 3 and out.stride(0) != out.shape[1] * out.stride(1))):
        raise RuntimeError("output slot must be a float32 [B,N,C] slice with 16 B aligned rows of one uniform stride")
    return out


class _JoinSlicesFn(torch.autograd.Function):
    """The concatenation of tensors that were WRITTEN into channel slices of one buffer: returns the buffer,
    the backward hands every producer the matching slice of the gradient (views, no copies)."""

    @staticmethod
    def forward(ctx, holder, *parts):
        ctx.widths = [p.shape[-1] for p in parts]
        return holder[0].view(holder[0].shape)

    @staticmethod
    def backward(ctx, g):
        outs, off = [], 0
        for w in ctx.widths:
            outs.append(g[..., off:off + w])
            off += w
        return (None, *outs)


def join_slices(buf, *parts):
    """``torch.cat(parts, dim=-1)`` for parts produced with ``out=buf[..., a:b]`` (consecutive slices of buf)."""
    off = 0
    for p in parts:
        if p.data_ptr() != buf.data_ptr() + 4 * off or p.shape[:-1] != buf.shape[:-1]:
            raise RuntimeError("join_slices: parts must be the consecutive channel slices of the buffer")
        off += p.shape[-1]
    if off != buf.shape[-1]:
        raise RuntimeError("join_slices: the parts do not cover the buffer")
    return _JoinSlicesFn.apply([buf], *parts)


# ------------------------------------------------------------------------------------------------
# shared MLP + BN + act + global pool
# ------------------------------------------------------------------------------------------------
POOL_MAX, POOL_MAX_AVG = 0, 1


class _MlpPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, training, eps, momentum, slope, pool):
        x = _rows(x)
        B, N, Cin = x.shape
        Cout = weight.shape[0]
        dev = x.device
        w2 = weight.detach().reshape(Cout, Cin).contiguous()
        y = torch.empty(B * N, Cout, dtype=torch.float32, device=dev)
        out = torch.empty(B, Cout * (2 if pool == POOL_MAX_AVG else 1), dtype=torch.float32, device=dev)
        lib = _lib.load()
        ws = _workspace(lib.sug_mlp_pool_ws_bytes(B, N, Cin, Cout), dev)
        argext = torch.empty(B, Cout, dtype=torch.int32, device=dev) if training else None
        save = torch.empty(2 * Cout, dtype=torch.float32, device=dev) if training else None
        running_mean, running_var = _running(running_mean, running_var, save, B * N, momentum, eps, training)
        with torch.cuda.device(dev):
            _lib.check(lib.sug_mlp_pool_fwd(_ptr(x), x.stride(1), _ptr(w2), _ptr(None if bias is None else bias.detach()),
                                            _ptr(gamma.detach()), _ptr(beta.detach()), _ptr(running_mean),
                                            _ptr(running_var), B, N, Cin, Cout, eps, momentum, slope, pool,
                                            int(training), _ptr(y), _ptr(out), _ptr(argext), _ptr(save), _ptr(ws),
                                            ws.numel(), _stream()), "sug_mlp_pool_fwd")
        if training:
            ctx.save_for_backward(x, w2, gamma, beta, y, argext, save)
            ctx.meta = (slope, pool, weight.shape, bias is not None)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, w2, gamma, beta, y, argext, save = ctx.saved_tensors
        slope, pool, wshape, has_bias = ctx.meta
        B, N, Cin = x.shape
        Cout = w2.shape[0]
        dev = x.device
        gout = gout.contiguous()
        lib = _lib.load()
        need_dx = ctx.needs_input_grad[0]
        dx = torch.empty(B, N, Cin, dtype=torch.float32, device=dev) if need_dx else None
        dw = torch.empty(Cout, Cin, dtype=torch.float32, device=dev)
        dbias = torch.empty(Cout, dtype=torch.float32, device=dev) if has_bias else None
        dgamma = torch.empty(Cout, dtype=torch.float32, device=dev)
        dbeta = torch.empty(Cout, dtype=torch.float32, device=dev)
        ws = _workspace(lib.sug_mlp_pool_ws_bytes(B, N, Cin, Cout), dev)
        with torch.cuda.device(dev):
            # y is consumed (overwritten with dL/dy): the tape is single-use, like any freed buffer
            _lib.check(lib.sug_mlp_pool_bwd(_ptr(gout), _ptr(x), x.stride(1), _ptr(w2), None, _ptr(gamma.detach()),
                                            _ptr(beta.detach()), _ptr(y), _ptr(argext), _ptr(save), B, N, Cin, Cout,
                                            slope, pool, _ptr(dx), Cin, 0, _ptr(dw), _ptr(dbias), _ptr(dgamma),
                                            _ptr(dbeta), _ptr(ws), ws.numel(), _stream()), "sug_mlp_pool_bwd")
        return dx, dw.view(wshape), dbias, dgamma, dbeta, None, None, None, None, None, None, None


def mlp_bn_act_pool(x, weight, bias, gamma, beta, running_mean, running_var, training: bool, slope: float,
                    pool: int, eps: float = 1e-5, momentum: float = 0.1):
    """x [B,N,Cin] -> [B,Cout] (max) or [B,2*Cout] (max || avg) of act(BN(x W^T + bias))."""
    _need_cuda(x, weight)
    return _MlpPoolFn.apply(x, weight, bias, gamma, beta, running_mean, running_var, bool(training), float(eps),
                            float(momentum), float(slope), int(pool))


# ------------------------------------------------------------------------------------------------
# per-point linear layers
# ------------------------------------------------------------------------------------------------
class _LinearBnActFn(torch.autograd.Function):
    """x [.., Cin] -> act(BN(x W^T + b)) [.., Cout], BatchNorm statistics over all leading rows."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, running_mean, running_var, training, eps, momentum, slope):
        lead = x.shape[:-1]
        Cin = x.shape[-1]
        x2 = x.reshape(-1, Cin)
        if x2.stride(1) != 1 or x2.dtype != torch.float32:
            x2 = x2.float().contiguous()
        P = x2.shape[0]
        Cout = weight.shape[0]
        dev = x.device
        w2 = weight.detach().reshape(Cout, Cin).contiguous()
        y = torch.empty(P, Cout, dtype=torch.float32, device=dev)
        out = torch.empty(P, Cout, dtype=torch.float32, device=dev)
        save = torch.empty(2 * Cout, dtype=torch.float32, device=dev) if training else None
        lib = _lib.load()
        ws = _workspace(32 * Cout + 4096, dev)
        running_mean, running_var = _running(running_mean, running_var, save, P, momentum, eps, training)
        with torch.cuda.device(dev):
            _lib.check(lib.sug_linear_bn_act_fwd(_ptr(x2), x2.stride(0), _ptr(w2), _ptr(None if bias is None else bias.detach()),
                                                 _ptr(gamma.detach()), _ptr(beta.detach()), _ptr(running_mean),
                                                 _ptr(running_var), P, Cin, Cout, eps, momentum, slope, int(training),
                                                 _ptr(y), _ptr(out), Cout, _ptr(save), _ptr(ws), ws.numel(), _stream()),
                       "sug_linear_bn_act_fwd")
        if training:
            ctx.save_for_backward(x2, w2, gamma, beta, y, save)
            ctx.meta = (slope, weight.shape, bias is not None, tuple(x.shape))
        return out.view(*lead, Cout)

    @staticmethod
    def backward(ctx, gout):
        x2, w2, gamma, beta, y, save = ctx.saved_tensors
        slope, wshape, has_bias, xshape = ctx.meta
        P, Cin = x2.shape
        Cout = w2.shape[0]
        dev = x2.device
        g2 = gout.reshape(P, Cout)
        if g2.stride(1) != 1 or g2.stride(0) % 4 != 0 or g2.data_ptr() % 16 != 0 or g2.dtype != torch.float32:
            g2 = g2.float().contiguous()
        need_dx = ctx.needs_input_grad[0]
        dx = torch.empty(P, Cin, dtype=torch.float32, device=dev) if need_dx else None
        dw = torch.empty(Cout, Cin, dtype=torch.float32, device=dev)
        dbias = torch.empty(Cout, dtype=torch.float32, device=dev) if has_bias else None
        dgamma = torch.empty(Cout, dtype=torch.float32, device=dev)
        dbeta = torch.empty(Cout, dtype=torch.float32, device=dev)
        lib = _lib.load()
        ws = _workspace(32 * Cout + 4096, dev)
        with torch.cuda.device(dev):
            _lib.check(lib.sug_linear_bn_act_bwd(_ptr(g2), g2.stride(0), _ptr(x2), x2.stride(0), _ptr(w2),
                                                 _ptr(gamma.detach()), _ptr(beta.detach()), _ptr(y), _ptr(save), P, Cin,
                                                 Cout, slope, _ptr(dx), Cin, _ptr(dw), _ptr(dbias), _ptr(dgamma),
                                                 _ptr(dbeta), _ptr(ws), ws.numel(), _stream()), "sug_linear_bn_act_bwd")
        return (None if dx is None else dx.view(xshape)), dw.view(wshape), dbias, dgamma, dbeta, None, None, None, None, None, None


def linear_bn_act(x, weight, bias, gamma, beta, running_mean, running_var, training: bool, slope: float,
                  eps: float = 1e-5, momentum: float = 0.1):
    """1x1 conv over points + BatchNorm + ReLU/LeakyReLU on point-major rows (model_utils.py:8-32)."""
    _need_cuda(x, weight)
    return _LinearBnActFn.apply(x, weight, bias, gamma, beta, running_mean, running_var, bool(training), float(eps),
                                float(momentum), float(slope))


def _gemm_auto(a, b, bias=None, out=None):
    """a [M,K] x b[N,K]^T (+bias) through the library's dispatcher (tensor cores when possible)."""
    M, K = a.shape
    N = b.shape[0]
    c = torch.empty(M, N, dtype=torch.float32, device=a.device) if out is None else out
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.sug_gemm_auto_f32(_ptr(a), a.stride(0), a.stride(1), _ptr(b), b.stride(0), b.stride(1), _ptr(bias),
                                         _ptr(c), c.stride(0), M, N, K, _stream()), "sug_gemm_auto_f32")
    return c


class _LinearFn(torch.autograd.Function):
    """x [.., K] -> x W^T + b  (fp32-accurate tensor-core GEMM)."""

    @staticmethod
    def forward(ctx, x, weight, bias, slot):
        K = x.shape[-1]
        x2 = x.reshape(-1, K)
        if x2.stride(1) != 1 or x2.dtype != torch.float32:
            x2 = x2.float().contiguous()
        w2 = weight.detach().reshape(weight.shape[0], K).contiguous()
        ctx.save_for_backward(x2, w2)
        ctx.meta = (tuple(x.shape), weight.shape, bias is not None)
        if slot is not None:  # [B,N,Cout] slice of a wider buffer: rows of one uniform stride
            out = _take_slot(slot, (*x.shape[:-1], w2.shape[0]), x.device)
            rows = out.as_strided((x2.shape[0], w2.shape[0]), (out.stride(-2), 1))
            _gemm_auto(x2, w2, None if bias is None else bias.detach(), out=rows)
            return out
        y = _gemm_auto(x2, w2, None if bias is None else bias.detach())
        return y.view(*x.shape[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, g):
        x2, w2 = ctx.saved_tensors
        xshape, wshape, has_bias = ctx.meta
        g2 = g.reshape(-1, w2.shape[0])
        if g2.stride(1) != 1 or g2.dtype != torch.float32:
            g2 = g2.float().contiguous()
        dx = _gemm_auto(g2, w2.t()).view(xshape) if ctx.needs_input_grad[0] else None
        dw = _gemm_auto(g2.t(), x2.t()).view(wshape) if ctx.needs_input_grad[1] else None
        db = g2.sum(0) if has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db, None


def linear(x, weight, bias=None, out=None):
    _need_cuda(x, weight)
    return _LinearFn.apply(x, weight, bias, None if out is None else [out])


# ------------------------------------------------------------------------------------------------
# focal loss
# ------------------------------------------------------------------------------------------------
class _FocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, labels, alpha_row, gamma, mean):
        preds = preds.float().contiguous()
        R, C = preds.shape
        loss = torch.empty((), dtype=torch.float32, device=preds.device)
        lib = _lib.load()
        with torch.cuda.device(preds.device):
            _lib.check(lib.sug_focal_loss_fwd(_ptr(preds), _ptr(labels), _ptr(alpha_row), R, C, float(gamma), int(mean),
                                              _ptr(loss), _stream()), "sug_focal_loss_fwd")
        ctx.save_for_backward(preds, labels, alpha_row)
        ctx.meta = (float(gamma), int(mean))
        return loss

    @staticmethod
    def backward(ctx, g):
        preds, labels, alpha_row = ctx.saved_tensors
        gamma, mean = ctx.meta
        R, C = preds.shape
        dp = torch.empty_like(preds)
        lib = _lib.load()
        with torch.cuda.device(preds.device):
            _lib.check(lib.sug_focal_loss_bwd(_ptr(g.float().contiguous()), _ptr(preds), _ptr(labels), _ptr(alpha_row), R, C,
                                              gamma, mean, _ptr(dp), _stream()), "sug_focal_loss_bwd")
        return dp, None, None, None, None


def focal_loss(preds, labels, alpha_row, gamma: float, mean: bool):
    """model_utils.py:164-176 in one launch: preds [R,C], labels int64 [R], alpha_row [R] -> scalar."""
    _need_cuda(preds, labels, alpha_row)
    return _FocalFn.apply(preds, labels.long().contiguous(), alpha_row.float().contiguous(), gamma, mean)


# ------------------------------------------------------------------------------------------------
# MMD + Chamfer
# ------------------------------------------------------------------------------------------------
def _mmd_forward(ctx, z, m, weights, sigmas, biased):
    """z [2m, D] (rows of X then Y) -> biased / unbiased mixture-RBF MMD^2; saves z and dL/dG for the backward."""
    D = z.shape[1]
    dev = z.device
    loss = torch.empty((), dtype=torch.float32, device=dev)
    coef = torch.empty(2 * m, 2 * m, dtype=torch.float32, device=dev)
    lib = _lib.load()
    ws = _workspace(lib.sug_mmd_ws_bytes(m, D), dev)
    sg = (ctypes.c_float * len(sigmas))(*[float(s) for s in sigmas])
    w = None
    if weights is not None:
        w = weights.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if w.numel() != m:
            raise RuntimeError(f"sample_weights has {w.numel()} entries, expected {m}")
    with torch.cuda.device(dev):
        _lib.check(lib.sug_mmd_rbf_fwd(_ptr(z), D, m, D, ctypes.cast(sg, ctypes.c_void_p), len(sigmas), _ptr(w),
                                       int(biased), _ptr(loss), _ptr(coef), _ptr(ws), ws.numel(), _stream()),
                   "sug_mmd_rbf_fwd")
    ctx.save_for_backward(z, coef)
    ctx.m = m
    return loss


def _mmd_backward(ctx, gloss):
    z, coef = ctx.saved_tensors
    m = ctx.m
    D = z.shape[1]
    dz = torch.empty_like(z)
    g = gloss.detach().float().contiguous()
    lib = _lib.load()
    with torch.cuda.device(z.device):
        _lib.check(lib.sug_mmd_rbf_bwd(_ptr(z), D, m, D, _ptr(coef), _ptr(g), _ptr(dz), D, _stream()),
                   "sug_mmd_rbf_bwd")
    return dz


class _MmdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, Y, weights, sigmas, biased):
        z = torch.cat((X.detach(), Y.detach()), 0).float().contiguous()
        return _mmd_forward(ctx, z, X.shape[0], weights, sigmas, biased)

    @staticmethod
    def backward(ctx, gloss):
        dz = _mmd_backward(ctx, gloss)
        return dz[:ctx.m], dz[ctx.m:], None, None, None


class _SoftMmdFn(torch.autograd.Function):
    """soft_mmd (mmd.py:56-66): the operand [feat | onehot(label) * scale] is assembled by one kernel."""

    @staticmethod
    def forward(ctx, feat_s, feat_t, label_s, label_t, scale, weights, sigmas):
        fs, ft = feat_s.detach().float(), feat_t.detach().float()
        if fs.stride(1) != 1:
            fs = fs.contiguous()
        if ft.stride(1) != 1:
            ft = ft.contiguous()
        m, D = fs.shape
        z = torch.empty(2 * m, D + 10, dtype=torch.float32, device=fs.device)
        lib = _lib.load()
        with torch.cuda.device(fs.device):
            _lib.check(lib.sug_soft_mmd_assemble(_ptr(fs), fs.stride(0), _ptr(ft), ft.stride(0), _ptr(label_s), _ptr(label_t),
                                                 m, D, 10, float(scale), _ptr(z), _stream()), "sug_soft_mmd_assemble")
        ctx.D = D
        return _mmd_forward(ctx, z, m, weights, sigmas, True)

    @staticmethod
    def backward(ctx, gloss):
        dz = _mmd_backward(ctx, gloss)
        return dz[:ctx.m, :ctx.D], dz[ctx.m:, :ctx.D], None, None, None, None, None


def soft_mmd(feat_s, feat_t, label_s, label_t, scale: float, sigma_list: Sequence[float], sample_weights=None):
    _need_cuda(feat_s, feat_t, label_s, label_t)
    if feat_s.shape != feat_t.shape:
        raise AssertionError("X and Y must have the same shape")  # mmd.py:240
    return _SoftMmdFn.apply(feat_s, feat_t, label_s.long().contiguous(), label_t.long().contiguous(), float(scale),
                            sample_weights, tuple(sigma_list))


def sda_sem_weights(pred_s, pred_t, label_s, label_t, label_weight: float):
    """prob_weights_soft with weighting 'mean2one' (mmd.py:134-148) in one launch -> [m]."""
    _need_cuda(pred_s, pred_t, label_s, label_t)
    ps = pred_s.detach().float().reshape(-1, 10).contiguous()
    pt = pred_t.detach().float().reshape(-1, 10).contiguous()
    m = ps.shape[0]
    w = torch.empty(m, dtype=torch.float32, device=ps.device)
    lib = _lib.load()
    with torch.cuda.device(ps.device):
        _lib.check(lib.sug_sda_sem_weights(_ptr(ps), _ptr(pt), _ptr(label_s.long().contiguous()), _ptr(label_t.long().contiguous()),
                                           m, 10, float(label_weight), _ptr(w), _stream()), "sug_sda_sem_weights")
    return w


def mix_rbf_mmd2(X, Y, sigma_list: Sequence[float], biased: bool = True, sample_weights=None):
    _need_cuda(X, Y)
    if X.shape != Y.shape:
        raise AssertionError("X and Y must have the same shape")  # mmd.py:240
    return _MmdFn.apply(X, Y, sample_weights, tuple(sigma_list), bool(biased))


def chamfer(p1: torch.Tensor, p2: torch.Tensor):
    """p1 [B,N,3], p2 [B,M,3] -> (dist1 [B,N], dist2 [B,M]) squared NN distances."""
    _need_cuda(p1, p2)
    p1 = p1.detach().float().contiguous()
    p2 = p2.detach().float().contiguous()
    B, N, _ = p1.shape
    M = p2.shape[1]
    d1 = torch.empty(B, N, dtype=torch.float32, device=p1.device)
    d2 = torch.empty(B, M, dtype=torch.float32, device=p1.device)
    lib = _lib.load()
    with torch.cuda.device(p1.device):
        _lib.check(lib.sug_chamfer_f32(_ptr(p1), _ptr(p2), B, N, M, _ptr(d1), _ptr(d2), _stream()), "sug_chamfer_f32")
    return d1, d2


# ------------------------------------------------------------------------------------------------
# index builders of the self-adaptive node layer (all take the reference's [B,3,N] layout)
# ------------------------------------------------------------------------------------------------
def _xyz(t):
    _need_cuda(t)
    return t.detach().float().contiguous()


def fps(xyz: torch.Tensor, npoint: int, start: torch.Tensor) -> torch.Tensor:
    xyz = _xyz(xyz)
    B, _, N = xyz.shape
    st = start.to(device=xyz.device, dtype=torch.int32).contiguous()
    out = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
    lib = _lib.load()
    with torch.cuda.device(xyz.device):
        _lib.check(lib.sug_fps(_ptr(xyz), B, N, npoint, _ptr(st), _ptr(out), _stream()), "sug_fps")
    return out


def ball_query(xyz, query, radius: float, nsample: int) -> torch.Tensor:
    xyz, query = _xyz(xyz), _xyz(query)
    B, _, N = xyz.shape
    S = query.shape[2]
    out = torch.empty(B, S, nsample, dtype=torch.int32, device=xyz.device)
    lib = _lib.load()
    with torch.cuda.device(xyz.device):
        _lib.check(lib.sug_ball_query(_ptr(xyz), _ptr(query), B, N, S, float(radius), nsample, _ptr(out), _stream()),
                   "sug_ball_query")
    return out


def knn_query(xyz, query, nsample: int, ordered: bool = True) -> torch.Tensor:
    """nsample nearest points of every query: ascending by distance (``ordered``) or as a set."""
    xyz, query = _xyz(xyz), _xyz(query)
    B, _, N = xyz.shape
    S = query.shape[2]
    out = torch.empty(B, S, nsample, dtype=torch.int32, device=xyz.device)
    lib = _lib.load()
    fn = lib.sug_knn_query if ordered else lib.sug_knn_query_set
    with torch.cuda.device(xyz.device):
        _lib.check(fn(_ptr(xyz), _ptr(query), B, N, S, nsample, _ptr(out), _stream()), "sug_knn_query")
    return out


class _GroupMaxFn(torch.autograd.Function):
    """x [B,N,C] point-major, idx int32 [B,S,K] -> max over the K gathered rows, [B,S,C]."""

    @staticmethod
    def forward(ctx, x, idx):
        x = x.float().contiguous()
        B, N, C = x.shape
        _, S, K = idx.shape
        out = torch.empty(B, S, C, dtype=torch.float32, device=x.device)
        arg = torch.empty(B, S, C, dtype=torch.int32, device=x.device)
        lib = _lib.load()
        with torch.cuda.device(x.device):
            _lib.check(lib.sug_group_max_fwd(_ptr(x), _ptr(idx), B, N, S, K, C, _ptr(out), _ptr(arg), _stream()),
                       "sug_group_max_fwd")
        ctx.save_for_backward(arg)
        ctx.shape = (B, N, S, C)
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        B, N, S, C = ctx.shape
        g = g.float().contiguous()
        dx = torch.zeros(B, N, C, dtype=torch.float32, device=g.device)
        lib = _lib.load()
        with torch.cuda.device(g.device):
            _lib.check(lib.sug_group_max_bwd(_ptr(g), _ptr(arg), B, N, S, C, _ptr(dx), _stream()), "sug_group_max_bwd")
        return dx, None


def group_max(x, idx):
    _need_cuda(x, idx)
    return _GroupMaxFn.apply(x, idx.int().contiguous())


class _InterpFn(torch.autograd.Function):
    """f [B,S,C], idx int32 [B,N,K], w [B,N,K] -> sum_k w f[idx]  [B,N,C]."""

    @staticmethod
    def forward(ctx, f, idx, w):
        f = f.float().contiguous()
        w = w.float().contiguous()
        B, S, C = f.shape
        _, N, K = idx.shape
        out = torch.empty(B, N, C, dtype=torch.float32, device=f.device)
        lib = _lib.load()
        with torch.cuda.device(f.device):
            _lib.check(lib.sug_interp_fwd(_ptr(f), _ptr(idx), _ptr(w), B, N, S, K, C, _ptr(out), _stream()),
                       "sug_interp_fwd")
        ctx.save_for_backward(f, idx, w)
        return out

    @staticmethod
    def backward(ctx, g):
        f, idx, w = ctx.saved_tensors
        B, S, C = f.shape
        _, N, K = idx.shape
        g = g.float().contiguous()
        df = torch.zeros_like(f)
        dw = torch.zeros_like(w)
        lib = _lib.load()
        with torch.cuda.device(g.device):
            _lib.check(lib.sug_interp_bwd(_ptr(g), _ptr(f), _ptr(idx), _ptr(w), B, N, S, K, C, _ptr(df), _ptr(dw),
                                          _stream()), "sug_interp_bwd")
        return df, None, dw


def interpolate(f, idx, w):
    _need_cuda(f, idx, w)
    return _InterpFn.apply(f, idx.int().contiguous(), w)


class _NodeOffsetFn(torch.autograd.Function):
    """h [B,N,3], xyz [B,3,N], fidx int32 [B,S], gidx int32 [B,S,G] -> mean_j tanh(h[g_j]-h[f]) * (xyz[g_j]-xyz[f])."""

    @staticmethod
    def forward(ctx, h, xyz, fidx, gidx):
        h = h.float().contiguous()
        B, N, _ = h.shape
        S, G = gidx.shape[1], gidx.shape[2]
        out = torch.empty(B, S, 3, dtype=torch.float32, device=h.device)
        lib = _lib.load()
        with torch.cuda.device(h.device):
            _lib.check(lib.sug_node_offset_fwd(_ptr(h), _ptr(xyz), _ptr(fidx), _ptr(gidx), B, N, S, G, _ptr(out), _stream()),
                       "sug_node_offset_fwd")
        ctx.save_for_backward(h, xyz, fidx, gidx)
        return out

    @staticmethod
    def backward(ctx, g):
        h, xyz, fidx, gidx = ctx.saved_tensors
        B, N, _ = h.shape
        S, G = gidx.shape[1], gidx.shape[2]
        dh = torch.zeros_like(h)
        lib = _lib.load()
        with torch.cuda.device(h.device):
            _lib.check(lib.sug_node_offset_bwd(_ptr(g.float().contiguous()), _ptr(h), _ptr(xyz), _ptr(fidx), _ptr(gidx), B, N, S, G,
                                               _ptr(dh), _stream()), "sug_node_offset_bwd")
        return dh, None, None, None


def node_offset(h, xyz, fidx, gidx):
    """model_utils.py:107-117 after pred_offset has been applied per point (it is linear and bias-free)."""
    _need_cuda(h, xyz, fidx, gidx)
    return _NodeOffsetFn.apply(h, _xyz(xyz), fidx.int().contiguous(), gidx.int().contiguous())


class _InterpWeightFn(torch.autograd.Function):
    """xyz [B,3,N], nodes [B,S,3], idx int32 [B,N,K] -> normalised inverse squared distances [B,N,K]."""

    @staticmethod
    def forward(ctx, xyz, nodes, idx):
        nodes = nodes.float().contiguous()
        B, _, N = xyz.shape
        S, K = nodes.shape[1], idx.shape[2]
        w = torch.empty(B, N, K, dtype=torch.float32, device=nodes.device)
        lib = _lib.load()
        with torch.cuda.device(nodes.device):
            _lib.check(lib.sug_interp_weight_fwd(_ptr(xyz), _ptr(nodes), _ptr(idx), B, N, S, K, _ptr(w), _stream()),
                       "sug_interp_weight_fwd")
        ctx.save_for_backward(xyz, nodes, idx)
        return w

    @staticmethod
    def backward(ctx, g):
        xyz, nodes, idx = ctx.saved_tensors
        B, _, N = xyz.shape
        S, K = nodes.shape[1], idx.shape[2]
        dn = torch.zeros_like(nodes)
        lib = _lib.load()
        with torch.cuda.device(nodes.device):
            _lib.check(lib.sug_interp_weight_bwd(_ptr(g.float().contiguous()), _ptr(xyz), _ptr(nodes), _ptr(idx), B, N, S, K,
                                                 _ptr(dn), _stream()), "sug_interp_weight_bwd")
        return None, dn, None


def interp_weights(xyz, nodes, idx):
    """point_utils.py:141-160: weights of the k-NN inverse-squared-distance interpolation; gradients flow to ``nodes``."""
    _need_cuda(xyz, nodes, idx)
    return _InterpWeightFn.apply(_xyz(xyz), nodes, idx.int().contiguous())


def three_nn(xyz, nodes, k: int = 3) -> torch.Tensor:
    xyz, nodes = _xyz(xyz), _xyz(nodes)
    B, _, N = xyz.shape
    M = nodes.shape[2]
    out = torch.empty(B, N, k, dtype=torch.int32, device=xyz.device)
    lib = _lib.load()
    with torch.cuda.device(xyz.device):
        _lib.check(lib.sug_three_nn(_ptr(xyz), _ptr(nodes), B, N, M, k, _ptr(out), _stream()), "sug_three_nn")
    return out


def gemm(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a [M,K] (any strides) x b [N,K]^T -> [M,N]; the library's fp32 GEMM, exposed for tests."""
    _need_cuda(a, b)
    M, K = a.shape
    N = b.shape[0]
    c = torch.empty(M, N, dtype=torch.float32, device=a.device)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.sug_gemm_f32(_ptr(a), a.stride(0), a.stride(1), _ptr(b), b.stride(0), b.stride(1), _ptr(bias),
                                    _ptr(c), N, M, N, K, 0, _stream()), "sug_gemm_f32")
    return c


def gemm_tc(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C = a @ b.T on the tcgen05 3xTF32 path.  a [M,K] / b [N,K]: either row-major (K-major) or the
    transpose of a row-major [K,M] / [K,N] tensor (MN-major), detected from the strides."""
    _need_cuda(a, b)
    M, K = a.shape
    N = b.shape[0]

    def major(t):
        if t.stride(1) == 1:
            return 0, t.stride(0)
        if t.stride(0) == 1:
            return 1, t.stride(1)
        raise RuntimeError("gemm_tc operands need a unit stride")
    a_mn, lda = major(a)
    b_mn, ldb = major(b)
    c = torch.empty(M, N, dtype=torch.float32, device=a.device)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.sug_gemm_tc_f32(_ptr(a), lda, a_mn, _ptr(b), ldb, b_mn, _ptr(bias), _ptr(c), N, M, N, K,
                                       _stream()), "sug_gemm_tc_f32")
    return c
