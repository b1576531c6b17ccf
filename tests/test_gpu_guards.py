"""Out-of-bounds WRITE check of the hand-written kernels (compute-sanitizer is closed on this GPU pool -- see
profiles/r02_sanitizer.md -- so this is the substitute the pool recommends: bounds checks of our own on small and ragged
cases).  Every output / scratch buffer handed to the C ABI is carved out of a larger allocation whose 64 KB guard bands
before and after it hold a byte pattern; after the call (tail tiles, ragged N, odd strides) the bands must be intact.
Results are compared with the oracle elsewhere (test_gpu_parity.py); here only memory safety is asserted."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GUARD = 64 * 1024
PAT = 0xA5


class Guarded:
    """A tensor of `shape` / `dtype` living between two guard bands of one larger uint8 allocation."""

    def __init__(self, shape, dtype=torch.float32):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        n_al = (n + 255) // 256 * 256
        self.raw = torch.full((2 * GUARD + n_al,), PAT, dtype=torch.uint8, device=DEV)
        self.n = n
        self.t = self.raw[GUARD:GUARD + n].view(dtype).view(*shape)

    def ptr(self):
        return ctypes.c_void_p(self.t.data_ptr())

    def check(self, what):
        torch.cuda.synchronize()
        lo = self.raw[:GUARD]
        hi = self.raw[GUARD + (self.n + 255) // 256 * 256:]
        assert bool((lo == PAT).all()), f"{what}: write BEFORE the buffer"
        assert bool((hi == PAT).all()), f"{what}: write AFTER the buffer"


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from sug_b200 import _lib
    return _lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("B,C,N,k", [(3, 3, 1000, 20), (2, 3, 130, 20), (3, 64, 1000, 20), (2, 128, 200, 40), (2, 20, 77, 20),
                                     (1, 7, 300, 13)])
def test_knn_writes_stay_in_bounds(lib, B, C, N, k):
    L = lib.load()
    x = torch.randn(B, N, C, device=DEV)
    idx = Guarded((B, N, k), torch.int32)
    nws = int(L.sug_knn_ws_bytes(B, C, N, k))
    ws = Guarded((max(nws, 256),), torch.uint8)
    lib.check(L.sug_knn_f32(_p(x), B, C, N, k, N * C, C, 1, idx.ptr(), ws.ptr(), nws, _stream()), "knn")
    idx.check("knn idx")
    ws.check("knn workspace")
    assert int(idx.t.min()) >= 0 and int(idx.t.max()) < N
    rp, re = Guarded((B, N + 1), torch.int32), Guarded((B, N * k), torch.int32)
    lib.check(L.sug_knn_reverse(idx.ptr(), B, N, k, rp.ptr(), re.ptr(), _stream()), "knn_reverse")
    rp.check("rev_ptr")
    re.check("rev_edge")


@pytest.mark.parametrize("B,C,N,Co,k", [(2, 3, 1000, 64, 20), (3, 64, 900, 128, 20), (2, 128, 333, 256, 20), (2, 16, 2500, 48, 24),
                                        (1, 8, 5000, 20, 20)])
def test_edgeconv_writes_stay_in_bounds(lib, B, C, N, Co, k):
    L = lib.load()
    P = B * N
    x = torch.randn(B, N, C, device=DEV)
    idx = torch.stack([torch.stack([torch.randperm(N, device=DEV)[:k] for _ in range(N)]) for _ in range(B)]).int() \
        if N <= 1000 else torch.randint(0, N, (B, N, k), device=DEV, dtype=torch.int32)
    w = torch.randn(Co, 2 * C, device=DEV) * 0.2
    gamma = torch.randn(Co, device=DEV)
    beta = torch.randn(Co, device=DEV)
    rm, rv = Guarded((Co,)), Guarded((Co,))
    rm.t.zero_(); rv.t.fill_(1.0)
    out, ab = Guarded((P, Co)), Guarded((P, 2 * Co))
    ext, ssum, arg, save = Guarded((P, Co)), Guarded((P, Co)), Guarded((P, Co), torch.uint8), Guarded((2 * Co,))
    nws = int(L.sug_edgeconv_ws_bytes(B, N, C, Co, k))
    ws = Guarded((nws,), torch.uint8)
    for training in (1, 0):
        lib.check(L.sug_edgeconv_fwd(_p(x), C, _p(idx), _p(w), _p(gamma), _p(beta), rm.ptr(), rv.ptr(), B, N, C, Co, k, 1e-5, 0.1,
                                     0.01, training, out.ptr(), Co, ab.ptr(), ext.ptr() if training else None,
                                     arg.ptr() if training else None, ssum.ptr() if training else None,
                                     save.ptr() if training else None, ws.ptr(), nws, _stream()), "edgeconv_fwd")
        for g, n in ((out, "out"), (ab, "ab"), (ext, "ext"), (ssum, "ssum"), (arg, "arg"), (save, "save"), (ws, "ws"), (rm, "running_mean"),
                     (rv, "running_var")):
            g.check(f"edgeconv_fwd training={training} {n}")
    assert int(arg.t.max()) < k
    # backward (uses the training-mode saves: run the training forward again)
    lib.check(L.sug_edgeconv_fwd(_p(x), C, _p(idx), _p(w), _p(gamma), _p(beta), rm.ptr(), rv.ptr(), B, N, C, Co, k, 1e-5, 0.1, 0.01, 1,
                                 out.ptr(), Co, ab.ptr(), ext.ptr(), arg.ptr(), ssum.ptr(), save.ptr(), ws.ptr(), nws, _stream()), "fwd")
    rp, re = Guarded((B, N + 1), torch.int32), Guarded((B, N * k), torch.int32)
    lib.check(L.sug_knn_reverse(_p(idx), B, N, k, rp.ptr(), re.ptr(), _stream()), "knn_reverse")
    gout = torch.randn(P, Co, device=DEV)
    dx, dw, dg, db, dab = Guarded((P, C)), Guarded((Co, 2 * C)), Guarded((Co,)), Guarded((Co,)), Guarded((P, 2 * Co))
    lib.check(L.sug_edgeconv_bwd(_p(gout), Co, _p(x), C, _p(idx), rp.ptr(), re.ptr(), _p(w), _p(gamma), _p(beta), ab.ptr(), ext.ptr(),
                                 arg.ptr(), ssum.ptr(), save.ptr(), B, N, C, Co, k, 0.01, dx.ptr(), C, 0, dw.ptr(), dg.ptr(), db.ptr(),
                                 dab.ptr(), ws.ptr(), nws, _stream()), "edgeconv_bwd")
    for g, n in ((dx, "dx"), (dw, "dw"), (dg, "dgamma"), (db, "dbeta"), (dab, "dab"), (ws, "ws"), (ab, "ab"), (ext, "ext")):
        g.check(f"edgeconv_bwd {n}")
    assert bool(torch.isfinite(dx.t).all()) and bool(torch.isfinite(dw.t).all())


@pytest.mark.parametrize("B,N,Cin,Co,pool", [(3, 1000, 512, 512, 1), (2, 333, 128, 1024, 0), (2, 130, 64, 64, 0)])
def test_mlp_pool_and_gemm_writes_stay_in_bounds(lib, B, N, Cin, Co, pool):
    L = lib.load()
    P = B * N
    x = torch.randn(P, Cin, device=DEV)
    w = torch.randn(Co, Cin, device=DEV) * 0.1
    gamma, beta = torch.randn(Co, device=DEV), torch.randn(Co, device=DEV)
    rm, rv = Guarded((Co,)), Guarded((Co,))
    rm.t.zero_(); rv.t.fill_(1.0)
    y, out = Guarded((P, Co)), Guarded((B, Co * (2 if pool else 1)))
    argext, save = Guarded((B, Co), torch.int32), Guarded((2 * Co,))
    nws = int(L.sug_mlp_pool_ws_bytes(B, N, Cin, Co))
    ws = Guarded((nws,), torch.uint8)
    lib.check(L.sug_mlp_pool_fwd(_p(x), Cin, _p(w), None, _p(gamma), _p(beta), rm.ptr(), rv.ptr(), B, N, Cin, Co, 1e-5, 0.1, 0.2, pool, 1,
                                 y.ptr(), out.ptr(), argext.ptr(), save.ptr(), ws.ptr(), nws, _stream()), "mlp_pool_fwd")
    for g, n in ((y, "y"), (out, "out"), (argext, "argext"), (save, "save"), (ws, "ws"), (rm, "rm"), (rv, "rv")):
        g.check(f"mlp_pool_fwd {n}")
    gout = torch.randn(B, Co * (2 if pool else 1), device=DEV)
    dx, dw, dg, db = Guarded((P, Cin)), Guarded((Co, Cin)), Guarded((Co,)), Guarded((Co,))
    lib.check(L.sug_mlp_pool_bwd(_p(gout), _p(x), Cin, _p(w), None, _p(gamma), _p(beta), y.ptr(), argext.ptr(), save.ptr(), B, N, Cin, Co,
                                 0.2, pool, dx.ptr(), Cin, 0, dw.ptr(), None, dg.ptr(), db.ptr(), ws.ptr(), nws, _stream()), "mlp_pool_bwd")
    for g, n in ((dx, "dx"), (dw, "dw"), (dg, "dgamma"), (db, "dbeta"), (y, "y"), (ws, "ws")):
        g.check(f"mlp_pool_bwd {n}")
    # the tensor-core GEMM on ragged shapes, all four operand layouts
    for M2, N2, K2 in ((1000, 200, 72), (130, 520, 40), (77, 24, 1000)):
        a, bt = torch.randn(M2, K2, device=DEV), torch.randn(N2, K2, device=DEV)
        at, btt = a.t().contiguous(), bt.t().contiguous()   # [K,M], [K,N] row-major = MN-major operands
        for (ap, lda, amn), (bp, ldb, bmn) in (((a, K2, 0), (bt, K2, 0)), ((at, M2, 1), (bt, K2, 0)), ((a, K2, 0), (btt, N2, 1)),
                                              ((at, M2, 1), (btt, N2, 1))):
            if (amn and M2 % 4) or (bmn and N2 % 4):
                continue
            c = Guarded((M2, N2))
            lib.check(L.sug_gemm_tc_f32(_p(ap), lda, amn, _p(bp), ldb, bmn, None, c.ptr(), N2, M2, N2, K2, _stream()), "gemm_tc")
            c.check(f"gemm_tc {M2}x{N2}x{K2} a_mn={amn} b_mn={bmn}")
            ref = a.double() @ bt.double().t()
            assert float((c.t.double() - ref).norm() / ref.norm()) < 1e-5
