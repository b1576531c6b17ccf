"""Parity of the sm_100a kernels (through the C ABI) against the CPU oracle and the fixtures that
were generated from the reference.  Tolerances (BASELINE.json north_star): neighbour sets
bit-exact except near-ties with relative gap < 1e-6 (counted and reported); outputs, losses and
gradients within rel 1e-3."""
import numpy as np
import pytest
import torch

from oracle import sug_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def S():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sug_b200  # noqa: F401
    from sug_b200 import Model, mmd, model_pointnet, model_utils, ops, point_utils, step
    import types
    return types.SimpleNamespace(Model=Model, mmd=mmd, model_pointnet=model_pointnet, model_utils=model_utils,
                                 ops=ops, point_utils=point_utils, step=step)


class teacher_forced:
    """Replay the oracle's neighbour lists inside the GPU model.  kNN is validated on its own
    (test_knn_*): an fp32 near-tie that flips one neighbour is legitimate but moves the downstream
    activations by ~1e-3 (chaotic amplification through max-over-k), so the end-to-end rel-1e-3 gate
    is applied with the neighbour graph teacher-forced, as SURVEY.md §7.3 prescribes; the
    free-running deviation is printed next to it."""

    def __init__(self, S, trace):
        self.S, self.trace, self.pos = S, trace, 0

    def _next(self, x, k):
        idx = self.trace[self.pos]
        self.pos += 1
        assert idx.shape[-1] == k
        return idx.to(DEV).int().contiguous()

    def __enter__(self):
        self.saved = (self.S.ops.knn_cm, self.S.ops.knn_pm)
        self.S.ops.knn_cm = self._next
        self.S.ops.knn_pm = self._next
        return self

    def __exit__(self, *a):
        self.S.ops.knn_cm, self.S.ops.knn_pm = self.saved


def oracle_trace(fn):
    """Run ``fn`` with the oracle recording every knn() result; returns (result, trace)."""
    O.KNN_TRACE = []
    try:
        out = fn()
        return out, O.KNN_TRACE
    finally:
        O.KNN_TRACE = None


def feat_input(B, C, N, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.standard_normal((B, C, N)).astype(np.float32))


def relerr(a, b):
    a = a.detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def assert_close(a, b, tol=1e-3, what=""):
    e = relerr(a, b)
    assert e <= tol, f"{what}: relative error {e:.3e} > {tol}"


def knn_report(x_cpu, idx_ours, k):
    """Compare neighbour sets with torch.topk on the reference fp32 distance matrix.
    Returns (rows_differing, worst gap relative to |d_k|, worst gap relative to |xi|^2+|xj|^2)."""
    D = O.pairwise_neg_sqdist(x_cpu)  # [B,N,N]
    B, N, _ = D.shape
    top = D.topk(k + 1 if N > k else k, dim=-1)
    ref = top[1][..., :k]
    ours = idx_ours.cpu().long()
    assert ours.min() >= 0 and ours.max() < N
    # no duplicates in a row
    so = ours.sort(dim=-1)[0]
    assert bool((so[..., 1:] != so[..., :-1]).all()), "duplicate neighbour in a row"
    sr = ref.sort(dim=-1)[0]
    diff_rows = (so != sr).any(-1)
    n_diff = int(diff_rows.sum())
    worst_d, worst_n = 0.0, 0.0
    if n_diff:
        dk = top[0][..., k - 1]  # k-th best key of the oracle
        xx = (x_cpu ** 2).sum(1)  # [B,N]
        for b, r in diff_rows.nonzero().tolist():
            mine = set(ours[b, r].tolist())
            theirs = set(ref[b, r].tolist())
            for j in mine - theirs:
                gap = abs(float(D[b, r, j] - dk[b, r]))
                worst_d = max(worst_d, gap / max(abs(float(dk[b, r])), 1e-30))
                worst_n = max(worst_n, gap / float(xx[b, r] + xx[b, j] + 1e-30))
    # ordering: keys along our list must be non-increasing up to fp32 noise
    keys = torch.gather(D, 2, ours)
    xx = (x_cpu ** 2).sum(1)
    scale = (xx.unsqueeze(-1) + torch.gather(xx.unsqueeze(1).expand(B, N, N), 2, ours))
    assert bool(((keys[..., 1:] - keys[..., :-1]) <= 2e-6 * scale[..., 1:] + 1e-12).all()), "list not sorted nearest-first"
    return n_diff, worst_d, worst_n


# ------------------------------------------------------------------------------------------------
def test_gemm(S):
    torch.manual_seed(0)
    for (M, N, K) in [(300, 70, 3), (257, 129, 64), (1024, 128, 6), (128, 128, 4106)]:
        a = torch.randn(M, K, device=DEV)
        b = torch.randn(N, K, device=DEV)
        ref = (a.double() @ b.double().t())
        assert_close(S.ops.gemm(a, b), ref, 2e-6, f"gemm NT {M}x{N}x{K}")
    # transposed operands + split-K (weight-gradient shape)
    P, Co, C = 20000, 96, 40
    dy = torch.randn(P, Co, device=DEV)
    x = torch.randn(P, C, device=DEV)
    ref = dy.double().t() @ x.double()
    assert_close(S.ops.gemm(dy.t(), x.t()), ref, 5e-6, "gemm TN split-K")
    w = torch.randn(Co, C, device=DEV)
    assert_close(S.ops.gemm(dy, w.t()), dy.double() @ w.double(), 2e-6, "gemm NN")
    bias = torch.randn(N, device=DEV)
    assert_close(S.ops.gemm(a, b, bias), a.double() @ b.double().t() + bias.double(), 2e-6, "gemm bias")


def test_gemm_tc(S):
    """tcgen05 3xTF32 GEMM (all operand majors, ragged edges, split-K) against fp64."""
    torch.manual_seed(1)
    for (M, N, K) in [(256, 128, 64), (1000, 200, 100), (65536, 128, 64), (8192, 512, 512), (4096, 64, 128), (130, 72, 36)]:
        a = torch.randn(M, K, device=DEV)
        b = torch.randn(N, K, device=DEV)
        bias = torch.randn(N, device=DEV)
        ref = a.double() @ b.double().t() + bias.double()
        assert_close(S.ops.gemm_tc(a, b, bias), ref, 1e-5, f"gemm_tc NT {M}x{N}x{K}")
    buf = torch.randn(4096, 512, device=DEV)
    a, b = buf[:, 128:256], torch.randn(256, 128, device=DEV)
    assert_close(S.ops.gemm_tc(a, b), a.double() @ b.double().t(), 1e-5, "gemm_tc strided A")
    for (P, Co, C) in [(4096, 128, 64), (65536, 512, 128), (20000, 96, 40)]:
        dy, x = torch.randn(P, Co, device=DEV), torch.randn(P, C, device=DEV)
        assert_close(S.ops.gemm_tc(dy.t(), x.t()), dy.double().t() @ x.double(), 1e-4, f"gemm_tc TN {P}x{Co}x{C}")
    dy, w = torch.randn(8192, 256, device=DEV), torch.randn(256, 128, device=DEV)
    assert_close(S.ops.gemm_tc(dy, w.t()), dy.double() @ w.double(), 1e-5, "gemm_tc NN")


@pytest.mark.parametrize("C,N", [(3, 256), (64, 256), (128, 128)])
def test_knn_golden(S, golden, C, N):
    x = O.synth_clouds(2, N, 10 + C)[0].squeeze(-1) if C == 3 else feat_input(2, C, N, 10 + C)
    idx = S.model_utils.knn(x.to(DEV), 20)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (2, N, 20)
    g = torch.from_numpy(golden(f"knn_c{C}")["idx"].astype(np.int64))
    same = (idx.cpu().sort(-1)[0] == g.sort(-1)[0]).all(-1)
    n_diff, wd, wn = knn_report(x, idx, 20)
    print(f"knn golden C={C}: rows differing from reference {int((~same).sum())}, near-tie gap {wd:.2e}/{wn:.2e}")
    assert wn < 1e-6
    assert int((~same).sum()) <= max(1, int(2e-3 * same.numel()))


@pytest.mark.parametrize("B,C,N,k,layout", [(4, 3, 1024, 20, "cm"), (4, 64, 1024, 20, "pm"), (2, 128, 1024, 40, "pm"),
                                            (3, 3, 1000, 20, "cm"), (2, 64, 333, 20, "pm"), (1, 7, 50, 50, "cm"),
                                            (2, 16, 128, 20, "pm"), (1, 32, 2000, 20, "pm"), (2, 64, 777, 40, "pm"),
                                            (2, 96, 640, 20, "pm"), (3, 64, 20, 20, "pm")])
def test_knn_oracle(S, B, C, N, k, layout):
    x = O.synth_clouds(B, N, 3)[0].squeeze(-1) if C == 3 else feat_input(B, C, N, 5)
    if layout == "cm":
        idx = S.ops.knn_cm(x.to(DEV), k)
    else:
        idx = S.ops.knn_pm(x.to(DEV).transpose(1, 2).contiguous(), k)
    n_diff, wd, wn = knn_report(x, idx, k)
    print(f"knn B={B} C={C} N={N} k={k}: {n_diff}/{B * N} rows differ (near-ties), worst gap rel d_k {wd:.2e}, "
          f"rel norms {wn:.2e}")
    assert wn < 1e-6, "a neighbour differs from torch.topk beyond an fp32 near-tie"
    assert n_diff <= max(2, int(2e-3 * B * N))
    if N >= k and C == 3:
        assert bool((idx[..., 0].cpu() == torch.arange(N).view(1, N)).all()), "self must be neighbour 0"


def test_knn_duplicates_and_strided_rows(S):
    """Exact ties (duplicated points) must still give k distinct neighbours with the top-k distances, and
    a channel slice of a wider point-major buffer is a legal operand."""
    B, C, N, k = 2, 64, 512, 20
    x = feat_input(B, C, N, 77)                      # [B,C,N]
    x[:, :, 100:140] = x[:, :, 60:61]                # 41 identical points
    xpm = x.transpose(1, 2).contiguous().to(DEV)
    idx = S.ops.knn_pm(xpm, k)
    assert int(idx.min()) >= 0 and int(idx.max()) < N
    srt = idx.sort(-1)[0]
    assert bool((srt[..., 1:] != srt[..., :-1]).all()), "duplicate neighbour index"
    D = O.pairwise_neg_sqdist(x)
    ref_d = D.topk(k, dim=-1)[0]
    got_d = torch.gather(D, 2, idx.cpu().long()).sort(-1, descending=True)[0]
    assert float((ref_d - got_d).abs().max()) <= 1e-4 * float(D.abs().max())
    wide = torch.randn(B, N, 256, device=DEV)
    wide[:, :, 64:128] = xpm
    idx2 = S.ops.knn_pm(wide[:, :, 64:128], k)
    assert torch.equal(idx2.sort(-1)[0], srt)


def test_knn_reverse(S):
    x = feat_input(3, 16, 200, 9)
    idx = S.ops.knn_pm(x.to(DEV).transpose(1, 2).contiguous(), 20)
    rp, re = S.ops.knn_reverse(idx)
    idx_c, rp, re = idx.cpu().numpy(), rp.cpu().numpy(), re.cpu().numpy()
    for b in range(3):
        assert rp[b, 0] == 0 and rp[b, -1] == 200 * 20
        exp = {j: [] for j in range(200)}
        for i in range(200):
            for s in range(20):
                exp[idx_c[b, i, s]].append((i << 8) | s)
        for j in range(200):
            got = re[b, rp[b, j]:rp[b, j + 1]].tolist()
            assert sorted(got) == sorted(exp[j]), (b, j)


def _edge_case(S, B, C, N, Co, k, seed, golden_dict=None):
    x = feat_input(B, C, N, seed)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    W = torch.from_numpy(rng.standard_normal((Co, 2 * C, 1, 1)).astype(np.float32) * 0.3)
    ga = rng.uniform(0.5, 1.5, Co).astype(np.float32)
    ga[::3] *= -1
    be = torch.from_numpy(rng.standard_normal(Co).astype(np.float32) * 0.1)
    Rw = torch.from_numpy(rng.standard_normal((B, Co, N)).astype(np.float32))
    # oracle
    xo = x.clone().requires_grad_(True)
    sd = {"b.conv.0.weight": W.clone().requires_grad_(True), "b.conv.1.weight": torch.from_numpy(ga).clone().requires_grad_(True),
          "b.conv.1.bias": be.clone().requires_grad_(True), "b.conv.1.running_mean": torch.zeros(Co),
          "b.conv.1.running_var": torch.ones(Co), "b.conv.1.num_batches_tracked": torch.zeros((), dtype=torch.int64)}
    idx = O.knn(x, k)
    out_o = O.edgeconv(xo, sd, "b", True, k=k, idx=idx)
    (out_o * Rw).sum().backward()
    # ours
    blk = S.model_utils.conv_2d(2 * C, Co, 1, activation="leakyrelu", bias=False).to(DEV)
    with torch.no_grad():
        blk.conv[0].weight.copy_(W)
        blk.conv[1].weight.copy_(torch.from_numpy(ga))
        blk.conv[1].bias.copy_(be)
    blk.train()
    xg = x.to(DEV).transpose(1, 2).contiguous().requires_grad_(True)
    out = blk.edgeconv(xg, idx.to(DEV).int())
    (out * Rw.to(DEV).transpose(1, 2)).sum().backward()
    assert_close(out.transpose(1, 2), out_o, 1e-4, "edgeconv out")
    assert_close(xg.grad.transpose(1, 2), xo.grad, 1e-3, "edgeconv dx")
    assert_close(blk.conv[0].weight.grad, sd["b.conv.0.weight"].grad, 1e-3, "edgeconv dW")
    assert_close(blk.conv[1].weight.grad, sd["b.conv.1.weight"].grad, 1e-3, "edgeconv dgamma")
    assert_close(blk.conv[1].bias.grad, sd["b.conv.1.bias"].grad, 1e-3, "edgeconv dbeta")
    assert_close(blk.conv[1].running_mean, sd["b.conv.1.running_mean"], 1e-4, "running_mean")
    assert_close(blk.conv[1].running_var, sd["b.conv.1.running_var"], 1e-4, "running_var")
    assert int(blk.conv[1].num_batches_tracked) == 1
    if golden_dict is not None:
        assert_close(out.transpose(1, 2), golden_dict["out"], 1e-4, "edgeconv out vs reference fixture")
        assert_close(xg.grad.transpose(1, 2), golden_dict["dx"], 1e-3, "edgeconv dx vs fixture")
        assert_close(blk.conv[0].weight.grad, golden_dict["dW"], 1e-3, "edgeconv dW vs fixture")
        assert_close(blk.conv[1].weight.grad, golden_dict["dgamma"], 1e-3, "dgamma vs fixture")
        assert_close(blk.conv[1].running_var, golden_dict["running_var"], 1e-4, "running_var vs fixture")
    # eval mode uses the running statistics
    blk.eval()
    sd_eval = {k_: v.detach() for k_, v in sd.items()}
    with torch.no_grad():
        out_e = blk.edgeconv(xg.detach(), idx.to(DEV).int())
        out_eo = O.edgeconv(x, sd_eval, "b", False, k=k, idx=idx)
    assert_close(out_e.transpose(1, 2), out_eo, 1e-4, "edgeconv eval")


def test_edgeconv_golden(S, golden):
    _edge_case(S, 2, 8, 128, 16, 20, 21, golden("edgeconv_block"))


@pytest.mark.parametrize("B,C,N,Co,k", [(2, 3, 512, 64, 20), (2, 64, 1024, 64, 20), (2, 64, 256, 128, 40),
                                        (1, 128, 300, 256, 20)])
def test_edgeconv_oracle(S, B, C, N, Co, k):
    _edge_case(S, B, C, N, Co, k, 100 + C + Co)


@pytest.mark.parametrize("pool,slope,bias", [(1, 0.2, False), (0, 0.0, True)])
def test_mlp_pool(S, pool, slope, bias):
    import torch.nn.functional as F
    B, N, Ci, Co = 3, 500, 72, 136
    rng = np.random.Generator(np.random.PCG64(77))
    x = torch.from_numpy(rng.standard_normal((B, N, Ci)).astype(np.float32))
    W = torch.from_numpy(rng.standard_normal((Co, Ci)).astype(np.float32) * 0.2)
    bi = torch.from_numpy(rng.standard_normal(Co).astype(np.float32)) if bias else None
    ga = rng.uniform(0.5, 1.5, Co).astype(np.float32)
    ga[::4] *= -1
    ga = torch.from_numpy(ga)
    be = torch.from_numpy(rng.standard_normal(Co).astype(np.float32) * 0.1)
    Rw = torch.from_numpy(rng.standard_normal((B, Co * (2 if pool else 1))).astype(np.float32))
    # oracle: conv1d -> BN -> act -> pool
    ps = [t.clone().requires_grad_(True) for t in (x, W, ga, be)] + ([bi.clone().requires_grad_(True)] if bias else [None])
    rm, rv = torch.zeros(Co), torch.ones(Co)
    y = F.conv1d(ps[0].transpose(1, 2), ps[1].unsqueeze(-1), ps[4])
    z = F.leaky_relu(F.batch_norm(y, rm, rv, ps[2], ps[3], True, 0.1, 1e-5), slope)
    o_ref = torch.cat((z.max(2)[0], z.mean(2)), 1) if pool else z.max(2)[0]
    (o_ref * Rw).sum().backward()
    # ours
    gs = [t.to(DEV).requires_grad_(True) for t in (x, W, ga, be)] + ([bi.to(DEV).requires_grad_(True)] if bias else [None])
    rmg, rvg = torch.zeros(Co, device=DEV), torch.ones(Co, device=DEV)
    o = S.ops.mlp_bn_act_pool(gs[0], gs[1], gs[4], gs[2], gs[3], rmg, rvg, True, slope, pool)
    (o * Rw.to(DEV)).sum().backward()
    assert_close(o, o_ref, 1e-4, "pool out")
    for a, b, n in zip(gs[:4], ps[:4], ("dx", "dW", "dgamma", "dbeta")):
        assert_close(a.grad, b.grad, 1e-3, "pool " + n)
    if bias:
        assert float(gs[4].grad.abs().max()) <= 1e-6 + 1e-3 * float(ps[4].grad.abs().max() + 1e-6)
    assert_close(rmg, rm, 1e-4, "pool running_mean")
    assert_close(rvg, rv, 1e-4, "pool running_var")
    with torch.no_grad():
        oe = S.ops.mlp_bn_act_pool(gs[0], gs[1], gs[4], gs[2], gs[3], rmg, rvg, False, slope, pool)
        ze = F.leaky_relu(F.batch_norm(F.conv1d(x.transpose(1, 2), W.unsqueeze(-1), bi), rm, rv, ga, be, False, 0.1, 1e-5), slope)
        oe_ref = torch.cat((ze.max(2)[0], ze.mean(2)), 1) if pool else ze.max(2)[0]
    assert_close(oe, oe_ref, 1e-4, "pool eval")


def _mmd_inputs():
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


def test_mmd_golden(S, golden):
    g = golden("mmd")
    X, Y, Xs, Ys, ls, lt, w, ps, pt, ds, dt = _mmd_inputs()
    d = lambda t: t.to(DEV)
    Xg, Yg = d(X).requires_grad_(True), d(Y).requires_grad_(True)
    assert_close(S.mmd.mix_rbf_mmd2(Xg, Yg, S.mmd.sigma_list), g["v_plain"], 1e-4, "mmd plain")
    v = S.mmd.mix_rbf_mmd2(Xg, Yg, S.mmd.sigma_list, sample_weights=d(w))
    v.backward()
    assert_close(v, g["v_w"], 1e-4, "mmd weighted")
    # Gradients: the reference's fp32 autograd cancels catastrophically on the Gram diagonal
    # (oracle/sug_oracle.py:mmd_cal docstring), so the fixture's dX is itself far from the exact
    # gradient.  The kernel must match the fp64 evaluation of the same loss; the fixture's own
    # error against fp64 is reported and bounds the kernel-vs-fixture distance (triangle).
    def truth(A, Bm, scale=1.0):
        A64, B64 = A.double().requires_grad_(True), Bm.double().requires_grad_(True)
        (scale * O.mix_rbf_mmd2(A64, B64, sample_weights=w.double())).backward()
        return A64.grad, B64.grad
    tX, tY = truth(X, Y)
    e_fix = relerr(torch.from_numpy(g["dX"]), tX)
    print(f"mmd dX (4096-d): kernel vs fp64 {relerr(Xg.grad, tX):.2e}; reference fp32 fixture vs fp64 {e_fix:.2e}")
    assert_close(Xg.grad, tX, 1e-3, "mmd dX vs fp64")
    assert_close(Yg.grad, tY, 1e-3, "mmd dY vs fp64")
    assert relerr(Xg.grad, g["dX"]) <= 1.1 * e_fix + 1e-3
    Xsg, Ysg = d(Xs).requires_grad_(True), d(Ys).requires_grad_(True)
    v = S.mmd.mix_rbf_mmd2(Xsg, Ysg, S.mmd.sigma_list, sample_weights=d(w))
    (3.0 * v).backward()
    assert_close(v, g["v_sem"], 1e-4, "mmd sem")
    tXs, tYs = truth(Xs, Ys, 3.0)
    e_fix = relerr(3.0 * torch.from_numpy(g["dXs"]), tXs)
    print(f"mmd dX (256-d): kernel vs fp64 {relerr(Xsg.grad, tXs):.2e}; reference fp32 fixture vs fp64 {e_fix:.2e}")
    assert_close(Xsg.grad, tXs, 1e-3, "mmd dXs vs fp64")
    assert_close(Ysg.grad, tYs, 1e-3, "mmd dYs vs fp64")
    assert relerr(Xsg.grad, 3.0 * g["dXs"]) <= 1.1 * e_fix + 1e-3
    assert_close(S.mmd.geometric_weights(d(ds), d(dt), weighting="mean2one"), g["geo_w"], 1e-4, "geo weights")
    assert_close(S.mmd.prob_weights_soft(d(ps), d(pt), d(ls), d(lt), 0.5, "mean2one"), g["sem_w"], 1e-3, "sem weights")
    assert_close(S.mmd.mmd_cal(d(ls), d(X), d(lt), d(Y), O.SUG_CFG["GEO_MMD"], data_s=d(ds), data_t=d(dt)), g["geo"],
                 1e-4, "mmd_cal geo")
    assert_close(S.mmd.mmd_cal(d(ls), d(Xs), d(lt), d(Ys), O.SUG_CFG["SEM_MMD"], data_s=d(ps), data_t=d(pt)), g["sem"],
                 1e-3, "mmd_cal sem")
    c1, c2 = S.ops.chamfer(d(ds).squeeze(-1).transpose(1, 2), d(dt).squeeze(-1).transpose(1, 2))
    assert_close(c1, g["cd1"], 1e-5, "chamfer d1")
    assert_close(c2, g["cd2"], 1e-5, "chamfer d2")


def test_api_functions_golden(S, golden):
    """The helper functions in the reference's own layouts, against outputs of the unmodified reference:
    get_graph_feature, farthest_point_sample (same CPU RNG draw), index_points, query_ball_point (radius and
    k-NN form), upsample_inter, square_distance, focal_loss (gamma = 2, sum reduction, alpha re-gathered)."""
    g = golden("api_funcs")
    d = lambda t: t.to(DEV)
    xa = feat_input(2, 8, 64, 71)
    assert_close(S.model_utils.get_graph_feature(d(xa), k=5), g["graph_feature"], 1e-5, "get_graph_feature")
    pc = O.synth_clouds(2, 128, 72)[0].squeeze(-1)
    torch.manual_seed(73)
    fi = S.point_utils.farthest_point_sample(d(pc), 16)
    assert fi.dtype == torch.int64 and np.array_equal(fi.cpu().numpy(), g["fps"])
    ctr = S.point_utils.index_points(d(pc), fi)
    assert_close(ctr, g["centres"], 1e-6, "index_points")
    qb = S.point_utils.query_ball_point(0.3, 8, d(pc), ctr)
    qk = S.point_utils.query_ball_point(None, 8, d(pc), ctr)
    assert np.array_equal(qb.cpu().numpy(), g["ball"]) and np.array_equal(qk.cpu().numpy(), g["knn8"])
    assert_close(S.point_utils.index_points(d(pc), qb), g["grouped"], 1e-6, "index_points grouped")
    nodes_f = feat_input(2, 32, 16, 74)
    assert_close(S.point_utils.upsample_inter(d(pc), ctr, None, d(nodes_f), 3), g["upsample"], 1e-4, "upsample_inter")
    assert_close(S.point_utils.square_distance(d(pc), ctr), g["sqdist"], 1e-5, "square_distance")
    fl = S.model_utils.focal_loss(num_classes=10, gamma=2, alpha=[0.05 * (i + 1) for i in range(10)], size_average=False)
    rngf = np.random.Generator(np.random.PCG64(75))
    lg = torch.from_numpy(rngf.standard_normal((12, 10)).astype(np.float32) * 2)
    lb = torch.from_numpy(rngf.integers(0, 10, 12).astype(np.int64))
    assert_close(fl(d(lg), d(lb)), g["focal1"], 1e-5, "focal loss")
    assert_close(fl(d(lg) * 0.5, (d(lb) + 3) % 10), g["focal2"], 1e-5, "focal loss, second call (stateful alpha)")
    # gradients of the fused kernel against the oracle's tensor-op spelling of model_utils.py:164-176, gamma = 2, 0, 0.5
    for gamma, mean in ((2.0, False), (0.0, True), (0.5, True)):
        a0 = [0.05 * (i + 1) for i in range(10)]
        f_gpu = S.model_utils.focal_loss(num_classes=10, gamma=gamma, alpha=a0, size_average=mean)
        zc = lg.clone().requires_grad_(True)
        zg = d(lg).requires_grad_(True)
        logsoft = torch.log_softmax(zc, dim=1)
        soft = torch.exp(logsoft).gather(1, lb.view(-1, 1))
        loss_c = torch.mul(torch.tensor(a0).gather(0, lb), (-torch.pow(1 - soft, gamma) * logsoft.gather(1, lb.view(-1, 1))).t())
        lc = loss_c.mean() if mean else loss_c.sum()
        lgp = f_gpu(zg, d(lb))
        (1.7 * lc).backward()
        (1.7 * lgp).backward()
        assert_close(lgp, lc, 1e-5, f"focal gamma={gamma}")
        assert_close(zg.grad, zc.grad, 1e-4, f"focal grad gamma={gamma}")


def test_mmd_modes_golden(S, golden):
    """The mmd_cal / weighting modes the SUG config does not exercise, against the unmodified reference:
    HARD_MMD, OFF, the unbiased estimator, weighting 'none' (mmd.py:25-41, 69-77, 178-202, 274-312)."""
    g = golden("mmd_modes")
    X, Y, Xs, Ys, ls, lt, w, ps, pt, ds, dt = _mmd_inputs()
    d = lambda t: t.to(DEV)
    lt_same = torch.from_numpy(g["lt_same"].astype(np.int64))
    assert_close(S.mmd.mmd_cal(d(ls), d(Xs), d(lt_same), d(Ys), {"NAME": "HARD_MMD"}), g["hard"], 1e-4, "hard mmd")
    assert_close(S.mmd.mmd_cal(d(ls), d(Xs), d(lt), d(Ys), {"NAME": "OFF"}), g["off"], 1e-4, "mmd off")
    assert_close(S.mmd.mix_rbf_mmd2(d(Xs), d(Ys), S.mmd.sigma_list, biased=False), g["unbiased"], 1e-4, "unbiased mmd")
    cd = S.mmd.cd_distance(d(ds).squeeze(-1).transpose(1, 2), d(dt).squeeze(-1).transpose(1, 2))
    assert_close(cd, g["cd"], 1e-5, "cd_distance")
    assert_close(S.mmd.distance2weights(cd, method="none"), g["w_none"], 1e-5, "weights none")
    assert_close(S.mmd.geometric_weights(d(ds), d(dt), weighting="none"), g["geo_none"], 1e-5, "geo weights none")
    with pytest.raises(RuntimeError):
        S.mmd.mmd_cal(d(ls), d(Xs), d(lt), d(Ys), {"NAME": "LINEAR"})
    # MAX_HARD_MMD (mmd.py:96-105) and the entropy distance / weights (mmd.py:155-166)
    ia, ib = S.mmd.get_most_overlapped_element(d(ls), d(lt))
    assert ia == g["mh_s"].tolist() and ib == g["mh_t"].tolist()
    assert_close(S.mmd.mmd_cal(d(ls), d(Xs), d(lt), d(Ys), {"NAME": "MAX_HARD_MMD"}), g["max_hard"], 1e-4, "max hard mmd")
    prob_s, prob_t = torch.softmax(Xs[:, :10], 1), torch.softmax(Ys[:, :10], 1)
    assert_close(S.mmd.entropy_dis(d(prob_s), d(prob_t)), g["ent_dis"], 1e-4, "entropy distance")
    assert_close(S.mmd.entropy_weights(d(prob_s), d(prob_t), weighting="mean2one"), g["ent_w"], 1e-4, "entropy weights")


def test_mmd_unbiased_and_sizes(S):
    rng = np.random.Generator(np.random.PCG64(5))
    for m, D in ((64, 4106), (64, 266), (7, 33)):
        X = torch.from_numpy(rng.standard_normal((m, D)).astype(np.float32) * 0.3)
        Y = torch.from_numpy(rng.standard_normal((m, D)).astype(np.float32) * 0.35)
        for biased in (True, False):
            # fp64 evaluation of the reference formula (see test_mmd_golden for why not fp32)
            Xo, Yo = X.double().requires_grad_(True), Y.double().requires_grad_(True)
            vo = O.mix_rbf_mmd2(Xo, Yo, biased=biased)
            vo.backward()
            Xg, Yg = X.to(DEV).requires_grad_(True), Y.to(DEV).requires_grad_(True)
            vg = S.mmd.mix_rbf_mmd2(Xg, Yg, S.mmd.sigma_list, biased=biased)
            vg.backward()
            assert_close(vg, vo, 1e-4 if biased else 1e-3, f"mmd m={m} D={D} biased={biased}")
            assert_close(Xg.grad, Xo.grad, 2e-3, "mmd dX")
            assert_close(Yg.grad, Yo.grad, 2e-3, "mmd dY")


def test_adapt_indices(S):
    B, N = 3, 1024
    loc = O.synth_clouds(B, N, 61)[0].squeeze(-1)
    start = torch.tensor([5, 1000, 333])
    f_o = O.farthest_point_sample(loc, 64, start)
    f_g = S.ops.fps(loc.to(DEV), 64, start).cpu().long()
    assert bool((f_o == f_g).all()), "FPS indices differ"
    f_loc = O.index_points(loc, f_o)
    b_o = O.query_ball_point(0.3, 64, loc, f_loc)
    b_g = S.ops.ball_query(loc.to(DEV), f_loc.to(DEV), 0.3, 64).cpu().long()
    frac = float((b_o != b_g).float().mean())
    print(f"ball query: fraction of differing entries {frac:.2e}")
    assert frac < 1e-3
    node = f_loc + 0.01 * feat_input(B, 3, 64, 62)
    k_o = O.query_ball_point(None, 64, loc, node)
    k_g = S.ops.knn_query(loc.to(DEV), node.to(DEV), 64).cpu().long()
    same = (k_o.sort(-1)[0] == k_g.sort(-1)[0]).all(-1).float().mean()
    assert float(same) > 0.995, f"64-NN sets agree on only {float(same):.4f} of the nodes"
    d, i3 = O.square_distance(loc, node).sort(dim=-1)
    t_g = S.ops.three_nn(loc.to(DEV), node.to(DEV), 3).cpu().long()
    assert float((i3[..., :3] == t_g).float().mean()) > 0.999


def test_adapt_layer_golden(S, golden):
    g = golden("adapt_layer")
    sd = O.synth_state("Net_MDA:DGCNN")
    ad = S.model_utils.adapt_layer_off()
    ad.load_state_dict({k[len("g.node_fea_adapt."):]: v for k, v in sd.items() if k.startswith("g.node_fea_adapt.")})
    ad = ad.to(DEV).train()
    loc = O.synth_clouds(2, 256, 31)[0].squeeze(-1)
    fea = feat_input(2, 64, 256, 32)
    torch.manual_seed(5)
    o, nf, no = ad(fea.unsqueeze(3).to(DEV), loc.to(DEV))
    assert_close(o, g["out"], 1e-3, "adapt out")
    assert_close(nf, g["node_fea"], 1e-3, "adapt node_fea")
    assert_close(no, g["node_off"], 1e-3, "adapt node_off")


def _load(mod, spec, seed=666):
    mod.load_state_dict(O.synth_state(spec, seed), strict=True)
    return mod.to(DEV)


def test_dgcnn_g_golden(S, golden):
    g = golden("dgcnn_g")
    x, _ = O.synth_clouds(2, 1024, 41)
    sd = O.synth_state("Net_MDA:DGCNN")

    def ora():
        torch.manual_seed(7)
        a = O.dgcnn_trunk(x, sd, "g.", True, adapt=True)
        torch.manual_seed(8)
        b = O.dgcnn_trunk(x, sd, "g.", False, adapt=True)
        return a, b
    _, trace = oracle_trace(ora)
    net = _load(S.Model.Net_MDA("DGCNN"), "Net_MDA:DGCNN").train()
    with teacher_forced(S, trace):
        torch.manual_seed(7)
        f, n, _ = net.g(x.to(DEV), node=True)
        assert_close(f, g["feat_train"], 1e-3, "DGCNN feat (train)")
        assert_close(n, g["node_train"], 1e-3, "DGCNN node_fea (train)")
        net.eval()
        torch.manual_seed(8)
        with torch.no_grad():
            f, n, _ = net.g(x.to(DEV), node=True)
        assert_close(f, g["feat_eval"], 1e-3, "DGCNN feat (eval)")
        assert_close(n, g["node_eval"], 1e-3, "DGCNN node_fea (eval)")
    with torch.no_grad():
        torch.manual_seed(8)
        f2, _, _ = net.g(x.to(DEV), node=True)
    print(f"free-running (own kNN) eval feat vs fixture: {relerr(f2, g['feat_eval']):.2e}")
    assert relerr(f2, g["feat_eval"]) < 5e-2
    assert_close(net.g.conv1.conv[1].running_mean, g["rm1"], 1e-4, "rm1")
    assert_close(net.g.conv4.conv[1].running_var, g["rv4"], 1e-3, "rv4")
    assert_close(net.g.bn5.running_mean, g["rm5"], 1e-3, "rm5")
    assert_close(net.g.bn5.running_var, g["rv5"], 1e-3, "rv5")
    assert int(net.g.bn5.num_batches_tracked) == 1 and int(net.g.conv3.conv[1].num_batches_tracked) == 1


def test_net_mda_golden(S, golden):
    g = golden("net_mda_dgcnn")
    x, _ = O.synth_clouds(2, 1024, 41)
    net = _load(S.Model.Net_MDA("DGCNN"), "Net_MDA:DGCNN").train()
    for hd in (net.c1, net.c2):  # CPU and CUDA dropout streams differ: compare with dropout off
        hd.dropout1.p = hd.dropout2.p = 0.0
    sd = O.synth_state("Net_MDA:DGCNN")

    def ora():
        torch.manual_seed(11)
        a = O.net_mda(x, sd, True, semantic_adaption=True, drop_p=0.0)
        torch.manual_seed(12)
        b = O.net_mda(x, sd, True, node_adaptation_s=True)
        torch.manual_seed(13)
        c = O.net_mda(x, sd, True, node_adaptation_t=True)
        return a, b, c
    ((o1, o2, t1, t2), _, _), trace = oracle_trace(ora)
    with teacher_forced(S, trace):
        torch.manual_seed(11)
        y1, y2, s1, s2 = net(x.to(DEV), semantic_adaption=True)
        for a, b, n in ((y1, o1, "y1"), (y2, o2, "y2"), (s1, t1, "s1"), (s2, t2, "s2")):
            assert_close(a, b, 1e-3, n)
        torch.manual_seed(12)
        ns = net(x.to(DEV), node_adaptation_s=True)
        assert_close(ns, g["node_s"], 1e-3, "node_s vs fixture")
        torch.manual_seed(13)
        nt = net(x.to(DEV), node_adaptation_t=True)
        assert_close(nt, g["node_t"], 1e-3, "node_t vs fixture")


def test_net_mda_pointnet_golden(S, golden):
    """Net_MDA('Pointnet') against the fixture of the unmodified reference: eval logits, node branches."""
    g = golden("net_mda_pointnet")
    x, _ = O.synth_clouds(2, 1024, 41)
    net = _load(S.Model.Net_MDA("Pointnet"), "Net_MDA:Pointnet", 668).eval()
    with torch.no_grad():
        torch.manual_seed(21)
        y1, y2 = net(x.to(DEV))
    assert_close(y1, g["y1_eval"], 1e-3, "pointnet y1 eval")
    assert_close(y2, g["y2_eval"], 1e-3, "pointnet y2 eval")
    net.train()
    torch.manual_seed(22)
    assert_close(net(x.to(DEV), node_adaptation_s=True), g["node_s"], 1e-3, "pointnet node_s")
    torch.manual_seed(23)
    assert_close(net(x.to(DEV), node_adaptation_t=True), g["node_t"], 1e-3, "pointnet node_t")


def test_dgcnn_cls_golden(S, golden):
    x, _ = O.synth_clouds(2, 1024, 41)
    cls = _load(S.model_pointnet.DGCNN(), "DGCNN_cls", 667).eval()
    _, trace = oracle_trace(lambda: O.dgcnn_cls(x, O.synth_state("DGCNN_cls", 667), False))
    with torch.no_grad(), teacher_forced(S, trace):
        lg = cls(x.to(DEV))
    assert_close(lg, golden("dgcnn_cls")["logits_eval"], 1e-3, "DGCNN_cls logits")
    with torch.no_grad():
        lg2 = cls(x.to(DEV))
    print(f"free-running (own kNN) logits vs fixture: {relerr(lg2, golden('dgcnn_cls')['logits_eval']):.2e}")


def test_pointnet_g_golden(S, golden):
    g = golden("pointnet_g")
    x, _ = O.synth_clouds(2, 1024, 41)
    pn = _load(S.Model.Net_MDA("Pointnet"), "Net_MDA:Pointnet", 668).train()
    torch.manual_seed(17)
    f, n, off = pn.g(x.to(DEV), node=True)
    assert_close(f, g["feat"], 1e-3, "Pointnet feat")
    assert_close(n, g["node_fea"], 1e-3, "Pointnet node_fea")
    assert_close(off, g["node_off"], 1e-3, "Pointnet node_off")
    assert_close(pn.g.conv5.conv[1].running_var, g["rv5"], 1e-3, "Pointnet rv5")


# Parameters that are NOT upstream of any MMD loss (train_dg_single_gpu.py:309-320: the geometric MMD sees g + attention_*,
# the semantic MMD sees g + c*.mlp1 + c*.mlp2): their gradients come from the classification loss alone.
NOT_UPSTREAM_OF_MMD = ("c1.mlp3.weight", "c1.mlp3.bias", "c2.mlp3.weight", "c2.mlp3.bias")


# A per-channel constant added in front of a train-mode BatchNorm has an identically zero gradient: what either side
# computes for these two biases is rounding noise (|g| ~ 1e-8 against 0.4 for the largest gradient).
ZERO_GRADIENTS = ("g.conv1d.bias", "g.node_fea_adapt.residual.conv.0.bias")


def _grad_rows(params, ref_grads):
    """(relative error of the full gradient, relative error of its norm, name, |g_ref|) per parameter.  A gradient that
    is mathematically zero (g.conv1d.bias, *.residual.conv.0.bias: a per-channel constant in front of train-mode
    BatchNorm) is rounding noise on both sides, so errors are measured against max(|g_ref|, 1e-4 * largest norm)."""
    gmax = max(float(v.norm()) for v in ref_grads.values())
    rows = []
    for k, go in ref_grads.items():
        gp = params[k].grad.detach().cpu().double()
        den = max(float(go.norm()), 1e-4 * gmax)
        rows.append((float((gp - go.double()).norm()) / den, abs(float(gp.norm()) - float(go.norm())) / den, k, float(go.norm())))
    return sorted(rows, reverse=True)


def test_sug_step_golden(S, golden):
    """One whole SUG step (B = 12 + 12) against the unmodified reference's fixture and against the oracle.

    Gates (north_star: losses, logits and gradients within rel 1e-3):
      * losses, logits: 1e-3 against the REFERENCE fixture;
      * gradients of the parameters that are not upstream of an MMD loss: 1e-3 (full gradient and norm) against the
        REFERENCE fixture;
      * every other parameter is upstream of `mix_rbf_mmd2`, whose fp32 autograd in the reference is dominated by
        cancellation noise on the Gram diagonal (mmd.py:245-247; test_mmd_golden): the reference's own gradient of
        those parameters is that noise plus the true gradient.  Their NORMS must agree with the reference fixture as
        well as the norms of the exact (fp64-MMD) gradient of the same loss do -- bound_k = |n64_k - n_ref_k| / n_ref_k
        + 1e-3, measured here from the oracle and printed -- and the gradients themselves must match that exact
        gradient (oracle, fp64 MMD autograd): norms at 1e-3, full vectors at 5e-3 (the step is chaotic at the 1e-3
        level: a 1-ulp weight perturbation moves the reference's own gradients by 3.4e-3, tools/reference_sensitivity.py).
    """
    g = golden("sug_step")
    Bs = 12
    data, label = O.synth_clouds(Bs, 1024, 0)
    data_t, label_t = O.synth_clouds(Bs, 1024, 1)
    net = _load(S.Model.Net_MDA("DGCNN"), "Net_MDA:DGCNN").train()
    for hd in (net.c1, net.c2):
        hd.dropout1.p = hd.dropout2.p = 0.0
    crit = S.model_utils.focal_loss(num_classes=10, gamma=0.0, alpha=[0.1] * 10)
    sd = O.clone_state(O.synth_state("Net_MDA:DGCNN"), requires_grad=True)

    def ora():
        torch.manual_seed(101)
        return O.sug_losses(sd, data, label, data_t, label_t, O.FocalLoss([0.1] * 10, 0.0), drop_p=0.0,
                            mmd_dtype=torch.float64)
    ro, trace = oracle_trace(ora)
    ro["loss"].backward()
    assert len(trace) == 16
    with teacher_forced(S, trace):
        torch.manual_seed(101)
        r = S.step.sug_losses(net, data.to(DEV), label.to(DEV), data_t.to(DEV), label_t.to(DEV), crit)
    r["loss"].backward()
    for k in ("loss", "loss_cls", "loss_geo", "loss_sem"):
        assert_close(r[k], g[k], 1e-3, k)
    assert_close(r["pred_s1"], g["pred_s1"], 1e-3, "pred_s1")
    assert_close(r["pred_t1"], g["pred_t1"], 1e-3, "pred_t1")
    params = dict(net.named_parameters())
    assert_close(r["loss"], ro["loss"], 1e-3, "loss vs oracle")
    for k, p in params.items():
        if sd[k].grad is None:
            assert p.grad is None, f"{k} has a gradient here but not in the reference"
    ref64 = {k: v.grad for k, v in sd.items() if v.grad is not None}
    assert len(ref64) >= 50
    assert {k for k in ref64} == {k[3:] for k in g.keys() if k.startswith("gn.")}

    # (1) parameters outside the MMD: the REFERENCE fixture itself, 1e-3
    for k in NOT_UPSTREAM_OF_MMD:
        assert_close(params[k].grad, g["gf." + k], 1e-3, f"gradient of {k} vs the reference")
        assert abs(float(params[k].grad.norm()) - float(g["gn." + k])) <= 1e-3 * float(g["gn." + k]), k
    # (2) against the exact gradient of the same loss
    rows = _grad_rows(params, ref64)
    print(f"{len(rows)} gradients vs oracle (fp64 MMD); worst full-vector errors: {[(f'{a:.1e}', k) for a, _, k, _ in rows[:4]]}")
    gmax64 = max(t[3] for t in rows)
    for t in rows:  # mathematically zero gradients: rounding noise on both sides
        if t[2] in ZERO_GRADIENTS:
            assert t[3] < 1e-6 * gmax64 and float(params[t[2]].grad.norm()) < 1e-5 * gmax64, t
    worst_norm = max((t for t in rows if t[2] not in ZERO_GRADIENTS), key=lambda t: t[1])
    print(f"worst norm error vs oracle (fp64 MMD): {worst_norm[1]:.2e} ({worst_norm[2]})")
    assert rows[0][0] < 5e-3, f"gradient of {rows[0][2]}: rel err {rows[0][0]:.2e}"
    assert worst_norm[1] < 1e-3, f"gradient norm of {worst_norm[2]}: rel err {worst_norm[1]:.2e}"
    # (3) norms against the REFERENCE fixture, within the reference's own MMD noise (measured: the bound column)
    gmax = max(float(v) for k, v in g.items() if k.startswith("gn."))
    table = []
    for k, go in ref64.items():
        if k in ZERO_GRADIENTS:
            continue  # rounding-noise gradients, handled above
        nref = max(float(g["gn." + k]), 1e-4 * gmax)
        bound = abs(float(go.norm()) - float(g["gn." + k])) / nref + 1e-3
        err = abs(float(params[k].grad.norm()) - float(g["gn." + k])) / nref
        table.append((bound, err, k))
        assert err <= bound, f"|grad {k}|: {err:.2e} from the reference fixture, exact-gradient bound {bound:.2e}"
    table.sort(reverse=True)
    print("gradient norms vs the reference fixture; the five largest reference-noise bounds (bound, ours, name): "
          + ", ".join(f"({b:.1e}, {e:.1e}, {k})" for b, e, k in table[:5]))
    assert params["g.input_transform_net.fc3.weight"].grad is None
    assert params["g.node_fea_adapt.trans.conv.0.weight"].grad is None


def test_sug_step_b64_golden(S, golden):
    """The batch bench.py times (BASELINE.json configs[1]: 64 + 64 clouds x 1024 points) against the UNMODIFIED
    reference's fixture (tests/golden/sug_step_b64.npz), FREE-RUNNING: the neighbour graphs are this library's own.
    The fixture holds a 16-bit hash of every row of the reference's 16 neighbour lists, so the rows whose neighbour SET
    differs are counted over all 16 x 65 536 rows (north_star: bit-exact except genuine near-ties, counted and
    reported); losses, logits, node / semantic features and BatchNorm statistics at 1e-3; the gradients of the
    parameters outside the MMD against the reference; every other gradient norm against the exact gradient of the same
    loss and, within the reference's own fp32-MMD noise (see test_sug_step_golden), against the reference."""
    g = golden("sug_step_b64")
    Bs = 64
    data, label = O.synth_clouds(Bs, 1024, 0)
    data_t, label_t = O.synth_clouds(Bs, 1024, 1)
    net = _load(S.Model.Net_MDA("DGCNN"), "Net_MDA:DGCNN").train()
    for hd in (net.c1, net.c2):
        hd.dropout1.p = hd.dropout2.p = 0.0
    crit = S.model_utils.focal_loss(num_classes=10, gamma=0.0, alpha=[0.1] * 10)
    hashes, xyz_lists = [], {}
    real_cm, real_pm = S.ops.knn_cm, S.ops.knn_pm

    def rec(fn):
        def f(x, k):
            idx = fn(x, k)
            if len(hashes) % 4 == 0:
                xyz_lists[len(hashes)] = idx.cpu().long()
            hashes.append(O.knn_row_hash(idx).cpu())
            return idx
        return f
    S.ops.knn_cm, S.ops.knn_pm = rec(real_cm), rec(real_pm)
    try:
        torch.manual_seed(101)
        r = S.step.sug_losses(net, data.to(DEV), label.to(DEV), data_t.to(DEV), label_t.to(DEV), crit)
    finally:
        S.ops.knn_cm, S.ops.knn_pm = real_cm, real_pm
    r["loss"].backward()
    ours = torch.stack(hashes).numpy()
    ref = np.asarray(g["knn_hash"])
    assert ours.shape == ref.shape == (16, Bs, 1024)
    per_call = (ours != ref).reshape(16, -1).sum(1)
    print(f"rows whose neighbour set differs from the reference's, per kNN call (of {Bs * 1024} each): {per_call.tolist()}")
    # xyz graphs (calls 0, 4, 8, 12) have bit-identical inputs on both sides: a differing row must be a genuine fp32
    # near-tie of the reference's own distance matrix (relative gap < 1e-6), checked row by row
    n_tie = 0
    for c in (0, 4, 8, 12):
        pts = (data if c % 8 == 0 else data_t).squeeze(-1)
        for bi, ri in np.argwhere(ours[c] != ref[c]).tolist():
            D = O.pairwise_neg_sqdist(pts[bi:bi + 1])[0, ri]  # the reference's fp32 keys of that row
            mine = D[xyz_lists[c][bi, ri]]
            kth = D.topk(20)[0][-1]
            xx = (pts[bi] ** 2).sum(0)
            gap = float((kth - mine.min()).abs() / (xx[ri] + xx.max()))
            assert gap < 1e-6, f"xyz kNN call {c}, cloud {bi}, row {ri}: not a near-tie (relative gap {gap:.2e})"
            n_tie += 1
    print(f"xyz rows differing from the reference: {n_tie}, all fp32 near-ties (relative gap < 1e-6)")
    # feature-space graphs: the inputs of layers 2-4 already differ at the 1e-6 level between the CPU reference and the
    # GPU, which flips fp32 near-ties (layer 2: <= 0.05 % of the rows, the share of rows with a relative k/(k+1) gap
    # below 1e-5), and every flip perturbs the next layer's input by ~1e-3 (measured cascade: ~0.01 %,
    # 0.05-0.4 %, 0.3-1.3 % of the rows at layers 2, 3, 4)
    frac = per_call.reshape(4, 4) / float(Bs * 1024)
    assert frac[:, 1].max() <= 5e-4 and frac[:, 2].max() <= 1e-2 and frac[:, 3].max() <= 3e-2, frac
    tol = 1e-3 if per_call.sum() == 0 else 5e-3
    for k in ("loss", "loss_cls", "loss_geo", "loss_sem"):
        assert_close(r[k], g[k], 1e-3, k)
    assert_close(r["pred_s1"], g["pred_s1"], tol, "pred_s1")
    assert_close(r["pred_t1"], g["pred_t1"], tol, "pred_t1")
    assert_close(net.g.conv4.conv[1].running_mean, g["rm_conv4"], 1e-3, "running_mean conv4")
    assert_close(net.g.bn5.running_var, g["rv_bn5"], 1e-3, "running_var bn5")
    params = dict(net.named_parameters())
    for k in NOT_UPSTREAM_OF_MMD:
        assert_close(params[k].grad, g["gf." + k], tol, f"gradient of {k} vs the reference")
    # gradient norms: against the exact gradient of the same loss (gn64.*: oracle with the MMD autograd in fp64, computed
    # with the fixture -- the oracle cannot run at this size on the GPU box), and against the REFERENCE's own norms (gn.*)
    # within the reference's fp32-MMD noise |gn64 - gn| / gn.  FREE_RUN = 3e-2 is the allowance for the neighbour rows
    # that legitimately differ at this size (up to 1.3 % of the rows at layer 4, see above): every such row reroutes
    # a max-pooled activation and its gradient.  The tight gates (norms 1e-3, vectors 5e-3) are those of
    # test_sug_step_golden, where the graphs are teacher-forced.
    FREE_RUN = 3e-2
    gmax = max(float(v) for k, v in g.items() if k.startswith("gn."))
    rows, table = [], []
    for k, v in g.items():
        if not k.startswith("gn.") or k[3:] in ZERO_GRADIENTS:
            continue
        name = k[3:]
        p = params[name]
        assert p.grad is not None, k
        n_ref, n64, n = float(v), float(g["gn64." + name]), float(p.grad.norm())
        rows.append((abs(n - n64) / max(n64, 1e-4 * gmax), name))
        bound = abs(n64 - n_ref) / max(n_ref, 1e-4 * gmax) + FREE_RUN
        err = abs(n - n_ref) / max(n_ref, 1e-4 * gmax)
        table.append((bound, err, name))
        assert err <= bound, f"|grad {name}|: {err:.2e} from the reference fixture, exact-gradient bound {bound:.2e}"
    rows.sort(reverse=True)
    table.sort(reverse=True)
    print(f"{len(rows)} gradient norms vs the exact (fp64-MMD) gradient, worst: " + ", ".join(f"({e:.1e}, {k})" for e, k in rows[:4]))
    print("vs the reference fixture; largest reference-noise bounds (bound, ours, name): "
          + ", ".join(f"({b:.1e}, {e:.1e}, {k})" for b, e, k in table[:4]))
    assert rows[0][0] <= FREE_RUN, rows[0]
    n_with_grad = sum(1 for pp in params.values() if pp.grad is not None)
    assert n_with_grad == len(rows) + len(ZERO_GRADIENTS)


def test_sug_step_pointnet_vs_oracle(S):
    """The same SUG step on the PointNet backbone (Net_MDA('Pointnet'): T-Nets, shared MLP + max pool, adapt
    layer) against the oracle, whose Pointnet modules are pinned by pointnet_g.npz / net_mda_pointnet.npz."""
    Bs = 12
    data, label = O.synth_clouds(Bs, 1024, 4)
    data_t, label_t = O.synth_clouds(Bs, 1024, 5)
    net = _load(S.Model.Net_MDA("Pointnet"), "Net_MDA:Pointnet", 668).train()
    for hd in (net.c1, net.c2):
        hd.dropout1.p = hd.dropout2.p = 0.0
    crit = S.model_utils.focal_loss(num_classes=10, gamma=0.0, alpha=[0.1] * 10)
    sd = O.clone_state(O.synth_state("Net_MDA:Pointnet", 668), requires_grad=True)
    torch.manual_seed(103)
    ro = O.sug_losses(sd, data, label, data_t, label_t, O.FocalLoss([0.1] * 10, 0.0), model_name="Pointnet", drop_p=0.0,
                      mmd_dtype=torch.float64)
    ro["loss"].backward()
    torch.manual_seed(103)
    r = S.step.sug_losses(net, data.to(DEV), label.to(DEV), data_t.to(DEV), label_t.to(DEV), crit)
    r["loss"].backward()
    for k in ("loss", "loss_cls", "loss_geo", "loss_sem"):
        assert_close(r[k], ro[k], 1e-3, "pointnet step " + k)
    assert_close(r["pred_s1"], ro["pred_s1"], 1e-3, "pointnet pred_s1")
    gmax = max(float(v.grad.norm()) for v in sd.values() if v.grad is not None)
    rows = []
    for k, p in net.named_parameters():
        go = sd[k].grad
        if go is None:
            assert p.grad is None, f"{k} has a gradient here but not in the reference"
            continue
        d = float((p.grad.detach().cpu().double() - go.double()).norm())
        rows.append((d / max(float(go.norm()), 1e-4 * gmax), k, float(go.norm())))
    rows.sort(reverse=True)
    print(f"{len(rows)} parameter gradients vs oracle; five worst (err, name, |g_ref|): {rows[:5]}")
    # Gate 8e-2: the PointNet step (global max-pools, T-Nets) is far more chaotic than the DGCNN one.
    # Perturbing the oracle's OWN weights by 2e-6 relative (the size of a 3xTF32 / summation-order
    # difference) moves its own gradients by up to 3.95e-2, with the same parameters on top
    # (g.conv1.conv.1.bias 3.95e-2, g.conv3.residual.conv.1.bias 3.3e-2; measured on CPU,
    # tools/reference_sensitivity.py pointnet 2e-6).  Losses and logits stay within 1e-3 (above).
    assert rows[0][0] < 8e-2, f"gradient of {rows[0][1]}: rel err {rows[0][0]:.2e}"
    med = sorted(r_[0] for r_ in rows)[len(rows) // 2]
    assert med < 5e-2, f"median gradient error {med:.2e}"  # the oracle's own median under that perturbation: 2.0e-2
    assert len(rows) >= 40


def test_full_size_properties(S):
    """BASELINE config-2 sizes (B=64, N=1024): properties that need no CPU oracle."""
    B, N, k = 64, 1024, 20
    x, _ = O.synth_clouds(B, N, 0)
    xg = x.to(DEV).squeeze(-1)
    idx = S.ops.knn_cm(xg, k).long()
    assert bool((idx[..., 0] == torch.arange(N, device=DEV)).all())
    D = -(torch.cdist(xg.transpose(1, 2), xg.transpose(1, 2)) ** 2)
    sel = torch.gather(D, 2, idx)
    kth = sel.min(-1)[0]
    mask = torch.ones_like(D, dtype=torch.bool).scatter_(2, idx, False)
    best_out = torch.where(mask, D, torch.full_like(D, -1e30)).max(-1)[0]
    assert bool((best_out <= kth + 2e-6).all()), "an excluded point is closer than the k-th neighbour"
    # a feature kNN at C=64 + EdgeConv: permuting the neighbour slots must not change the output
    f = feat_input(B, 64, N, 3).to(DEV).transpose(1, 2).contiguous()
    idf = S.ops.knn_pm(f, k)
    blk = S.model_utils.conv_2d(128, 64, 1, activation="leakyrelu", bias=False).to(DEV).train()
    with torch.no_grad():
        o1 = blk.edgeconv(f, idf)
        o2 = blk.edgeconv(f, idf.flip(-1).contiguous())
    assert float((o1 - o2).abs().max()) <= 1e-5 * float(o1.abs().max())
    # BatchNorm: per-channel statistics of the pre-activation are (0, 1) => running_mean moved by 0.1*mean
    assert int(blk.conv[1].num_batches_tracked) == 2


def test_node_offset_and_interp_weights(S):
    """Fused adapt-layer pieces against their tensor-op spelling (model_utils.py:107-117,
    point_utils.py:141-160) in fp64, values and gradients."""
    torch.manual_seed(11)
    B, N, Sn, G = 3, 200, 16, 24
    xyz = torch.rand(B, 3, N, device=DEV) * 2 - 1
    h = torch.randn(B, N, 3, device=DEV, requires_grad=True)
    fidx = torch.stack([torch.randperm(N, device=DEV)[:Sn] for _ in range(B)])
    gidx = torch.randint(0, N, (B, Sn, G), device=DEV)
    out = S.ops.node_offset(h, xyz, fidx, gidx)
    go = torch.randn_like(out)
    out.backward(go)
    hd = h.detach().double().requires_grad_(True)
    loc = xyz.double().transpose(1, 2)
    bi = torch.arange(B, device=DEV).view(B, 1)
    ref = (torch.tanh(hd[bi.view(B, 1, 1), gidx] - hd[bi, fidx].unsqueeze(2))
           * (loc[bi.view(B, 1, 1), gidx] - loc[bi, fidx].unsqueeze(2))).mean(dim=2)
    ref.backward(go.double())
    assert_close(out.detach(), ref.detach(), 1e-5, "node_offset")
    assert_close(h.grad, hd.grad, 1e-5, "node_offset dh")

    nodes = (torch.rand(B, Sn, 3, device=DEV) * 2 - 1).requires_grad_(True)
    idx3 = S.ops.three_nn(xyz, nodes.detach().transpose(1, 2).contiguous(), 3)
    w = S.ops.interp_weights(xyz, nodes, idx3)
    gw = torch.randn_like(w)
    w.backward(gw)
    nd = nodes.detach().double().requires_grad_(True)
    nb = nd[bi.view(B, 1, 1), idx3.long()]
    d = -2 * (loc.unsqueeze(2) * nb).sum(-1) + (loc ** 2).sum(-1, keepdim=True) + (nb ** 2).sum(-1)
    d = torch.where(d < 1e-10, torch.full_like(d, 1e-10), d)
    wr = 1.0 / d
    wr = wr / wr.sum(-1, keepdim=True)
    wr.backward(gw.double())
    assert_close(w.detach(), wr.detach(), 1e-4, "interp weights")
    assert_close(nodes.grad, nd.grad, 2e-4, "interp weights dnodes")


def test_fused_adam_matches_torch(S):
    """optim.FusedAdam (one multi-tensor launch) against torch.optim.Adam: weight decay, a parameter
    without gradient, odd sizes, an LR change between steps."""
    from sug_b200.optim import FusedAdam
    torch.manual_seed(5)
    shapes = [(64, 6, 1, 1), (1024,), (513, 7), (3,), (256, 512), (10, 10)]
    ref = [torch.randn(sh, device=DEV).requires_grad_(True) for sh in shapes]
    mine = [p.detach().clone().requires_grad_(True) for p in ref]
    o_ref = torch.optim.Adam([{"params": ref[:3]}, {"params": ref[3:], "lr": 3e-3}], lr=1e-3, weight_decay=5e-4)
    o_mine = FusedAdam([{"params": mine[:3]}, {"params": mine[3:], "lr": 3e-3}], lr=1e-3, weight_decay=5e-4)
    for it in range(6):
        if it == 3:
            for o in (o_ref, o_mine):
                o.param_groups[0]["lr"] = 2.5e-4
        for k, (a, b) in enumerate(zip(ref, mine)):
            if k == 5:
                continue  # never receives a gradient: both optimizers must leave it alone
            g = torch.randn_like(a) * (10.0 ** (k - 2))
            a.grad, b.grad = g.clone(), g.clone()
        o_ref.step()
        o_mine.step()
        o_ref.zero_grad()
        o_mine.zero_grad()
    for k, (a, b) in enumerate(zip(ref, mine)):
        assert_close(b.detach(), a.detach(), 2e-6, f"fused adam param {k}")
    assert torch.equal(mine[5].detach(), ref[5].detach())
    assert float(o_mine.state[mine[0]]["step"]) == 6.0
    # the opt-in switch of the compatibility layer hands CUDA parameters to FusedAdam
    from sug_b200 import compat
    real = torch.optim.Adam
    try:
        compat.install(fused_adam=True)
        o = torch.optim.Adam([{"params": [mine[1]]}], lr=1e-3, weight_decay=5e-4)
        assert isinstance(o, FusedAdam)
        assert not isinstance(torch.optim.Adam([torch.nn.Parameter(torch.zeros(2))], lr=1e-3), FusedAdam)  # CPU tensor
    finally:
        torch.optim.Adam = real


def test_fused_adam_save_load_continue(S):
    """FusedAdam must survive ``load_state_dict`` -- moments AND step counter restored, device tables
    rebuilt -- and keep tracking torch.optim.Adam when training continues from the checkpoint."""
    import copy
    import io
    from sug_b200.optim import FusedAdam
    torch.manual_seed(11)
    shapes = [(64, 6), (1000,), (33, 5)]
    ref = [torch.randn(sh, device=DEV).requires_grad_(True) for sh in shapes]
    mine = [p.detach().clone().requires_grad_(True) for p in ref]
    mk_ref = lambda ps: torch.optim.Adam([{"params": ps[:2]}, {"params": ps[2:]}], lr=1e-2, weight_decay=5e-4)
    mk_mine = lambda ps: FusedAdam([{"params": ps[:2]}, {"params": ps[2:]}], lr=1e-2, weight_decay=5e-4)
    o_ref, o_mine = mk_ref(ref), mk_mine(mine)
    gen = torch.Generator(device=DEV).manual_seed(3)

    def steps(o_a, ps_a, o_b, ps_b, n):
        for _ in range(n):
            for a, b in zip(ps_a, ps_b):
                g = torch.randn(a.shape, device=DEV, generator=gen)
                a.grad, b.grad = g.clone(), g.clone()
            o_a.step()
            o_b.step()
    steps(o_ref, ref, o_mine, mine, 4)
    buf = io.BytesIO()
    torch.save(o_mine.state_dict(), buf)
    buf.seek(0)
    sd = torch.load(buf, map_location=DEV)
    assert float(sd["state"][0]["step"]) == 4.0
    # a NEW optimizer over NEW parameter tensors (a resumed process), and an in-place reload on the old one
    mine2 = [p.detach().clone().requires_grad_(True) for p in mine]
    o_new = mk_mine(mine2)
    o_new.load_state_dict(sd)
    o_mine.load_state_dict(copy.deepcopy(sd))
    assert float(o_new.state[mine2[0]]["step"]) == 4.0
    ref2 = [p.detach().clone().requires_grad_(True) for p in ref]
    o_ref2 = mk_ref(ref2)
    o_ref2.load_state_dict(copy.deepcopy(o_ref.state_dict()))
    gen_state = gen.get_state()
    steps(o_ref2, ref2, o_new, mine2, 3)
    gen.set_state(gen_state)
    steps(o_ref, ref, o_mine, mine, 3)
    for k in range(len(shapes)):
        assert_close(mine2[k].detach(), ref2[k].detach(), 2e-6, f"resumed (new optimizer) param {k}")
        assert_close(mine[k].detach(), ref[k].detach(), 2e-6, f"resumed (in place) param {k}")
    assert float(o_new.state[mine2[0]]["step"]) == 7.0
    assert float(o_new.state_dict()["state"][2]["step"]) == 7.0   # a later state_dict() saves the live counter


def test_graphed_step_matches_eager(S):
    """The CUDA-graph step (step.GraphedTrainStep) must train exactly like the eager step: same RNG
    consumption for FPS, same losses and weights after a few steps."""
    B = 12
    batches = []
    for i in range(2):
        d, l = O.synth_clouds(B, 1024, 10 + 2 * i)
        dt, lt = O.synth_clouds(B, 1024, 11 + 2 * i)
        batches.append(tuple(t.to(DEV) for t in (d, l, dt, lt)))

    def build():
        net = _load(S.Model.Net_MDA("DGCNN"), "Net_MDA:DGCNN").train()
        for hd in (net.c1, net.c2):
            hd.dropout1.p = hd.dropout2.p = 0.0
        opts = S.step.make_optimizers(net, capturable=True)
        crit = S.model_utils.focal_loss(num_classes=10, gamma=0.0, alpha=[0.1] * 10)
        return net, opts, crit

    # schedule: three warm-up steps on batch 0 (eager, inside warm()), then the capture (records the
    # graph, executes nothing, but draws one set of FPS starts), then three replays
    order = [0, 0, 0, 0, 1, 0]
    net_e, opts_e, crit_e = build()
    torch.manual_seed(5)
    losses_e = []
    for n, i in enumerate(order):
        if n == 3:
            for _ in range(4):  # the RNG draws of the capture pass
                torch.randint(0, 1024, (B,), dtype=torch.long)
        out = S.step.train_step(net_e, opts_e, *batches[i], crit_e)
        losses_e.append(float(out["loss"]))
    net_g, opts_g, crit_g = build()
    torch.manual_seed(5)
    gs = S.step.GraphedTrainStep(net_g, opts_g, crit_g, B, 1024, DEV)
    gs.warm(*batches[0], iters=3)
    gs.capture()
    losses_g = [float(gs(*batches[i])["loss"]) for i in order[3:]]
    losses_e_cmp = losses_e[3:]
    print("eager  :", [f"{v:.6f}" for v in losses_e])
    print("graphed:", [f"{v:.6f}" for v in losses_g])
    # not bit-identical: split-K / BatchNorm-sum atomics reorder fp32 additions from run to run, and
    # the model amplifies ulp-level differences; same trajectory within 2e-3
    for a, b in zip(losses_e_cmp, losses_g):
        assert abs(a - b) <= 2e-3 * abs(a), (losses_e, losses_g)
    pe, pg = dict(net_e.named_parameters()), dict(net_g.named_parameters())
    for k in ("g.conv1.conv.0.weight", "g.conv5.weight", "c1.mlp3.weight", "attention_s.bn.weight"):
        assert relerr(pg[k], pe[k]) < 1e-2, k


def test_shared_trunk_matches_four_full_forwards(S):
    """Opt-in DGCNN.share_trunk: the second pass on a batch reuses conv1 / conv2 of the first.  Losses,
    gradients, BatchNorm running statistics and counters must equal the four full forwards."""
    B = 12
    d, l = O.synth_clouds(B, 512, 30)
    dt, lt = O.synth_clouds(B, 512, 31)
    batch = tuple(t.to(DEV) for t in (d, l, dt, lt))
    res = []
    for share in (False, True):
        net = _load(S.Model.Net_MDA("DGCNN"), "Net_MDA:DGCNN").train()
        net.g.share_trunk = share
        crit = S.model_utils.focal_loss(num_classes=10, gamma=0.0, alpha=[0.1] * 10)
        torch.manual_seed(9)
        out = S.step.sug_losses(net, *batch, crit)
        out["loss"].backward()
        res.append((net, out))
    (n0, o0), (n1, o1) = res
    for key in ("loss", "loss_cls", "loss_geo", "loss_sem"):
        assert abs(float(o0[key]) - float(o1[key])) <= 2e-5 * max(1e-3, abs(float(o0[key]))), key
    g0 = {k: p.grad for k, p in n0.named_parameters() if p.grad is not None}
    g1 = {k: p.grad for k, p in n1.named_parameters() if p.grad is not None}
    assert g0.keys() == g1.keys()
    gmax = max(float(v.abs().max()) for v in g0.values())
    for k in g0:
        err = float((g0[k] - g1[k]).abs().max()) / max(float(g0[k].abs().max()), 1e-4 * gmax)
        assert err < 2e-3, (k, err)
    for name in ("conv1", "conv2", "conv3"):
        b0, b1 = getattr(n0.g, name).conv[1], getattr(n1.g, name).conv[1]
        assert_close(b1.running_mean, b0.running_mean, 1e-4, name + " running_mean")
        assert_close(b1.running_var, b0.running_var, 1e-4, name + " running_var")
        assert int(b1.num_batches_tracked) == int(b0.num_batches_tracked) == 4


def test_lidar_scale_inference(S):
    """BASELINE configs[4]: DGCNN inference at LiDAR scale.  N = 4096 against the CPU oracle (teacher
    forced), N = 16384 through properties only (the reference needs four 1.07 GB N x N tensors per
    cloud there; the tiled kNN never forms them)."""
    cls = _load(S.model_pointnet.DGCNN(), "DGCNN_cls", 667).eval()
    x, _ = O.synth_clouds(1, 4096, 77)
    ref, trace = oracle_trace(lambda: O.dgcnn_cls(x, O.synth_state("DGCNN_cls", 667), False))
    with torch.no_grad(), teacher_forced(S, trace):
        lg = cls(x.to(DEV))
    assert_close(lg, ref, 1e-3, "DGCNN_cls logits at N=4096")
    n_diff, wd, wn = knn_report(x.squeeze(-1), S.ops.knn_cm(x.squeeze(-1).to(DEV), 20), 20)
    print(f"kNN N=4096: {n_diff} rows differ, gap {wn:.1e}")
    assert wn < 1e-6
    torch.cuda.reset_peak_memory_stats()
    xb, _ = O.synth_clouds(1, 16384, 78)
    xg = xb.to(DEV)
    with torch.no_grad():
        out = cls(xg)
    assert out.shape == (1, 10) and bool(torch.isfinite(out).all())
    peak = torch.cuda.max_memory_allocated() / 2 ** 20
    print(f"N=16384 inference peak memory {peak:.0f} MiB (one N x N fp32 matrix alone would be 1024 MiB)")
    assert peak < 900
    idx = S.ops.knn_cm(xg.squeeze(-1), 20).long()
    assert bool((idx[..., 0] == torch.arange(16384, device=DEV)).all())
    # exactness on a slice of rows: brute force on the GPU for 256 query rows
    q = xg.squeeze(-1)[0].t()[:256]                     # [256,3]
    allp = xg.squeeze(-1)[0].t()                        # [N,3]
    D = -((q[:, None, :] - allp[None, :, :]) ** 2).sum(-1)
    kth = torch.gather(D, 1, idx[0, :256]).min(-1)[0]
    mask = torch.ones_like(D, dtype=torch.bool).scatter_(1, idx[0, :256], False)
    assert bool((torch.where(mask, D, torch.full_like(D, -1e30)).max(-1)[0] <= kth + 1e-6).all())


def test_graphed_eval_matches_eager(S):
    """step.GraphedEval: the eval forward as one CUDA graph returns what the eager forward returns --
    model_pointnet.DGCNN (no RNG) up to the order of the fp32 additions (the batch-sized head GEMMs split K over the
    SMs and reduce with atomics, so two runs of the SAME path differ in the last bit too), Net_MDA('DGCNN') with the
    FPS start fed from the device buffer under the same CPU RNG stream."""
    x = O.synth_clouds(2, 2048, 77)[0].to(DEV)
    net = S.model_pointnet.DGCNN().to(DEV).eval()
    with torch.no_grad():
        ref = net(x)
    fwd = S.step.GraphedEval(net, x)
    assert_close(fwd(x), ref, 1e-6, "graphed DGCNN logits")
    x2 = O.synth_clouds(2, 2048, 78)[0]
    with torch.no_grad():
        ref2 = net(x2.to(DEV))
    assert_close(fwd(x2.pin_memory()), ref2, 1e-6, "graphed DGCNN logits, second input")
    mda = _load(S.Model.Net_MDA("DGCNN"), "Net_MDA:DGCNN").eval()
    x3 = O.synth_clouds(3, 1024, 79)[0].to(DEV)
    g = S.step.GraphedEval(mda, x3, fps_points=1024)
    torch.manual_seed(9)
    a1, a2 = (t.clone() for t in g(x3))
    torch.manual_seed(9)
    with torch.no_grad():
        b1, b2 = mda(x3)
    assert_close(a1, b1, 1e-6, "graphed Net_MDA head 1")
    assert_close(a2, b2, 1e-6, "graphed Net_MDA head 2")
