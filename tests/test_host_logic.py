"""Host-side logic of the step that needs no GPU: the deferred BatchNorm side effects (ops.BNRecorder), which let the
four encoder passes of a step run concurrently (step.ConcurrentPasses), must reproduce nn.BatchNorm's running
statistics when the passes are applied afterwards in the reference's order (train_dg_single_gpu.py:260-264,
309-310: source pass, target pass, source-node pass, target-node pass)."""
import torch

from sug_b200 import ops, step


def _save_of(x, eps):
    """What the fused kernels leave in their `save` buffer: [mean | invstd] of the batch (biased variance)."""
    mean = x.mean(0)
    var = x.var(0, unbiased=False)
    return torch.cat([mean, (var + eps).rsqrt()])


def test_bn_recorder_reproduces_sequential_running_stats():
    torch.manual_seed(0)
    C = (7, 16, 5)
    mom = (0.1, 0.1, 0.03)
    bns = [torch.nn.BatchNorm1d(c, momentum=m) for c, m in zip(C, mom)]
    for bn in bns:  # non-trivial starting buffers
        bn.running_mean.normal_()
        bn.running_var.uniform_(0.5, 2.0)
    mine = [(bn.running_mean.clone(), bn.running_var.clone(), bn.num_batches_tracked.clone()) for bn in bns]
    rec = ops.BNRecorder()
    # four "passes", each through all layers; recorded OUT of order (the passes run concurrently) ...
    batches = {p: [torch.randn(33 + 5 * p, c) * (1 + p) + p for c in C] for p in range(4)}
    for p in (2, 0, 3, 1):
        rec.current = p
        for (rm, rv, nbt), bn, x in zip(mine, bns, batches[p]):
            rec.add(rm, rv, _save_of(x, bn.eps), x.shape[0], bn.momentum, bn.eps)
            rec.tick(nbt)
    # ... the reference feeds them in order
    for p in range(4):
        for bn, x in zip(bns, batches[p]):
            bn.train()
            bn(x)
    rec.apply()
    for (rm, rv, nbt), bn in zip(mine, bns):
        assert torch.allclose(rm, bn.running_mean, rtol=1e-5, atol=1e-6)
        assert torch.allclose(rv, bn.running_var, rtol=1e-5, atol=1e-6)
        assert int(nbt) == int(bn.num_batches_tracked) == 4
    assert rec.passes == {}  # a recorder is reusable after apply()


def test_bn_recorder_single_row_batch_keeps_biased_variance():
    """count == 1: nn.BatchNorm cannot train on it, the recorder must at least not divide by zero."""
    rec = ops.BNRecorder()
    rm, rv = torch.zeros(3), torch.ones(3)
    x = torch.randn(1, 3)
    rec.add(rm, rv, _save_of(x, 1e-5), 1, 0.1, 1e-5)
    rec.apply()
    assert torch.isfinite(rm).all() and torch.isfinite(rv).all()
    assert torch.allclose(rm, 0.1 * x[0], atol=1e-6)


def test_running_buffers_are_deferred_only_while_a_recorder_is_installed():
    rm, rv, save = torch.zeros(4), torch.ones(4), torch.zeros(8)
    assert ops.BN_RECORDER is None
    assert ops._running(rm, rv, save, 10, 0.1, 1e-5, True) == (rm, rv)       # the kernel updates them itself
    rec = ops.BNRecorder()
    ops.BN_RECORDER = rec
    try:
        assert ops._running(rm, rv, save, 10, 0.1, 1e-5, True) == (None, None)  # deferred
        assert ops._running(rm, rv, save, 10, 0.1, 1e-5, False) == (rm, rv)     # eval mode never records
        assert len(rec.passes[0]["bn"]) == 1
    finally:
        ops.BN_RECORDER = None


def test_concurrent_passes_are_a_no_op_without_cuda():
    """On CPU tensors (and in eval mode) the context runs the passes in place, installs no recorder and has no lanes."""
    model = torch.nn.Linear(2, 2)
    with step.ConcurrentPasses(model, torch.device("cpu")) as cp:
        assert not cp.on and ops.BN_RECORDER is None
        assert cp.run(0, lambda: 5) == 5
    assert cp.lanes(3) is None
    with cp.lane(None, 0):
        pass
    cp.merge(None, ())
    cp.join()


def _bench(args, env_extra, timeout=600):
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SUG_BENCH_CPU_BUDGET_S="1", **env_extra)
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), *args], cwd=root, env=env, capture_output=True,
                       text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    return [json.loads(l) for l in lines]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU oracle on the host cores) prints ONE JSON line with the GPU arm's metric,
    unit and config, `impl: reference`, a cpu_baseline describing itself and a zero-copy e2e; ranks other than 0 of a
    torchrun launch exit 0 without work and without output.  (The sample is bounded to 4+4 clouds here.)"""
    out = _bench(["--impl", "reference", "--steps", "1", "--warmup", "0"], {})
    assert len(out) == 1
    d = out[0]
    assert d["impl"] == "reference" and d["metric"] == "DGCNN SUG train clouds/sec" and d["unit"] == "clouds/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    assert d["config"]["workload"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["execution"]["sample_batch_per_subdomain"] == 4
    other = _bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                   {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert other == []
