#!/usr/bin/env python
"""Benchmark of the SUG hot path on B200: one "step" = one SUG domain-generalisation training step
(DGCNN backbone, class-weighted CE on two heads x two sub-domains, geometric + semantic MMD with SDA
weights, backward, three Adam updates) on B=64 source + 64 target synthetic PointDA-10-shaped clouds
per GPU — BASELINE.json configs[1] (and configs[2] for --gpus > 1, one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W]          this repository's CUDA path
    python bench.py --impl reference [...]                        the reference's algorithm on host cores

Prints ONE JSON line (rank 0).  `value` = clouds/s with inputs resident in HBM; `e2e` = the same
through the public API with HOST (pinned) inputs copied in and the loss read back every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 64
N_POINTS = 1024
METRIC = "DGCNN SUG train clouds/sec"
UNIT = "clouds/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="clouds per sub-domain per GPU")
    ap.add_argument("--mmd-scope", default="local", choices=["local", "global"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-reference", action="store_true",
                    help="skip the torch_gpu_reference block (the reference algorithm as PyTorch ops on this GPU)")
    ap.add_argument("--profile-all", action="store_true", help="time every kernel class (diagnostics)")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of as one CUDA graph")
    ap.add_argument("--nccl-in-graph", type=int, default=-1,
                    help="data parallel: 1 (default, -1) = capture the NCCL collectives into the step's CUDA graph (one graph "
                         "per step; needed for a graphed --mmd-scope global), 0 = graph A, eager coalesced all-reduce, graph B")
    ap.add_argument("--no-overlap", action="store_true",
                    help="data parallel with NCCL in the graph: one all-reduce after the backward instead of the grouped "
                         "all-reduce that overlaps the backward")
    ap.add_argument("--dp-no-comm", action="store_true",
                    help="DIAGNOSTIC ONLY (the line is marked invalid): N independent replicas without the gradient "
                         "all-reduce, to separate communication cost from multi-process interference")
    ap.add_argument("--no-serial-roofline", action="store_true",
                    help="skip timed region 3 (single-stream replay that times the dominant kernel class alone)")
    ap.add_argument("--share-trunk", action="store_true",
                    help="NOT the headline configuration: let the second encoder pass on a batch reuse conv1/conv2 of the "
                         "first (DGCNN.share_trunk); the default times the four full forwards the reference runs")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor": d["bf16_tflops_sustained"], "tensor_burst": d["bf16_tflops"],
                "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "src": "measured"}
    return {"hbm": 6650.0, "tensor": 1400.0, "tensor_burst": 1590.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
def workload_config(B, world, mmd_scope="local", shared_trunk=False):
    """The `config` object -- IDENTICAL for both arms (ours and --impl reference) given the same command line:
    BASELINE.json configs[1] (N = 1) / configs[2] (N > 1).  How an arm executes it is in the line's `execution` key."""
    return {"workload": "SUG DG train step: Net_MDA(DGCNN, k=20) x4 forwards + backward + 3 Adam, "
                        "CE(2 heads x 2 sub-domains) + GEO/SEM soft-MMD with SDA weights "
                        "(DG_unified_loss_onedataset_shapenet.yaml) = BASELINE.json configs[1]",
            "clouds_per_step_per_gpu": 2 * B, "batch_per_subdomain": B, "points": N_POINTS, "k": 20,
            "classes": 10, "parallelism": f"dp{world}", "mmd_scope": mmd_scope, "shared_trunk": bool(shared_trunk),
            "l2": "per-step working set (several GB of activations) exceeds the 126 MB L2; no flush needed"}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------------
class CpuStep:
    """One SUG step of the CPU oracle (oracle/sug_oracle.py: the reference's algorithm in functional PyTorch, pinned to
    the unmodified reference by tests/golden) at `batch` + `batch` clouds on all host threads."""

    def __init__(self, batch: int):
        from oracle import sug_oracle as O
        self.O = O
        torch.set_num_threads(os.cpu_count() or 1)
        self.batch = batch
        self.sd = O.clone_state(O.synth_state("Net_MDA:DGCNN"), requires_grad=True)
        self.opt = torch.optim.Adam([v for v in self.sd.values() if v.requires_grad], lr=1e-4, weight_decay=5e-4)
        self.data, label = O.synth_clouds(batch, N_POINTS, 0)
        self.data_t, self.label_t = O.synth_clouds(batch, N_POINTS, 1)
        self.label = label % batch if batch < 10 else label
        self.crit = O.FocalLoss([0.1] * 10, 0.0)
        self.threads = torch.get_num_threads()

    def __call__(self):
        self.crit.alpha = torch.full((10,), 0.1)  # keep the reference's re-gathered alpha valid for small batches
        t0 = time.perf_counter()
        out = self.O.sug_losses(self.sd, self.data, self.label, self.data_t, self.label_t, self.crit)
        out["loss"].backward()
        self.opt.step()
        self.opt.zero_grad()
        float(out["loss"].detach())
        return time.perf_counter() - t0


def host_mem_gib():
    try:
        import psutil
        return psutil.virtual_memory().available / 2 ** 30
    except Exception:
        return 16.0


def pick_cpu_batch(total_steps: int, budget_s: float):
    """Largest per-sub-domain batch in {64, 32, 16, 8, 4} whose `total_steps` steps fit the time budget and whose
    activations (~0.45 GiB per cloud pair: the reference materialises every [B,2C,N,k] edge tensor) fit the host."""
    probe = CpuStep(4)
    probe()
    per_cloud = min(probe(), probe()) / 8.0
    del probe
    mem = host_mem_gib()
    for bs in (64, 32, 16, 8, 4):
        if total_steps * per_cloud * 2 * bs <= budget_s and 0.5 * bs + 6.0 <= mem:
            return bs, per_cloud
    return 4, per_cloud


def run_reference(args, rank, emit):
    """--impl reference: the reference's own CPU path (there is no compiled reference: it is Python / PyTorch, and
    /root/reference does not travel to the GPU box, so the arm times the oracle port), all host threads, EXACTLY
    `--warmup` untimed and `--steps` timed steps of the GPU arm's workload.  Each step is a bounded sample: the
    same SUG step at the largest batch that keeps the whole run within ~3 minutes (stated in cpu_baseline.sample);
    clouds/s = clouds of the sample / its time."""
    if rank != 0:
        return
    # SUG_BENCH_CPU_BUDGET_S: the CPU test-suite bounds the sample further (default: ~3 minutes for the whole run)
    bs, _ = pick_cpu_batch(args.steps + args.warmup, float(os.environ.get("SUG_BENCH_CPU_BUDGET_S", "170")))
    step = CpuStep(bs)
    for _ in range(args.warmup):
        step()
    ts = [step() for _ in range(args.steps)]
    sec = sum(ts) / len(ts)
    val = 2 * bs / sec
    cfg = workload_config(args.batch, max(1, args.gpus), args.mmd_scope, args.share_trunk)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "execution": {"mode": "cpu_eager", "host_threads": step.threads, "sample_batch_per_subdomain": bs,
                          "note": "one process on rank 0's host cores whatever --gpus says"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": step.threads, "kind": "port",
                             "sample": f"{args.steps} timed + {args.warmup} warm-up SUG steps at {bs}+{bs} clouds x {N_POINTS} pts "
                                       f"(a bounded sample of the 64+64 step: same network, losses, backward, Adam), "
                                       f"oracle/sug_oracle.py on {step.threads} host threads; best step {2 * bs / min(ts):.1f} clouds/s"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def finish(world):
    """End of a multi-rank run.  CUDA graphs that hold captured NCCL kernels keep the communicator busy in
    ProcessGroupNCCL's eyes, and destroy_process_group() then waits forever (measured: the process had to be killed);
    the JSON line is already on the real stdout (os.write), so the ranks leave without the collective teardown."""
    if world > 1:
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def torch_gpu_reference(dev, B, steps=3):
    """The reference ALGORITHM on this GPU: the oracle's plain PyTorch ops (torch.topk on the materialised N x N
    matrix, the [B,2C,N,k] edge tensor, cuDNN / cuBLAS convolutions, Python FPS loop with its host syncs) moved to the
    device -- the kernels the reference's own modules would launch.  This is the denominator of north_star's
    ">= 3x the reference's PyTorch-CUDA path"; timed with cuDNN TF32 on (PyTorch's default, what the reference gets)
    and off (fp32, the accuracy this library delivers).  Checker code (oracle/), never on the product path."""
    from oracle import sug_oracle as O
    sd = {k: v.to(dev) for k, v in O.clone_state(O.synth_state("Net_MDA:DGCNN")).items()}
    for k, v in sd.items():
        if v.is_floating_point() and v.dim() > 0 and "running_" not in k:
            v.requires_grad_(True)
    opt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=1e-4, weight_decay=5e-4)
    data, label = (t.to(dev) for t in O.synth_clouds(B, N_POINTS, 0))
    data_t, label_t = (t.to(dev) for t in O.synth_clouds(B, N_POINTS, 1))
    crit = O.FocalLoss([0.1] * 10, 0.0)
    out = {}
    saved = torch.backends.cudnn.allow_tf32
    try:
        for name, tf32 in (("tf32_default", True), ("fp32", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            ts = []
            for it in range(steps + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                o = O.sug_losses(sd, data, label, data_t, label_t, crit)
                o["loss"].backward()
                opt.step()
                opt.zero_grad()
                float(o["loss"].detach())
                e1.record()
                torch.cuda.synchronize()
                if it >= 1:
                    ts.append(e0.elapsed_time(e1))
            ms = sum(ts) / len(ts)
            out[name] = {"ms_per_step": ms, "clouds_per_s": 2 * B / ms * 1e3}
    finally:
        torch.backends.cudnn.allow_tf32 = saved
    out["peak_mem_gib"] = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    out["what"] = (f"oracle/sug_oracle.py (the reference's algorithm as plain PyTorch ops) on cuda, {B}+{B} clouds, "
                   f"1 warm-up + {steps} timed steps per mode, CUDA events")
    return out


def measure_tf32_peak(dev):
    """Dense TF32 matmul throughput of this GPU the way MEASURED_PEAKS.json measures bf16 (torch.matmul 8192^3, best
    of 10): the tensor-pipe denominator of the fp32-accurate kernels (3xTF32 => a third of it)."""
    n = 8192
    a = torch.randn(n, n, device=dev)
    b = torch.randn(n, n, device=dev)
    saved = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        best = 1e9
        for it in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            if it >= 2:
                best = min(best, e0.elapsed_time(e1))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = saved
    del a, b
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    # stdout carries exactly one JSON line: anything libraries print (e.g. NCCL's version banner) goes
    # to stderr instead
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    try:
        _main(args, rank, emit)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)


def _main(args, rank, emit):
    if args.impl == "reference":
        run_reference(args, rank, emit)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")

    from sug_b200 import Model, _lib, dist as sdist, model_utils, step, synth
    import torch.distributed as tdist
    trace_on = os.environ.get("SUG_BENCH_TRACE") == "1"

    def trace(msg):
        if trace_on:
            print(f"[bench rank {os.environ.get('RANK', '0')}] {time.strftime('%H:%M:%S')} {msg}", file=sys.stderr, flush=True)
    trace("init process group")
    rank, world, local = sdist.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B = args.batch

    torch.manual_seed(666)  # train_dg_single_gpu.py:65
    model = Model.Net_MDA("DGCNN").to(dev).train()
    model.g.share_trunk = bool(args.share_trunk)
    opts = step.make_optimizers(model, capturable=not args.no_graph)
    crit = model_utils.focal_loss(num_classes=10, gamma=0.0, alpha=[0.1] * 10)  # ClassWeighting / DLSA, uniform counts
    mmd_fn = sdist.global_mmd_cal if (args.mmd_scope == "global" and world > 1) else None
    hook = sdist.allreduce_grads if (world > 1 and not args.dp_no_comm) else None

    # host batches (pinned) -- a small pool so that every step copies fresh memory
    pool = []
    for i in range(4):
        d, l = synth.synth_clouds(B, N_POINTS, 1000 * rank + 2 * i)
        dt, lt = synth.synth_clouds(B, N_POINTS, 1000 * rank + 2 * i + 1)
        pool.append(tuple(t.pin_memory() for t in (d, l, dt, lt)))
    dev_batches = [tuple(t.to(dev) for t in hb) for hb in pool]
    h2d_bytes = sum(t.numel() * t.element_size() for t in pool[0])

    def run_step(d, l, dt, lt):
        if mmd_fn is None:
            return step.train_step(model, opts, d, l, dt, lt, crit, grad_hook=hook)
        out = step.sug_losses(model, d, l, dt, lt, crit, mmd_fn=mmd_fn)
        out["loss"].backward()
        if hook:
            hook(model)
        for o in opts:
            o.step()
        for o in opts:
            o.zero_grad()
        return out

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (untimed) + one instrumented step to find the dominant kernel class --------------
    trace("eager warm-up")
    for i in range(args.warmup):
        run_step(*dev_batches[i % len(dev_batches)])
    prof1 = None
    for i in range(2):  # two instrumented steps, per-class minimum: one-off stalls must not pick the class
        _lib.prof_reset(mask=0xFFFFFFFF)
        run_step(*dev_batches[i % len(dev_batches)])
        pr = _lib.prof_collect()
        if prof1 is None:
            prof1 = pr
        else:
            for k, v in pr.items():
                if v["launches"] and v["ms"] < prof1[k]["ms"]:
                    prof1[k] = v
    dom = max(prof1.items(), key=lambda kv: kv[1]["ms"])[0]
    breakdown = {k: round(v["ms"], 4) for k, v in prof1.items() if v["launches"]}

    # ---- the step as one CUDA graph (public API: step.GraphedTrainStep); eager fallback ---------------
    prof_mask = 0xFFFFFFFF if args.profile_all else (1 << prof1[dom]["index"])
    _lib.prof_reset(mask=prof_mask)
    graphed, mode = None, "eager"
    trace(f"dominant class {dom}; building the graphed step")
    nccl_in_graph = True if args.nccl_in_graph < 0 else bool(args.nccl_in_graph)
    nccl_in_graph = nccl_in_graph and world > 1
    if not args.no_graph and (mmd_fn is None or nccl_in_graph):  # a collective inside the forward needs NCCL capture
        try:
            for o in opts:
                o.zero_grad(set_to_none=True)
            graphed = step.GraphedTrainStep(model, opts, crit, B, N_POINTS, dev, mmd_fn=mmd_fn, grad_hook=hook,
                                            nccl_in_graph=nccl_in_graph, overlap=not args.no_overlap)
            graphed.warm(*dev_batches[0])
            _lib.prof_reset(mask=prof_mask)  # count / time exactly the launches recorded into the graph
            graphed.capture()
            mode = "cuda_graph"
        except Exception as e:  # keep measuring: the eager path is the same kernels
            print(f"[bench] CUDA-graph capture failed ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
            graphed = None
            _lib.prof_reset(mask=prof_mask)
    launches_per_capture = int(sum(v["launches"] for v in _lib.prof_collect().values())) if graphed else 0
    if graphed is not None:
        run_step = graphed
        for i in range(3):
            run_step(*dev_batches[i % len(dev_batches)])

    # ---- timed region 1: inputs resident in HBM ------------------------------------------------------
    trace(f"mode {mode}; timed region 1")
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for i in range(args.steps):
            run_step(*dev_batches[i % len(dev_batches)])
        e1.record()
        barrier()
        if not clk.lines:
            time.sleep(0.25)  # very short timed regions: make sure nvidia-smi emitted at least one sample
    ms = e0.elapsed_time(e1)
    prof = _lib.prof_collect()
    launches = launches_per_capture * args.steps if graphed else int(sum(v["launches"] for v in prof.values()))
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    ms = float(t.item())
    clouds_per_step = 2 * B * world
    value = clouds_per_step * args.steps / (ms / 1e3)

    # ---- timed region 2: end to end from pinned host memory, loss read back every step --------------
    if graphed is None:
        _lib.prof_reset(mask=0)
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()
    trace("timed region 2 (e2e)")
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for i in range(args.steps):
        hb = pool[i % len(pool)]
        batch = hb if graphed is not None else tuple(t.to(dev, non_blocking=True) for t in hb)
        out = run_step(*batch)
        host_loss.copy_(out["loss"].detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        float(host_loss)
    e3.record()
    barrier()
    t = torch.tensor([e2.elapsed_time(e3)], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    e2e_value = clouds_per_step * args.steps / (ms_e2e / 1e3)

    # ---- timed region 3 (1 GPU): the same graphed step on ONE stream, for the roofline of the dominant kernel ----------
    # In regions 1 / 2 the four encoder passes share the SMs (up to four streams inside the graph), so an event pair
    # around a launch measures the launch PLUS whatever the other streams ran in its window.  The roofline wants the
    # kernel's own duration: replay the step once more with the concurrency switched off and time the launches there.
    prof_serial, ms_serial = None, None
    if graphed is not None and world == 1 and step.ENABLE_CONCURRENT_PASSES and not args.no_serial_roofline:
        trace("timed region 3 (single-stream replay for the roofline)")
        try:
            step.ENABLE_CONCURRENT_PASSES = False
            fps_flag, model.g.overlap_fps = model.g.overlap_fps, False
            for o in opts:
                o.zero_grad(set_to_none=True)
            serial = step.GraphedTrainStep(model, opts, crit, B, N_POINTS, dev, mmd_fn=mmd_fn, grad_hook=hook)
            serial.warm(*dev_batches[0])
            _lib.prof_reset(mask=prof_mask)
            serial.capture()
            for i in range(3):
                serial(*dev_batches[i % len(dev_batches)])
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(args.steps):
                serial(*dev_batches[i % len(dev_batches)])
            s1.record()
            barrier()
            ms_serial = s0.elapsed_time(s1) / args.steps
            prof_serial = _lib.prof_collect()
            del serial
        except Exception as e:
            print(f"[bench] single-stream replay failed ({type(e).__name__}: {e}); roofline from the concurrent replay",
                  file=sys.stderr)
            prof_serial = None
        finally:
            step.ENABLE_CONCURRENT_PASSES = True
            model.g.overlap_fps = fps_flag

    trace("done timing")
    # ---- data parallel: every replica must hold the same weights after the timed steps ---------------------
    replicas_in_sync = None
    if world > 1:
        with torch.no_grad():
            cs = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
            cs2 = torch.stack([p.detach().double().abs().sum() for p in model.parameters()]).sum().reshape(1)
            v = torch.cat([cs, cs2])
            lo, hi = v.clone(), v.clone()
            tdist.all_reduce(lo, op=tdist.ReduceOp.MIN)
            tdist.all_reduce(hi, op=tdist.ReduceOp.MAX)
            replicas_in_sync = bool(torch.equal(lo, hi))
    if rank != 0:
        finish(world)
        return

    # ---- roofline of the dominant kernel class --------------------------------------------------------
    pk = peaks()
    tf32_peak = measure_tf32_peak(dev)
    d = prof[dom]
    if graphed is not None:
        # graph replays re-record the same event pairs: the totals are those of the LAST timed step
        d = dict(d, launches=d["timed"], flops=d["flops"], bytes=d["bytes"])
    per_launch_s = (d["ms"] / 1e3) / max(1, d["timed"])
    step_ms_for_share = ms / args.steps if graphed is not None else ms
    concurrent = None
    if prof_serial is not None and prof_serial[dom]["timed"]:
        # the dominant class as the concurrent replay saw it (kept for the record), then switch to the single-stream one
        concurrent = {"avg_launch_us": per_launch_s * 1e6, "launches_timed": int(d["timed"]),
                      "sum_of_launch_windows_over_step": d["ms"] / step_ms_for_share,
                      "what": "event pairs around the same launches in timed region 1: four streams share the SMs, a "
                              "window holds other streams' work too"}
        ds = prof_serial[dom]
        d = dict(ds, launches=ds["timed"])
        per_launch_s = (d["ms"] / 1e3) / max(1, d["timed"])
        step_ms_for_share = ms_serial

    # Every tensor-core kernel of this library is fp32-accurate 3xTF32 (hi/lo split, three kind::tf32 products per
    # algorithmic product: the parity gate is fp32), so its tensor roofline is a third of the dense TF32 rate -- measured
    # above with the method MEASURED_PEAKS.json uses for bf16 (for comparison: bf16 sustained / 6 = the same ceiling
    # derived from the driver-written file).  FLOPs are counted once (algorithmic), never three times.
    tensor_ceiling = tf32_peak / 3.0

    TENSOR_CORE = ("gemm_tc", "knn_tc")          # tcgen05 kernels
    SIMT_MATH = ("gemm_simt", "knn_simt")        # CUDA-core distance / GEMM kernels: fp32 pipe, 2 x 128 lanes x SMs x clock
    fp32_simt_peak = 2.0 * 128 * 148 * (pk.get("sm_max_mhz", 1965.0) * 1e6) / 1e12

    def judge(name, flops_pl, bytes_pl, sec):
        """A kernel is judged against whichever roofline binds it harder at the measured peaks: HBM copy bandwidth,
        the 3xTF32 tensor ceiling (tcgen05 kernels) or -- by-class block only -- the nominal fp32 CUDA-core rate."""
        t_hbm = bytes_pl / (pk["hbm"] * 1e9)
        tfl, gbs = flops_pl / sec / 1e12, bytes_pl / sec / 1e9
        if name in TENSOR_CORE and flops_pl / (tensor_ceiling * 1e12) > t_hbm:
            r = {"bound": "tensor", "achieved": tfl, "peak": tensor_ceiling, "unit": "TFLOP/s", "frac": tfl / tensor_ceiling,
                 "peak_source": "measured here: dense TF32 torch.matmul / 3 (fp32-accurate 3xTF32 split); "
                                f"MEASURED_PEAKS bf16 sustained / 6 = {pk['tensor'] / 6.0:.0f}"}
        elif name in SIMT_MATH and flops_pl / (fp32_simt_peak * 1e12) > t_hbm:
            r = {"bound": "fp32_simt", "achieved": tfl, "peak": fp32_simt_peak, "unit": "TFLOP/s", "frac": tfl / fp32_simt_peak,
                 "peak_source": "nominal: 148 SMs x 128 lanes x 2 flop x max SM clock (not a measured peak)"}
        else:
            r = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                 "peak_source": pk["src"] + " copy"}
        r.update({"tflops": tfl, "gbps": gbs, "frac_of_hbm": gbs / pk["hbm"]})
        if name in TENSOR_CORE:
            r["frac_of_3xtf32_ceiling"] = tfl / tensor_ceiling
            r["frac_of_bf16_sustained"] = tfl / pk["tensor"]
        return r
    flops_pl, bytes_pl = d["flops"] / max(1, d["launches"]), d["bytes"] / max(1, d["launches"])
    roof = judge(dom, flops_pl, bytes_pl, per_launch_s)
    # DRAM traffic of the class's representative launch from the committed ncu --set full capture (profiles/)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        t = json.load(open(tp)).get(dom)
        if t:
            traffic = t["dram_bytes_per_launch"]
            roof["traffic_capture"] = {k: t[k] for k in t if k != "dram_bytes_per_launch"}
    roof["traffic"] = traffic
    roof.update({"kernel": dom, "launches_timed": int(d["timed"]), "avg_launch_us": per_launch_s * 1e6,
                 "algorithmic_bytes_per_launch": bytes_pl, "algorithmic_flops_per_launch": flops_pl,
                 "share_of_step": d["ms"] / step_ms_for_share})
    if concurrent is not None:
        roof["timed_in"] = ("single-stream replay of the same graphed step (timed region 3, %.3f ms per step): CUDA events "
                            "on the launch stream around every launch of the class" % ms_serial)
        roof["concurrent_replay"] = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in concurrent.items()}
    else:
        roof["timed_in"] = "timed region 1: CUDA events on the launch stream around every launch of the class"

    # every kernel class of the step (the metric names the kNN / EdgeConv kernels, which are not the largest class):
    # launches, average device time and algorithmic work per launch from the instrumented eager step that precedes
    # the timed region (same kernels, same shapes; the per-class minimum of two steps)
    by_class = {}
    for name, v in prof1.items():
        if not v["launches"] or not v["timed"] or v["ms"] <= 0:
            continue
        sec = v["ms"] / 1e3 / v["timed"]
        r = judge(name, v["flops"] / v["launches"], v["bytes"] / v["launches"], sec)
        r.update({"launches_per_step": int(v["launches"]), "avg_launch_us": sec * 1e6, "ms_per_step": v["ms"]})
        by_class[name] = {k: (round(x, 4) if isinstance(x, float) else x) for k, x in r.items()}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(B, world, args.mmd_scope, args.share_trunk),
            "execution": {"mode": mode, "nccl_in_graph": bool(nccl_in_graph), "grad_allreduce_overlap": bool(nccl_in_graph and not args.no_overlap)},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "roofline": roof, "roofline_by_class": by_class,
            "peaks": {"hbm_gbs": pk["hbm"], "bf16_tflops_sustained": pk["tensor"], "bf16_tflops_burst": pk["tensor_burst"],
                      "source": pk["src"], "tf32_tflops_measured_here": tf32_peak,
                      "tf32_how": "torch.matmul fp32 8192^3 with allow_tf32, best of 10, CUDA events (same method as "
                                  "MEASURED_PEAKS.json's bf16 burst figure)"},
            "kernel_ms_one_step": breakdown}
    if replicas_in_sync is not None:
        line["replicas_in_sync"] = replicas_in_sync
    if args.dp_no_comm:
        line["diagnostic"] = "no gradient all-reduce: NOT a valid data-parallel number"

    if world == 1 and not args.no_torch_reference:
        # free this arm's graph pools first: the reference algorithm keeps ~50 GB of activations at 64+64
        del graphed
        run_step = None
        torch.cuda.empty_cache()
        try:
            ref = torch_gpu_reference(dev, B)
            ref["speedup_of_this_library"] = {"vs_tf32_default": ref["tf32_default"]["ms_per_step"] / (ms / args.steps),
                                              "vs_fp32": ref["fp32"]["ms_per_step"] / (ms / args.steps)}
            line["torch_gpu_reference"] = ref
        except Exception as e:  # never lose the measured line over the comparison arm
            line["torch_gpu_reference"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}

    if world == 1 and not args.no_cpu_baseline:  # rank 0 at N = 1 only (the contract); N > 1 lines carry none
        # the reference's CPU path on this box's host cores: BASELINE.json configs[0]-sized sample of the same step
        # (32+32 clouds when the host has the memory), one warm-up step, best of two
        bs = 32 if host_mem_gib() >= 24.0 else (16 if host_mem_gib() >= 14.0 else 8)
        step_cpu = CpuStep(bs)
        step_cpu()
        sec = min(step_cpu(), step_cpu())
        line["cpu_baseline"] = {"value": 2 * bs / sec, "unit": UNIT, "cores": step_cpu.threads, "kind": "port",
                                "sample": f"SUG step at {bs}+{bs} clouds x {N_POINTS} pts (sample of the 64+64 step), "
                                          f"1 warm-up + best of 2, oracle/sug_oracle.py on {step_cpu.threads} host threads"}
    emit(line)
    finish(world)


if __name__ == "__main__":
    main()
