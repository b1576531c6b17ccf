"""Data-parallel numerics (reference intent train_dg.py:216-217,357-368): an N-rank SUG step must equal a
SINGLE-PROCESS step on the concatenated global batch with per-shard BatchNorm.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tools/dp_parity.py [--batch 16] [--points 1024]

For both MMD scopes:
  * local  (the reference's DDP intent: every rank evaluates the MMD on its own shard): the averaged gradient must equal
    the mean over the shards of the single-process gradients;
  * global (north_star: the sub-domain features of all ranks are all-gathered, m = N * B): the averaged gradient must equal
    the gradient of   mean_r cls_r + MMD(all shards' features)   evaluated in one process.
Rank r draws its FPS start indices from torch.manual_seed(1000 + r); the single-process run reseeds the same way before
each shard, so both sides consume the RNG identically.  Dropout is switched off (its CUDA RNG stream cannot be matched).
After the comparison of the gradients the optimizers step once and the updated weights are compared as well, and all
ranks must hold bit-identical weights.  Prints one JSON line per scope on rank 0; exit code 1 on a mismatch."""
import argparse
import copy
import json
import os
import sys

import torch
import torch.distributed as tdist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sug_b200 import Model, dist as sdist, mmd, model_utils, step, synth  # noqa: E402


def relerr(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--points", type=int, default=1024)
    ap.add_argument("--tol", type=float, default=1e-3)
    args = ap.parse_args()
    rank, world, local = sdist.init_from_env()
    assert world > 1, "run under torchrun with at least 2 ranks"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    B, N = args.batch, args.points

    def shard(r):
        d, l = synth.synth_clouds(B, N, 100 + 2 * r)
        dt, lt = synth.synth_clouds(B, N, 101 + 2 * r)
        return tuple(t.to(dev) for t in (d, l, dt, lt))

    def fresh():
        torch.manual_seed(666)
        net = Model.Net_MDA("DGCNN").to(dev).train()
        for hd in (net.c1, net.c2):
            hd.dropout1.p = hd.dropout2.p = 0.0
        return net

    def crit():
        return model_utils.focal_loss(num_classes=10, gamma=0.0, alpha=[0.1] * 10)

    ok = True
    for scope in ("local", "global"):
        # ---------------- N ranks ----------------
        net = fresh()
        opts = step.make_optimizers(net)
        torch.manual_seed(1000 + rank)
        out = step.sug_losses(net, *shard(rank), crit(), mmd_fn=sdist.global_mmd_cal if scope == "global" else mmd.mmd_cal)
        out["loss"].backward()
        sdist.allreduce_grads(net)
        g_dp = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
        for o in opts:
            o.step()
        w_dp = {k: p.detach().clone() for k, p in net.named_parameters()}
        # all replicas identical?
        flat = torch.cat([p.reshape(-1) for p in w_dp.values()])
        lo, hi = flat.clone(), flat.clone()
        tdist.all_reduce(lo, op=tdist.ReduceOp.MIN)
        tdist.all_reduce(hi, op=tdist.ReduceOp.MAX)
        in_sync = bool(torch.equal(lo, hi))
        loss_dp = out["loss"].detach().clone()
        tdist.all_reduce(loss_dp)
        loss_dp /= world

        # ---------------- one process, concatenated batch, per-shard BatchNorm ----------------
        ref = fresh()
        ropts = step.make_optimizers(ref)
        if scope == "local":
            total = 0.0
            for r in range(world):
                torch.manual_seed(1000 + r)
                o = step.sug_losses(ref, *shard(r), crit())
                (o["loss"] / world).backward()
                total = total + o["loss"].detach() / world
        else:
            cfg = step.SUG_CFG
            feats = []
            cls = 0.0
            for r in range(world):
                d, l, dt, lt = shard(r)
                torch.manual_seed(1000 + r)
                c = crit()
                ps1, ps2, ss1, ss2 = ref(d, semantic_adaption=True)
                pt1, pt2, st1, st2 = ref(dt, semantic_adaption=True)
                ls = 0.5 * c(ps1, l) + 0.5 * c(ps2, l)
                ltg = 0.5 * c(pt1, l) + 0.5 * c(pt2, l)
                cls = cls + cfg["CLS_WEIGHT"] * (0.5 * ls + 0.5 * ltg) / world
                fs = ref(d, node_adaptation_s=True)
                ft = ref(dt, node_adaptation_t=True)
                feats.append((l, lt, fs, ft, ss1, st1, ss2, st2, ps1, pt1, ps2, pt2, d, dt))
            cat = lambda i: torch.cat([f[i] for f in feats], 0)  # noqa: E731
            geo, sem = cfg["GEO_MMD"][0], cfg["SEM_MMD"][0]
            L, LT = cat(0), cat(1)
            lg = cfg["MMD_WEIGHT"] * geo["GEO_SCALE"] * mmd.mmd_cal(L, cat(2), LT, cat(3), geo, data_s=cat(12), data_t=cat(13))
            l1 = sem["SEM_SCALE"] * mmd.mmd_cal(L, cat(4), LT, cat(5), sem, data_s=cat(8).detach(), data_t=cat(9).detach())
            l2 = sem["SEM_SCALE"] * mmd.mmd_cal(L, cat(6), LT, cat(7), sem, data_s=cat(10).detach(), data_t=cat(11).detach())
            total = cls + lg + cfg["MMD_WEIGHT"] * (0.5 * l1 + 0.5 * l2)
            total.backward()
            total = total.detach()
        g_ref = {k: p.grad.detach().clone() for k, p in ref.named_parameters() if p.grad is not None}
        for o in ropts:
            o.step()
        w_ref = {k: p.detach().clone() for k, p in ref.named_parameters()}

        assert set(g_dp) == set(g_ref), set(g_dp) ^ set(g_ref)
        gmax = max(float(v.norm()) for v in g_ref.values())
        gerr = sorted(((float((g_dp[k].double() - g_ref[k].double()).norm()) / max(float(g_ref[k].norm()), 1e-4 * gmax), k)
                       for k in g_ref), reverse=True)
        werr = max(relerr(w_dp[k], w_ref[k]) for k in w_ref)
        lerr = abs(float(loss_dp) - float(total)) / abs(float(total))
        rec = {"scope": scope, "world": world, "batch_per_rank": B, "points": N, "loss_dp_mean": float(loss_dp),
               "loss_single_process": float(total), "loss_rel_err": lerr, "grad_worst_rel_err": gerr[0][0],
               "grad_worst_param": gerr[0][1], "n_grads": len(gerr), "weights_after_step_max_rel_err": werr,
               "replicas_in_sync": in_sync, "tol": args.tol}
        good = lerr <= args.tol and gerr[0][0] <= args.tol and werr <= args.tol and in_sync
        rec["ok"] = bool(good)
        ok = ok and good
        if rank == 0:
            print(json.dumps(rec), flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    tdist.all_reduce(flag)
    tdist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
