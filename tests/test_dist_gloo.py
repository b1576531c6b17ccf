"""Data-parallel plumbing on CPU: world size 2 over gloo (the GPU box uses NCCL with the same code)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from sug_b200 import dist as sdist
    r, w, _ = sdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    unused = torch.nn.Linear(2, 2)  # never receives a gradient, like g.input_transform_net in the reference
    model = torch.nn.ModuleList([net, unused])
    x = torch.arange(8, dtype=torch.float32).view(2, 4) + 10 * rank
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    sdist.allreduce_grads(model)
    assert all(p.grad is None for p in unused.parameters())
    # expected: mean over ranks of the per-rank gradients
    exp = []
    for rr in range(world):
        n2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
        n2.load_state_dict(net.state_dict())
        xx = torch.arange(8, dtype=torch.float32).view(2, 4) + 10 * rr
        n2(xx).sum().backward()
        exp.append([p.grad for p in n2.parameters()])
    for i, p in enumerate(net.parameters()):
        want = sum(e[i] for e in exp) / world
        assert torch.allclose(p.grad, want, atol=1e-6), (rank, i)
    # autograd-aware all_gather: every rank evaluates the same global loss
    f = (torch.arange(6, dtype=torch.float32).view(3, 2) + rank).requires_grad_(True)
    g = sdist.all_gather_rows(f)
    assert g.shape == (3 * world, 2)
    wts = torch.arange(1, 3 * world + 1, dtype=torch.float32).view(-1, 1)
    (g * wts).sum().backward()
    # d(global loss)/d(local rows) * world (undone by the gradient average)
    assert torch.allclose(f.grad, wts[3 * rank:3 * rank + 3].expand(3, 2) * world)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, "ok"))


def test_allreduce_and_allgather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(timeout=180)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "ok"), (1, "ok")]
