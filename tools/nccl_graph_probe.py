"""Does an NCCL collective captured inside a CUDA graph replay correctly here?  torchrun --nproc-per-node 2 tools/nccl_graph_probe.py"""
import os, sys, time
import torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
x = torch.ones(10 * 1024 * 1024, device=dev) * (rank + 1)
y = torch.ones(64, 4096, device=dev) * (rank + 1)
out = torch.empty(world * 64, 4096, device=dev)
# warm the communicator and the collectives eagerly on a side stream, like any graph capture
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        dist.all_reduce(x)
        dist.all_gather_into_tensor(out, y)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
dist.barrier()
x.fill_(rank + 1)
g = torch.cuda.CUDAGraph()
print(f"[{rank}] capturing", flush=True)
with torch.cuda.graph(g, capture_error_mode=os.environ.get('CAPTURE_MODE', 'thread_local')):
    z = x * 2
    dist.all_reduce(z)
    dist.all_gather_into_tensor(out, y)
    w = z + out.sum()
torch.cuda.synchronize()
print(f"[{rank}] captured", flush=True)
t0 = time.time()
for i in range(200):
    g.replay()
torch.cuda.synchronize()
exp = 2 * sum(r + 1 for r in range(world))
print(f"[{rank}] replayed 200x in {time.time() - t0:.3f}s; z[0]={float(z[0])} expected {exp}; gather ok {bool((out[64 * (world - 1)] == world).all())}", flush=True)
dist.barrier()
dist.destroy_process_group()
