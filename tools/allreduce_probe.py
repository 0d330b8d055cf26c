"""Latency of the per-step gradient all-reduce (flat vector of the model's size) under torch.distributed/NCCL."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for n in (2226, 8835, 331779):
    x = torch.ones(n, device="cuda")
    for _ in range(20): dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(200): dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0: print(f"all_reduce {n} floats x200: device {e0.elapsed_time(e1)/200*1e3:.1f} us/op, host wall {(t1-t0)/200*1e6:.1f} us/op")
    # with a dependent tiny kernel before/after (as in the training step)
    y = torch.zeros(n, device="cuda")
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(200):
        y.add_(1.0); dist.all_reduce(y); y.mul_(0.5)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"  kernel + all_reduce + kernel: device {e0.elapsed_time(e1)/200*1e3:.1f} us/iter")
dist.destroy_process_group()
