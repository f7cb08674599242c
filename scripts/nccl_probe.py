"""Minimal NCCL sanity probe: init, all_reduce, barrier (run under torchrun)."""
import os, time, torch, torch.distributed as dist
t0 = time.time()
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
print(f"rank {rank} init {time.time()-t0:.1f}s", flush=True)
x = torch.ones(1 << 20, device="cuda") * (rank + 1)
dist.all_reduce(x)
torch.cuda.synchronize()
print(f"rank {rank} allreduce ok {float(x[0])} {time.time()-t0:.1f}s", flush=True)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    dist.all_reduce(x, op=dist.ReduceOp.AVG)
torch.cuda.current_stream().wait_stream(s)
dist.barrier()
torch.cuda.synchronize()
print(f"rank {rank} done {time.time()-t0:.1f}s", flush=True)
dist.destroy_process_group()
