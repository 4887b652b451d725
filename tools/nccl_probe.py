"""2-rank NCCL probe: all-reduce of a 1.1 MB fp32 bucket with nothing else in the process (isolates NCCL from our kernels)."""
import os, torch, torch.distributed as dist
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for n in (1, 1024, 279664, 4 << 20):
    x = torch.ones(n, device="cuda") * (rank + 1)
    dist.all_reduce(x)
    torch.cuda.synchronize()
    if rank == 0:
        print("n", n, "ok", float(x[0]), flush=True)
dist.destroy_process_group()
