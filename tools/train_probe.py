"""Debug probe: one training step of the bench configuration with / without NCCL initialised in the process."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nylon_amt_b200 as hft
mode = sys.argv[1]
lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if mode != "none":
    import torch.distributed as dist
    if mode == "nccl_eager":
        dist.init_process_group("nccl", device_id=dev)
    elif mode == "nccl_lazy":
        dist.init_process_group("nccl")
    elif mode == "nccl_used":
        dist.init_process_group("nccl", device_id=dev)
        x = torch.ones(1000, device=dev); dist.all_reduce(x); torch.cuda.synchronize()
    elif mode == "gloo":
        dist.init_process_group("gloo")
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
model = hft.build_model(hft.default_config(), 64, 128, 2, 2, dropout=0.0, seed=1234, device=dev)
opt = hft.training.Adam(model, lr=1e-4, batch_size=B)
g = torch.Generator(device=dev).manual_seed(2000)
spec = -9.0 + 3.0 * torch.randn((B, 256, 192), device=dev, generator=g)
u = torch.rand((3, B, 128, 88), device=dev, generator=g)
lab = [(u[i] > 0.9).float() for i in range(3)] + [torch.zeros((B, 128, 88), dtype=torch.int64, device=dev)]
try:
    opt.forward_backward(spec, *lab)
    torch.cuda.synchronize()
    print("rank", lr, mode, "B", B, "OK loss", float(opt.loss.item()), flush=True)
except Exception as e:
    print("rank", lr, mode, "B", B, "FAIL", str(e)[:200], flush=True)
