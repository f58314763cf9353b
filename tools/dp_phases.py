"""Per-phase CUDA-event times of the replicated data-parallel step (rank 0), bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from ncf_b200 import ops
from ncf_b200.models import NCF
from ncf_b200.trainer import FusedTrainStep
from ncf_b200.dist import ReplicatedDataParallel, gather_indices, average_
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
U, I, f, L, B = 138493, 26744, 32, 3, 65536
torch.manual_seed(0)
model = NCF(U, I, f, L, 0.0, "NeuMF-end").to(dev)
ts = FusedTrainStep(model, "adam", 1e-3, max_batch=B)
dp = ReplicatedDataParallel(ts)
g = torch.Generator(device=dev).manual_seed(1 + rank)
K = 12
names = ["all_gather", "adam_prepare", "train_step_grads", "all_reduce", "adam_step"]
acc = [0.0] * 5
for k in range(K):
    u = torch.randint(0, U, (B,), device=dev, generator=g); i = torch.randint(0, I, (B,), device=dev, generator=g)
    y = (torch.rand(B, device=dev, generator=g) < 0.2).float()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    dist.barrier(); torch.cuda.synchronize()
    ev[0].record()
    gu, gi = gather_indices(u, i, world); ev[1].record()
    ops.adam_prepare(ts._m, ts._g, ts._s, gu, gi, ts.lr); ev[2].record()
    ops.train_step_grads(ts._m, ts._g, u, i, y, None, 1.0, ts.loss_accum, ts.workspace); ev[3].record()
    average_(ts.grads.flat, world); ev[4].record()
    ops.adam_step(ts._m, ts._g, ts._s, ts.lr); ev[5].record()
    torch.cuda.synchronize()
    if k >= 4:
        for j in range(5): acc[j] += ev[j].elapsed_time(ev[j + 1]) / (K - 4)
if rank == 0:
    print(f"world {world}: " + "  ".join(f"{n} {a*1e3:.0f} us" for n, a in zip(names, acc)) + f"  total {sum(acc)*1e3:.0f} us  (flat buffer {ts.grads.flat.numel()*4/2**20:.0f} MiB)")
dist.destroy_process_group()
