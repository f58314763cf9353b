"""A few training steps of the bench workload through the tcgen05 path (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ncf_b200.models import NCF
from ncf_b200.trainer import FusedTrainStep
f = int(os.environ.get("F", "32")); L = 3; B = int(os.environ.get("B", "65536"))
U, I = 138493, 26744
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = NCF(U, I, f, L, 0.0, "NeuMF-end").to(dev)
ts = FusedTrainStep(model, "adam", 1e-3, max_batch=B)
g = torch.Generator(device=dev).manual_seed(1)
steps = int(os.environ.get("STEPS", "4"))
for k in range(steps):
    u = torch.randint(0, U, (B,), device=dev, generator=g)
    i = torch.randint(0, I, (B,), device=dev, generator=g)
    y = (torch.rand(B, device=dev, generator=g) < 0.2).float()
    ts.step(u, i, y)
torch.cuda.synchronize()
print("loss", ts.pop_loss())
