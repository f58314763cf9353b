"""Small run of every tcgen05 kernel (fused tower train + inference, per-layer fallback, wgrad) for compute-sanitizer."""
import os, sys
os.environ["NCF_UMMA_MIN_B"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ncf_b200.models import NCF
from ncf_b200.trainer import FusedTrainStep
dev = torch.device("cuda:0")
for f, L, B in ((32, 3, 700), (64, 3, 300), (32, 1, 257)):
    torch.manual_seed(0)
    m = NCF(500, 400, f, L, 0.0, "NeuMF-end").to(dev)
    ts = FusedTrainStep(m, "adam", 1e-3, max_batch=B)
    g = torch.Generator(device=dev).manual_seed(1)
    for _ in range(2):
        u = torch.randint(0, 500, (B,), device=dev, generator=g); i = torch.randint(0, 400, (B,), device=dev, generator=g)
        y = (torch.rand(B, device=dev, generator=g) < 0.3).float()
        ts.step(u, i, y)
    ts.flush()
    with torch.no_grad():
        out = m.eval()(u, i)
    torch.cuda.synchronize()
    print(f, L, B, "loss", ts.pop_loss(), "logit0", float(out[0]))
print("done")
