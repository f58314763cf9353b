"""Times the leave-one-out evaluation (forward over users x 100 candidates + ranking) repeatedly."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ncf_b200.models import NCF
from ncf_b200.metrics import evaluate
dev = torch.device("cuda:0")
torch.manual_seed(0)
U, I = 138493, 26744
model = NCF(U, I, 32, 3, 0.0, "NeuMF-end").to(dev).eval()
users = torch.arange(U, device=dev)
cands = torch.randint(0, I, (U, 100), device=dev)
with torch.no_grad():
    evaluate(model, users, cands, 10)
    torch.cuda.synchronize()
    for rep in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); evaluate(model, users, cands, 10); e1.record(); torch.cuda.synchronize()
        print(f"eval {rep}: {e0.elapsed_time(e1):.2f} ms", flush=True)
