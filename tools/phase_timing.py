"""Debug: per-phase cycle breakdown of the mma tile kernel (build with NCF_EXTRA_NVCC_FLAGS=-DNCF_PHASE_TIMING)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from ncf_b200 import _lib, ops
from ncf_b200.models import NCF
from ncf_b200.trainer import FusedTrainStep
lib = _lib.load()
dev = torch.device("cuda:0")
U, I, f, L, B = 138493, 26744, 32, 3, 65536
model = NCF(U, I, f, L, 0.0, "NeuMF-end").to(dev)
model.tower_math = sys.argv[1] if len(sys.argv) > 1 else "fp32"
print("tower_math", model.tower_math)
ts = FusedTrainStep(model, "adam", 1e-3, max_batch=B)
g = torch.Generator(device=dev).manual_seed(0)
u = torch.randint(0, U, (B,), device=dev, generator=g); i = torch.randint(0, I, (B,), device=dev, generator=g)
y = (torch.rand(B, device=dev, generator=g) < 0.2).float()
for _ in range(3): ts.step(u, i, y)
buf = (ctypes.c_ulonglong * 16)()
lib.ncf_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.ncf_debug_phase_cycles(buf, 1)
n = 5
for _ in range(n): ts.step(u, i, y)
lib.ncf_debug_phase_cycles(buf, 0)
names = ["gather", "fwd tower", "predict+loss+pgrads", "gmf scatter+deltaL", "bias grads", "wgrad k=0", "wgrad k=1", "wgrad k=2",
         "bwd dX", "bwd k=1", "bwd k=2"]
tot = sum(buf[:11])
tiles = n * ((B + 63) // 64 + 147) // 148
for nm, c in zip(names, buf):
    print(f"{nm:22s} {c/ n:12.0f} cyc/step  {100*c/tot:5.1f}%")
print("total cycles/step (block 0):", tot / n, "=> us @1.9GHz:", tot / n / 1900)
