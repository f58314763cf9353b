"""Scratch diagnostics for the tcgen05 path: where do the gradients differ from the oracle?"""
import os, sys
os.environ["NCF_UMMA_MIN_B"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import ncf_numpy as onp
from ncf_b200 import ops, _lib
from ncf_b200.models import NCF

def run(model_type, f, L, B, U=700, I=500, seed=0, bias=True):
    dev = torch.device("cuda:0")
    torch.manual_seed(seed); rng = np.random.default_rng(seed)
    model = NCF(U, I, f, L, 0.0, model_type).to(dev)
    if bias:
        with torch.no_grad():
            for lin in model.linears(): lin.bias.uniform_(-0.1, 0.1)
    params = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    u = rng.integers(0, U, B); i = rng.integers(0, I, B); y = (rng.random(B) < 0.3).astype(np.float32)
    ud, idd, yd = (torch.from_numpy(a).to(dev) for a in (u, i, y))
    g = ops.GradBuffers.allocate(model.abi_type(), f, L, U, I, B, dev)
    m = model.abi_struct()
    ws = torch.empty(ops.train_workspace_bytes(m, B), dtype=torch.uint8, device=dev)
    loss = torch.zeros(1, dtype=torch.float64, device=dev); logits = torch.empty(B, device=dev)
    ops.mark_rows(m, g.struct(), ud, idd)
    ops.train_step_grads(m, g.struct(), ud, idd, yd, None, 1.0, loss, ws, logits)
    torch.cuda.synchronize()
    print(f"--- {model_type} f={f} L={L} B={B} bias={bias} path={_lib.load().ncf_last_tile_path()}")
    rl = onp.forward(params, u, i, model_type); _, dl = onp.loss_and_dlogit(rl, y)
    ref = onp.backward(params, u, i, model_type, dl)
    print("logits err", np.abs(logits.cpu().numpy() - rl).max() / np.abs(rl).max())
    tabs = {"embed_user_GMF.weight": g.g_user_gmf, "embed_item_GMF.weight": g.g_item_gmf,
            "embed_user_MLP.weight": g.g_user_mlp, "embed_item_MLP.weight": g.g_item_mlp}
    for k, buf in tabs.items():
        if k not in ref or buf is None: continue
        a = buf.cpu().numpy(); e = np.abs(a - ref[k]); mx = np.abs(ref[k]).max()
        r, c = np.unravel_index(e.argmax(), e.shape)
        badrows = np.where(e.max(1) > 1e-5 * mx)[0]
        idx = u if "user" in k else i
        print(f"{k}: rel {e.max()/mx:.2e} at row {r} col {c}; bad rows {len(badrows)}/{a.shape[0]}; cols of worst row bad: {np.where(e[r] > 1e-5*mx)[0][:16]}")
        if len(badrows):
            samples = np.where(np.isin(idx, badrows[:5]))[0]
            print("   samples touching first bad rows:", samples[:24], " (mod 128:", samples[:24] % 128, ")")
    flat = g.g_tower.cpu().numpy(); off = 0
    keys = [f"MLP_layers.{3*k+1}.{s}" for k in range(L) for s in ("weight", "bias")] + ["predict_layer.weight", "predict_layer.bias"]
    sd = model.state_dict()
    for k in keys:
        n = sd[k].numel(); piece = flat[off:off+n].reshape(tuple(sd[k].shape)); off += n
        e = np.abs(piece - ref[k]); mx = np.abs(ref[k]).max()
        print(f"{k}: rel {e.max()/mx:.2e} argmax {np.unravel_index(e.argmax(), e.shape)}")

for cfg in [("NeuMF-end", 32, 3, 3000), ("NeuMF-end", 32, 3, 128), ("NeuMF-end", 32, 3, 256), ("MLP", 32, 2, 1111), ("NeuMF-end", 64, 3, 700), ("NeuMF-end", 32, 1, 257)]:
    run(*cfg)
run("NeuMF-end", 32, 3, 3000, bias=False)
