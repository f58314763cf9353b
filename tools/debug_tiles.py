"""Debug aid: one fused training step on the tcgen05 path with DISTINCT users and items per sample, so that
row s of the embedding-gradient tables is exactly sample s's gradient; compares with the numpy oracle and
prints the error per 128-sample tile (tile t runs on CTA t % 148 as that CTA's (t // 148)-th tile)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ncf_b200 import ops  # noqa: E402
from ncf_b200.models import NCF  # noqa: E402
from oracle import ncf_numpy as onp  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
f, L = 32, 3
U = I = B
torch.manual_seed(0)
rng = np.random.default_rng(0)
model = NCF(U, I, f, L, 0.0, "NeuMF-end").to(dev)
with torch.no_grad():
    for lin in model.linears():
        lin.bias.uniform_(-0.1, 0.1)
    for p in (model.embed_user_MLP.weight, model.embed_item_MLP.weight, model.embed_user_GMF.weight, model.embed_item_GMF.weight):
        p.mul_(20.0)
u = rng.permutation(U).astype(np.int64)
i = rng.permutation(I).astype(np.int64)
y = (rng.random(B) < 0.3).astype(np.float32)
params = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
g = ops.GradBuffers.allocate(model.abi_type(), f, L, U, I, B, dev)
m = model.abi_struct()
ws = torch.empty(ops.train_workspace_bytes(m, B), dtype=torch.uint8, device=dev)
loss = torch.zeros(1, dtype=torch.float64, device=dev)
logits = torch.empty(B, device=dev)
tu, ti, ty = (torch.from_numpy(a).to(dev) for a in (u, i, y))
ref_logits, acts = onp.forward(params, u, i, "NeuMF-end", return_acts=True)
ref_loss, dl = onp.loss_and_dlogit(ref_logits, y)
ref = onp.backward(params, u, i, "NeuMF-end", dl)
want_u = ref["embed_user_MLP.weight"][u]
scale_u = np.abs(want_u).max()
for rep in range(4):
    g.flat.zero_(); loss.zero_()
    ops.train_step_grads(m, g.struct(), tu, ti, ty, None, 1.0, loss, ws, logits)
    torch.cuda.synchronize()
    got_u = g.g_user_mlp.cpu().numpy()[u]
    err = np.abs(got_u - want_u) / scale_u                     # [B, d]
    rows = np.nonzero(err.max(axis=1) > 1e-5)[0]
    tiles = sorted(set((rows // 128).tolist()))
    print(f"rep {rep}: bad rows {len(rows)} in tiles {tiles}")
    for t in tiles[:3]:
        r = rows[rows // 128 == t]
        e = err[r]
        cols = np.nonzero(e.max(axis=0) > 1e-5)[0]
        print(f"   tile {t}: rows-in-tile {(r % 128).tolist()[:40]}{'...' if len(r) > 40 else ''}; bad cols {len(cols)} "
              f"[{cols.min()}..{cols.max()}]; panels {sorted(set((cols // 32).tolist()))}; max err {e.max():.3e}")
        r0 = r[0]
        ratio = got_u[r0] / np.where(want_u[r0] == 0, 1, want_u[r0])
        print(f"      sample {r0}: got/want first cols {np.round(ratio[:8], 4).tolist()} label {y[r0]} logit {ref_logits[r0]:.4f}")
print("tile path", ops._lib.load().ncf_last_tile_path(), "B", B, "tiles", (B + 127) // 128)
print("loss", loss.item(), float(ref_loss))


def per_tile(name, got, want, idx):
    got, want = got.cpu().numpy()[idx], want[idx]          # row s = sample s
    scale = np.abs(want).max()
    err = np.abs(got - want).max(axis=1) / scale
    nt = (B + 127) // 128
    e = np.array([err[t * 128:(t + 1) * 128].max() for t in range(nt)])
    bad = np.nonzero(e > 1e-5)[0]
    print(f"{name}: max rel err {e.max():.3e}; tiles over 1e-5: {len(bad)} of {nt}",
          ("" if len(bad) == 0 else f"first {bad[:12].tolist()} local-tile index {sorted(set((bad // 148).tolist()))}"))


lg = np.abs(logits.cpu().numpy() - ref_logits) / np.abs(ref_logits).max()
print("logits max rel err", lg.max())
per_tile("g_user_mlp", g.g_user_mlp, ref["embed_user_MLP.weight"], u)
per_tile("g_item_mlp", g.g_item_mlp, ref["embed_item_MLP.weight"], i)
per_tile("g_user_gmf", g.g_user_gmf, ref["embed_user_GMF.weight"], u)
per_tile("g_item_gmf", g.g_item_gmf, ref["embed_item_GMF.weight"], i)
# tower gradients
flat = g.g_tower.cpu().numpy()
off = 0
for k in range(L):
    for nm in (f"MLP_layers.{3 * k + 1}.weight", f"MLP_layers.{3 * k + 1}.bias"):
        w = ref[nm]
        got = flat[off:off + w.size].reshape(w.shape)
        print(f"{nm}: rel err {np.abs(got - w).max() / np.abs(w).max():.3e}")
        off += w.size
