"""Debug aid: emulates the user-partitioned data-parallel step of ncf_b200.dist on ONE GPU (two model
replicas stepped in turn, the all-reduce done by hand) and prints per-tensor differences against the
single-process run at the global batch."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ncf_b200 import ops  # noqa: E402
from ncf_b200.models import NCF  # noqa: E402
from ncf_b200.trainer import FusedTrainStep  # noqa: E402

dev = torch.device("cuda:0")
U, I, f, L, B, T, W = 3000, 2000, 32, 3, 20000, 3, 2
if len(sys.argv) > 1 and sys.argv[1] == "small":
    U, I, f, L, B, T, W = 300, 200, 16, 2, 128, 3, 2
rng = np.random.default_rng(0)
users = torch.from_numpy(rng.integers(0, U, (T, B))).to(dev)
items = torch.from_numpy(rng.integers(0, I, (T, B))).to(dev)
labels = torch.from_numpy((rng.random((T, B)) < 0.3).astype(np.float32)).to(dev)


def make():
    torch.manual_seed(0)
    m = NCF(U, I, f, L, 0.0, "NeuMF-end").to(dev)
    return m, FusedTrainStep(m, "adam", 1e-3, max_batch=B)


ref_m, ref = make()
reps = [make() for _ in range(W)]
bounds = [(r * U // W, (r + 1) * U // W) for r in range(W)]
for t in range(T):
    ref.step(users[t], items[t], labels[t])
    dense = ref.dense_adam(B)
    tails = []
    for r, (m, ts) in enumerate(reps):
        lo, hi = bounds[r]
        mine = (users[t] >= lo) & (users[t] < hi)
        u, i, y = users[t][mine].contiguous(), items[t][mine].contiguous(), labels[t][mine].contiguous()
        if not dense:
            ops.mark_rows_side(ts._m, ts._g, u, 0)
            ops.mark_rows_side(ts._m, ts._g, items[t].contiguous(), 1)
            ops.adam_catchup(ts._m, ts._g, ts._s, ts.lr)
        ops.train_step_grads_norm(ts._m, ts._g, u, i, y, B, ts.loss_accum, ts.workspace)
        g = ts.grads
        n_user = (g.g_item_gmf.data_ptr() - g.flat.data_ptr()) // 4
        tails.append(g.flat[n_user:])
    total = sum(t_.clone() for t_ in tails)
    # single-process gradient check: user grads of replica r on its rows, item + tower = total
    torch.cuda.synchronize()
    print(f"step {t}: dense={dense} path={ops._lib.load().ncf_last_tile_path()}")
    for r, (m, ts) in enumerate(reps):
        tails[r].copy_(total)
        lo, hi = bounds[r]
        if dense:
            ops.adam_step_dense_range(ts._m, ts._g, ts._s, lo, hi, ts.lr)
        else:
            ops.adam_step(ts._m, ts._g, ts._s, ts.lr)
    ref.flush()
    for r, (m, ts) in enumerate(reps):
        if not dense:
            ops.adam_flush(ts._m, ts._s, ts.lr)
    torch.cuda.synchronize()
    for (k, a) in ref_m.state_dict().items():
        scale = max(a.abs().max().item(), 1e-30)
        if "user" in k:
            full = torch.cat([reps[r][0].state_dict()[k][bounds[r][0]:bounds[r][1]] for r in range(W)])
            print(f"   {k:28s} {(full - a).abs().max().item() / scale:.3e}")
        else:
            print(f"   {k:28s} " + " ".join(f"{(reps[r][0].state_dict()[k] - a).abs().max().item() / scale:.3e}" for r in range(W)))
    print("   loss ref", ref.loss_accum.item(), "dp", sum(ts.loss_accum.item() for _, ts in reps))
