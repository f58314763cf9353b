"""ORACLE — test infrastructure, not product code.

A CPU port of the reference's training / evaluation step on stock PyTorch CPU ops, used ONLY as
the timed CPU baseline of bench.py (`cpu_baseline`, kind "port", and `--impl reference`) and by
tests that cross-check it against the numpy restatement.  The reference itself is Python and
cannot travel to the GPU box (/root/reference does not exist there), so this port stands in for
"the reference's CPU PyTorch path on the box's own host cores": the same op sequence as
reference src/ncf/models.py:97-118 (embedding gathers, GMF product, Linear+ReLU tower, concat,
predict), nn.BCEWithLogitsLoss (scripts/train_neumf.py:86), autograd backward producing DENSE
embedding gradients, and dense torch.optim.Adam over all parameters (train_neumf.py:90) — i.e. it
keeps the reference's cost structure (O(table size) per step), multi-threaded through torch's
intra-op pool.  Validated against oracle/ncf_numpy.py in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def init_params(user_num, item_num, f, L, model_type, seed=0):
    """Random parameters of the reference's shapes (state_dict keys), as leaf tensors."""
    g = torch.Generator().manual_seed(seed)
    d = f << (L - 1)
    P = {
        "embed_user_GMF.weight": torch.randn(user_num, f, generator=g) * 0.01,
        "embed_item_GMF.weight": torch.randn(item_num, f, generator=g) * 0.01,
        "embed_user_MLP.weight": torch.randn(user_num, d, generator=g) * 0.01,
        "embed_item_MLP.weight": torch.randn(item_num, d, generator=g) * 0.01,
    }
    w = f << L
    for k in range(L):
        bound = (6.0 / (w + w // 2)) ** 0.5
        P[f"MLP_layers.{3 * k + 1}.weight"] = (torch.rand(w // 2, w, generator=g) * 2 - 1) * bound
        P[f"MLP_layers.{3 * k + 1}.bias"] = (torch.rand(w // 2, generator=g) * 2 - 1) / w ** 0.5
        w //= 2
    ps = f if model_type in ("GMF", "MLP") else 2 * f
    P["predict_layer.weight"] = (torch.rand(1, ps, generator=g) * 2 - 1) * (3.0 / ps) ** 0.5
    P["predict_layer.bias"] = torch.zeros(1)
    return {k: v.requires_grad_(True) for k, v in P.items()}


def forward(P, user, item, model_type):
    L = sum(1 for k in P if k.startswith("MLP_layers.") and k.endswith(".weight"))
    parts = []
    if model_type != "MLP":
        parts.append(F.embedding(user, P["embed_user_GMF.weight"]) * F.embedding(item, P["embed_item_GMF.weight"]))
    if model_type != "GMF":
        h = torch.cat((F.embedding(user, P["embed_user_MLP.weight"]),
                       F.embedding(item, P["embed_item_MLP.weight"])), -1)
        for k in range(L):
            h = F.relu(F.linear(h, P[f"MLP_layers.{3 * k + 1}.weight"], P[f"MLP_layers.{3 * k + 1}.bias"]))
        parts.append(h)
    x = parts[0] if len(parts) == 1 else torch.cat(parts, -1)
    return F.linear(x, P["predict_layer.weight"], P["predict_layer.bias"]).view(-1)


def used_params(P, model_type):
    keep = {"GMF": ("embed_user_GMF", "embed_item_GMF", "predict"),
            "MLP": ("embed_user_MLP", "embed_item_MLP", "MLP_layers", "predict")}.get(model_type)
    return [v for k, v in P.items() if keep is None or k.startswith(keep)]


class CpuTrainer:
    """The reference inner loop (scripts/train_neumf.py:111-117) on CPU tensors."""

    def __init__(self, P, model_type, lr=1e-3):
        self.P, self.model_type = P, model_type
        self.opt = torch.optim.Adam(list(P.values()), lr=lr)  # model.parameters(): all of them

    def step(self, user, item, label):
        self.opt.zero_grad()
        loss = F.binary_cross_entropy_with_logits(forward(self.P, user, item, self.model_type), label)
        loss.backward()
        self.opt.step()
        return loss.item()


def evaluate(P, model_type, users, cands, k):
    """metrics() (reference src/training/metrics.py:4-25): one forward + topk per user."""
    hits, ndcg = 0, 0.0
    C = cands.shape[1]
    with torch.no_grad():
        for u, row in zip(users.tolist(), cands):
            s = forward(P, torch.full((C,), u, dtype=torch.int64), row, model_type)
            _, idx = torch.topk(s, k)
            rec = row[idx].tolist()
            if row[0].item() in rec:
                hits += 1
                ndcg += 1.0 / torch.log2(torch.tensor(rec.index(row[0].item()) + 2.0)).item()
    return hits / len(cands), ndcg / len(cands)
