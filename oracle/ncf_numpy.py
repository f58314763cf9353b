"""ORACLE — test infrastructure, not product code.

A plain-numpy CPU restatement of the reference's NCF arithmetic.  Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import this; the product
(ncf_b200/) never does.

Parity status: PINNED.  The reference is pure Python on PyTorch and has no golden vectors of its
own, so this restatement is pinned against outputs of the reference itself,
generated in the build container by oracle/make_golden.py (which imports /root/reference) and
committed under tests/golden/; tests/test_oracle_golden.py checks every function below against
them.  The arithmetic lives in a third-party dependency of the reference (PyTorch, pinned
torch==1.0.1 in the reference's setup.py:10; the goldens were produced with torch 2.11.0 CPU): the
formulas restated here are the published semantics of nn.Embedding, nn.Linear, nn.ReLU,
BCEWithLogitsLoss, F.mse_loss and optim.Adam/SGD at the reference's call sites, cited per function.

Parameters are passed as a dict with the reference's state_dict keys
(`embed_user_GMF.weight`, ..., `MLP_layers.{3k+1}.weight`, `predict_layer.bias`).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def num_layers_of(params) -> int:
    return sum(1 for k in params if k.startswith("MLP_layers.") and k.endswith(".weight"))


def _lin(params, k):
    return params[f"MLP_layers.{3 * k + 1}.weight"], params[f"MLP_layers.{3 * k + 1}.bias"]


def forward(params, user, item, model_type, return_acts=False):
    """reference src/ncf/models.py:97-118 (dropout = 0).  Returns logits[B] (float32)."""
    L = num_layers_of(params)
    acts = {}
    pieces = []
    if model_type != "MLP":
        gu = params["embed_user_GMF.weight"][user]
        gi = params["embed_item_GMF.weight"][item]
        acts["gu"], acts["gi"] = gu, gi
        pieces.append((gu * gi).astype(F32))  # models.py:101,110
    if model_type != "GMF":
        h = np.concatenate([params["embed_user_MLP.weight"][user],
                            params["embed_item_MLP.weight"][item]], axis=-1).astype(F32)  # :105,113
        hs = [h]
        for k in range(L):
            w, b = _lin(params, k)
            h = np.maximum(h @ w.T + b, F32(0)).astype(F32)  # Linear + ReLU, models.py:24-25
            hs.append(h)
        acts["h"] = hs
        pieces.append(h)
    feat = np.concatenate(pieces, axis=-1).astype(F32)  # models.py:115
    acts["feat"] = feat
    logits = (feat @ params["predict_layer.weight"].T + params["predict_layer.bias"]).reshape(-1)
    logits = logits.astype(F32)
    return (logits, acts) if return_acts else logits


def bce_with_logits(x, y):
    """nn.BCEWithLogitsLoss elementwise term: max(x,0) - x*y + log1p(exp(-|x|))
    (reference scripts/train_neumf.py:86,113)."""
    x = x.astype(np.float64)
    y = y.astype(np.float64)
    return np.maximum(x, 0) - x * y + np.log1p(np.exp(-np.abs(x)))


def loss_and_dlogit(logits, label, teacher_logits=None, alpha=0.5):
    """Plain BCE mean, or ResponseDistillation.combined_loss = alpha*BCE + (1-alpha)*mse(x, t)
    (reference src/distillation/base.py:40-50, response.py:28-32).  Returns (loss, dloss/dx)."""
    B = logits.shape[0]
    x = logits.astype(np.float64)
    y = label.astype(np.float64)
    sig = 1.0 / (1.0 + np.exp(-x))
    bce = bce_with_logits(logits, label).mean()
    if teacher_logits is None:
        return F32(bce), ((sig - y) / B).astype(F32)
    t = teacher_logits.astype(np.float64)
    kd = ((x - t) ** 2).mean()
    loss = alpha * bce + (1 - alpha) * kd
    d = (alpha * (sig - y) + (1 - alpha) * 2.0 * (x - t)) / B
    return F32(loss), d.astype(F32)


def backward(params, user, item, model_type, dlogit):
    """Gradients autograd produces for `loss.backward()` (reference scripts/train_neumf.py:114):
    dense zero-initialised table gradients with the per-sample rows summed, and tower dW / db.
    Returns a dict keyed like the state_dict (absent key = parameter unused by this model_type)."""
    L = num_layers_of(params)
    _, acts = forward(params, user, item, model_type, return_acts=True)
    f = params["embed_user_GMF.weight"].shape[1]
    pw = params["predict_layer.weight"].reshape(-1).astype(np.float64)
    dl = dlogit.astype(np.float64)
    g = {}
    feat = acts["feat"].astype(np.float64)
    g["predict_layer.weight"] = (dl[:, None] * feat).sum(0).reshape(1, -1).astype(F32)
    g["predict_layer.bias"] = np.array([dl.sum()], dtype=F32)
    off = 0
    if model_type != "MLP":
        gu, gi = acts["gu"].astype(np.float64), acts["gi"].astype(np.float64)
        dprod = dl[:, None] * pw[None, :f]
        for key, idx, val in (("embed_user_GMF.weight", user, dprod * gi),
                              ("embed_item_GMF.weight", item, dprod * gu)):
            gt = np.zeros(params[key].shape, dtype=np.float64)
            np.add.at(gt, idx, val)
            g[key] = gt.astype(F32)
        off = f
    if model_type != "GMF":
        hs = [h.astype(np.float64) for h in acts["h"]]
        delta = dl[:, None] * pw[None, off:] * (hs[L] > 0)
        for k in range(L - 1, -1, -1):
            w, _ = _lin(params, k)
            g[f"MLP_layers.{3 * k + 1}.weight"] = (delta.T @ hs[k]).astype(F32)
            g[f"MLP_layers.{3 * k + 1}.bias"] = delta.sum(0).astype(F32)
            delta = delta @ w.astype(np.float64)
            if k > 0:
                delta = delta * (hs[k] > 0)
        d = params["embed_user_MLP.weight"].shape[1]
        for key, idx, val in (("embed_user_MLP.weight", user, delta[:, :d]),
                              ("embed_item_MLP.weight", item, delta[:, d:])):
            gt = np.zeros(params[key].shape, dtype=np.float64)
            np.add.at(gt, idx, val)
            g[key] = gt.astype(F32)
    return g


class DenseAdam:
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8) as the reference uses it
    (scripts/train_neumf.py:90,115), i.e. DENSE over every parameter that has a gradient: a row
    touched once keeps moving through its momentum.  Math of torch 2.11
    optim/adam.py::_single_tensor_adam; state per parameter: step, exp_avg, exp_avg_sq."""

    def __init__(self, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.state = {}

    def step(self, params, grads):
        for key, g in grads.items():
            st = self.state.setdefault(key, {"t": 0, "m": np.zeros_like(params[key]),
                                             "v": np.zeros_like(params[key])})
            st["t"] += 1
            t = st["t"]
            m, v = st["m"], st["v"]
            m += (g - m) * F32(1 - self.b1)                      # exp_avg.lerp_(grad, 1-beta1)
            v *= F32(self.b2)
            v += F32(1 - self.b2) * g * g                        # mul_(beta2).addcmul_(g, g, 1-beta2)
            bc1 = 1 - self.b1 ** t
            bc2 = 1 - self.b2 ** t
            step_size = self.lr / bc1
            denom = np.sqrt(v) / F32(np.sqrt(bc2)) + F32(self.eps)
            params[key] -= (F32(step_size) * (m / denom)).astype(F32)


def sgd_step(params, grads, lr):
    """optim.SGD(lr) without momentum (reference scripts/train_neumf.py:88)."""
    for key, g in grads.items():
        params[key] -= F32(lr) * g


def topk_indices(scores, k):
    """Descending top-k of one score vector; ties -> lower index first (our documented rule;
    torch.topk leaves tie order unspecified)."""
    order = np.lexsort((np.arange(scores.shape[0]), -scores.astype(np.float64)))
    return order[:k]


def metrics_from_scores(scores, items, k):
    """reference src/training/metrics.py:4-25 on precomputed scores [n, C] / items [n, C]:
    HR list (ints) and NDCG list (python floats, 1/log2(index+2))."""
    HR, NDCG = [], []
    for s, it in zip(scores, items):
        idx = topk_indices(s, k)
        rec = it[idx]
        gt = it[0]
        hit = int(gt in rec)
        HR.append(hit)
        nd = 0.0
        if hit:
            index = np.where(rec == gt)[0][0]
            nd = 1.0 / np.log2(index + 2)
        NDCG.append(nd)
    return HR, NDCG


def metrics(params, model_type, users, cands, k):
    """Full leave-one-out evaluation: users [n], cands [n, C] with column 0 the held-out item."""
    n, C = cands.shape
    u = np.repeat(users, C)
    scores = forward(params, u, cands.reshape(-1), model_type).reshape(n, C)
    return metrics_from_scores(scores, cands, k), scores
