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


def backward(params, user, item, model_type, dlogit, dfeat=None):
    """Gradients autograd produces for `loss.backward()` (reference scripts/train_neumf.py:114):
    dense zero-initialised table gradients with the per-sample rows summed, and tower dW / db.
    Returns a dict keyed like the state_dict (absent key = parameter unused by this model_type).
    `dfeat`: optional {feature name: dloss/dfeature [B, width]} for the intermediate features of
    FeatureDistillation.extract_features (reference src/distillation/feature.py:48-81): gmf_features,
    mlp_input, mlp_linear_k (pre-activation of tower layer k), mlp_relu_k (its output)."""
    dfeat = dfeat or {}
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
        if "gmf_features" in dfeat:
            dprod = dprod + dfeat["gmf_features"].astype(np.float64)
        for key, idx, val in (("embed_user_GMF.weight", user, dprod * gi),
                              ("embed_item_GMF.weight", item, dprod * gu)):
            gt = np.zeros(params[key].shape, dtype=np.float64)
            np.add.at(gt, idx, val)
            g[key] = gt.astype(F32)
        off = f
    if model_type != "GMF":
        hs = [h.astype(np.float64) for h in acts["h"]]
        dh = dl[:, None] * pw[None, off:]                      # dloss / d(relu output of the last layer)
        for k in range(L - 1, -1, -1):
            if f"mlp_relu_{k}" in dfeat:
                dh = dh + dfeat[f"mlp_relu_{k}"].astype(np.float64)
            delta = dh * (hs[k + 1] > 0)                        # dloss / d(pre-activation of layer k)
            if f"mlp_linear_{k}" in dfeat:
                delta = delta + dfeat[f"mlp_linear_{k}"].astype(np.float64)
            w, _ = _lin(params, k)
            g[f"MLP_layers.{3 * k + 1}.weight"] = (delta.T @ hs[k]).astype(F32)
            g[f"MLP_layers.{3 * k + 1}.bias"] = delta.sum(0).astype(F32)
            dh = delta @ w.astype(np.float64)
        delta = dh
        if "mlp_input" in dfeat:
            delta = delta + dfeat["mlp_input"].astype(np.float64)
        d = params["embed_user_MLP.weight"].shape[1]
        for key, idx, val in (("embed_user_MLP.weight", user, delta[:, :d]),
                              ("embed_item_MLP.weight", item, delta[:, d:])):
            gt = np.zeros(params[key].shape, dtype=np.float64)
            np.add.at(gt, idx, val)
            g[key] = gt.astype(F32)
    return g


def kd_loss_and_dlogit(logits, label, teacher_logits, w_task, w_kd, temperature=1.0, kd_mode=0):
    """w_task * BCE(x, y) + w_kd * KD(x, t), means over the batch; returns (loss, dloss/dx).
    kd_mode 0: KD = mse(x, t)  (ResponseDistillation.knowledge_distillation_loss, response.py:28-32)
    kd_mode 1: KD = T^2 * mse(sigmoid(x/T), sigmoid(t/T))  (BaseDistillation, base.py:26-33; also the soft
    term of SoftTargetDistillation, response.py:48-60, and the response term of Feature- /
    AttentionDistillation)."""
    B = logits.shape[0]
    x, y = logits.astype(np.float64), label.astype(np.float64)
    sig = 1.0 / (1.0 + np.exp(-x))
    loss = w_task * bce_with_logits(logits, label).mean()
    d = w_task * (sig - y)
    if teacher_logits is not None and w_kd != 0:
        t = teacher_logits.astype(np.float64)
        if kd_mode == 0:
            loss += w_kd * ((x - t) ** 2).mean()
            d = d + w_kd * 2.0 * (x - t)
        else:
            T = float(temperature)
            ss, st = 1.0 / (1.0 + np.exp(-x / T)), 1.0 / (1.0 + np.exp(-t / T))
            loss += w_kd * T * T * ((ss - st) ** 2).mean()
            d = d + w_kd * T * T * 2.0 * (ss - st) * ss * (1.0 - ss) / T
    return F32(loss), (d / B).astype(F32)


def extract_features(params, user, item):
    """FeatureDistillation.extract_features (reference src/distillation/feature.py:48-81) in numpy."""
    _, acts = forward(params, user, item, "NeuMF-end", return_acts=True)
    L = num_layers_of(params)
    feats = {"gmf_features": (acts["gu"] * acts["gi"]).astype(F32), "mlp_input": acts["h"][0]}
    x = acts["h"][0].astype(np.float64)
    for k in range(L):
        w, b = _lin(params, k)
        z = x @ w.T.astype(np.float64) + b
        feats[f"mlp_linear_{k}"] = z.astype(F32)
        x = np.maximum(z, 0.0)
        feats[f"mlp_relu_{k}"] = x.astype(F32)
    return feats


def feature_matching(student_params, teacher_params, user, item, adapters, beta):
    """beta * FeatureDistillation.feature_matching_loss (feature.py:83-123) and its gradient with respect to
    every matched student feature.  adapters: {feature name: (W [teacher width, student width], b)} for the
    features whose widths differ; a feature with different widths and no adapter is skipped, like the
    reference does.  Returns (beta * loss, {feature name: dloss/dfeature})."""
    fs, ft = extract_features(student_params, user, item), extract_features(teacher_params, user, item)
    matched = []
    for key in ft:
        if key not in fs:
            continue
        if ft[key].shape != fs[key].shape and key not in adapters:
            continue
        matched.append(key)
    loss, dfeat = 0.0, {}
    for key in matched:
        s_, t_ = fs[key].astype(np.float64), ft[key].astype(np.float64)
        if ft[key].shape != fs[key].shape:
            W, b = (a.astype(np.float64) for a in adapters[key])
            r = s_ @ W.T + b - t_
            loss += (r ** 2).mean() / len(matched)
            dfeat[key] = (beta / len(matched)) * 2.0 / r.size * (r @ W)
        else:
            r = s_ - t_
            loss += (r ** 2).mean() / len(matched)
            dfeat[key] = (beta / len(matched)) * 2.0 / r.size * r
    return beta * loss, dfeat


def attention_transfer(student_params, teacher_params, user, item):
    """AttentionDistillation.attention_transfer_loss (reference src/distillation/attention.py:16-79): the
    attention map is softmax over the batch of the L2 norm of L2-normalised feature rows, i.e. 1/B for every
    row with a non-zero feature vector on either side, so the KL term is ~0 (fp32 rounding) and has zero
    gradient.  Evaluated here literally."""
    def amap(x):
        x = x.astype(np.float64)
        n = np.linalg.norm(x, axis=-1, keepdims=True)
        xn = x / np.maximum(n, 1e-12)
        a = np.linalg.norm(xn, axis=-1, keepdims=True)
        e = np.exp(a - a.max())
        return (e / e.sum(axis=0, keepdims=True)).reshape(-1)
    fs, ft = extract_features(student_params, user, item), extract_features(teacher_params, user, item)
    total = 0.0
    for key in ("gmf_features", "mlp_input"):
        t, s_ = amap(ft[key]) + 1e-8, amap(fs[key]) + 1e-8
        t, s_ = t / t.sum(), s_ / s_.sum()
        total += float((t * (np.log(t) - np.log(s_))).sum() / t.shape[0])      # F.kl_div(log s, t, 'batchmean')
    return total / 2


def metrics_at_k(params, model_type, users, cands, ks):
    """HR@k / NDCG@k means for every k (reference scripts/evaluate_models.py:22-32 calls metrics() per k)."""
    (_, _), scores = metrics(params, model_type, users, cands, 1)
    out = {}
    for k in ks:
        HR, NDCG = metrics_from_scores(scores, cands, k)
        out[k] = (float(np.mean(HR)), float(np.mean(NDCG)))
    return out


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
