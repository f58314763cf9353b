"""Knowledge distillation with the reference's class names and constructor signatures
(reference src/distillation/base.py:6-50, response.py:6-61, feature.py:6-146, attention.py:6-101).

Every class is an `nn.Module` whose `forward(user, item, label)` returns the scalar loss, so the
reference's loop `loss = distillation(user, item, label); loss.backward(); optimizer.step()`
(scripts/train_student.py:148-156) runs unchanged:

  ResponseDistillation    alpha * BCE + (1 - alpha) * mse(student logits, teacher logits); the
                          temperature is accepted and unused exactly as in the reference
                          (response.py:28-32).
  SoftTargetDistillation  alpha * BCE + (1 - alpha) * T^2 * mse(sigmoid(s/T), sigmoid(t/T))  (response.py:34-61)
  FeatureDistillation     alpha * BCE + max(0, 1 - alpha - beta) * KD_soft + beta * feature matching over
                          gmf_features / mlp_input (through fixed adapter Linears when the widths differ)
                          and the tower activations whose shapes agree (feature.py:48-146)
  AttentionDistillation   alpha * BCE + (1 - alpha - gamma) * KD_soft + gamma * attention transfer; the
                          "attention map" of the reference is a softmax over the batch of the L2 norm of
                          L2-normalised rows, i.e. the constant 1/B for teacher and student alike, so
                          the transfer term is ~1e-9 with zero gradient (attention.py:16-28, SURVEY.md §2
                          row 5).  It is evaluated all the same so that the loss value matches.
  UnifiedDistillation     the reference imports it (scripts/train_student.py:19) but ships an empty
                          unified.py; provided as the sum of the three terms above.

Logit-level losses come from the library (ncf_loss_grad / ncf_loss_grad_kd); under autograd the
student's backward is the fused kernel (ncf_b200.autograd).  The feature terms under autograd go
through the module's own nn.Embedding / nn.Linear submodules like the reference's extract_features
does.  Training scripts do not use autograd: `FusedTrainStep(student, distillation=...)` runs the same
arithmetic on the fused path (ncf_feature_kd for the embedding-level features).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class _KDLoss(torch.autograd.Function):
    """w_task * BCE(x, y) + w_kd * KD(x, t) evaluated by the library; backward = its dlogit."""

    @staticmethod
    def forward(ctx, student_logits, teacher_logits, label, w_task, w_kd, temperature, kd_mode):
        acc = torch.zeros(1, dtype=torch.float64, device=student_logits.device)
        dl = torch.empty_like(student_logits)
        ops.loss_grad_kd(student_logits.contiguous(), label.contiguous().float(),
                         None if teacher_logits is None else teacher_logits.contiguous(), float(w_task),
                         float(w_kd), float(temperature), int(kd_mode), acc, dl)
        ctx.save_for_backward(dl)
        return acc.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (dl,) = ctx.saved_tensors
        return dl * grad_out, None, None, None, None, None, None


class BaseDistillation(nn.Module):
    def __init__(self, teacher_model, student_model, temperature=2.0, alpha=0.5):
        super().__init__()
        self.teacher_model = teacher_model
        self.student_model = student_model
        self.temperature = temperature
        self.alpha = alpha
        for p in self.teacher_model.parameters():  # frozen teacher (base.py:16-18)
            p.requires_grad = False
        self.teacher_model.eval()

    def forward(self, user, item, label):
        raise NotImplementedError("Subclasses must implement forward method")

    def knowledge_distillation_loss(self, teacher_logits, student_logits):
        """T^2 * mse(sigmoid(student / T), sigmoid(teacher / T))  (base.py:26-33)."""
        return _KDLoss.apply(student_logits, teacher_logits, torch.zeros_like(student_logits), 0.0, 1.0,
                             self.temperature, 1)

    def task_loss(self, predictions, labels):
        return _KDLoss.apply(predictions, None, labels, 1.0, 0.0, 1.0, 0)

    def combined_loss(self, teacher_logits, student_logits, labels):
        """alpha * task + (1 - alpha) * knowledge_distillation_loss  (base.py:40-50)."""
        return _KDLoss.apply(student_logits, teacher_logits, labels, self.alpha, 1.0 - self.alpha,
                             self.temperature, self._kd_mode)

    _kd_mode = 1   # the base class' KD term is the soft-target one

    # what FusedTrainStep needs to run this objective without autograd
    def fused_spec(self) -> dict:
        return {"w_task": self.alpha, "w_kd": 1.0 - self.alpha, "kd_mode": self._kd_mode,
                "temperature": float(self.temperature), "features": []}

    def _logits(self, user, item):
        with torch.no_grad():
            teacher_logits = self.teacher_model(user, item)
        return teacher_logits, self.student_model(user, item)


class ResponseDistillation(BaseDistillation):
    _kd_mode = 0   # mse on the raw logits (response.py:28-32)

    def forward(self, user, item, label):
        teacher_logits, student_logits = self._logits(user, item)
        return self.combined_loss(teacher_logits, student_logits, label)

    def knowledge_distillation_loss(self, teacher_logits, student_logits):
        return _KDLoss.apply(student_logits, teacher_logits, torch.zeros_like(student_logits), 0.0, 1.0, 1.0, 0)


class SoftTargetDistillation(BaseDistillation):
    def __init__(self, teacher_model, student_model, temperature=4.0, alpha=0.7):
        super().__init__(teacher_model, student_model, temperature, alpha)

    def forward(self, user, item, label):
        teacher_logits, student_logits = self._logits(user, item)
        return _KDLoss.apply(student_logits, teacher_logits, label, self.alpha, 1.0 - self.alpha,
                             self.temperature, 1)


def _embedding_features(model, user, item):
    """The features of reference feature.py:51-81, through the module's own submodules."""
    features = {}
    if hasattr(model, "embed_user_GMF"):
        features["gmf_features"] = model.embed_user_GMF(user) * model.embed_item_GMF(item)
    if hasattr(model, "embed_user_MLP"):
        x = torch.cat((model.embed_user_MLP(user), model.embed_item_MLP(item)), -1)
        features["mlp_input"] = x
        k = 0
        for layer in model.MLP_layers:
            if isinstance(layer, nn.Linear):
                x = layer(x)
                features[f"mlp_linear_{k}"] = x
                k += 1
            elif isinstance(layer, nn.ReLU):
                x = layer(x)
                features[f"mlp_relu_{k - 1}"] = x
    return features


class FeatureDistillation(BaseDistillation):
    def __init__(self, teacher_model, student_model, temperature=2.0, alpha=0.5, beta=0.3):
        super().__init__(teacher_model, student_model, temperature, alpha)
        self.beta = beta
        self.adaptation_layers = nn.ModuleDict()
        t, s = teacher_model, student_model
        # adapters student -> teacher width (feature.py:20-46); never handed to an optimiser by the
        # reference script (train_student.py:131), i.e. fixed random projections
        if t.embed_user_GMF.embedding_dim != s.embed_user_GMF.embedding_dim:
            self.adaptation_layers["gmf_features"] = nn.Linear(s.embed_user_GMF.embedding_dim,
                                                               t.embed_user_GMF.embedding_dim)
        t_in = t.embed_user_MLP.embedding_dim + t.embed_item_MLP.embedding_dim
        s_in = s.embed_user_MLP.embedding_dim + s.embed_item_MLP.embedding_dim
        if t_in != s_in:
            self.adaptation_layers["mlp_input"] = nn.Linear(s_in, t_in)

    def extract_features(self, model, user, item):
        return _embedding_features(model, user, item)

    def matched_keys(self):
        """Feature names that enter the loss: same shape, or an adapter exists (feature.py:88-108)."""
        t, s = self.teacher_model, self.student_model
        keys = []
        if t.embed_user_GMF.embedding_dim == s.embed_user_GMF.embedding_dim or "gmf_features" in self.adaptation_layers:
            keys.append("gmf_features")
        keys.append("mlp_input")   # equal widths or adapter: always one of the two
        tl, sl = t.linears(), s.linears()
        for k in range(min(len(tl), len(sl))):
            if tl[k].out_features == sl[k].out_features:
                keys += [f"mlp_linear_{k}", f"mlp_relu_{k}"]
        return keys

    def feature_matching_loss(self, teacher_features, student_features):
        total, count = 0, 0
        for key in teacher_features:
            if key not in student_features:
                continue
            tf, sf = teacher_features[key], student_features[key]
            if tf.shape != sf.shape:
                if key not in self.adaptation_layers:
                    continue            # the reference prints a warning per batch and skips (feature.py:104-106)
                sf = self.adaptation_layers[key](sf)
            total = total + F.mse_loss(sf, tf)
            count += 1
        if count == 0:
            return torch.tensor(0.0, device=next(iter(teacher_features.values())).device)
        return total / count

    def forward(self, user, item, label):
        with torch.no_grad():
            teacher_features = self.extract_features(self.teacher_model, user, item)
        teacher_logits, student_logits = self._logits(user, item)
        student_features = self.extract_features(self.student_model, user, item)
        remaining = max(0, 1 - self.alpha - self.beta)
        logit_loss = _KDLoss.apply(student_logits, teacher_logits, label, self.alpha, remaining, self.temperature, 1)
        return logit_loss + self.beta * self.feature_matching_loss(teacher_features, student_features)

    def fused_spec(self) -> dict:
        keys = self.matched_keys()
        tower = [k for k in keys if k.startswith("mlp_linear") or k.startswith("mlp_relu")]
        if tower:
            raise NotImplementedError(
                f"FeatureDistillation: tower activations {tower} have equal shapes in teacher and student; the fused "
                "path matches the embedding-level features only — run this pair through the autograd path "
                "(loss = distillation(user, item, label); loss.backward())")
        feats = []
        for kind, key in enumerate(("gmf_features", "mlp_input")):
            if key in keys:
                ad = self.adaptation_layers[key] if key in self.adaptation_layers else None
                feats.append({"kind": kind, "weight": self.beta / len(keys),
                              "w": None if ad is None else ad.weight.detach().contiguous(),
                              "b": None if ad is None else ad.bias.detach().contiguous()})
        return {"w_task": self.alpha, "w_kd": max(0, 1 - self.alpha - self.beta), "kd_mode": 1,
                "temperature": float(self.temperature), "features": feats}


def _attention_map(features):
    """attention.py:16-28: softmax over the batch of the L2 norm of L2-normalised rows."""
    features_norm = F.normalize(features, p=2, dim=-1)
    attention = torch.norm(features_norm, p=2, dim=-1, keepdim=True)
    return F.softmax(attention, dim=0)


def _attention_features(model, user, item):
    out = {}
    if hasattr(model, "embed_user_GMF"):
        out["gmf_attention"] = _attention_map(model.embed_user_GMF(user) * model.embed_item_GMF(item))
    if hasattr(model, "embed_user_MLP"):
        out["mlp_attention"] = _attention_map(torch.cat((model.embed_user_MLP(user), model.embed_item_MLP(item)), -1))
    return out


def _attention_transfer_loss(teacher_attention, student_attention):
    """attention.py:50-79: KL between the (re-normalised) attention vectors, mean over the matched keys."""
    total, count, eps = 0, 0, 1e-8
    for key in teacher_attention:
        if key not in student_attention:
            continue
        t = teacher_attention[key].view(-1) + eps
        s = student_attention[key].view(-1) + eps
        t, s = t / t.sum(), s / s.sum()
        total = total + F.kl_div(torch.log(s), t, reduction="batchmean")
        count += 1
    return total / max(count, 1)


class AttentionDistillation(BaseDistillation):
    def __init__(self, teacher_model, student_model, temperature=2.0, alpha=0.5, gamma=0.2):
        super().__init__(teacher_model, student_model, temperature, alpha)
        self.gamma = gamma

    def compute_attention_map(self, features):
        return _attention_map(features)

    def extract_attention_features(self, model, user, item):
        return _attention_features(model, user, item)

    def attention_transfer_loss(self, teacher_attention, student_attention):
        return _attention_transfer_loss(teacher_attention, student_attention)

    def forward(self, user, item, label):
        with torch.no_grad():
            teacher_attention = self.extract_attention_features(self.teacher_model, user, item)
        teacher_logits, student_logits = self._logits(user, item)
        student_attention = self.extract_attention_features(self.student_model, user, item)
        logit_loss = _KDLoss.apply(student_logits, teacher_logits, label, self.alpha, 1 - self.alpha - self.gamma,
                                   self.temperature, 1)
        return logit_loss + self.gamma * self.attention_transfer_loss(teacher_attention, student_attention)

    def fused_spec(self) -> dict:
        # the transfer term is a constant (~1e-9) with zero gradient: the fused path leaves it out
        return {"w_task": self.alpha, "w_kd": 1 - self.alpha - self.gamma, "kd_mode": 1,
                "temperature": float(self.temperature), "features": []}


class UnifiedDistillation(FeatureDistillation):
    """alpha * BCE + (1 - alpha - beta - gamma) * KD_soft + beta * feature matching + gamma * attention
    transfer.  The reference's own unified.py is empty at HEAD (its train_student.py cannot even be
    imported); this is the combination its CLI flags (--alpha --beta --gamma) describe."""

    def __init__(self, teacher_model, student_model, temperature=2.0, alpha=0.5, beta=0.3, gamma=0.2):
        super().__init__(teacher_model, student_model, temperature, alpha, beta)
        self.gamma = gamma

    def forward(self, user, item, label):
        with torch.no_grad():
            teacher_features = self.extract_features(self.teacher_model, user, item)
            t_att = _attention_features(self.teacher_model, user, item)
        teacher_logits, student_logits = self._logits(user, item)
        student_features = self.extract_features(self.student_model, user, item)
        s_att = _attention_features(self.student_model, user, item)
        remaining = max(0, 1 - self.alpha - self.beta - self.gamma)
        logit_loss = _KDLoss.apply(student_logits, teacher_logits, label, self.alpha, remaining, self.temperature, 1)
        return (logit_loss + self.beta * self.feature_matching_loss(teacher_features, student_features)
                + self.gamma * _attention_transfer_loss(t_att, s_att))

    def fused_spec(self) -> dict:
        spec = super().fused_spec()
        spec["w_kd"] = max(0, 1 - self.alpha - self.beta - self.gamma)
        return spec
