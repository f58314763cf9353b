"""The epoch loop shared by the four entry points (scripts/train_neumf.py, pretrain.py,
train_teacher.py, train_student.py): the reference copies the same loop into each script
(reference scripts/train_neumf.py:98-144, pretrain.py:60-106, train_teacher.py:54-100,
train_student.py:141-181).  Per epoch: ng_sample on the GPU -> shuffled windows of fused steps
(CUDA-graph replayed) -> Adam flush -> one batched evaluation launch -> checkpoint on HR gain.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np
import torch

from .metrics import evaluate
from .models import NCF
from .trainer import EpochStream, FusedTrainStep, train_epoch


@dataclass
class TrainResult:
    best_hr: float = 0.0
    best_ndcg: float = 0.0
    best_epoch: int = 0
    best_loss: float = 0.0
    history: list = field(default_factory=list)


def load_dataset(device, synthetic: Optional[str] = None):
    """(train pairs [P,2] tensor, test users [n], test candidates [n,C], user_num, item_num, train_mat).
    Reads the reference's files through datasets.load_all() unless a synthetic shape is named."""
    if synthetic:
        from .synth import make_interactions
        d = make_interactions(synthetic, device=device)
        train = torch.stack([d.pos_user, d.pos_item], 1)
        return train, d.test_users, d.test_cands, d.user_num, d.item_num, None
    from .datasets import load_all_device
    from .config import config
    train, users, cands, user_num, item_num = load_all_device(device, config.test_num_ng + 1)
    return train, users, cands, user_num, item_num, None


def fit(model: NCF, train_pairs: torch.Tensor, test_users: torch.Tensor, test_cands: torch.Tensor,
        *, epochs: int, batch_size: int, lr: float, num_ng: int, top_k: int,
        optimizer: str = "adam", teacher: Optional[NCF] = None, alpha: float = 0.5, seed: int = 0,
        on_epoch: Optional[Callable] = None, on_best: Optional[Callable] = None,
        use_graph: bool = True, distillation=None) -> TrainResult:
    """`teacher` + `alpha`: response KD; `distillation`: any ncf_b200.distillation object (its teacher,
    weights and adapters are used; reference scripts/train_student.py:96-127)."""
    device = next(model.parameters()).device
    ts = FusedTrainStep(model, optimizer=optimizer, lr=lr, max_batch=batch_size, teacher=teacher,
                        alpha=alpha, distillation=distillation)
    stream = EpochStream(train_pairs[:, 0].contiguous(), train_pairs[:, 1].contiguous(),
                         model.user_num, model.item_num, num_ng, seed=seed)
    res = TrainResult()
    cache = {}
    for epoch in range(epochs):
        model.train()
        t0 = time.time()
        avg_loss, _ = train_epoch(ts, stream, epoch, batch_size, use_graph=use_graph, cache=cache)
        ts.flush()                      # lazy Adam rows -> dense-equivalent state before reading weights
        model.eval()
        with torch.no_grad():
            ev = evaluate(model, test_users, test_cands, top_k)
        HR, NDCG = ev.lists()
        hr, ndcg = float(np.mean(HR)), float(np.mean(NDCG))
        elapsed = time.time() - t0
        res.history.append({"epoch": epoch + 1, "loss": avg_loss, "hr": hr, "ndcg": ndcg, "time": elapsed})
        if on_epoch:
            on_epoch(epoch, avg_loss, hr, ndcg, elapsed)
        if hr > res.best_hr:
            res.best_hr, res.best_ndcg, res.best_epoch, res.best_loss = hr, ndcg, epoch, avg_loss
            if on_best:
                on_best(model)
    return res


def count_parameters(model) -> int:
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
