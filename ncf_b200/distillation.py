"""Teacher-score (response) knowledge distillation with the reference's class names
(reference src/distillation/base.py:6-50, response.py:6-32).

`ResponseDistillation(teacher, student, temperature, alpha).forward(user, item, label)` returns the
scalar `alpha * BCE(student, label) + (1 - alpha) * mse(student, teacher)`; the temperature is
accepted and unused on this path exactly as in the reference (response.py:28-32).  The loss value
comes from the ncf_loss_grad kernel; under autograd the student's backward is the fused kernel.
Training scripts use `FusedTrainStep(student, teacher=teacher, alpha=...)`, which runs the same
arithmetic without autograd.  Feature / attention distillation are out of scope.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _KDLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student_logits, teacher_logits, label, alpha):
        acc = torch.zeros(1, dtype=torch.float64, device=student_logits.device)
        dl = torch.empty_like(student_logits)
        ops.loss_grad(student_logits.contiguous(), label.contiguous().float(),
                      None if teacher_logits is None else teacher_logits.contiguous(), float(alpha), acc, dl)
        ctx.save_for_backward(dl)
        return acc.to(torch.float32).reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (dl,) = ctx.saved_tensors
        return dl * grad_out, None, None, None


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

    def task_loss(self, predictions, labels):
        return _KDLoss.apply(predictions, None, labels, 1.0)

    def forward(self, user, item, label):
        raise NotImplementedError("Subclasses must implement forward method")


class ResponseDistillation(BaseDistillation):
    def forward(self, user, item, label):
        with torch.no_grad():
            teacher_logits = self.teacher_model(user, item)
        student_logits = self.student_model(user, item)
        return self.combined_loss(teacher_logits, student_logits, label)

    def combined_loss(self, teacher_logits, student_logits, labels):
        return _KDLoss.apply(student_logits, teacher_logits, labels, self.alpha)
