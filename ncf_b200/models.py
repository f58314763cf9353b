"""`NCF` — the GMF / MLP / NeuMF module with the reference's constructor, attribute names and
state_dict layout (reference src/ncf/models.py:4-118), computing through the sm_100a kernels.

What is kept from the reference so that it drops in:
  * `NCF(user_num, item_num, factor_num, num_layers, dropout, model_type)`;
  * attributes `embed_{user,item}_{GMF,MLP}` (nn.Embedding), `MLP_layers` (nn.Sequential of
    Dropout/Linear/ReLU triples, so checkpoints keep the keys `MLP_layers.{3k+1}.{weight,bias}`),
    `predict_layer`, `model_type`, `dropout`;
  * `forward(user, item) -> logits[B]` and `load_pretrain_weights(gmf_state, mlp_state)`;
  * same construction + initialisation order, so `torch.manual_seed(s); NCF(...)` yields the same
    initial weights as the reference (models.py:38-46).

What differs: the arithmetic.  `forward` is one fused CUDA kernel (ncf_forward); under autograd it
is a `torch.autograd.Function` whose backward is the fused backward kernel writing straight into
dense-addressed gradient buffers, so an unmodified reference loop (`loss.backward();
optimizer.step()`) still works.  The fast path for training is `ncf_b200.trainer.FusedTrainStep`.
Parameters must live on a CUDA device to compute; there is no CPU forward.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops

_TYPES = ("GMF", "MLP", "NeuMF-end", "NeuMF-pre")


class NCF(nn.Module):
    def __init__(self, user_num, item_num, factor_num, num_layers, dropout, model_type):
        super().__init__()
        if model_type not in _TYPES:
            raise ValueError(f"model_type must be one of {_TYPES}, got {model_type!r}")
        if not 1 <= int(num_layers) <= _lib.NCF_MAX_LAYERS:
            raise ValueError(f"num_layers must be in [1, {_lib.NCF_MAX_LAYERS}]")
        self.model_type = model_type
        self.dropout = dropout
        # "fp32": tower contractions keep fp32 accuracy (3xTF32 / FMA) — the parity mode;
        # "tf32": single-pass TF32 tensor-core math (opt-in, looser stated tolerance).
        self.tower_math = "fp32"
        user_num, item_num = int(user_num), int(item_num)
        factor_num, num_layers = int(factor_num), int(num_layers)
        self.user_num, self.item_num = user_num, item_num
        self.factor_num, self.num_layers = factor_num, num_layers
        mlp_dim = factor_num << (num_layers - 1)

        # Construction order matters for RNG parity with the reference (models.py:11-34).
        self.embed_user_GMF = nn.Embedding(user_num, factor_num)
        self.embed_item_GMF = nn.Embedding(item_num, factor_num)
        self.embed_user_MLP = nn.Embedding(user_num, mlp_dim)
        self.embed_item_MLP = nn.Embedding(item_num, mlp_dim)
        widths = ops.tower_widths(factor_num, num_layers)
        blocks = []
        for w_in, w_out in zip(widths[:-1], widths[1:]):
            blocks += [nn.Dropout(p=dropout), nn.Linear(w_in, w_out), nn.ReLU()]
        self.MLP_layers = nn.Sequential(*blocks)
        predict_size = factor_num if model_type in ("GMF", "MLP") else 2 * factor_num
        self.predict_layer = nn.Linear(predict_size, 1)
        self._init_weight()

    # -- initialisation (reference models.py:38-46) --------------------------------------------
    def _init_weight(self):
        for emb in (self.embed_user_GMF, self.embed_item_GMF, self.embed_user_MLP, self.embed_item_MLP):
            nn.init.normal_(emb.weight, std=0.01)
        for lin in self.linears():
            nn.init.xavier_uniform_(lin.weight)
        nn.init.kaiming_uniform_(self.predict_layer.weight, a=1, nonlinearity="sigmoid")

    def linears(self):
        return [m for m in self.MLP_layers if isinstance(m, nn.Linear)]

    # -- NeuMF-pre initialisation (reference models.py:48-95) ------------------------------------
    def load_pretrain_weights(self, gmf_state, mlp_state):
        """Copies the four tables and the tower from pretrained GMF / MLP state dicts and
        re-draws the predict layer (zero bias); a no-op unless model_type == "NeuMF-pre"."""
        if self.model_type != "NeuMF-pre":
            return
        try:
            with torch.no_grad():
                self.embed_user_GMF.weight.copy_(gmf_state["embed_user_GMF.weight"])
                self.embed_item_GMF.weight.copy_(gmf_state["embed_item_GMF.weight"])
                self.embed_user_MLP.weight.copy_(mlp_state["embed_user_MLP.weight"])
                self.embed_item_MLP.weight.copy_(mlp_state["embed_item_MLP.weight"])
                for idx, layer in enumerate(self.MLP_layers):
                    if not isinstance(layer, nn.Linear):
                        continue
                    wk, bk = f"MLP_layers.{idx}.weight", f"MLP_layers.{idx}.bias"
                    if wk in mlp_state and bk in mlp_state:
                        layer.weight.copy_(mlp_state[wk])
                        layer.bias.copy_(mlp_state[bk])
                    else:  # reference models.py:77-82: missing layer => fresh xavier, zero bias
                        nn.init.xavier_uniform_(layer.weight)
                        nn.init.zeros_(layer.bias)
                nn.init.kaiming_uniform_(self.predict_layer.weight, a=1, nonlinearity="sigmoid")
                nn.init.zeros_(self.predict_layer.bias)
        except Exception as e:  # same error convention as the reference (models.py:91-95)
            raise RuntimeError(f"Failed to load pretrained weights: {e}") from e

    # -- C-ABI view -------------------------------------------------------------------------------
    def abi_type(self) -> int:
        return _lib.MODEL_TYPES[self.model_type]

    def abi_struct(self) -> "_lib.NcfModel":
        tables = (self.embed_user_GMF.weight, self.embed_item_GMF.weight,
                  self.embed_user_MLP.weight, self.embed_item_MLP.weight)
        lin = [(l.weight, l.bias) for l in self.linears()]
        return ops.model_struct(self.abi_type(), self.factor_num, self.num_layers, self.user_num,
                                self.item_num, [t.detach() for t in tables],
                                [(w.detach(), b.detach()) for w, b in lin],
                                (self.predict_layer.weight.detach(), self.predict_layer.bias.detach()),
                                tower_math=self.tower_math)

    def _check_dropout(self):
        if self.training and self.dropout and self.dropout > 0:
            raise NotImplementedError(
                "ncf_b200 kernels implement dropout=0.0 (the reference default, neumf.yaml:11); "
                "call .eval() or construct with dropout=0.0")

    # -- forward (reference models.py:97-118) -------------------------------------------------------
    def forward(self, user, item):
        self._check_dropout()
        user = user.reshape(-1).to(torch.int64)
        item = item.reshape(-1).to(torch.int64)
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if needs_grad:
            from .autograd import ncf_apply
            return ncf_apply(self, user, item)
        return ops.forward(self.abi_struct(), user.contiguous(), item.contiguous())
