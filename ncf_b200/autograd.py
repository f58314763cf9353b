"""Autograd compatibility path: lets an UNMODIFIED reference loop
(`loss = criterion(model(user, item), label); loss.backward(); optimizer.step()`,
reference scripts/train_neumf.py:111-115) run on the fused kernels.

forward  = ncf_forward (one kernel);
backward = ncf_backward (one fused kernel + the tower weight-gradient GEMM) accumulating straight
into the parameters' `.grad` tensors, which for the embedding tables are dense zero-initialised
[rows, dim] buffers exactly like autograd's `embedding_dense_backward` output — so torch.optim.Adam
/ SGD see what they would see with the reference model.  This path keeps the reference's dense
optimiser cost; `ncf_b200.trainer.FusedTrainStep` is the fast path.
"""
from __future__ import annotations

import torch

from . import _lib, ops


class _NCFFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, module, user, item):
        ctx.module = module
        ctx.save_for_backward(user, item)
        return ops.forward(module.abi_struct(), user, item)

    @staticmethod
    def backward(ctx, grad_out):
        module = ctx.module
        user, item = ctx.saved_tensors
        dev = user.device
        mt = module.abi_type()
        gmf, mlp = mt != _lib.NCF_MLP, mt != _lib.NCF_GMF

        def dense_grad(p, used):
            if not used or not p.requires_grad:
                return None
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            elif not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
            return p.grad

        g_ug = dense_grad(module.embed_user_GMF.weight, gmf)
        g_ig = dense_grad(module.embed_item_GMF.weight, gmf)
        g_um = dense_grad(module.embed_user_MLP.weight, mlp)
        g_im = dense_grad(module.embed_item_MLP.weight, mlp)
        # Frozen tables still need somewhere to scatter: a throw-away buffer.
        scratch = lambda p: torch.zeros_like(p)
        B = user.numel()
        bufs = ops.GradBuffers(
            g_ug if g_ug is not None or not gmf else scratch(module.embed_user_GMF.weight),
            g_ig if g_ig is not None or not gmf else scratch(module.embed_item_GMF.weight),
            g_um if g_um is not None or not mlp else scratch(module.embed_user_MLP.weight),
            g_im if g_im is not None or not mlp else scratch(module.embed_item_MLP.weight),
            torch.zeros(ops.tower_param_count(mt, module.factor_num, module.num_layers), device=dev),
            torch.zeros(module.user_num, dtype=torch.int32, device=dev),
            torch.zeros(module.item_num, dtype=torch.int32, device=dev),
            torch.empty(min(B, module.user_num), dtype=torch.int64, device=dev),
            torch.empty(min(B, module.item_num), dtype=torch.int64, device=dev),
            torch.zeros(2, dtype=torch.int32, device=dev))
        m = module.abi_struct()
        ws = torch.empty(ops.train_workspace_bytes(m, B), dtype=torch.uint8, device=dev)
        ops.backward(m, bufs.struct(), user, item, grad_out.contiguous().to(torch.float32), ws)

        # hand the flat tower gradient back to the individual parameters
        flat, off = bufs.g_tower, 0
        tower = []
        if mlp:
            for lin in module.linears():
                tower += [lin.weight, lin.bias]
        else:
            off = flat.numel() - (module.predict_layer.weight.numel() + 1)
        tower += [module.predict_layer.weight, module.predict_layer.bias]
        for p in tower:
            n = p.numel()
            if p.requires_grad:
                piece = flat[off:off + n].view_as(p)
                p.grad = piece.clone() if p.grad is None else p.grad.add_(piece)
            off += n
        return None, None, None, None


def ncf_apply(module, user, item):
    anchor = next(p for p in module.parameters() if p.requires_grad)
    return _NCFFunction.apply(anchor, module, user.contiguous(), item.contiguous())
