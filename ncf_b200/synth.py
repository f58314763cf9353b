"""Seeded synthetic MovieLens-shaped interaction data.

There is no network and the reference ships no data (its .gitignore:28-31), so every workload is
synthetic: per-user interaction counts are log-normal, items are drawn without replacement with
probability proportional to exp(<z_u, z_i>/tau + b_i) (planted low-rank structure + Zipf
popularity, via the Gumbel top-k trick), one interaction per user is held out, and 99 evaluation
negatives are drawn uniformly from the user's non-interacted items and stored in ascending order
like the reference writes them (src/data/preprocessing.py:132).  Data generation is not part of
the hot path; it uses torch ops on whichever device it is given.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

SHAPES = {
    # name: (user_num, item_num, total interactions incl. the held-out one per user, seed)
    "tiny": (400, 300, 14_000, 20250602),        # test-sized
    "ml100k": (943, 1682, 100_000, 20250603),
    "ml1m": (6040, 3706, 1_000_209, 20250604),
    "ml20m": (138_493, 26_744, 20_000_263, 20250605),
    "big": (10_000_000, 1_000_000, 1_000_000_000, 20250606),   # table shape only (bench draws batches directly)
}


@dataclass
class Interactions:
    user_num: int
    item_num: int
    pos_user: torch.Tensor    # int64 [P]   training positives
    pos_item: torch.Tensor    # int64 [P]
    test_users: torch.Tensor  # int64 [n]
    test_cands: torch.Tensor  # int64 [n, 1 + n_neg]; column 0 = held-out item


def make_interactions(name_or_shape, device="cpu", n_test_neg: int = 99, dim: int = 16,
                      tau: float = 2.0, seed=None, chunk: int = 2048) -> Interactions:
    if isinstance(name_or_shape, str):
        U, I, total, default_seed = SHAPES[name_or_shape]
    else:
        U, I, total = name_or_shape
        default_seed = 1
    seed = default_seed if seed is None else seed
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    rnd = lambda *s: torch.rand(*s, generator=g, device=dev)
    zu = torch.randn(U, dim, generator=g, device=dev)
    zi = torch.randn(I, dim, generator=g, device=dev)
    pop_rank = torch.randperm(I, generator=g, device=dev).float() + 1.0
    bias = -torch.log(pop_rank)
    # per-user counts: lognormal(4.6, 1.0) clipped to [20, I//2], rescaled to the target total
    n_u = torch.exp(4.6 + torch.randn(U, generator=g, device=dev))
    hi = max(21, min(I // 2, I - n_test_neg - 1))
    n_u = n_u.clamp(20, hi)
    for _ in range(8):
        n_u = (n_u * (total / n_u.sum())).clamp(20, hi)
    n_u = n_u.round().long()
    pu, pi, tu, tc = [], [], [], []
    ar = torch.arange(I, device=dev)
    for lo in range(0, U, chunk):
        hi_u = min(U, lo + chunk)
        k = n_u[lo:hi_u]
        score = zu[lo:hi_u] @ zi.T / tau + bias
        gumbel = -torch.log(-torch.log(rnd(hi_u - lo, I).clamp_min(1e-20)))
        order = torch.argsort(score + gumbel, dim=1, descending=True)
        taken = ar[None, :] < k[:, None]                      # first k items of each row's order
        held = order[:, 0]                                    # the held-out interaction
        train_mask = taken.clone()
        train_mask[:, 0] = False
        rows = torch.arange(lo, hi_u, device=dev)[:, None].expand(-1, I)
        pu.append(rows[train_mask])
        pi.append(order[train_mask])
        # evaluation negatives: uniform over the items the user never touched
        seen = torch.zeros(hi_u - lo, I, dtype=torch.bool, device=dev)
        seen.scatter_(1, order, taken)
        r = rnd(hi_u - lo, I).masked_fill(seen, -1.0)
        neg = torch.topk(r, n_test_neg, dim=1).indices.sort(dim=1).values
        tu.append(torch.arange(lo, hi_u, device=dev))
        tc.append(torch.cat([held[:, None], neg], dim=1))
    return Interactions(U, I, torch.cat(pu).contiguous(), torch.cat(pi).contiguous(),
                        torch.cat(tu).contiguous(), torch.cat(tc).contiguous())
