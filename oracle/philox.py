"""ORACLE — test infrastructure, not product code.

CPU restatement (numpy, bit-exact) of the integer side of the hot path: the Philox4x32-10
counter RNG, the negative sampler with rejection against the observed pairs, the keyed epoch
permutation, and the sorted CSR.  The reference's own sampler (src/data/datasets.py:53-69) draws
from numpy's global MT19937 stream sequentially, which no parallel sampler can reproduce
(SURVEY.md §7 H6); this file restates the reference's *semantics* (uniform over [0, item_num),
redraw while the pair is observed, independent draws, positives-then-negatives sample order of
datasets.py:68) on the RNG the CUDA sampler uses, so the two can be compared bit for bit.  The
statistical agreement with the reference sampler is tested separately against goldens.

Philox4x32-10 is the published algorithm of Salmon et al., "Parallel Random Numbers: As Easy as
1, 2, 3" (SC'11); philox4x32_10() below is pinned against the Random123 known-answer vectors in
tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
U64 = np.uint64
M0, M1 = U64(0xD2511F53), U64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = U64(0xFFFFFFFF)
MAX_ATTEMPTS = 1 << 16


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Counter words / key words are uint32 arrays (broadcastable)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=U64) & MASK32 for c in (c0, c1, c2, c3))
    k0 = np.asarray(k0, dtype=U64) & MASK32
    k1 = np.asarray(k1, dtype=U64) & MASK32
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> U64(32), p0 & MASK32
        hi1, lo1 = p1 >> U64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0), lo1, (hi0 ^ c3 ^ k1), lo0
        k0 = (k0 + U64(W0)) & MASK32
        k1 = (k1 + U64(W1)) & MASK32
    return c0.astype(U32), c1.astype(U32), c2.astype(U32), c3.astype(U32)


def csr_build(pos_user, pos_item, user_num):
    """Sorted-column CSR of the observed pairs (replaces the dok fill of datasets.py:20-24)."""
    order = np.lexsort((pos_item, pos_user))
    rowptr = np.zeros(user_num + 1, dtype=np.int64)
    np.add.at(rowptr, pos_user + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, pos_item[order].astype(np.int32)


def _observed(rowptr, col, users, items):
    """Vectorised membership test (u, j) in train_mat."""
    out = np.zeros(users.shape[0], dtype=bool)
    lo, hi = rowptr[users], rowptr[users + 1]
    # search in the globally sorted key space user * BIG + item
    for i in range(users.shape[0]):
        a, b = lo[i], hi[i]
        k = np.searchsorted(col[a:b], items[i])
        out[i] = k < b - a and col[a + k] == items[i]
    return out


def sample_neg(rowptr, col, pos_user, num_ng, item_num, seed, epoch, p_offset=0):
    """out[p*num_ng + t]: see include/ncf_b200.h (ncf_sample_neg) for the exact draw rule."""
    P = pos_user.shape[0]
    n = P * num_ng
    out = np.zeros(n, dtype=np.int64)
    if n == 0:
        return out
    g = np.arange(n, dtype=np.uint64)
    sid = g + U64(p_offset * num_ng)
    users = np.repeat(pos_user, num_ng)
    # key-space trick for a vectorised membership test
    big = np.int64(item_num)
    keys = np.repeat(np.arange(rowptr.shape[0] - 1, dtype=np.int64), np.diff(rowptr)) * big + col
    pending = np.arange(n)
    k0, k1 = U32(seed & 0xFFFFFFFF), U32((seed >> 32) & 0xFFFFFFFF)
    for a in range(MAX_ATTEMPTS):
        if pending.size == 0:
            break
        s = sid[pending]
        w = philox4x32_10(s & MASK32, s >> U64(32), U32(a >> 2), U32(epoch & 0xFFFFFFFF), k0, k1)[a & 3]
        j = ((w.astype(U64) * U64(item_num)) >> U64(32)).astype(np.int64)
        out[pending] = j
        q = users[pending] * big + j
        pos = np.searchsorted(keys, q)
        hit = (pos < keys.shape[0]) & (keys[np.minimum(pos, keys.shape[0] - 1)] == q)
        pending = pending[hit]
    out[pending] = -1   # every draw was an observed pair (the reference loops forever): reported, not trained on
    bad = (users < 0) | (users >= rowptr.shape[0] - 1)
    out[bad] = -1
    return out


def fmix32(h):
    h = np.asarray(h, dtype=U64) & MASK32
    h ^= h >> U64(16)
    h = (h * U64(0x85EBCA6B)) & MASK32
    h ^= h >> U64(13)
    h = (h * U64(0xC2B2AE35)) & MASK32
    h ^= h >> U64(16)
    return h


def shuffle_perm(S, seed, epoch, q):
    """perm(q) for positions q (array): balanced 6-round Feistel over 2^(2*half) >= S with cycle
    walking, round keys = Philox words of (seed, epoch).  Mirrors csrc/rng.cuh."""
    bits = 2
    while (1 << bits) < S:
        bits += 1
    half = (bits + 1) // 2
    mask = U64((1 << half) - 1)
    k0, k1 = U32(seed & 0xFFFFFFFF), U32((seed >> 32) & 0xFFFFFFFF)
    a = philox4x32_10(U32(0), U32(0), U32(0x53485546), U32(epoch & 0xFFFFFFFF), k0, k1)
    b = philox4x32_10(U32(1), U32(0), U32(0x53485546), U32(epoch & 0xFFFFFFFF), k0, k1)
    rk = [U64(int(a[0])), U64(int(a[1])), U64(int(a[2])), U64(int(a[3])), U64(int(b[0])), U64(int(b[1]))]
    x = np.asarray(q, dtype=U64).copy()
    todo = np.arange(x.shape[0])
    while todo.size:
        v = x[todo]
        Lh = (v >> U64(half)) & mask
        R = v & mask
        for r in range(6):
            t = Lh ^ (fmix32((R + rk[r]) & MASK32) & mask)
            Lh, R = R, t
        v = (Lh << U64(half)) | R
        x[todo] = v
        todo = todo[v >= U64(S)]
    return x.astype(np.int64)


def shuffle_epoch(pos_user, pos_item, neg_item, num_ng, seed, epoch, q_begin, count):
    """Stream positions [q_begin, q_begin+count) -> (user, item, label): sample s < P is
    positive s, otherwise negative s-P of positive (s-P)//num_ng (datasets.py:65-69)."""
    P = pos_user.shape[0]
    S = P * (1 + num_ng)
    s = shuffle_perm(S, seed, epoch, np.arange(q_begin, q_begin + count))
    is_pos = s < P
    n = np.where(is_pos, 0, s - P)
    user = np.where(is_pos, pos_user[np.minimum(s, P - 1)], pos_user[n // max(num_ng, 1)])
    item = np.where(is_pos, pos_item[np.minimum(s, P - 1)],
                    neg_item[n] if num_ng > 0 else 0)
    return user.astype(np.int64), item.astype(np.int64), is_pos.astype(np.float32)
