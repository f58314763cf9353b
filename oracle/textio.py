"""ORACLE — test infrastructure, not product code.

CPU restatement of the reference's on-disk formats and of its leave-one-out preprocessing, pinned by
tests/golden/load_all_small.npz and preprocess_small.npz (outputs of the reference itself,
oracle/make_golden_r2.py).

  parse_train_rating / parse_test_negative   reference src/data/datasets.py:9-36 (`load_all`):
      u.train.rating = `user<TAB>item` per line (pandas read_csv, usecols=[0, 1]);
      u.test.negative = `(user,pos)<TAB>neg1<TAB>...` per line, each line expanded to [user, item] rows with
      the held-out item first.
  temporal_split                              reference src/data/preprocessing.py:23-90: sort by (user,
      timestamp), last interaction of every user with >= 2 interactions = test, everything else = train in
      that order.  Ties in a user's timestamps keep file order here (the reference's order under ties is
      whatever pandas' unstable per-group sort produces; the goldens are tie-free).
  eval_negatives                              reference preprocessing.py:92-135: up to `num_negatives` DISTINCT
      items per test user, uniform over [0, num_items), none of the user's train or test items, at most
      10 * num_negatives draws, written in ascending order.  The reference draws from numpy's global
      MT19937; this restatement (and the kernel) draws from Philox4x32-10 keyed like the training sampler:
      draw a of test row r uses word a % 4 of Philox(counter = {r, 0, a / 4, 0x4e454753 'NEGS'}, key = seed).
"""
from __future__ import annotations

import re

import numpy as np

from . import philox as ph


def parse_train_rating(text: bytes) -> np.ndarray:
    rows = [ln.split("\t")[:2] for ln in text.decode().splitlines() if ln.strip()]
    return np.array(rows, dtype=np.int64).reshape(-1, 2)


def parse_test_negative(text: bytes) -> np.ndarray:
    """-> [n_lines * (1 + K), 2] rows [user, item], held-out item first (datasets.py:26-35)."""
    out = []
    for ln in text.decode().split("\n"):
        if ln == "":
            break                                   # `while line != None and line != ''`
        ints = [int(x) for x in re.findall(r"\d+", ln)]
        u = ints[0]
        out += [[u, it] for it in ints[1:]]
    return np.array(out, dtype=np.int64).reshape(-1, 2)


def temporal_split(user, item, ts):
    order = np.lexsort((np.arange(user.shape[0]), ts, user))       # stable: user, then timestamp, then file order
    u, i = user[order], item[order]
    last = np.ones(u.shape[0], dtype=bool)
    last[:-1] = u[1:] != u[:-1]
    first = np.ones(u.shape[0], dtype=bool)
    first[1:] = u[1:] != u[:-1]
    is_test = last & ~first                                          # single-interaction users stay in train
    train = np.stack([u[~is_test], i[~is_test]], 1)
    test = np.stack([u[is_test], i[is_test]], 1)
    return train.astype(np.int64), test.astype(np.int64)


NEGS_TAG = 0x4E454753


def eval_negatives(rowptr, col, test_user, num_items, num_negatives, seed):
    """rowptr / col: sorted CSR over ALL of a user's items (train and test).  Returns
    (negatives [n, num_negatives] ascending, padded with -1, count [n])."""
    n = test_user.shape[0]
    out = np.full((n, num_negatives), -1, dtype=np.int64)
    cnt = np.zeros(n, dtype=np.int64)
    k0, k1 = ph.U32(seed & 0xFFFFFFFF), ph.U32((seed >> 32) & 0xFFFFFFFF)
    for r in range(n):
        u = int(test_user[r])
        mine = col[rowptr[u]:rowptr[u + 1]]
        got = []
        for a in range(10 * num_negatives):
            if len(got) == num_negatives:
                break
            w = ph.philox4x32_10(ph.U32(r & 0xFFFFFFFF), ph.U32(r >> 32), ph.U32(a >> 2), ph.U32(NEGS_TAG), k0, k1)[a & 3]
            j = (int(w) * int(num_items)) >> 32
            pos = np.searchsorted(mine, j)
            if (pos < mine.shape[0] and mine[pos] == j) or j in got:
                continue
            got.append(j)
        got.sort()
        out[r, :len(got)] = got
        cnt[r] = len(got)
    return out, cnt
