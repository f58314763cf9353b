"""`load_all` and `NCFData` with the reference's signatures (reference src/data/datasets.py:9-83),
backed by arrays and device kernels instead of Python lists and a dok_matrix.

  * load_all(test_num=100) -> (train_data, test_data, user_num, item_num, train_mat): the two files
    are read as bytes, copied to the GPU and parsed there (ncf_text_line_starts / ncf_text_parse_ints:
    a byte-parallel line index, then one thread per line) where the reference loops in Python and
    `eval`s every line (:22-35); train_data / test_data are int64 arrays of [user, item] rows
    (indexable like the reference's list of lists), train_mat is a `TrainMatrix` answering
    `(u, j) in train_mat`.  `load_all_device()` keeps everything on the GPU for the training scripts.
  * NCFData(features, num_item, train_mat=None, num_ng=0, is_training=None).ng_sample() draws the
    negatives on the GPU (CSR rejection + Philox, ncf_sample_neg); `features_fill` / `labels_fill`
    keep the positives-then-negatives layout of datasets.py:65-69; `__len__` / `__getitem__` serve
    a DataLoader exactly like the reference.  The fast path is `epoch_stream()`, which keeps
    everything on the device.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.utils.data as data

from . import _lib
from .config import config


class TrainMatrix:
    """Observed (user, item) pairs; the slice of the dok_matrix API the hot path uses."""

    def __init__(self, pairs: np.ndarray, user_num: int, item_num: int):
        self.pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
        self.shape = (int(user_num), int(item_num))
        self._keys = None

    def _sorted_keys(self):
        if self._keys is None:
            self._keys = np.unique(self.pairs[:, 0] * self.shape[1] + self.pairs[:, 1])
        return self._keys

    def __contains__(self, key):
        u, j = key
        k = int(u) * self.shape[1] + int(j)
        keys = self._sorted_keys()
        pos = np.searchsorted(keys, k)
        return bool(pos < keys.shape[0] and keys[pos] == k)

    def nonzero(self):
        keys = self._sorted_keys()
        return keys // self.shape[1], keys % self.shape[1]

    def __len__(self):
        return int(self._sorted_keys().shape[0])


def parse_train_rating(path) -> np.ndarray:
    """`user\\titem[\\t...]` per line, no header (reference preprocessing.py:142-143) -> [P, 2]."""
    raw = np.loadtxt(path, dtype=np.int64, delimiter="\t", usecols=(0, 1), ndmin=2)
    return np.ascontiguousarray(raw)


def parse_test_negative(path) -> np.ndarray:
    """`(u, pos)\\tneg1\\t...\\tnegK` per line (reference preprocessing.py:131-133) -> flat
    [n*(1+K), 2] rows `[u, item]`, held-out item first — the layout datasets.py:26-35 builds."""
    rows = []
    with open(path, "r") as fd:
        for line in fd:
            line = line.strip()
            if not line:
                continue
            head, *negs = line.split("\t")
            u, pos = head.strip("() ").split(",")
            u = int(u)
            rows.append([u, int(pos)])
            rows.extend([u, int(x)] for x in negs)
    return np.asarray(rows, dtype=np.int64).reshape(-1, 2)


def _file_to_device(path, device) -> torch.Tensor:
    raw = np.fromfile(path, dtype=np.uint8)
    return torch.from_numpy(raw).to(device) if raw.size else torch.empty(0, dtype=torch.uint8, device=device)


def parse_train_rating_device(text: torch.Tensor) -> torch.Tensor:
    """uint8 CUDA tensor of a u.train.rating file -> int64 [P, 2] (user, item) on the device."""
    from . import ops
    pairs, status = ops.text_parse_ints(text, 2, exact=False)     # further columns (rating, timestamp) are ignored
    if status & 1:
        raise _lib.NcfError("u.train.rating: a line holds fewer than two integers")
    return pairs


def parse_test_negative_device(text: torch.Tensor, test_num: int = 100):
    """uint8 CUDA tensor of a u.test.negative file -> (users int64 [n], cands int64 [n, test_num]) with the
    held-out item in column 0.  A line with another number of candidates is an error: the reference
    silently shifts every later user in that case."""
    from . import ops
    vals, status = ops.text_parse_ints(text, 1 + test_num, exact=True)
    if status:
        raise _lib.NcfError(f"u.test.negative: every line must hold (user, item) + {test_num - 1} negatives "
                            f"({'fewer' if status & 1 else 'more'} found on some line)")
    return vals[:, 0].contiguous(), vals[:, 1:].contiguous()


def load_all_device(device="cuda", test_num=100):
    """The files of load_all(), parsed on the GPU and left there:
    (train pairs int64 [P, 2], test users [n], test candidates [n, test_num], user_num, item_num)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.NcfError("load_all_device parses on the GPU; use load_all(host=True) for host-side tooling")
    train = parse_train_rating_device(_file_to_device(config.train_rating, dev))
    users, cands = parse_test_negative_device(_file_to_device(config.test_negative, dev), test_num)
    mx = train.max(dim=0).values
    return train, users, cands, int(mx[0]) + 1, int(mx[1]) + 1


def load_all(test_num=100, host=False):
    """Reference signature (src/data/datasets.py:9).  host=False (default): GPU ingestion, needs a CUDA
    device.  host=True: the numpy parser, for host-only tooling and CPU tests of the file writers."""
    if host:
        train_data = parse_train_rating(config.train_rating)
        test_data = parse_test_negative(config.test_negative)
    else:
        if not torch.cuda.is_available():
            raise _lib.NcfError("load_all parses the data files on the GPU and no CUDA device is visible "
                                "(load_all(host=True) is the host-side parser)")
        train, users, cands, _, _ = load_all_device("cuda", test_num)
        train_data = train.cpu().numpy()
        C = cands.shape[1]
        test_data = torch.stack([users[:, None].expand(-1, C), cands], 2).reshape(-1, 2).cpu().numpy()
    user_num = int(train_data[:, 0].max()) + 1
    item_num = int(train_data[:, 1].max()) + 1
    train_mat = TrainMatrix(train_data, user_num, item_num)
    return train_data, test_data, user_num, item_num, train_mat


def write_reference_files(inter, train_path, test_negative_path):
    """Writes a synthetic `Interactions` in the reference's on-disk format."""
    pu, pi = inter.pos_user.cpu().numpy(), inter.pos_item.cpu().numpy()
    np.savetxt(train_path, np.stack([pu, pi], 1), fmt="%d", delimiter="\t")
    users, cands = inter.test_users.cpu().numpy(), inter.test_cands.cpu().numpy()
    lines = [f"({u}, {row[0]})\t" + "\t".join(str(x) for x in row[1:]) for u, row in zip(users, cands)]
    with open(test_negative_path, "w") as fd:
        fd.write("\n".join(lines))  # no trailing newline, like preprocessing.py:152-153


class NCFData(data.Dataset):
    def __init__(self, features, num_item, train_mat=None, num_ng=0, is_training=None,
                 device=None, seed=0):
        super().__init__()
        self.features_ps = np.ascontiguousarray(np.asarray(features, dtype=np.int64)).reshape(-1, 2)
        self.num_item = int(num_item)
        self.train_mat = train_mat
        self.num_ng = int(num_ng)
        self.is_training = is_training
        self.labels = np.zeros(self.features_ps.shape[0], dtype=np.int64)
        self.device = device
        self.seed = int(seed)
        self._epoch = -1
        self._stream = None
        self.features_fill = None
        self.labels_fill = None

    # -- device side ----------------------------------------------------------------------------
    def epoch_stream(self, device=None):
        """The on-device sampler/shuffler over this dataset's positives."""
        if self._stream is None:
            from .trainer import EpochStream
            dev = torch.device(device or self.device or "cuda")
            if dev.type != "cuda":
                raise _lib.NcfError("negative sampling runs on the GPU; there is no CPU path")
            pos = torch.from_numpy(self.features_ps).to(dev)
            if self.train_mat is not None and hasattr(self.train_mat, "pairs"):
                seen = torch.from_numpy(self.train_mat.pairs).to(dev)
                user_num = self.train_mat.shape[0]
            else:
                seen, user_num = pos, int(self.features_ps[:, 0].max()) + 1
            self._stream = EpochStream(pos[:, 0], pos[:, 1], user_num, self.num_item, self.num_ng,
                                       seed=self.seed, observed=(seen[:, 0], seen[:, 1]))
        return self._stream

    def ng_sample(self):
        assert self.is_training, "no need to sampling when testing"
        st = self.epoch_stream()
        self._epoch += 1
        st.begin_epoch(self._epoch)
        neg = st.neg_item.cpu().numpy()
        users = np.repeat(self.features_ps[:, 0], self.num_ng)
        self.features_ng = np.stack([users, neg], 1)
        self.features_fill = np.concatenate([self.features_ps, self.features_ng], 0)
        self.labels_fill = np.concatenate([np.ones(self.features_ps.shape[0], dtype=np.int64),
                                           np.zeros(self.features_ng.shape[0], dtype=np.int64)])

    def __len__(self):
        return (self.num_ng + 1) * len(self.labels)

    def __getitem__(self, idx):
        features = self.features_fill if self.is_training else self.features_ps
        labels = self.labels_fill if self.is_training else self.labels
        return int(features[idx][0]), int(features[idx][1]), int(labels[idx])
