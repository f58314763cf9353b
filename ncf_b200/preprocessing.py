"""`LeaveOneOutPreprocessor` with the reference's constructor, method names and output files
(reference src/data/preprocessing.py:10-265), computing on the GPU.

    raw u.data (`user<TAB>item<TAB>rating<TAB>timestamp`)
      -> temporal leave-one-out split: per user, last interaction = test, the rest = train   (:45-90)
      -> 99 evaluation negatives per test user: distinct, never one of the user's items       (:92-135)
      -> u.train.rating, u.test.rating, u.test.negative in the reference's format              (:137-154)

The reference does this with pandas group-bys, Python sets and `np.random.randint`; here the file is
parsed by ncf_text_parse_ints, split by ncf_leave_one_out_split (bucket by user + shared-memory sort of
each user's (timestamp, file position) keys) and the negatives come from ncf_eval_negatives (one warp per
user, Philox draws, CSR rejection).  What differs observably: the negatives follow a Philox stream keyed
by `seed` instead of numpy's global MT19937 (same distribution and contract), ties between equal
timestamps of one user keep file order (the reference's order there is unspecified), and the run
metadata / log files of `save_results` are not written.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from . import _lib, ops


class LeaveOneOutPreprocessor:
    def __init__(self, raw_path="data/raw/u.data", processed_dir="data/processed", num_negatives=99,
                 device="cuda", seed=0):
        self.raw_path = Path(raw_path)
        self.processed_dir = Path(processed_dir)
        self.num_negatives = int(num_negatives)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.NcfError("LeaveOneOutPreprocessor computes on the GPU; there is no CPU path")
        self.seed = int(seed)
        self.results = {}
        self.train_file = self.processed_dir / "u.train.rating"
        self.test_rating_file = self.processed_dir / "u.test.rating"
        self.test_negative_file = self.processed_dir / "u.test.negative"
        self.processed_dir.mkdir(parents=True, exist_ok=True)

    # -- steps, named like the reference's -------------------------------------------------------------------
    def load_and_prepare_data(self) -> torch.Tensor:
        """-> int64 [n, 3] (user, item, timestamp) on the device, in file order."""
        raw = np.fromfile(self.raw_path, dtype=np.uint8)
        vals, status = ops.text_parse_ints(torch.from_numpy(raw).to(self.device), 4, exact=False)
        if status & 1:
            raise _lib.NcfError(f"{self.raw_path}: a line holds fewer than four integers (user item rating timestamp)")
        return vals[:, [0, 1, 3]].contiguous()

    def temporal_split(self, df: torch.Tensor):
        """-> (train_data [n_train, 2], test_data [n_test, 2]) device tensors of [user, item] rows."""
        user_num = int(df[:, 0].max()) + 1
        return ops.leave_one_out_split(df[:, 0].contiguous(), df[:, 1].contiguous(), df[:, 2].contiguous(), user_num)

    def generate_test_negatives(self, train_data: torch.Tensor, test_data: torch.Tensor, num_items: int):
        """-> negatives int64 [n_test, num_negatives] (ascending; -1 pads a user that ran out of draws)."""
        allp = torch.cat([train_data, test_data])           # the user's train AND test items are excluded (:105-109)
        user_num = int(allp[:, 0].max()) + 1
        rowptr, col = ops.csr_build(allp[:, 0].contiguous(), allp[:, 1].contiguous(), user_num)
        negs, cnt = ops.eval_negatives(rowptr, col, test_data[:, 0].contiguous(), num_items, self.num_negatives, self.seed)
        short = int((cnt < self.num_negatives).sum())
        if short:
            print(f"Warning: {short} users got fewer than {self.num_negatives} negatives")
        return negs

    def verify_split(self, train_data: torch.Tensor, test_data: torch.Tensor):
        """No (user, item) pair may be both a training and a test interaction (:156-181)."""
        big = int(max(train_data[:, 1].max(), test_data[:, 1].max())) + 1
        tk = torch.unique(train_data[:, 0] * big + train_data[:, 1])
        ek = test_data[:, 0] * big + test_data[:, 1]
        pos = torch.searchsorted(tk, ek).clamp_max(tk.numel() - 1)
        leaked = int((tk[pos] == ek).sum())
        if leaked:
            raise RuntimeError("Data leakage detected in train/test split!")

    def save_splits(self, train_data, test_data, test_negatives):
        tr, te, ng = (t.cpu().numpy() for t in (train_data, test_data, test_negatives))
        np.savetxt(self.train_file, tr, fmt="%d", delimiter="\t")
        np.savetxt(self.test_rating_file, te, fmt="%d", delimiter="\t")
        lines = [f"({u},{i})\t" + "\t".join(str(x) for x in row[row >= 0]) for (u, i), row in zip(te, ng)]
        with open(self.test_negative_file, "w") as f:
            f.write("\n".join(lines))                        # no trailing newline, like the reference (:152-153)

    def run(self):
        df = self.load_and_prepare_data()
        train_data, test_data = self.temporal_split(df)
        num_users = int(train_data[:, 0].max()) + 1
        num_items = int(train_data[:, 1].max()) + 1          # build_interaction_matrix: max train id + 1 (:189-190)
        negs = self.generate_test_negatives(train_data, test_data, num_items)
        self.verify_split(train_data, test_data)
        self.save_splits(train_data, test_data, negs)
        self.results["preprocessing"] = {
            "num_users": num_users, "num_items": num_items, "total_original_interactions": int(df.shape[0]),
            "train_interactions": int(train_data.shape[0]), "test_interactions": int(test_data.shape[0]),
            "users_with_test": int(test_data.shape[0]), "test_coverage": float(test_data.shape[0] / num_users * 100),
            "sparsity": float(1 - train_data.shape[0] / (num_users * num_items)),
            "split_method": "temporal_leave_one_out"}
        return self.results["preprocessing"]
